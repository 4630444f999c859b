"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and against the golden vectors
made from the reference itself.  Tolerances are BASELINE.json's: fp32 mode 1e-4 relative on logits, bf16 mode
2e-2 absolute on logits, identical 0.5-threshold decisions / F1 wherever the reference logit is not within the
tolerance of 0."""
import os

import numpy as np
import pytest
import torch

import avformer_b200 as A
from oracle import avformer_oracle as O

pytestmark = pytest.mark.gpu
AF = A.functional

FP32_RTOL = 1e-4          # north_star: "within 1e-4 relative in the TF32/FP32 mode"
BF16_ATOL = 2e-2          # north_star: "within 2e-2 absolute in BF16"


def _maxerr(a, b):
    return (a.double().cpu() - b.double().cpu()).abs().max().item()


def _assert_rel(a, b, rtol):
    b = b.double().cpu()
    scale = max(1.0, b.abs().max().item())
    assert _maxerr(a, b) <= rtol * scale, f"max err {_maxerr(a, b):.3e} > {rtol} * {scale:.3g}"


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _golden(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def _model(seed, T, precision):
    m = A.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU").set_clip_length(T)
    m.load_state_dict(O.make_state_dict(seed, T), strict=True)
    return m.cuda().eval().set_precision(precision)


# ---------------------------------------------------------------------------------------------
# kernel level
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [128, 256, 512])
def test_layernorm(dim):
    torch.manual_seed(dim)
    x = torch.randn(777, dim, device="cuda") * 3 + 1
    g, b = torch.randn(dim, device="cuda"), torch.randn(dim, device="cuda")
    ref = O.layer_norm(x.double().cpu(), g.double().cpu(), b.double().cpu())
    assert _maxerr(AF.layernorm_fwd(x, g, b, "fp32"), ref) < 1e-5 * max(1, ref.abs().max().item())
    assert _maxerr(AF.layernorm_fwd(x, g, b, "bf16").float(), ref) < 2 ** -8 * ref.abs().max().item()


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (300, 768, 256), (77, 128, 512), (1000, 64, 1024), (8704, 1536, 512), (1, 256, 128)])
def test_linear_tcgen05_against_fp64(m, n, k):
    """bf16 operands are exact inputs here, so the only error is fp32 accumulation order + output rounding."""
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") / k ** 0.5).bfloat16()
    b, r = torch.randn(n, device="cuda"), torch.randn(m, n, device="cuda")
    ref = a.double() @ w.double().t()
    _assert_rel(AF.linear_fwd(a, w, precision="bf16"), ref, 2e-5)
    _assert_rel(AF.linear_fwd(a, w, b, r, precision="bf16"), ref + b.double() + r.double(), 2e-5)
    _assert_rel(AF.linear_fwd(a, w, out_dtype=torch.bfloat16, precision="bf16").float(), ref, 2 ** -8)
    gl = torch.nn.functional.gelu(ref + b.double(), approximate="tanh")
    _assert_rel(AF.linear_fwd(a, w, b, gelu=True, out_dtype=torch.bfloat16, precision="bf16").float(), gl, 2 ** -7)


@pytest.mark.parametrize("m,n,k", [(200, 768, 256), (131, 128, 512), (3, 1536, 512)])
def test_linear_fp32(m, n, k):
    torch.manual_seed(m)
    a, w = torch.randn(m, k, device="cuda"), torch.randn(n, k, device="cuda") / k ** 0.5
    b, r = torch.randn(n, device="cuda"), torch.randn(m, n, device="cuda")
    ref = a.double() @ w.double().t()
    _assert_rel(AF.linear_fwd(a, w, b, r, precision="fp32"), ref + b.double() + r.double(), 1e-5)
    gl = torch.nn.functional.gelu(ref + b.double(), approximate="tanh")
    _assert_rel(AF.linear_fwd(a, w, b, gelu=True, precision="fp32"), gl, 1e-5)


@pytest.mark.parametrize("ns,nt,h,dh", [(5, 49, 8, 32), (7, 17, 8, 64), (9, 12, 8, 32), (3, 33, 8, 64), (4, 9, 8, 64), (1, 1, 8, 32)])
def test_attention(ns, nt, h, dh):
    torch.manual_seed(nt)
    qkv = torch.randn(ns * nt, 3 * h * dh, device="cuda")

    def ref_of(t):
        q, k, v = (u.reshape(ns, nt, h, dh).permute(0, 2, 1, 3).double() for u in t.split(h * dh, -1))
        return (torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, -1) @ v).permute(0, 2, 1, 3).reshape(ns * nt, h * dh)

    _assert_rel(AF.attention_fwd(qkv, ns, nt, h, dh), ref_of(qkv), 1e-5)
    qb = qkv.bfloat16()
    _assert_rel(AF.attention_fwd(qb, ns, nt, h, dh).float(), ref_of(qb), 2 ** -8)


def test_token_glue_and_loss():
    torch.manual_seed(0)
    L = A._lib.lib()
    fm = torch.randn(5, 256, 7, 7, device="cuda")
    pos = torch.randn(49, 256, device="cuda")
    x = torch.empty(5 * 49, 256, device="cuda")
    AF.check(L.avf_sformer_tokens_pack(0, AF._ptr(fm), AF._ptr(pos), AF._ptr(x), 5, 256, 49, AF._stream()))
    assert torch.equal(x.view(5, 49, 256), fm.reshape(5, 256, 49).permute(0, 2, 1) + pos)       # bit exact: one fp32 add
    back = torch.empty_like(fm)
    AF.check(L.avf_sformer_tokens_unpack(0, AF._ptr(x), AF._ptr(back), 5, 256, 49, AF._stream()))
    assert torch.equal(back, x.view(5, 49, 256).permute(0, 2, 1).reshape(5, 256, 7, 7))
    fr = torch.randn(3 * 16, 512, device="cuda")
    cls, p2 = torch.randn(512, device="cuda"), torch.randn(17, 512, device="cuda")
    t = AF.tformer_embed(fr, cls, p2, 16)
    ref = torch.cat([cls.expand(3, 1, 512), fr.view(3, 16, 512)], 1) + p2
    assert torch.equal(t.view(3, 17, 512), ref)
    assert torch.equal(AF.tformer_cls_extract(t, 3, 17), ref[:, 0])
    logits = torch.randn(37, 21, device="cuda") * 2
    y = (torch.rand(37, 12, device="cuda") < 0.3).float()
    y[3, 0] = -1
    y[8, 5] = -1                                           # only column 0 drops a row (models/loss.py:85-86)
    pw = torch.tensor(O.AU_POS_WEIGHT, device="cuda")
    loss, nv, grad = AF.au_bce_loss(logits, y, pw, want_grad=True)
    assert int(nv.item()) == 36
    assert abs(loss.item() - O.au_loss(logits[:, :12].double().cpu(), y.double().cpu()).item()) < 1e-5
    assert _maxerr(grad, O.au_loss_grad(logits[:, :12].double().cpu(), y.double().cpu())) < 1e-7
    assert float(grad[3].abs().max()) == 0.0
    xt, w = torch.randn(37 * 12, 256, device="cuda"), torch.randn(12, 256, device="cuda")
    out, dec = AF.au_logits(xt, w, 37, want_decisions=True)
    ref = (xt.view(37, 12, 256).double() * w.double()).sum(-1)
    assert _maxerr(out[:, :12], ref) < 1e-4 and float(out[:, 12:].abs().max()) == 0.0
    sure = ref.abs() > 1e-4
    assert bool(((dec.bool() == (ref > 0)) | ~sure).all())


# ---------------------------------------------------------------------------------------------
# block level, against goldens made from the reference (tests/golden/make_golden.py)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T", [16, 8, 32])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_hot_path_blocks_vs_reference_golden(golden_dir, T, precision):
    g = _golden(golden_dir, f"hot_T{T}.npz")
    seed, B = int(g["seed"]), int(g["batch"])
    m = _model(seed, T, precision)
    stage3, frame, audio = O.synth_hot_path_inputs(seed, B, T)
    with torch.no_grad():
        vm = m.video_model.video_model
        s_out = vm.s_former.sformer(stage3.cuda())
        cls = vm.t_former(frame.cuda())
        au_v, vt = m.video_model.au_head(torch.from_numpy(g["tformer_cls"]).cuda())      # reference cls in: isolates the block
        au_a, at = m.audio_model.au_head(audio.cuda())
        logits = m.au_head(torch.cat([torch.from_numpy(g["audio_tokens"]), torch.from_numpy(g["video_tokens"])], 2).cuda())
        labels = torch.from_numpy(g["labels"]).cuda()
        loss = m.get_au_loss(torch.cat([logits, torch.zeros(B, 9, device="cuda")], 1), labels)
    ref_logits = torch.from_numpy(g["logits"])
    if precision == "fp32":
        _assert_rel(s_out[:6], torch.from_numpy(g["sformer_out_head"]), FP32_RTOL)
        _assert_rel(cls, torch.from_numpy(g["tformer_cls"]), FP32_RTOL)
        _assert_rel(vt, torch.from_numpy(g["video_tokens"]), FP32_RTOL)
        _assert_rel(at, torch.from_numpy(g["audio_tokens"]), FP32_RTOL)
        _assert_rel(au_v, torch.from_numpy(g["video_au_out"]), FP32_RTOL)
        _assert_rel(logits, ref_logits, FP32_RTOL)
        assert abs(loss.item() - float(g["loss"])) < 1e-4
        assert _maxerr(s_out.double().sum(dim=(1, 2, 3)), torch.from_numpy(g["sformer_out_sum"])) < 0.2
    else:
        # intermediate activations are O(1..30); the bar is on logits, intermediates get a proportional budget
        assert _maxerr(s_out[:6], torch.from_numpy(g["sformer_out_head"])) < 0.15
        assert _maxerr(cls, torch.from_numpy(g["tformer_cls"])) < 0.15
        assert _maxerr(vt, torch.from_numpy(g["video_tokens"])) < 0.08
        assert _maxerr(at, torch.from_numpy(g["audio_tokens"])) < 0.08
        assert _maxerr(logits, ref_logits) < BF16_ATOL
    tol = FP32_RTOL * max(1.0, ref_logits.abs().max().item()) if precision == "fp32" else BF16_ATOL
    sure = ref_logits.abs() > tol
    assert bool((((logits.cpu() > 0) == (ref_logits > 0)) | ~sure).all())


@pytest.mark.parametrize("T", [8, 16])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_whole_model_vs_reference_golden(golden_dir, T, precision):
    """BASELINE config 1 (B=2, T=8) and the native T=16: drop-in forward(x: dict) vs the reference's logits."""
    g = _golden(golden_dir, f"full_T{T}.npz")
    seed, B = int(g["seed"]), int(g["batch"])
    m = _model(seed, T, precision)
    clip, audio, labels = O.synth_inputs(seed, B, T)
    x = {"clip": clip.cuda(), "audio_features": audio.cuda(), "AU": labels.cuda(), "Index": torch.arange(B)}   # extra keys ignored
    with torch.no_grad():
        out = m(x)
        loss = m.get_au_loss(out, labels.cuda())
    assert out.shape == (B, 21) and out.dtype == torch.float32 and out.is_cuda
    assert float(out[:, 12:].abs().max()) == 0.0
    ref = torch.from_numpy(g["logits"])
    tol = 3e-4 * max(1.0, ref.abs().max().item()) if precision == "fp32" else BF16_ATOL   # fp32: conv stages add cuDNN-vs-MKL noise
    assert _maxerr(out[:, :12], ref) < tol
    sure = ref.abs() > tol
    dec = np.round(torch.sigmoid(out[:, :12]).cpu().numpy())                               # train.py:155
    assert bool(((dec == g["decisions"]) | ~sure.numpy()).all())
    if bool(sure.all()):
        acc, f1, _ = O.multilabel_acc_f1(labels.numpy(), dec, ignore_index=-1)
        assert abs(acc - float(g["acc"])) < 1e-12 and abs(f1 - float(g["f1"])) < 1e-12
    assert abs(loss.item() - float(g["loss"])) < (1e-3 if precision == "fp32" else 2e-2)


def test_larger_batch_against_oracle_and_idempotence():
    """B=24 clips, T=16 (1 176-row SFormer problem with ragged last tiles): oracle parity, run-to-run
    bit-reproducibility, and batch-composition independence (clip i's logits do not depend on its batch mates)."""
    T, B, seed = 16, 24, 77
    p = O.make_state_dict(seed, T, hot_path_only=True)
    stage3, frame, audio = O.synth_hot_path_inputs(seed, B, T)
    ref = O.hot_path_forward(stage3.double(), frame.double(), audio.double(), O.cast_params(p, torch.float64), T)
    m = _model(seed, T, "bf16")

    def run(sel):
        with torch.no_grad():
            vm = m.video_model.video_model
            fsel = (sel[:, None] * T + torch.arange(T)).reshape(-1)
            s_out = vm.s_former.sformer(stage3[fsel].cuda())
            cls = vm.t_former(frame[fsel].cuda())
            _, vt = m.video_model.au_head(cls)
            _, at = m.audio_model.au_head(audio[sel].cuda())
            return s_out, m.au_head(torch.cat([at, vt], 2))

    allc = torch.arange(B)
    s1, l1 = run(allc)
    s2, l2 = run(allc)
    assert torch.equal(s1, s2) and torch.equal(l1, l2)
    assert _maxerr(l1, ref["logits"]) < BF16_ATOL
    assert _maxerr(s1, ref["sformer_out"]) < 0.2
    _, l3 = run(torch.tensor([5, 17, 3]))
    assert _maxerr(l3, l1[[5, 17, 3]]) < 1e-5


def test_hot_path_from_host_matches_device_path():
    """The overlapped end-to-end entry (chunked H2D on a copy stream, kernels waiting per chunk) returns exactly what the
    device-resident call returns, also when called back to back on changing inputs and with ragged chunking."""
    T, B, seed = 16, 7, 41
    m = _model(seed, T, "bf16")
    out_host = torch.empty((B, 21), dtype=torch.float32).pin_memory()
    dec_host = torch.empty((B, 12), dtype=torch.int32).pin_memory()
    with torch.no_grad():
        for it in range(3):
            stage3, frame, audio = O.synth_hot_path_inputs(seed + it, B, T)
            h = (stage3.bfloat16().pin_memory(), frame.bfloat16().pin_memory(), audio.pin_memory())
            s_ref, o_ref, d_ref = m.hot_path(h[0].cuda(), h[1].cuda(), h[2].cuda(), want_decisions=True)
            s_out, o, d = m.hot_path_from_host(*h, out_host, dec_host, chunks=3)
            torch.cuda.synchronize()
            assert torch.equal(s_out, s_ref) and torch.equal(o, o_ref) and torch.equal(d, d_ref)
            assert torch.equal(out_host, o_ref.cpu()) and torch.equal(dec_host, d_ref.cpu())
        # back to back WITHOUT a synchronize in between (the two staging sets alternate; a result stays valid for one more call)
        hs, refs = [], []
        for it in range(4):
            stage3, frame, audio = O.synth_hot_path_inputs(seed + 10 + it, B, T)
            hs.append((stage3.bfloat16().pin_memory(), frame.bfloat16().pin_memory(), audio.pin_memory()))
            refs.append(m.hot_path(hs[-1][0].cuda(), hs[-1][1].cuda(), hs[-1][2].cuda(), want_decisions=True))
        torch.cuda.synchronize()
        prev = None
        for it in range(4):
            got = m.hot_path_from_host(*hs[it], chunks=5)
            if prev is not None:                       # the previous call's outputs are still intact
                assert all(torch.equal(a, b) for a, b in zip(prev, refs[it - 1]))
            prev = got
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(prev, refs[3]))


@pytest.mark.parametrize("B", [1, 2, 5, 33])
def test_eval_sweep_shape_T32_decisions_and_f1(B):
    """BASELINE config 5 (32-frame clips, batch 1..): logits within the bf16 tolerance of the oracle, and the 0.5-threshold
    decisions / accuracy / per-AU F1 (metrics/accf1.py via the oracle) identical to the reference's wherever the reference
    logit is not within the tolerance of the threshold."""
    T, seed = 32, 500 + B
    p = O.make_state_dict(seed, T, hot_path_only=True)
    stage3, frame, audio = O.synth_hot_path_inputs(seed, B, T)
    labels = O.synth_inputs(seed, B, T, image=8)[2]
    ref = O.hot_path_forward(stage3.double(), frame.double(), audio.double(), O.cast_params(p, torch.float64), T)
    m = _model(seed, T, "bf16")
    with torch.no_grad():
        _, out21, dec = m.hot_path(stage3.bfloat16().cuda(), frame.bfloat16().cuda(), audio.cuda(), want_decisions=True)
    # inputs were rounded to bf16 (the SFormer / TFormer feeds of the bench): compare against the oracle on the same rounded inputs
    ref = O.hot_path_forward(stage3.bfloat16().double(), frame.bfloat16().double(), audio.double(), O.cast_params(p, torch.float64), T)
    assert _maxerr(out21[:, :12], ref["logits"]) < BF16_ATOL
    ref_dec = O.decisions(ref["logits"])
    sure = (ref["logits"].abs() > BF16_ATOL).numpy()
    got = dec.cpu().numpy()
    assert np.array_equal(got[sure], ref_dec[sure])
    assert np.array_equal(got, (out21[:, :12].cpu().numpy() > 0).astype(got.dtype))
    if sure.all():
        a0, f0, _ = O.multilabel_acc_f1(labels.numpy(), ref_dec, ignore_index=-1)
        a1, f1, _ = O.multilabel_acc_f1(labels.numpy(), got.astype(np.int64), ignore_index=-1)
        assert a0 == a1 and f0 == f1


def test_multilabel_acc_f1_counters_match_reference_metric():
    """metrics/accf1.py:45-77 from 48 device-side counters: same accuracy / mean binary F1 as the oracle restatement (and as
    sklearn, which the reference calls), with per-entry ignore labels, an AU without positives, several updates and both entry points."""
    from sklearn.metrics import f1_score
    torch.manual_seed(4)
    metric = A.MultiLabelAccF1(ignore_index=-1)
    all_true, all_pred = [], []
    for n in (1, 37, 1000):
        logits = torch.randn(n, 21, device="cuda") * 2
        y = (torch.rand(n, 12, device="cuda") < 0.3).float()
        y[torch.rand(n, 12, device="cuda") < 0.1] = -1.0             # unlabeled entries
        y[:, 5] = torch.where(y[:, 5] == 1, torch.zeros_like(y[:, 5]), y[:, 5])     # AU 5 never positive
        logits[:, 5] = -1.0
        pred = O.decisions(logits[:, :12].cpu())
        if n == 37:
            metric.update(pred.astype(np.float32), y.cpu().numpy())   # the reference's call: rounded predictions, numpy
        else:
            metric.update_from_logits(logits, y)
        all_true.append(y.cpu().numpy())
        all_pred.append(pred)
    yt, yp = np.vstack(all_true), np.vstack(all_pred)
    acc, f1 = metric.get()
    acc0, f10, _ = O.multilabel_acc_f1(yt, yp, ignore_index=-1)
    assert abs(acc - acc0) < 1e-12 and abs(f1 - f10) < 1e-12
    sk = np.mean([f1_score(yt[:, i][yt[:, i] != -1], yp[:, i][yt[:, i] != -1], average="binary", zero_division=0) for i in range(12)])
    assert abs(f1 - sk) < 1e-12
    c = metric.confusion()
    assert c.sum() == int((yt != -1).sum()) and c[5, 0] == 0 and c[5, 1] == 0
    metric.clear()
    assert metric.confusion().sum() == 0


def test_inference_engine_matches_module_forward():
    """Whole model through InferenceEngine (BN folded into channels_last convs, one CUDA graph per shape): fp32 backbones reproduce
    model(x) up to the re-association of the fold; bf16 backbones (opt-in) stay within 0.25 on the logits and agree on every decision
    whose fp32 logit is farther than that from the threshold; replays track changing inputs."""
    T, B, seed = 8, 3, 208
    m = _model(seed, T, "bf16")
    eng32 = A.InferenceEngine(m, backbone_dtype=torch.float32)
    eng16 = A.InferenceEngine(m, backbone_dtype=torch.bfloat16)
    for it in range(2):
        clip, audio, _ = O.synth_inputs(seed + it, B, T)
        x = {"clip": clip.cuda(), "audio_features": audio.cuda(), "Index": torch.arange(B).cuda()}
        with torch.no_grad():
            ref = m(x)
        o32 = eng32(x).clone()
        o16 = eng16(x).clone()
        assert o32.shape == (B, 21) and float(o32[:, 12:].abs().max()) == 0.0
        assert _maxerr(o32, ref) < 5e-3, _maxerr(o32, ref)
        assert _maxerr(o16, ref) < 0.25, _maxerr(o16, ref)
        sure = (ref[:, :12].abs() > 0.25)
        assert bool((((o16[:, :12] > 0) == (ref[:, :12] > 0)) | ~sure).all())
    # new weights: the folded convolutions and the captured graphs belong to the old ones and have to be rebuilt
    m.load_state_dict(O.make_state_dict(seed + 7, T), strict=True)
    with torch.no_grad():
        ref2 = m(x)
    o32b = eng32(x).clone()
    assert eng32.refolds == 1 and _maxerr(o32b, ref2) < 5e-3 and _maxerr(o32b, o32) > 1e-2
