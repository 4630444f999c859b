"""GPU parity tests of the training step: backward kernels, the whole-path gradients and the fused Adam against the
CPU oracle (autograd over the restated forward, pinned by tests/golden/grad_T16.npz made from the reference).

Tolerances: fp32 mode — gradients within 1e-4 relative (of each tensor's max-abs) of the oracle / golden values;
bf16 mode — operands of every GEMM are rounded to bf16 (relative 2^-9 per element), so gradients are compared at
3e-2 of each tensor's max-abs and 2e-2 on its norm."""
import os

import numpy as np
import pytest
import torch

import avformer_b200 as A
from oracle import avformer_oracle as O

pytestmark = pytest.mark.gpu
AF = A.functional

FP32_RTOL = 1e-4
BF16_RTOL = 3e-2


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _rel(a, b):
    b = b.double().cpu()
    return (a.double().cpu() - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def _model(seed, T, precision, dropout=None):
    m = A.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU").set_clip_length(T)
    m.load_state_dict(O.make_state_dict(seed, T), strict=True)
    m = m.cuda().set_precision(precision)
    if dropout is not None:
        m.set_dropout(dropout)
    return m


# ---------------------------------------------------------------------------------------------
# kernel level
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ta,tb,m,n,k", [(0, 1, 300, 512, 256), (0, 1, 77, 256, 768), (0, 1, 8704, 512, 1024),
                                          (1, 1, 256, 512, 300), (1, 1, 1536, 512, 4), (1, 1, 768, 256, 50176 // 8),
                                          (1, 1, 128, 256, 48), (1, 1, 512, 1024, 8704)])
def test_gemm_operand_modes_tcgen05(ta, tb, m, n, k):
    """dgrad (NN) and wgrad (TN, split-K) forms against fp64 on bf16-exact inputs."""
    torch.manual_seed(m + n + k)
    a = torch.randn((k, m) if ta else (m, k), device="cuda").bfloat16()
    b = (torch.randn((k, n) if tb else (n, k), device="cuda") / k ** 0.5).bfloat16()
    A_ = a.double().t() if ta else a.double()
    B_ = b.double() if tb else b.double().t()
    ref = A_ @ B_
    got = AF.gemm(a, b, bool(ta), bool(tb), precision="bf16")
    assert _rel(got, ref) < 3e-5
    got32 = AF.gemm(a.float(), b.float(), bool(ta), bool(tb), precision="fp32")
    assert _rel(got32, ref) < 1e-5


def test_gemm_gelu_epilogues():
    torch.manual_seed(3)
    m, n, k = 333, 512, 256
    a = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") / k ** 0.5).bfloat16()
    bias = torch.randn(n, device="cuda")
    pre = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    g = AF.gemm(a, w, bias=bias, aux=pre, flags=1 | 2 | 16, out_dtype=torch.bfloat16, precision="bf16")
    ref_pre = a.double() @ w.double().t() + bias.double()
    assert _rel(pre.float(), ref_pre) < 2 ** -8
    assert _rel(g.float(), torch.nn.functional.gelu(ref_pre, approximate="tanh")) < 2 ** -7
    # DGELU: (dy @ W2) * gelu'(pre), W2 stored [K=n2, N=n]
    n2 = 256
    dy = torch.randn(m, n2, device="cuda").bfloat16()
    w2 = (torch.randn(n2, n, device="cuda") / n2 ** 0.5).bfloat16()
    x = pre.double().requires_grad_(True)
    torch.nn.functional.gelu(x, approximate="tanh").sum().backward()
    ref = (dy.double() @ w2.double()) * x.grad
    got = AF.gemm(dy, w2, False, True, aux=pre, flags=8, out_dtype=torch.float32, precision="bf16")
    assert _rel(got, ref) < 1e-4


@pytest.mark.parametrize("rows,cols", [(1, 64), (777, 512), (50176, 256), (64, 12544), (5, 6144)])
def test_colsum(rows, cols):
    torch.manual_seed(rows)
    x = torch.randn(rows, cols, device="cuda")
    assert _rel(AF.colsum(x), x.double().sum(0)) < 1e-5
    xb = x.bfloat16()
    assert _rel(AF.colsum(xb), xb.double().sum(0)) < 1e-5


@pytest.mark.parametrize("dim", [128, 256, 512])
def test_layernorm_bwd(dim):
    torch.manual_seed(dim)
    rows = 1234
    x = (torch.randn(rows, dim, device="cuda") * 2 + 0.5)
    gamma = torch.randn(dim, device="cuda")
    dyn = torch.randn(rows, dim, device="cuda")
    dres = torch.randn(rows, dim, device="cuda")
    xr = x.double().cpu().requires_grad_(True)
    gr = gamma.double().cpu().requires_grad_(True)
    br = torch.zeros(dim, dtype=torch.float64, requires_grad=True)
    y = O.layer_norm(xr, gr, br)
    y.backward(dyn.double().cpu())
    got, xb, dg, db, dbias = AF.layernorm_bwd_(x, gamma, dyn, dres.clone(), want_bf16=True)
    assert _rel(got, dres.double().cpu() + xr.grad) < 1e-5
    assert _rel(xb.float(), dres.double().cpu() + xr.grad) < 2 ** -8
    assert _rel(dg, gr.grad) < 1e-5 and _rel(db, br.grad) < 1e-5 and _rel(dbias, dres.double().sum(0)) < 1e-5


@pytest.mark.parametrize("n_tok,dh", [(12, 32), (17, 64), (49, 32), (33, 64)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_attention_bwd(n_tok, dh, dtype):
    torch.manual_seed(n_tok * dh)
    n_seq, heads = 5, 8
    inner = heads * dh
    qkv = torch.randn(n_seq * n_tok, 3 * inner, device="cuda").to(dtype)
    dout = torch.randn(n_seq * n_tok, inner, device="cuda").to(dtype)
    q = qkv.double().cpu().requires_grad_(True)
    qq, kk, vv = (t.reshape(n_seq, n_tok, heads, dh).permute(0, 2, 1, 3) for t in q.chunk(3, dim=-1))
    att = torch.softmax(qq @ kk.transpose(-1, -2) * dh ** -0.5, dim=-1) @ vv
    att.permute(0, 2, 1, 3).reshape(n_seq * n_tok, inner).backward(dout.double().cpu())
    got = AF.attention_bwd(qkv, dout, n_seq, n_tok, heads, dh)
    assert _rel(got.float(), q.grad) < (1e-5 if dtype == torch.float32 else 2 ** -7)


@pytest.mark.parametrize("decoupled", [False, True])
def test_fused_adam_matches_reference_rule(decoupled):
    torch.manual_seed(9)
    n = 100_003
    p0, m0, v0 = torch.randn(n), torch.zeros(n), torch.zeros(n)
    p, m, v = p0.cuda(), m0.cuda(), v0.cuda()
    pr, mr, vr = p0.double(), m0.double(), v0.double()
    for step in range(1, 4):
        g = torch.randn(n) * 0.1
        AF.adam_step_(p, g.cuda(), m, v, step, 5e-4, 0.9, 0.999, 1e-8, 5e-5, decoupled=decoupled)
        pr, mr, vr = O.adam_update(pr, g.double(), mr, vr, step, 5e-4, (0.9, 0.999), 1e-8, 5e-5, decoupled)
    assert (p.double().cpu() - pr).abs().max().item() < 1e-6
    # and torch's own optimiser agrees with the restated rule (pins the oracle)
    q = torch.nn.Parameter(p0.clone())
    opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)([q], lr=5e-4, weight_decay=5e-5)
    torch.manual_seed(9)
    torch.randn(n)
    for step in range(1, 4):
        q.grad = torch.randn(n) * 0.1
        opt.step()
    assert (q.detach().double() - pr).abs().max().item() < 1e-6


# ---------------------------------------------------------------------------------------------
# whole hot path: loss.backward() against the oracle and the golden vectors from the reference
# ---------------------------------------------------------------------------------------------
def _hot_path_backward(m, stage3, frame, audio, labels, w_sf):
    for p_ in m.parameters():
        p_.grad = None
    s3 = stage3.cuda().requires_grad_(True)
    fr = frame.cuda().requires_grad_(True)
    au = audio.cuda().requires_grad_(True)
    s_out, out21 = m.hot_path_train(s3, fr, au)
    loss = m.get_au_loss(out21, labels.cuda())
    if w_sf:
        loss = loss + w_sf * (s_out * O.sformer_probe(s_out.shape).cuda()).mean()
    loss.backward()
    torch.cuda.synchronize()
    return loss.detach(), {"stage3": s3.grad, "frame_feat": fr.grad, "audio_feat": au.grad}, out21.detach()


@pytest.mark.parametrize("precision,batch_stats", [("fp32", False), ("fp32", True), ("bf16", False), ("bf16", True)])
def test_hot_path_gradients_against_oracle(precision, batch_stats):
    T, B, seed, w_sf = 16, 6, 77, 3.0
    sd = O.make_state_dict(seed, T, hot_path_only=True)
    stage3, frame, audio = O.synth_hot_path_inputs(seed, B, T)
    labels = O.synth_inputs(seed, B, T, image=8)[2]
    labels[1, 0] = -1.0                                                  # one ignored row
    ref_loss, ref_g, ref_in, ref_out = O.hot_path_grads(stage3.double(), frame.double(), audio.double(), labels.double(),
                                                        O.cast_params(sd, torch.float64), T, batch_stats, w_sf)
    m = _model(seed, T, precision, dropout=0.0)
    m.train(batch_stats)
    run0 = m.video_model.au_head.AU_BN1.running_mean.clone()
    loss, gin, out21 = _hot_path_backward(m, stage3, frame, audio, labels, w_sf)
    tol = FP32_RTOL if precision == "fp32" else BF16_RTOL
    assert abs(loss.item() - ref_loss.item()) < (1e-4 if precision == "fp32" else 2e-2) * max(1.0, abs(ref_loss.item()))
    for k, g in gin.items():
        assert _rel(g, ref_in[k]) < tol, f"d{k}: {_rel(g, ref_in[k]):.3e}"
    named = dict(m.named_parameters())
    hot = {k: rg for k, rg in ref_g.items() if not O.is_backbone_key(k)}
    # BatchNorm with batch statistics makes sum_clips d(emb) vanish identically (dx = g rstd (dy - mean dy - xhat mean(dy xhat))),
    # and the TFormer's cls rows (cls_token + pos, both N(0,1)) are nearly the same for every clip.  Every TFormer parameter
    # gradient is a sum over rows in which those dominant cls-row terms cancel (the last net.3.bias is exactly 0 analytically),
    # so what is left carries the rounding of the cancelled terms: with 6 clips the TFormer tensors are compared at 5x (the
    # tensors downstream of the 6-row batch statistics at 2x) the tolerance per tensor (vectors against the scale of the summed terms, max|d tformer_cls|), and the gradient of the WHOLE
    # model is additionally held to the plain tolerance in the L2 norm.  Eval mode (above) has no such allowance.
    summand = ref_in["tformer_cls"].abs().max().item()
    checked, num, den = 0, 0.0, 0.0
    for k, rg in hot.items():
        if rg.abs().max() == 0:
            continue
        assert named[k].grad is not None, k
        diff = named[k].grad.double().cpu() - rg
        num += float((diff * diff).sum())
        den += float((rg * rg).sum())
        in_t = batch_stats and ".t_former." in k
        floor = summand if (in_t and (rg.dim() == 1 or "cls_token" in k or "pos_embedding" in k)) else 0.0
        r = diff.abs().max().item() / max(rg.abs().max().item(), floor)
        assert r < (5 * tol if in_t else (2 * tol if batch_stats else tol)), f"{k}: rel err {r:.3e} (|ref|max {rg.abs().max().item():.3e}, floor {floor:.3e})"
        checked += 1
    assert (num / den) ** 0.5 < (tol if precision == "fp32" else 2e-2), f"whole-model gradient: relative L2 error {(num / den) ** 0.5:.3e}"
    assert checked >= 150
    # parameters the loss never reaches (per-modality AU_linear_last*, models/avformer.py:53,70) get no gradient
    assert named["video_model.au_head.AU_linear_last1.weight"].grad is None
    if batch_stats:      # running statistics moved with momentum 0.1 towards the (unbiased) batch statistics
        cls = ref_out["tformer_cls"]
        rm, rv = O.bn_running_update(cls, sd["video_model.au_head.AU_BN1.running_mean"].double(), sd["video_model.au_head.AU_BN1.running_var"].double())
        bn = m.video_model.au_head.AU_BN1
        assert _rel(bn.running_mean, rm) < (1e-4 if precision == "fp32" else 2e-2) and _rel(bn.running_var, rv) < (1e-4 if precision == "fp32" else 2e-2)
        assert int(bn.num_batches_tracked) == 1 and not torch.equal(run0, bn.running_mean)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_gradients_against_reference_golden(golden_dir, precision):
    """tests/golden/grad_T16.npz: loss.backward() of the REFERENCE modules (eval mode, fusion head + both AU_formers + TFormer)."""
    d = dict(np.load(os.path.join(golden_dir, "grad_T16.npz")))
    T, B, seed = int(d["n_frames"]), int(d["batch"]), int(d["seed"])
    stage3, frame, audio = O.synth_hot_path_inputs(seed, B, T)
    m = _model(seed, T, precision, dropout=0.0).eval()
    loss, gin, out21 = _hot_path_backward(m, stage3, frame, audio, torch.from_numpy(d["labels"]), 0.0)
    tol = FP32_RTOL if precision == "fp32" else BF16_RTOL
    assert abs(loss.item() - float(d["loss"])) < (1e-4 if precision == "fp32" else 2e-2)
    assert _rel(out21[:, :12], torch.from_numpy(d["logits"])) < (1e-4 if precision == "fp32" else 2e-2)
    named = dict(m.named_parameters())
    n_full = n_norm = 0
    for k, v in d.items():
        if k.startswith("g:"):
            r = _rel(named[k[2:]].grad, torch.from_numpy(v))
            assert r < tol, f"{k}: {r:.3e}"
            n_full += 1
        elif k.startswith("gnorm:"):
            got = named[k[6:]].grad.double().norm().item()
            assert abs(got - float(v)) <= (1e-4 if precision == "fp32" else 2e-2) * float(v), f"{k}: {got} vs {float(v)}"
            n_norm += 1
    assert n_full > 100 and n_norm > 150


def test_frozen_submodels_take_inference_kernels_and_get_no_grads():
    """Reference default (models/avformer.py:78-85): pretrained sub-models frozen, only the fusion head trains."""
    T, B, seed = 16, 4, 5
    m = _model(seed, T, "bf16", dropout=0.0).eval()
    for p_ in list(m.video_model.parameters()) + list(m.audio_model.parameters()):
        p_.requires_grad = False
    stage3, frame, audio = O.synth_hot_path_inputs(seed, B, T)
    labels = O.synth_inputs(seed, B, T, image=8)[2]
    s_out, out21 = m.hot_path_train(stage3.cuda(), frame.cuda(), audio.cuda())
    assert not s_out.requires_grad and out21.requires_grad
    m.get_au_loss(out21, labels.cuda()).backward()
    for k, p_ in m.named_parameters():
        assert (p_.grad is not None) == (k.startswith("au_head.")), k


def test_training_steps_follow_the_oracle_trajectory():
    """Three optimiser steps on the hot path (fp32 mode, FusedAdam with coupled L2 as train.py:334) against the oracle's
    own loop: autograd over the restated forward + the restated Adam rule."""
    T, B, seed, lr, wd = 8, 4, 21, 5e-4, 5e-5
    sd = O.make_state_dict(seed, T, hot_path_only=True)
    stage3, frame, audio = O.synth_hot_path_inputs(seed, B, T)
    labels = O.synth_inputs(seed, B, T, image=8)[2]
    m = _model(seed, T, "fp32", dropout=0.0).eval()
    hot = [p_ for k, p_ in m.named_parameters() if not O.is_backbone_key(k)]
    opt = A.FusedAdam(hot, lr=lr, weight_decay=wd)
    p = {k: v.double() for k, v in sd.items()}
    state = {}
    losses, ref_losses = [], []
    for step in range(1, 4):
        opt.zero_grad()
        _, out21 = m.hot_path_train(stage3.cuda(), frame.cuda(), audio.cuda())
        loss = m.get_au_loss(out21, labels.cuda())
        loss.backward()
        opt.step()
        losses.append(loss.item())
        rl, rg, _, _ = O.hot_path_grads(stage3.double(), frame.double(), audio.double(), labels.double(), p, T)
        ref_losses.append(rl.item())
        for k, g in rg.items():
            if O.is_backbone_key(k):
                continue
            ea, es = state.get(k, (torch.zeros_like(g), torch.zeros_like(g)))
            p[k], ea, es = O.adam_update(p[k], g, ea, es, step, lr, (0.9, 0.999), 1e-8, wd)
            state[k] = (ea, es)
    assert np.allclose(losses, ref_losses, rtol=2e-4, atol=2e-4), (losses, ref_losses)
    assert losses[-1] < losses[0]
    named = dict(m.named_parameters())
    worst = max(_rel(named[k], p[k]) for k in state)
    assert worst < 2e-3, worst          # Adam's sign-like first steps amplify 1e-6 gradient noise where |g| ~ eps
    # parameters without a gradient were not touched (torch.optim.Adam skips grad None)
    assert torch.equal(named["audio_model.au_head.AU_linear_last3.weight"].cpu(), sd["audio_model.au_head.AU_linear_last3.weight"])


def test_full_model_training_step_end_to_end():
    """model(x) -> get_au_loss -> backward -> FusedAdam.step on the complete drop-in module (conv backbones in torch):
    gradients reach the backbones through the SFormer / TFormer / AU_former backward kernels, the loss goes down."""
    T, B, seed = 8, 2, 33
    m = _model(seed, T, "bf16", dropout=0.0).train()
    clip, audio, labels = O.synth_inputs(seed, B, T)
    x = {"clip": clip.cuda(), "audio_features": audio.cuda(), "Index": torch.arange(B).cuda()}
    opt = A.FusedAdam(m.parameters(), lr=1e-4, weight_decay=5e-5)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        out = m(x)
        assert out.shape == (B, 21) and float(out[:, 12:].abs().max()) == 0.0
        loss = m.get_au_loss(out, labels.cuda())
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    conv_w = m.video_model.video_model.s_former.conv1.weight
    assert conv_w.grad is not None and float(conv_w.grad.abs().max()) > 0
    assert m.audio_model.audio_model.resnet.conv1.weight.grad is not None


# ---------------------------------------------------------------------------------------------
# dropout (train() mode of the audio AU_former and the fusion head: p = 0.2, models/avformer.py:48,87)
# ---------------------------------------------------------------------------------------------
def test_dropout_mask_statistics_and_determinism():
    p, seed = 0.2, 123456789
    m = AF.dropout_mask(p, seed, 1, 2, 4096, 256, "cuda")
    vals = torch.unique(m)
    assert vals.numel() == 2 and float(vals[0]) == 0.0 and abs(float(vals[1]) - 1.25) < 1e-6
    keep = (m > 0).float().mean().item()
    assert abs(keep - 0.8) < 3e-3                                   # 1M samples: 4 sigma = 1.6e-3
    assert abs((m > 0).float().mean(0).std().item() - (0.8 * 0.2 / 4096) ** 0.5) < 2e-3     # no column structure
    assert torch.equal(m, AF.dropout_mask(p, seed, 1, 2, 4096, 256, "cuda"))
    for other in (AF.dropout_mask(p, seed + 1, 1, 2, 4096, 256, "cuda"), AF.dropout_mask(p, seed, 0, 2, 4096, 256, "cuda"),
                  AF.dropout_mask(p, seed, 1, 0, 4096, 256, "cuda")):
        agree = ((other > 0) == (m > 0)).float().mean().item()
        assert abs(agree - (0.8 * 0.8 + 0.2 * 0.2)) < 5e-3          # independent masks agree 68 % of the time


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("dim,depth,mlp,n_tok", [(256, 3, 256, 12), (128, 2, 256, 12)])
def test_encoder_stack_with_dropout_against_oracle_with_the_same_masks(precision, dim, depth, mlp, n_tok):
    """Forward and backward of a whole stack in train() mode: the kernels' counter-based masks are read back through
    avf_dropout_mask and injected into the oracle, so dropout is checked exactly (values, scaling, placement, gradients)."""
    torch.manual_seed(dim + depth)
    n_seq, heads, dh, p_drop, seed = 9, 8, 32, 0.2, 987654321
    tr = A.Transformer(dim, depth, heads, dh, mlp, dropout=p_drop).cuda().train()
    tr.precision = precision
    tr.fixed_dropout_seed = seed
    with torch.no_grad():
        for q in tr.parameters():
            if q.dim() == 1:
                q.add_(torch.randn_like(q) * 0.1)
    x = torch.randn(n_seq, n_tok, dim, device="cuda", requires_grad=True)
    dy = torch.randn(n_seq, n_tok, dim, device="cuda")
    y = tr(x)
    y.backward(dy)
    R = n_seq * n_tok
    masks = {(l, s): AF.dropout_mask(p_drop, seed, l, s, R, mlp if s == 1 else dim, "cuda").double().cpu() for l in range(depth) for s in range(3)}
    pr = {"t." + k: v.detach().double().cpu().requires_grad_(True) for k, v in tr.named_parameters()}
    xr = x.detach().double().cpu().requires_grad_(True)
    yr = O.transformer(xr, pr, "t.", depth, heads, masks)
    yr.backward(dy.double().cpu())
    tol_y, tol_g = (1e-4, 1e-4) if precision == "fp32" else (2e-2, 3e-2)
    assert _rel(y, yr.detach()) < tol_y
    assert _rel(x.grad, xr.grad) < tol_g
    for k, v in tr.named_parameters():
        assert _rel(v.grad, pr["t." + k].grad) < tol_g, k
    # eval() switches dropout off (same kernels, p = 0)
    tr.eval()
    y_eval = tr(x.detach())
    assert _rel(y_eval, O.transformer(xr.detach(), {k: v.detach() for k, v in pr.items()}, "t.", depth, heads)) < tol_y


def test_train_mode_model_uses_dropout_and_stays_finite():
    T, B, seed = 16, 8, 12
    m = _model(seed, T, "bf16").train()          # reference rates: 0.2 in the audio AU_former and the fusion head
    stage3, frame, audio = O.synth_hot_path_inputs(seed, B, T)
    labels = O.synth_inputs(seed, B, T, image=8)[2]
    torch.manual_seed(0)
    _, a = m.hot_path_train(stage3.cuda(), frame.cuda(), audio.cuda())
    torch.manual_seed(0)
    _, b = m.hot_path_train(stage3.cuda(), frame.cuda(), audio.cuda())
    torch.manual_seed(1)
    _, c = m.hot_path_train(stage3.cuda(), frame.cuda(), audio.cuda())
    assert torch.equal(a, b) and not torch.equal(a, c)            # reproducible under torch.manual_seed, random otherwise
    m.get_au_loss(c, labels.cuda()).backward()
    assert all(torch.isfinite(q.grad).all() for q in m.au_head.parameters())


# ---------------------------------------------------------------------------------------------
# CUDA-graph front ends
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sm_split", [None, (100, 40), (64, 84, 64)])
def test_graphed_hot_path_equals_eager(sm_split):
    """sm_split: the persistent SFormer kernel on its own share of the SMs next to the TFormer / head chain (avf_set_sm_cap);
    results must not depend on the partition, and the cap must be restored afterwards."""
    T, B, seed = 16, 9, 61
    m = _model(seed, T, "bf16").eval()
    stage3, frame, audio = O.synth_hot_path_inputs(seed, B, T)
    dev = (stage3.bfloat16().cuda(), frame.bfloat16().cuda(), audio.cuda())
    with torch.no_grad():
        g = A.GraphedHotPath(m, *dev, sm_split=sm_split)
        assert A._lib.lib().avf_set_sm_cap(0) == 0
        for it in range(3):
            s3, fr, au = O.synth_hot_path_inputs(seed + it, B, T)
            dev = (s3.bfloat16().cuda(), fr.bfloat16().cuda(), au.cuda())
            s_ref, o_ref, d_ref = m.hot_path(*dev, want_decisions=True)
            s_out, o, d = g.replay(*dev)
            torch.cuda.synchronize()
            assert torch.equal(s_out, s_ref) and torch.equal(o, o_ref) and torch.equal(d, d_ref)


def test_graphed_hot_path_with_the_logit_gather_inside_the_graph():
    """GraphedHotPath(gather_into=...): the NCCL all-gather is a branch of the captured graph (here: a one-rank communicator on this
    GPU).  The gathered table holds this rank's logits after every replay, nothing else changes, and release() lets the process
    group go away afterwards."""
    import socket
    import torch.distributed as dist
    if dist.is_initialized():
        pytest.skip("a process group is already up in this process")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=torch.device("cuda", torch.cuda.current_device()))
    try:
        T, B, seed = 16, 6, 71
        m = _model(seed, T, "bf16", dropout=0.0).eval()
        s3, fr, au = O.synth_hot_path_inputs(seed, B, T)
        dev = (s3.bfloat16().cuda(), fr.bfloat16().cuda(), au.cuda())
        gathered = torch.full((B, 21), float("nan"), device="cuda")
        with torch.no_grad():
            g = A.GraphedHotPath(m, *dev, gather_into=gathered)
            for it in range(3):
                s3, fr, au = O.synth_hot_path_inputs(seed + it, B, T)
                dev = (s3.bfloat16().cuda(), fr.bfloat16().cuda(), au.cuda())
                s_ref, o_ref, d_ref = m.hot_path(*dev, want_decisions=True)
                gathered.fill_(float("nan"))
                s_out, o, d = g.replay(*dev)
                torch.cuda.synchronize()
                assert torch.equal(s_out, s_ref) and torch.equal(o, o_ref) and torch.equal(d, d_ref)
                assert torch.equal(gathered, o_ref)
        g.release()
        with pytest.raises(RuntimeError, match="release"):
            g.replay()

        # the same with the gather as pushes over peer memory (csrc/avf_peer.cu); one rank: its own block is the only peer
        peer = A.dp.PeerLogitGather(B, 21, timeout_s=2.0)
        o_eager = torch.randn(B, 21, device="cuda")
        peer.push(o_eager)
        peer.wait()
        assert torch.equal(peer.table(), o_eager)
        with torch.no_grad():
            g = A.GraphedHotPath(m, *dev, gather_into=peer)
            tables = []
            for it in range(4):
                s3, fr, au = O.synth_hot_path_inputs(seed + 10 + it, B, T)
                dev = (s3.bfloat16().cuda(), fr.bfloat16().cuda(), au.cuda())
                s_ref, o_ref, d_ref = m.hot_path(*dev, want_decisions=True)
                s_out, o, d = g.replay(*dev)
                torch.cuda.synchronize()
                assert torch.equal(s_out, s_ref) and torch.equal(o, o_ref) and torch.equal(d, d_ref)
                assert torch.equal(peer.table(), o_ref)
                tables.append(peer.table())
            assert tables[-1].data_ptr() != tables[-2].data_ptr() and tables[-1].data_ptr() == tables[-3].data_ptr()      # two slots, alternating
        peer.check()
        g.release()
        # a block that never arrives: the wait gives up after its timeout and reports the missing rank instead of hanging the GPU
        lone = A.dp.PeerLogitGather(B, 21, timeout_s=0.05)
        lone.wait()
        with pytest.raises(RuntimeError, match="rank 0 did not arrive"):
            lone.check()

        # the gradient all-reduce kernel on a one-rank communicator: the sum over one bucket is the bucket (ragged length: padding and
        # flag rows behind it stay out of the way), twice in a row (sequence numbers), and FusedAdam takes its bucket from it
        ar = A.dp.PeerAllReduce(100003)
        ar.grad.copy_(torch.randn(100003, device="cuda"))
        before = ar.grad.clone()
        ar.reduce_()
        ar.reduce_()
        assert ar.check() == 2 and torch.equal(ar.grad, before)
        # adam_allreduce_step: the reduction and the Adam update as ONE kernel == reduce_() then avf_adam_step, bit for bit
        n = 100000
        ar2 = A.dp.PeerAllReduce(n)
        g0 = torch.randn(n, device="cuda") * 1e-2
        state0 = [torch.randn(n, device="cuda"), torch.randn(n, device="cuda") * 1e-3, torch.rand(n, device="cuda") * 1e-4]
        hyper = dict(lr=5e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=5e-5)
        for decoupled in (False, True):
            pa, ma, va = (t.clone() for t in state0)
            sha = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
            ar2.grad.copy_(g0)
            ar2.reduce_()
            AF.adam_step_(pa, ar2.grad, ma, va, 3, hyper["lr"], hyper["beta1"], hyper["beta2"], hyper["eps"], hyper["weight_decay"], decoupled=decoupled,
                          grad_scale=1.0, shadow=sha)
            pb, mb, vb = (t.clone() for t in state0)
            shb = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
            ar2.grad.copy_(g0)
            ar2.reduce_adam_(pb, mb, vb, 3, decoupled=decoupled, shadow=shb, **hyper)
            torch.cuda.synchronize()
            assert torch.equal(pa, pb) and torch.equal(ma, mb) and torch.equal(va, vb) and torch.equal(sha, shb)
            assert not torch.equal(pb, state0[0])
        assert ar2.check() == 4
        m.train()
        params = [q for q in m.au_head.parameters() if q.requires_grad]
        opt = A.FusedAdam(params, lr=1e-3)
        assert opt.reduce_mode == "peer"
    finally:
        dist.destroy_process_group()


def test_graphed_hot_path_notices_weight_changes():
    """A captured graph points at the packed (bf16 / stacked) copies of the weights.  After optimizer.step(), load_state_dict() or an
    in-place edit the eager path re-packs and frees them; the graph has to notice and capture again instead of replaying with stale
    weights or recycled memory."""
    T, B, seed = 16, 4, 63
    m = _model(seed, T, "bf16", dropout=0.0)
    stage3, frame, audio = O.synth_hot_path_inputs(seed, B, T)
    dev = (stage3.bfloat16().cuda(), frame.bfloat16().cuda(), audio.cuda())
    labels = O.synth_inputs(seed, B, T, image=8)[2].cuda()
    m.eval()
    with torch.no_grad():
        g = A.GraphedHotPath(m, *dev)
        first = [t.clone() for t in g.replay(*dev)]
        assert g.recaptures == 0

    def check(expect_new_capture):
        m.eval()
        with torch.no_grad():
            ref = [t.clone() for t in m.hot_path(*dev, want_decisions=True)]
            n0 = g.recaptures
            got = g.replay(*dev)
            torch.cuda.synchronize()
            assert all(torch.equal(a, b) for a, b in zip(ref, got))
            assert (g.recaptures > n0) == expect_new_capture
        return ref

    check(False)
    # (1) one training step of the fusion head through FusedAdam (raw-pointer update + bf16 shadow)
    m.train()
    opt = A.FusedAdam(list(m.au_head.parameters()) + list(m.video_model.video_model.s_former.spatial_transformer.parameters()), lr=1e-2)
    for _ in range(2):
        opt.zero_grad()
        s_out, out21 = m.hot_path_train(dev[0].clone().requires_grad_(True), dev[1].float().requires_grad_(True), dev[2].clone().requires_grad_(True))
        (m.get_au_loss(out21, labels) + s_out.float().mean()).backward()
        opt.step()
    after_step = check(True)
    assert not torch.equal(after_step[1], first[1]) and not torch.equal(after_step[0], first[0])
    check(False)
    # (2) load_state_dict of different weights
    m.load_state_dict(O.make_state_dict(seed + 1, T), strict=True)
    after_load = check(True)
    assert not torch.equal(after_load[1], after_step[1])
    # (3) an in-place edit of one parameter
    with torch.no_grad():
        m.video_model.video_model.t_former.cls_token.mul_(1.5)
    check(True)
    check(False)


def test_fused_adam_checkpoint_hooks_and_per_parameter_staleness():
    """(1) FusedAdam.state_dict() / load_state_dict() carry exp_avg, exp_avg_sq and the step, so a resumed run continues the trajectory
    bit for bit; (2) a parameter with a tensor hook gets its gradient through autograd (the hook fires) instead of the direct
    accumulation into the bucket; (3) stepping some parameters does not invalidate the packed copies of frozen ones."""
    T, B, seed = 8, 6, 73
    stage3, frame, audio = O.synth_hot_path_inputs(seed, B, T)
    labels = O.synth_inputs(seed, B, T, image=8)[2].cuda()
    ins = (stage3.cuda().requires_grad_(True), frame.cuda().requires_grad_(True), audio.cuda().requires_grad_(True))

    def hot(m):
        return [q for k, q in m.named_parameters() if not O.is_backbone_key(k)]

    def step(m, opt):
        opt.zero_grad()
        _, out21 = m.hot_path_train(*ins)
        loss = m.get_au_loss(out21, labels)
        loss.backward()
        opt.step()
        return loss.item()

    m1 = _model(seed, T, "bf16", dropout=0.0).train()
    o1 = A.FusedAdam(hot(m1), lr=5e-4, weight_decay=5e-5)
    for _ in range(3):
        step(m1, o1)
    sd_model = {k: v.clone() for k, v in m1.state_dict().items()}
    sd_opt = o1.state_dict()
    assert len(sd_opt["state"]) > 100 and all("exp_avg_sq" in v for v in sd_opt["state"].values())
    ref = [step(m1, o1) for _ in range(2)]
    # resume in a fresh model + optimiser (moments restored BEFORE the flat bucket exists)
    m2 = _model(seed, T, "bf16", dropout=0.0).train()
    m2.load_state_dict(sd_model, strict=True)
    o2 = A.FusedAdam(hot(m2), lr=5e-4, weight_decay=5e-5)
    o2.load_state_dict(sd_opt)
    got = [step(m2, o2) for _ in range(2)]
    assert got == ref, (got, ref)
    n1, n2 = dict(m1.named_parameters()), dict(m2.named_parameters())
    assert all(torch.equal(n1[k], n2[k]) for k in n1 if not O.is_backbone_key(k))
    # ... and into an optimiser whose bucket already exists
    o2.load_state_dict(o1.state_dict())
    assert step(m2, o2) == step(m1, o1)

    # (2) hooks
    seen = []
    w = m1.au_head.corr_transformer.layers[0][0].fn.fn.to_qkv.weight
    h = w.register_hook(lambda g: seen.append(float(g.abs().sum())))
    before = w.detach().clone()
    step(m1, o1)
    h.remove()
    assert len(seen) == 1 and seen[0] > 0 and not torch.equal(before, w.detach())

    # (3) frozen sub-model keeps its packed weights across optimiser steps of the others
    m3 = _model(seed, T, "bf16", dropout=0.0).train()
    for q in m3.video_model.video_model.t_former.parameters():
        q.requires_grad_(False)
    o3 = A.FusedAdam([q for q in hot(m3) if q.requires_grad], lr=5e-4)
    step(m3, o3)
    packed = m3.video_model.video_model.t_former.spatial_transformer.packed()
    step(m3, o3)
    assert m3.video_model.video_model.t_former.spatial_transformer.packed() is packed
    assert m3.au_head.corr_transformer.packed() is not None


def test_programmatic_dependent_launch_does_not_change_results():
    """Every kernel is launched with the programmatic-stream-serialization attribute and waits (griddepcontrol.wait) for its
    predecessor before touching memory: results must be bit-identical to plain launches (avf_set_pdl_enabled(0))."""
    T, B, seed = 16, 5, 67
    m = _model(seed, T, "bf16").eval()
    stage3, frame, audio = O.synth_hot_path_inputs(seed, B, T)
    dev = (stage3.bfloat16().cuda(), frame.bfloat16().cuda(), audio.cuda())
    L = A._lib.lib()
    with torch.no_grad():
        old = L.avf_set_pdl_enabled(0)
        try:
            ref = [t.clone() for t in m.hot_path(*dev, want_decisions=True)]
        finally:
            L.avf_set_pdl_enabled(old)
        for _ in range(3):
            got = m.hot_path(*dev, want_decisions=True)
            torch.cuda.synchronize()
            assert all(torch.equal(a, b) for a, b in zip(ref, got))


def test_graphed_train_step_follows_eager_training():
    """Same data, same start: N graph replays + FusedAdam steps land on the parameters N eager steps land on (dropout off so
    that both runs are deterministic), and with dropout on the replays draw different masks each step."""
    T, B, seed = 8, 8, 71
    stage3, frame, audio = O.synth_hot_path_inputs(seed, B, T)
    labels = O.synth_inputs(seed, B, T, image=8)[2].cuda()
    ins = (stage3.cuda().requires_grad_(True), frame.cuda().requires_grad_(True), audio.cuda().requires_grad_(True))

    def hot(m):
        return [q for k, q in m.named_parameters() if not O.is_backbone_key(k)]

    m1 = _model(seed, T, "bf16", dropout=0.0).train()
    opt1 = A.FusedAdam(hot(m1), lr=5e-4, weight_decay=5e-5)
    eager_losses = []
    for _ in range(3 + 4):                         # GraphedTrainStep spends 3 warm-up steps before capture
        opt1.zero_grad()
        _, out21 = m1.hot_path_train(*ins)
        loss = m1.get_au_loss(out21, labels)
        loss.backward()
        opt1.step()
        eager_losses.append(loss.item())
    m2 = _model(seed, T, "bf16", dropout=0.0).train()
    opt2 = A.FusedAdam(hot(m2), lr=5e-4, weight_decay=5e-5)
    g = A.GraphedTrainStep(m2, opt2, *ins, labels)
    graph_losses = [g.step().item() for _ in range(4)]
    assert np.allclose(graph_losses, eager_losses[3:], rtol=1e-5, atol=1e-6), (graph_losses, eager_losses)
    n1, n2 = dict(m1.named_parameters()), dict(m2.named_parameters())
    assert all(torch.equal(n1[k], n2[k]) for k in n1 if not O.is_backbone_key(k))
    bn1, bn2 = m1.video_model.au_head.AU_BN1, m2.video_model.au_head.AU_BN1
    assert torch.equal(bn1.running_mean, bn2.running_mean) and int(bn2.num_batches_tracked) == 7
    # dropout: fresh masks per replay
    m3 = _model(seed, T, "bf16").train()
    opt3 = A.FusedAdam(hot(m3), lr=0.0)            # lr 0: parameters frozen, so only the masks can change the loss
    g3 = A.GraphedTrainStep(m3, opt3, *ins, labels)
    ls = [g3.step().item() for _ in range(4)]
    assert len(set(ls)) == 4 and all(np.isfinite(ls)), ls


def test_runner_reproduces_the_reference_training_loop(tmp_path):
    """runner.train(): Adam over model.parameters(), per-step model(x)/get_au_loss/backward/step on dict batches with extra keys,
    per-epoch evaluate + latest.pth/best.pth; the checkpoint has the reference's 462 keys and reloads strictly."""
    from avformer_b200 import runner
    args = runner.parse_args(["--epochs", "2", "--steps-per-epoch", "3", "--eval-steps", "1", "--batch", "4", "--frames", "8",
                              "--checkpoint-path", str(tmp_path), "--learning-rate", "2e-4"])
    out = runner.train(args)
    hist = out["history"]
    assert len(hist) == 2 and all(np.isfinite(h["train_loss"]) and np.isfinite(h["loss"]) for h in hist)
    assert 0.0 <= hist[-1]["AU:acc"] <= 1.0 and 0.0 <= hist[-1]["f1"] <= 1.0
    sd = torch.load(os.path.join(tmp_path, "latest.pth"), map_location="cpu")
    spec = {k: s for k, s, _ in O.state_dict_spec(8)}
    assert set(sd) == set(spec) and all(tuple(sd[k].shape) == tuple(spec[k]) for k in sd)
    assert os.path.exists(os.path.join(tmp_path, "best.pth"))
    m = A.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU").set_clip_length(8)
    m.load_state_dict(sd, strict=True)
