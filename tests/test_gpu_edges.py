"""Edge cases and error behaviour of the CUDA path: ragged / minimal / maximal shapes, ignored labels, and the loud failures the
boundary promises (no silent fallbacks)."""
import ctypes

import numpy as np
import pytest
import torch

import avformer_b200 as A
from oracle import avformer_oracle as O

pytestmark = pytest.mark.gpu
AF = A.functional


def _maxerr(a, b):
    return (a.double().cpu() - b.double().cpu()).abs().max().item()


def _model(seed, T, precision):
    m = A.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU").set_clip_length(T)
    m.load_state_dict(O.make_state_dict(seed, T), strict=True)
    return m.cuda().eval().set_precision(precision)


@pytest.mark.parametrize("n_frames", [1, 2, 3, 149, 297])
def test_sformer_ragged_frame_counts(n_frames):
    """1 frame (half-empty single tile), odd counts (ragged last tile), one more than a full wave of 148 CTAs x 2 frames."""
    seed = 900 + n_frames
    p = O.make_state_dict(seed, 16, hot_path_only=True)
    rng = np.random.default_rng(seed)
    stage3 = torch.from_numpy(np.maximum(rng.standard_normal((n_frames, 256, 7, 7)) * 1.7 + 0.6, 0)).float()
    ref = O.sformer_tokens(stage3.bfloat16().double(), O.cast_params(p, torch.float64), "video_model.video_model.s_former.")
    m = _model(seed, 16, "bf16")
    with torch.no_grad():
        out = m.video_model.video_model.s_former.sformer(stage3.bfloat16().cuda())
    assert out.shape == stage3.shape and out.dtype == torch.bfloat16
    assert _maxerr(out, ref) < 0.2 and (out.double().cpu() - ref).abs().mean().item() < 2e-2
    with torch.no_grad():
        m.set_precision("fp32")
        out32 = m.video_model.video_model.s_former.sformer(stage3.cuda())
    ref32 = O.sformer_tokens(stage3.double(), O.cast_params(p, torch.float64), "video_model.video_model.s_former.")
    assert _maxerr(out32, ref32) < 1e-4 * max(1.0, ref32.abs().max().item())


@pytest.mark.parametrize("n_tok", [1, 2, 16, 17, 33, 64])
def test_encoder_stack_token_count_extremes(n_tok):
    """Sequences from 1 to the maximum of 64 tokens through a dim-256 stack (fused kernel) and a dim-512 stack (kernel per op)."""
    torch.manual_seed(n_tok)
    for dim, dh, mlp in ((256, 32, 512), (512, 64, 1024)):
        tr = A.Transformer(dim, 2, 8, dh, mlp).cuda().eval()
        x = torch.randn(5, n_tok, dim, device="cuda")
        pr = {"t." + k: v.detach().double().cpu() for k, v in tr.named_parameters()}
        ref = O.transformer(x.double().cpu(), pr, "t.", 2, 8)
        with torch.no_grad():
            tr.precision = "fp32"
            assert _maxerr(tr(x), ref) < 1e-4 * max(1.0, ref.abs().max().item())
            tr.precision = "bf16"
            assert _maxerr(tr(x), ref) < 6e-2


def test_tensor_maps_are_cached_by_pointer_and_shape():
    """The second eager launch of the same GEMM / fused stack on the same buffers encodes no new CUtensorMap."""
    L = A._lib.lib()

    def stats():
        h, m = ctypes.c_uint64(), ctypes.c_uint64()
        assert L.avf_debug_tmap_cache(ctypes.byref(h), ctypes.byref(m)) == 0
        return h.value, m.value

    torch.manual_seed(3)
    tr = A.Transformer(512, 2, 8, 64, 1024).cuda().eval()
    tr.precision = "bf16"
    x = torch.randn(4, 17, 512, device="cuda")
    with torch.no_grad():
        y0 = tr(x).clone()
        _, m0 = stats()
        y1 = tr(x).clone()
        h1, m1 = stats()
        y2 = tr(x).clone()
        h2, m2 = stats()
    assert torch.equal(y0, y1) and torch.equal(y1, y2)
    assert m2 == m1 and h2 > h1, (m0, m1, m2, h1, h2)


def test_single_clip_and_all_rows_ignored():
    T, seed = 16, 31
    m = _model(seed, T, "bf16")
    stage3, frame, audio = O.synth_hot_path_inputs(seed, 1, T)
    ref = O.hot_path_forward(stage3.double(), frame.double(), audio.double(), O.cast_params(O.make_state_dict(seed, T, hot_path_only=True), torch.float64), T)
    with torch.no_grad():
        _, out21, dec = m.hot_path(stage3.cuda(), frame.cuda(), audio.cuda(), want_decisions=True)
    assert out21.shape == (1, 21) and dec.shape == (1, 12) and _maxerr(out21[:, :12], ref["logits"]) < 2e-2
    labels = torch.zeros(1, 12, device="cuda")
    labels[0, 0] = -1.0
    loss = m.get_au_loss(out21, labels)
    assert torch.isnan(loss)                       # mean over an empty selection, like the reference (models/loss.py:85-102)
    labels[0, 0] = 1.0
    assert torch.isfinite(m.get_au_loss(out21, labels))


def test_errors_are_loud():
    m = A.Transformer(256, 1, 8, 32, 512).cuda().eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"), torch.no_grad():
        m(torch.zeros(1, 12, 256))
    with pytest.raises(NotImplementedError, match="mask"), torch.no_grad():
        m(torch.zeros(1, 12, 256, device="cuda"), mask=torch.ones(1, 12, dtype=torch.bool, device="cuda"))
    with pytest.raises(TypeError):
        AF.linear_fwd(torch.zeros(4, 64, device="cuda"), torch.zeros(64, 64, device="cuda"), precision="bf16")     # fp32 operands in bf16 mode
    with pytest.raises(RuntimeError, match="multiple of 64"):
        AF.linear_fwd(torch.zeros(4, 64, device="cuda").bfloat16(), torch.zeros(40, 64, device="cuda").bfloat16(), precision="bf16")
    bad = A.Transformer(192, 1, 8, 32, 256).cuda().eval()                                                            # dim not a multiple of 128
    with pytest.raises(RuntimeError, match="multiple of 128"), torch.no_grad():
        bad(torch.zeros(1, 12, 192, device="cuda"))
    with pytest.raises(RuntimeError, match="sequences longer than 64"), torch.no_grad():
        m(torch.zeros(1, 65, 256, device="cuda"))
    # workspace contract of the C ABI: too small -> AVF_EWORKSPACE, nothing computed
    L = A._lib.lib()
    shape = AF.make_shape(4, 17, 512, 8, 64, 1024, 1)
    tr = A.Transformer(512, 1, 8, 64, 1024).cuda().eval()
    x = torch.zeros(4 * 17, 512, device="cuda")
    ws = torch.empty(1024, dtype=torch.uint8, device="cuda")
    rc = L.avf_encoder_stack_fwd(AF.AVF_BF16, ctypes.byref(shape), tr.packed().array, AF._ptr(x), 512, None, 0, AF._ptr(ws), ws.numel(), AF._stream())
    assert rc == -3 and b"workspace too small" in L.avf_last_error()
    # BatchNorm1d in train() needs more than one clip, like torch
    head = A.AU_former().cuda().train()
    with pytest.raises(RuntimeError, match="more than one row"):
        head(torch.randn(1, 512, device="cuda"))
    # an evaluation engine refuses a model in train() mode
    with pytest.raises(RuntimeError, match="eval"):
        A.InferenceEngine(A.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU").cuda().train())


@pytest.mark.parametrize("n_tok,dim,dh,mlp", [(1, 128, 32, 256), (64, 512, 64, 1024), (49, 256, 32, 512), (3, 384, 32, 192)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_encoder_stack_backward_shape_extremes(n_tok, dim, dh, mlp, precision):
    torch.manual_seed(n_tok + dim)
    n_seq = 3
    tr = A.Transformer(dim, 1, 8, dh, mlp).cuda().eval()
    tr.precision = precision
    x = torch.randn(n_seq, n_tok, dim, device="cuda", requires_grad=True)
    dy = torch.randn(n_seq, n_tok, dim, device="cuda")
    tr(x).backward(dy)
    pr = {"t." + k: v.detach().double().cpu().requires_grad_(True) for k, v in tr.named_parameters()}
    xr = x.detach().double().cpu().requires_grad_(True)
    O.transformer(xr, pr, "t.", 1, 8).backward(dy.double().cpu())
    tol = 1e-4 if precision == "fp32" else 3e-2
    rel = lambda a, b: (a.double().cpu() - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
    assert rel(x.grad, xr.grad) < tol
    for k, v in tr.named_parameters():
        assert rel(v.grad, pr["t." + k].grad) < tol, k
