"""CPU oracle for the AVFormer transformer hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (torch CPU ops, fp64 for checking or fp32
for timing) of the reference's transformer-encoder path.  It is *not* product
code: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs may import it.  The product (package
``multi-modal-multi-label-facial-action-unit-detection-with-transformer_b200``)
never touches it and has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the reference itself, generated in
the build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference`` through four import shims) and committed as
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` replays them.

The arithmetic of the reference lives in third-party PyTorch
(nn.Linear/LayerNorm/BatchNorm/einsum/softmax/BCEWithLogits, README.md:13-19 pins
torch 1.6) and einops; this restatement therefore uses the same primitive
library ops on CPU, but is written function-by-function against the reference
lines cited in each docstring (paths relative to /root/reference).

Weights are a flat ``dict[str, Tensor]`` keyed by the reference's own
state-dict names (462 entries for ``TwoStreamAuralVisualFormer``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

P = Dict[str, torch.Tensor]

AU_POS_WEIGHT = (1., 1., 1., 1., 1., 1., 1., 3., 3., 3., 1., 2.)   # models/loss.py:73

# (prefix-relative) geometry of the five encoder stacks, SURVEY.md appendix B
STACKS = {
    "sformer": dict(dim=256, depth=1, heads=8, dim_head=32, mlp=512, tokens=49),
    "tformer": dict(dim=512, depth=3, heads=8, dim_head=64, mlp=1024, tokens=None),
    "au_former": dict(dim=128, depth=2, heads=8, dim_head=32, mlp=256, tokens=12),
    "fusion": dict(dim=256, depth=3, heads=8, dim_head=32, mlp=256, tokens=12),
}


# --------------------------------------------------------------------------
# a1..a5: encoder block primitives
# --------------------------------------------------------------------------
def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    """models/heads.py:164-166 — 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))."""
    c = math.sqrt(2.0 / math.pi)
    return 0.5 * x * (1.0 + torch.tanh(c * (x + 0.044715 * x * x * x)))


def layer_norm(x, g, b, eps: float = 1e-5):
    """nn.LayerNorm(dim) inside PreNorm, models/heads.py:178-185 (biased variance, eps 1e-5)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def attention(x, p: P, pre: str, heads: int, drop=None):
    """models/heads.py:203-239.  x [B,N,D]; to_qkv has no bias, its output columns are
    [q | k | v], each split head-major '(h d)' (:221-222); scale dh**-0.5 (:210,:224);
    softmax over keys (:234); heads merged 'b h n d -> b n (h d)' (:237); to_out.0 with bias, then
    Dropout (:216; ``drop`` is a callable applying an explicit scaled keep-mask, identity when None)."""
    B, N, _ = x.shape
    wqkv = p[pre + "fn.fn.to_qkv.weight"]
    inner = wqkv.shape[0] // 3
    dh = inner // heads
    qkv = x @ wqkv.t()
    q, k, v = (t.reshape(B, N, heads, dh).permute(0, 2, 1, 3) for t in qkv.split(inner, dim=-1))
    dots = torch.matmul(q, k.transpose(-1, -2)) * (dh ** -0.5)
    attn = torch.softmax(dots, dim=-1)
    out = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(B, N, inner)
    y = out @ p[pre + "fn.fn.to_out.0.weight"].t() + p[pre + "fn.fn.to_out.0.bias"]
    return y if drop is None else drop(0, y)


def feed_forward(x, p: P, pre: str, drop=None):
    """models/heads.py:188-200 — Linear(D,M)+b, GELU, Dropout, Linear(M,D)+b, Dropout (identity in eval / when
    ``drop`` is None; otherwise drop(site, tensor) with site 1 after the GELU and 2 after net.3)."""
    h = gelu_tanh(x @ p[pre + "fn.fn.net.0.weight"].t() + p[pre + "fn.fn.net.0.bias"])
    if drop is not None:
        h = drop(1, h)
    y = h @ p[pre + "fn.fn.net.3.weight"].t() + p[pre + "fn.fn.net.3.bias"]
    return y if drop is None else drop(2, y)


def encoder_layer(x, p: P, pre: str, heads: int, drop=None):
    """One (Residual(PreNorm(Attention)), Residual(PreNorm(FeedForward))) pair,
    models/heads.py:169-185,246-250.  ``pre`` ends in 'layers.L.'."""
    a = pre + "0."
    x = x + attention(layer_norm(x, p[a + "fn.norm.weight"], p[a + "fn.norm.bias"]), p, a, heads, drop)
    f = pre + "1."
    x = x + feed_forward(layer_norm(x, p[f + "fn.norm.weight"], p[f + "fn.norm.bias"]), p, f, drop)
    return x


def transformer(x, p: P, pre: str, depth: int, heads: int, masks=None):
    """models/heads.py:242-256 — depth x encoder_layer, no final norm.  ``pre`` ends in
    'spatial_transformer.' or 'corr_transformer.'.  ``masks[(layer, site)]`` are explicit scaled keep-masks
    (0 or 1/(1-p), shaped like the tensor flattened to [B*N, width]) for train()-mode dropout; the reference draws
    them from torch's Philox stream, which cannot be reproduced bit-for-bit, so tests inject the masks instead."""
    for l in range(depth):
        drop = None
        if masks is not None:
            drop = (lambda site, t, l=l: t * masks[(l, site)].reshape(t.shape).to(t.dtype))
        x = encoder_layer(x, p, f"{pre}layers.{l}.", heads, drop)
    return x


# --------------------------------------------------------------------------
# a6: SFormer token region of ResFormer.forward
# --------------------------------------------------------------------------
def sformer_tokens(fmap, p: P, pre: str):
    """models/vformer.py:245-259.  fmap [F,256,7,7] (stage-3 map) -> same shape.
    token n = 7*h + w, + pos_embedding[1,49,256], 1 encoder layer, transposed back."""
    Fn, C, H, W = fmap.shape
    x = fmap.reshape(Fn, C, H * W).permute(0, 2, 1)
    x = x + p[pre + "pos_embedding"][:, : H * W]
    x = transformer(x, p, pre + "spatial_transformer.", STACKS["sformer"]["depth"], STACKS["sformer"]["heads"])
    return x.permute(0, 2, 1).reshape(Fn, C, H, W)


# --------------------------------------------------------------------------
# a7: TFormer
# --------------------------------------------------------------------------
def tformer(frames, p: P, pre: str, n_frames: int):
    """models/vformer.py:279-293.  frames [B*T,512] -> view [B,T,512]; cls token prepended;
    + pos_embedding[1,T+1,512]; 3 layers; returns row 0 -> [B,512]."""
    x = frames.reshape(-1, n_frames, frames.shape[-1])
    B = x.shape[0]
    cls = p[pre + "cls_token"].expand(B, -1, -1)
    x = torch.cat([cls, x], dim=1) + p[pre + "pos_embedding"][:, : n_frames + 1]
    x = transformer(x, p, pre + "spatial_transformer.", STACKS["tformer"]["depth"], STACKS["tformer"]["heads"])
    return x[:, 0]


# --------------------------------------------------------------------------
# a8: AU_former
# --------------------------------------------------------------------------
def au_former(emb, p: P, pre: str, batch_stats: bool = False):
    """models/heads.py:291-339.  emb [B,512] -> BatchNorm1d (:293; running stats in eval,
    batch stats when ``batch_stats``) -> 12 x Linear(512,128)+b, token i = AU_linear_p{i+1}
    (:294-319) -> + pos (:323) -> 2 encoder layers (:324).  Returns (AU_out[B,12], tokens[B,12,128])
    like the reference; avformer discards AU_out (models/avformer.py:53,70)."""
    g, b = p[pre + "AU_BN1.weight"], p[pre + "AU_BN1.bias"]
    if batch_stats:
        mu, var = emb.mean(0), emb.var(0, unbiased=False)
    else:
        mu, var = p[pre + "AU_BN1.running_mean"], p[pre + "AU_BN1.running_var"]
    e = (emb - mu) / torch.sqrt(var + 1e-5) * g + b
    toks = [e @ p[f"{pre}AU_linear_p{i}.weight"].t() + p[f"{pre}AU_linear_p{i}.bias"] for i in range(1, 13)]
    x = torch.stack(toks, dim=1) + p[pre + "pos_embedding"][:, :12]
    x = transformer(x, p, pre + "corr_transformer.", STACKS["au_former"]["depth"], STACKS["au_former"]["heads"])
    au_out = torch.stack([(x[:, i] * p[f"{pre}AU_linear_last{i + 1}.weight"][0]).sum(-1) for i in range(12)], dim=1)
    return au_out, x


# --------------------------------------------------------------------------
# a9: fusion head
# --------------------------------------------------------------------------
def fusion_head(tokens, p: P, pre: str):
    """models/tformer.py:381-403 (aliased as former_AU_head, models/avformer.py:19,87).
    tokens [B,12,256] + pos -> 3 encoder layers -> logit_i = <x[:,i,:], AU_linear_last{i+1}.weight>."""
    B = tokens.shape[0]
    x = tokens.reshape(B, 12, -1) + p[pre + "pos_embedding"][:, :12]
    x = transformer(x, p, pre + "corr_transformer.", STACKS["fusion"]["depth"], STACKS["fusion"]["heads"])
    return torch.stack([(x[:, i] * p[f"{pre}AU_linear_last{i + 1}.weight"][0]).sum(-1) for i in range(12)], dim=1)


# --------------------------------------------------------------------------
# a11 / a12: loss, decision rule, metric
# --------------------------------------------------------------------------
def au_loss(logits, y_true, ignore: float = -1.0):
    """models/loss.py:75-103.  Rows whose FIRST label equals ``ignore`` are dropped (:85-88);
    BCE-with-logits, pos_weight AU_POS_WEIGHT on the positive term, mean over rows*12 (:102)."""
    keep = y_true[:, 0] != ignore
    x, y = logits[keep], y_true[keep]
    w = torch.tensor(AU_POS_WEIGHT, dtype=x.dtype)
    loss = -(w * y * F.logsigmoid(x) + (1.0 - y) * F.logsigmoid(-x))
    return loss.mean()


def au_loss_grad(logits, y_true, ignore: float = -1.0):
    """d au_loss / d logits, closed form: (sigma(x) (1 + (w-1) y) - w y) / (12 N_valid); 0 on dropped rows."""
    keep = (y_true[:, 0] != ignore)
    w = torch.tensor(AU_POS_WEIGHT, dtype=logits.dtype)
    s = torch.sigmoid(logits)
    g = (s * (1.0 + (w - 1.0) * y_true) - w * y_true) / (12.0 * keep.sum().clamp(min=1))
    return g * keep[:, None].to(g.dtype)


def decisions(logits) -> np.ndarray:
    """train.py:155 / test_aff2.py:112-113 — np.round(sigmoid(logit)) (half-to-even => logit > 0)."""
    return np.round(torch.sigmoid(logits.double()).numpy()).astype(np.int64)


def multilabel_acc_f1(y_true: np.ndarray, y_pred: np.ndarray, ignore_index=None):
    """metrics/accf1.py:45-77 — per-AU binary F1 (positive class 1; 0 when there is no TP/FP/FN,
    sklearn's zero_division default) averaged over the 12 AUs, and accuracy over labelled entries."""
    y_true, y_pred = np.asarray(y_true), np.asarray(y_pred)
    f1s, correct, labelled = [], 0, 0
    for i in range(y_pred.shape[1]):
        t, q = y_true[:, i], y_pred[:, i]
        if ignore_index is not None:
            m = t != ignore_index
            t, q = t[m], q[m]
        tp = int(np.sum((t == 1) & (q == 1)))
        fp = int(np.sum((t != 1) & (q == 1)))
        fn = int(np.sum((t == 1) & (q != 1)))
        f1s.append(0.0 if 2 * tp + fp + fn == 0 else 2.0 * tp / (2 * tp + fp + fn))
        correct += int(np.sum(t == q))
        labelled += t.size
    return correct / max(labelled, 1), float(np.mean(f1s)), f1s


# --------------------------------------------------------------------------
# Conv backbones (outside the hot path; needed only for the whole-model oracle)
# --------------------------------------------------------------------------
def _bn2d(x, p: P, pre: str):
    return F.batch_norm(x, p[pre + "running_mean"], p[pre + "running_var"], p[pre + "weight"], p[pre + "bias"],
                        training=False, eps=1e-5)


def _basic_block(x, p: P, pre: str, stride: int):
    """ResNet BasicBlock, models/vformer.py:128-165 (== torchvision)."""
    idt = x
    y = F.relu(_bn2d(F.conv2d(x, p[pre + "conv1.weight"], stride=stride, padding=1), p, pre + "bn1."))
    y = _bn2d(F.conv2d(y, p[pre + "conv2.weight"], padding=1), p, pre + "bn2.")
    if (pre + "downsample.0.weight") in p:
        idt = _bn2d(F.conv2d(x, p[pre + "downsample.0.weight"], stride=stride), p, pre + "downsample.1.")
    return F.relu(y + idt)


def _res_stage(x, p: P, pre: str, stride: int):
    return _basic_block(_basic_block(x, p, pre + "0.", stride), p, pre + "1.", 1)


def resnet_to_stage3(img, p: P, pre: str):
    """models/vformer.py:238-244 — conv7x7/2, BN, ReLU, maxpool3/2, layer1..3 -> [F,256,7,7] at 112x112."""
    x = F.relu(_bn2d(F.conv2d(img, p[pre + "conv1.weight"], stride=2, padding=3), p, pre + "bn1."))
    x = F.max_pool2d(x, 3, 2, 1)
    x = _res_stage(x, p, pre + "layer1.", 1)
    x = _res_stage(x, p, pre + "layer2.", 2)
    return _res_stage(x, p, pre + "layer3.", 2)


def resnet_stage4_pool(x, p: P, pre: str):
    """models/vformer.py:261-265 — layer4, global average pool, flatten -> [F,512]."""
    return _res_stage(x, p, pre + "layer4.", 2).mean(dim=(2, 3))


def audio_backbone(mel, p: P, pre: str):
    """models/audio.py:22-39 — torchvision resnet18 with 1-channel conv1 and fc = identity."""
    return resnet_stage4_pool(resnet_to_stage3(mel, p, pre), p, pre)


def avformer_forward(clip, audio_features, p: P, batch_stats: bool = False) -> Dict[str, torch.Tensor]:
    """models/avformer.py:93-106 with VideoModel.forward (models/vformer.py:303-311).  Returns the
    hot-path boundary tensors as well as the [B,21] output (logits in [:, :12], zeros elsewhere)."""
    B, _, T = clip.shape[:3]
    out: Dict[str, torch.Tensor] = {}
    vs = "video_model.video_model.s_former."
    frames = clip[:, -3:].permute(0, 2, 1, 3, 4).reshape(B * T, 3, clip.shape[3], clip.shape[4])
    out["stage3"] = resnet_to_stage3(frames, p, vs)
    out["sformer_out"] = sformer_tokens(out["stage3"], p, vs)
    out["frame_feat"] = resnet_stage4_pool(out["sformer_out"], p, vs)
    out["tformer_cls"] = tformer(out["frame_feat"], p, "video_model.video_model.t_former.", T)
    _, out["video_tokens"] = au_former(out["tformer_cls"], p, "video_model.au_head.", batch_stats)
    out["audio_feat"] = audio_backbone(audio_features, p, "audio_model.audio_model.resnet.")
    _, out["audio_tokens"] = au_former(out["audio_feat"], p, "audio_model.au_head.", batch_stats)
    out["fused_tokens"] = torch.cat([out["audio_tokens"], out["video_tokens"]], dim=2)   # avformer.py:100
    out["logits"] = fusion_head(out["fused_tokens"], p, "au_head.")
    y = torch.zeros(B, 21, dtype=out["logits"].dtype)
    y[:, :12] = out["logits"]
    out["output"] = y
    return out


def hot_path_forward(stage3, frame_feat, audio_feat, p: P, n_frames: int) -> Dict[str, torch.Tensor]:
    """The transformer stack alone (what bench.py times): SFormer on given stage-3 maps, then
    TFormer / AU_former x2 / fusion head on given frame and audio features (the conv stages that
    sit between them in the real model are outside the hot path and are bypassed here)."""
    out = {"sformer_out": sformer_tokens(stage3, p, "video_model.video_model.s_former.")}
    out["tformer_cls"] = tformer(frame_feat, p, "video_model.video_model.t_former.", n_frames)
    _, out["video_tokens"] = au_former(out["tformer_cls"], p, "video_model.au_head.")
    _, out["audio_tokens"] = au_former(audio_feat, p, "audio_model.au_head.")
    out["logits"] = fusion_head(torch.cat([out["audio_tokens"], out["video_tokens"]], dim=2), p, "au_head.")
    return out


# --------------------------------------------------------------------------
# Training step of the hot path (train.py:206-236): loss, gradients, Adam
# --------------------------------------------------------------------------
def hot_path_forward_train(stage3, frame_feat, audio_feat, p: P, n_frames: int, batch_stats: bool = False) -> Dict[str, torch.Tensor]:
    """hot_path_forward with the AU_BN1 layers in train() mode when ``batch_stats`` (nn.BatchNorm1d: biased batch
    variance for normalisation, models/heads.py:263,293)."""
    out = {"sformer_out": sformer_tokens(stage3, p, "video_model.video_model.s_former.")}
    out["tformer_cls"] = tformer(frame_feat, p, "video_model.video_model.t_former.", n_frames)
    _, out["video_tokens"] = au_former(out["tformer_cls"], p, "video_model.au_head.", batch_stats)
    _, out["audio_tokens"] = au_former(audio_feat, p, "audio_model.au_head.", batch_stats)
    out["logits"] = fusion_head(torch.cat([out["audio_tokens"], out["video_tokens"]], dim=2), p, "au_head.")
    return out


def hot_path_grads(stage3, frame_feat, audio_feat, labels, p: P, n_frames: int, batch_stats: bool = False, sformer_loss_weight: float = 0.0):
    """loss.backward() of train.py:235 on the hot path: AULoss of the fusion-head logits (models/avformer.py:114-117)
    differentiated by torch autograd THROUGH THE RESTATED FORWARD above (the reference does exactly that with its own
    modules; pinned by tests/golden/grad_T16.npz).  In the full model the SFormer output reaches the loss through conv
    stage 4; on the isolated hot path a synthetic term  sformer_loss_weight * mean(sformer_out * probe)  stands in for it
    (probe = a fixed pseudo-random +-1 pattern) so that the SFormer backward is exercised too.
    Returns (loss, grads by state-dict name, input grads dict)."""
    q = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v) for k, v in p.items()}
    ins = {"stage3": stage3.clone().requires_grad_(True), "frame_feat": frame_feat.clone().requires_grad_(True),
           "audio_feat": audio_feat.clone().requires_grad_(True)}
    out = hot_path_forward_train(ins["stage3"], ins["frame_feat"], ins["audio_feat"], q, n_frames, batch_stats)
    out["tformer_cls"].retain_grad()
    loss = au_loss(out["logits"], labels)
    if sformer_loss_weight != 0.0:
        loss = loss + sformer_loss_weight * (out["sformer_out"] * sformer_probe(out["sformer_out"].shape, out["sformer_out"].dtype)).mean()
    loss.backward()
    grads = {k: v.grad for k, v in q.items() if torch.is_tensor(v) and v.requires_grad and v.grad is not None}
    gin = {k: v.grad for k, v in ins.items()}
    gin["tformer_cls"] = out["tformer_cls"].grad          # gradient entering the TFormer's cls rows (scale of the per-clip terms)
    return loss.detach(), grads, gin, {k: v.detach() for k, v in out.items()}


def sformer_probe(shape, dtype=torch.float32):
    """Fixed +-1 pattern used as d(loss)/d(sformer_out) direction in hot-path training tests."""
    n = int(np.prod(shape))
    idx = np.arange(n, dtype=np.int64)
    return torch.from_numpy((((idx * 2654435761) >> 7) & 1).astype(np.float64) * 2.0 - 1.0).reshape(shape).to(dtype)


def bn_running_update(emb, run_mean, run_var, momentum: float = 0.1):
    """nn.BatchNorm1d in train(): running = (1-m) running + m batch, with the UNBIASED batch variance."""
    return ((1 - momentum) * run_mean + momentum * emb.mean(0), (1 - momentum) * run_var + momentum * emb.var(0, unbiased=True))


def adam_update(param, grad, exp_avg, exp_avg_sq, step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                decoupled: bool = False):
    """One step of torch.optim.Adam as train.py:334 configures it (weight decay added to the gradient), or AdamW when
    ``decoupled``.  Returns the new (param, exp_avg, exp_avg_sq)."""
    b1, b2 = betas
    if decoupled:
        param = param * (1.0 - lr * weight_decay)
    else:
        grad = grad + weight_decay * param
    exp_avg = b1 * exp_avg + (1 - b1) * grad
    exp_avg_sq = b2 * exp_avg_sq + (1 - b2) * grad * grad
    denom = exp_avg_sq.sqrt() / math.sqrt(1 - b2 ** step) + eps
    return param - (lr / (1 - b1 ** step)) * exp_avg / denom, exp_avg, exp_avg_sq


# --------------------------------------------------------------------------
# Deterministic synthetic weights (numpy PCG64: identical in the build container and on the GPU box)
# --------------------------------------------------------------------------
def _encoder_spec(pre: str, dim: int, depth: int, inner: int, mlp: int):
    for l in range(depth):
        a, f = f"{pre}layers.{l}.0.fn.", f"{pre}layers.{l}.1.fn."
        yield a + "norm.weight", (dim,), "gamma"
        yield a + "norm.bias", (dim,), "beta"
        yield a + "fn.to_qkv.weight", (3 * inner, dim), "linear_w"
        yield a + "fn.to_out.0.weight", (dim, inner), "linear_w"
        yield a + "fn.to_out.0.bias", (dim,), ("linear_b", inner)
        yield f + "norm.weight", (dim,), "gamma"
        yield f + "norm.bias", (dim,), "beta"
        yield f + "fn.net.0.weight", (mlp, dim), "linear_w"
        yield f + "fn.net.0.bias", (mlp,), ("linear_b", dim)
        yield f + "fn.net.3.weight", (dim, mlp), "linear_w"
        yield f + "fn.net.3.bias", (dim,), ("linear_b", mlp)


def _bn_spec(pre: str, c: int):
    yield pre + "weight", (c,), "gamma"
    yield pre + "bias", (c,), "beta"
    yield pre + "running_mean", (c,), "beta"
    yield pre + "running_var", (c,), "var"
    yield pre + "num_batches_tracked", (), "count"


def _resnet_spec(pre: str, in_ch: int):
    yield pre + "conv1.weight", (64, in_ch, 7, 7), "conv"
    yield from _bn_spec(pre + "bn1.", 64)
    cin = 64
    for li, c in enumerate((64, 128, 256, 512), start=1):
        for b in range(2):
            bp = f"{pre}layer{li}.{b}."
            yield bp + "conv1.weight", (c, cin if b == 0 else c, 3, 3), "conv"
            yield from _bn_spec(bp + "bn1.", c)
            yield bp + "conv2.weight", (c, c, 3, 3), "conv"
            yield from _bn_spec(bp + "bn2.", c)
            if b == 0 and li > 1:
                yield bp + "downsample.0.weight", (c, cin, 1, 1), "conv"
                yield from _bn_spec(bp + "downsample.1.", c)
        cin = c


def _au_former_spec(pre: str):
    yield pre + "pos_embedding", (1, 12, 128), "normal"
    yield from _bn_spec(pre + "AU_BN1.", 512)
    for i in range(1, 13):
        yield f"{pre}AU_linear_p{i}.weight", (128, 512), "linear_w"
        yield f"{pre}AU_linear_p{i}.bias", (128,), ("linear_b", 512)
    yield from _encoder_spec(pre + "corr_transformer.", 128, 2, 256, 256)
    for i in range(1, 13):
        yield f"{pre}AU_linear_last{i}.weight", (1, 128), "linear_w"


def state_dict_spec(n_frames: int = 16):
    """(key, shape, kind) for all 462 entries of TwoStreamAuralVisualFormer.state_dict()
    (SURVEY.md §8b); ``n_frames`` sizes t_former.pos_embedding (models/vformer.py:271-276)."""
    yield from _resnet_spec("audio_model.audio_model.resnet.", 1)
    yield from _au_former_spec("audio_model.au_head.")
    s = "video_model.video_model.s_former."
    yield s + "pos_embedding", (1, 49, 256), "normal"
    yield from _resnet_spec(s, 3)
    yield from _encoder_spec(s + "spatial_transformer.", 256, 1, 256, 512)
    t = "video_model.video_model.t_former."
    yield t + "cls_token", (1, 1, 512), "normal"
    yield t + "pos_embedding", (1, n_frames + 1, 512), "normal"
    yield from _encoder_spec(t + "spatial_transformer.", 512, 3, 512, 1024)
    yield from _au_former_spec("video_model.au_head.")
    yield "au_head.pos_embedding", (1, 12, 256), "normal"
    yield from _encoder_spec("au_head.corr_transformer.", 256, 3, 256, 256)
    for i in range(1, 13):
        yield f"au_head.AU_linear_last{i}.weight", (1, 256), "linear_w"
    yield "loss_AU.loss_fn.pos_weight", (12,), "pos_weight"


# --------------------------------------------------------------------------
# the other instantiations of the same block
# --------------------------------------------------------------------------
def variant_spec(name: str, n_frames: int = 16):
    """"tformer1536": TFormer(dim=128*12) of models/tformer.py:301 (depth 3, 8 x 64, mlp 1024);  "sformer512": the spatial transformer
    of VGGFormer, models/vggformer.py:252-258 (49 tokens, dim 512, depth 1, 8 x 32, mlp 512);  "va_former": models/heads.py:341-353."""
    if name == "tformer1536":
        yield "cls_token", (1, 1, 1536), "normal"
        yield "pos_embedding", (1, n_frames + 1, 1536), "normal"
        yield from _encoder_spec("spatial_transformer.", 1536, 3, 512, 1024)
    elif name == "sformer512":
        yield "pos_embedding", (1, 49, 512), "normal"
        yield from _encoder_spec("spatial_transformer.", 512, 1, 256, 512)
    elif name == "va_former":
        yield "pos_embedding", (1, 2, 128), "normal"
        yield from _bn_spec("VA_BN1.", 512)
        for i in (1, 2):
            yield f"VA_linear_p{i}.weight", (128, 512), "linear_w"
            yield f"VA_linear_p{i}.bias", (128,), ("linear_b", 512)
        yield from _encoder_spec("corr_transformer.", 128, 2, 256, 128)
        for i in (1, 2):
            yield f"VA_linear_last{i}.weight", (1, 128), "linear_w"
    else:  # pragma: no cover
        raise ValueError(name)


def make_variant_params(name: str, seed: int, n_frames: int = 16, dtype=torch.float32) -> P:
    return params_from_spec(variant_spec(name, n_frames), seed, dtype)


def va_former(emb, p: P, pre: str = ""):
    """models/heads.py:354-372.  emb [B,512] -> BatchNorm1d (running statistics) -> 2 x Linear(512,128)+b, token i = VA_linear_p{i+1}
    -> + pos -> 2 encoder layers (128, 8 x 32, mlp 128) -> VA_linear_last{i+1} on token i.  Returns (VA_out [B,2], tokens [B,2,128])."""
    g, b = p[pre + "VA_BN1.weight"], p[pre + "VA_BN1.bias"]
    mu, var = p[pre + "VA_BN1.running_mean"], p[pre + "VA_BN1.running_var"]
    e = (emb - mu) / torch.sqrt(var + 1e-5) * g + b
    toks = [e @ p[f"{pre}VA_linear_p{i}.weight"].t() + p[f"{pre}VA_linear_p{i}.bias"] for i in (1, 2)]
    x = torch.stack(toks, dim=1) + p[pre + "pos_embedding"][:, :2]
    x = transformer(x, p, pre + "corr_transformer.", 2, 8)
    out = torch.stack([(x[:, i] * p[f"{pre}VA_linear_last{i + 1}.weight"][0]).sum(-1) for i in range(2)], dim=1)
    return out, x


def is_backbone_key(key: str) -> bool:
    """True for the conv-backbone entries (outside the hot path): the audio ResNet18 and the conv
    stages of ResFormer (conv1, bn1, layer1..4)."""
    if key.startswith("audio_model.audio_model.resnet."):
        return True
    s = "video_model.video_model.s_former."
    return key.startswith(s) and key[len(s):].split(".")[0] in ("conv1", "bn1", "layer1", "layer2", "layer3", "layer4")


def make_state_dict(seed: int = 0, n_frames: int = 16, dtype=torch.float32, hot_path_only: bool = False) -> P:
    """Synthetic weights with the reference's default-init *distributions* (Linear: U(+-fan_in^-0.5),
    pos/cls: N(0,1), conv: Kaiming-normal fan_out — SURVEY.md §8c) but non-trivial LayerNorm/BatchNorm
    affine terms and running statistics so that no fused term can hide behind an identity.
    One numpy Generator per key (seeded by (seed, index)) so a subset reproduces the same values."""
    return params_from_spec(state_dict_spec(n_frames), seed, dtype, skip=is_backbone_key if hot_path_only else None)


def params_from_spec(spec, seed: int, dtype=torch.float32, skip=None) -> P:
    """Values for any (key, shape, kind) list: one numpy Generator per key, seeded by (seed, index)."""
    sd: P = {}
    for idx, (key, shape, kind) in enumerate(spec):
        if skip is not None and skip(key):
            continue
        rng = np.random.default_rng([seed, idx])
        fan = kind[1] if isinstance(kind, tuple) else None
        kind = kind[0] if isinstance(kind, tuple) else kind
        if kind == "linear_w":
            bound = 1.0 / math.sqrt(shape[-1])
            a = rng.uniform(-bound, bound, size=shape)
        elif kind == "linear_b":
            bound = 1.0 / math.sqrt(fan)
            a = rng.uniform(-bound, bound, size=shape)
        elif kind == "gamma":
            a = 1.0 + 0.1 * rng.standard_normal(size=shape)
        elif kind == "beta":
            a = 0.1 * rng.standard_normal(size=shape)
        elif kind == "var":
            a = rng.uniform(0.5, 1.5, size=shape)
        elif kind == "normal":
            a = rng.standard_normal(size=shape)
        elif kind == "conv":
            a = rng.standard_normal(size=shape) * math.sqrt(2.0 / (shape[0] * shape[2] * shape[3]))
        elif kind == "count":
            sd[key] = torch.tensor(0, dtype=torch.int64)
            continue
        elif kind == "pos_weight":
            a = np.asarray(AU_POS_WEIGHT)
        else:  # pragma: no cover
            raise ValueError(kind)
        sd[key] = torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
    return sd


def cast_params(p: P, dtype) -> P:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in p.items()}


def synth_inputs(seed: int, batch: int, n_frames: int, dtype=torch.float32, image: int = 112):
    """Aff-Wild2-shaped synthetic batch: clip [B,3,T,112,112], log-mel [B,1,64,1001] ~ N(0,1), and
    labels ~ Bernoulli(0.3) [B,12] (dataloader/aff2compdataset.py:48-65,114-175 for the shapes)."""
    rng = np.random.default_rng([seed, 7])
    clip = torch.from_numpy(rng.standard_normal((batch, 3, n_frames, image, image))).to(dtype)
    audio = torch.from_numpy(rng.standard_normal((batch, 1, 64, 1001))).to(dtype)
    labels = torch.from_numpy((rng.uniform(size=(batch, 12)) < 0.3).astype(np.float64)).to(dtype)
    return clip, audio, labels


def synth_hot_path_inputs(seed: int, batch: int, n_frames: int, dtype=torch.float32):
    """Stage-3-like maps (post-ReLU: mean ~1.06, std ~1.36), frame features and audio
    features with the statistics the real conv stages produce for N(0,1) inputs."""
    rng = np.random.default_rng([seed, 11])
    stage3 = np.maximum(rng.standard_normal((batch * n_frames, 256, 7, 7)) * 1.7 + 0.6, 0.0)
    frame = np.abs(rng.standard_normal((batch * n_frames, 512))) * 1.2
    audio = np.abs(rng.standard_normal((batch, 512))) * 1.0
    t = lambda a: torch.from_numpy(a).to(dtype)
    return t(stage3), t(frame), t(audio)
