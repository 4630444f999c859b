"""Pins oracle/avformer_oracle.py against outputs of the reference itself (tests/golden/*.npz,
made by tests/golden/make_golden.py in the build container).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import avformer_oracle as O


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def _close(a, b, atol, rtol=1e-5):
    a = a.detach().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    np.testing.assert_allclose(a, np.asarray(b, dtype=np.float64), atol=atol, rtol=rtol)


def test_state_dict_contract():
    spec = list(O.state_dict_spec(16))
    assert len(spec) == 462 and len({k for k, _, _ in spec}) == 462
    sd = O.make_state_dict(3, 16)
    assert sum(v.numel() for k, v in sd.items() if v.is_floating_point() and "running" not in k
               and "pos_weight" not in k) == 32_764_416                      # SURVEY.md §6 parameter count
    assert sd["video_model.video_model.t_former.pos_embedding"].shape == (1, 17, 512)
    assert O.make_state_dict(3, 8)["video_model.video_model.t_former.pos_embedding"].shape == (1, 9, 512)
    hp = O.make_state_dict(3, 16, hot_path_only=True)
    assert all(torch.equal(hp[k], sd[k]) for k in hp) and not any(O.is_backbone_key(k) for k in hp)


@pytest.mark.parametrize("T", [16, 8, 32])
def test_hot_path_blocks_match_reference(golden_dir, T):
    g = _load(golden_dir, f"hot_T{T}.npz")
    seed, B = int(g["seed"]), int(g["batch"])
    p = O.cast_params(O.make_state_dict(seed, T, hot_path_only=True), torch.float64)
    stage3, frame, audio = O.synth_hot_path_inputs(seed, B, T, torch.float64)
    out = O.hot_path_forward(stage3, frame, audio, p, T)
    _close(out["sformer_out"][:6], g["sformer_out_head"], 2e-5)
    _close(out["sformer_out"].sum(dim=(1, 2, 3)), g["sformer_out_sum"], 5e-2, 1e-6)
    _close(out["sformer_out"].abs().sum(dim=(1, 2, 3)), g["sformer_out_abssum"], 5e-2, 1e-6)
    _close(out["tformer_cls"], g["tformer_cls"], 2e-5)
    _close(out["video_tokens"], g["video_tokens"], 2e-5)
    _close(out["audio_tokens"], g["audio_tokens"], 2e-5)
    _close(out["logits"], g["logits"], 2e-5)
    au_v, _ = O.au_former(out["tformer_cls"], p, "video_model.au_head.")
    _close(au_v, g["video_au_out"], 2e-5)
    _close(O.au_loss(out["logits"], torch.from_numpy(g["labels"]).double()), g["loss"], 1e-6)
    # single attention sub-layer (x + Attn(LN(x))) of the SFormer layer on the first 4 frames
    s = "video_model.video_model.s_former."
    x0 = stage3[:4].reshape(4, 256, 49).permute(0, 2, 1) + p[s + "pos_embedding"]
    a = s + "spatial_transformer.layers.0.0."
    sub = x0 + O.attention(O.layer_norm(x0, p[a + "fn.norm.weight"], p[a + "fn.norm.bias"]), p, a, 8)
    _close(sub, g["sformer_attn_sublayer"], 2e-5)


@pytest.mark.parametrize("T", [8, 16])
def test_whole_model_matches_reference(golden_dir, T):
    g = _load(golden_dir, f"full_T{T}.npz")
    seed, B = int(g["seed"]), int(g["batch"])
    p = O.cast_params(O.make_state_dict(seed, T), torch.float64)
    clip, audio, labels = O.synth_inputs(seed, B, T, torch.float64)
    out = O.avformer_forward(clip, audio, p)
    _close(out["stage3"][:1], g["stage3_frame0"], 5e-5)
    _close(out["sformer_out"][:1], g["sformer_out_frame0"], 1e-4)
    _close(out["frame_feat"], g["frame_feat"], 1e-4)
    _close(out["tformer_cls"], g["tformer_cls"], 1e-4)
    _close(out["audio_feat"], g["audio_feat"], 1e-4)
    _close(out["audio_tokens"], g["audio_tokens"], 1e-4)
    _close(out["video_tokens"], g["video_tokens"], 1e-4)
    _close(out["logits"], g["logits"], 1e-4)
    assert out["output"].shape == (B, 21) and float(out["output"][:, 12:].abs().max()) == 0.0
    _close(O.au_loss(out["logits"], labels), g["loss"], 1e-5)
    dec = O.decisions(out["logits"])
    assert np.array_equal(dec, g["decisions"])
    assert np.array_equal(dec, (out["logits"].numpy() > 0).astype(np.int64))     # round(sigmoid) == logit > 0
    acc, f1, _ = O.multilabel_acc_f1(labels.numpy(), dec, ignore_index=-1)
    assert abs(acc - float(g["acc"])) < 1e-12 and abs(f1 - float(g["f1"])) < 1e-12


def test_loss_gradient_closed_form(golden_dir):
    g = _load(golden_dir, "grad_T16.npz")
    logits = torch.from_numpy(g["logits"]).double()
    labels = torch.from_numpy(g["labels"]).double()
    _close(O.au_loss_grad(logits, labels), g["dlogits"], 1e-7)
    _close(O.au_loss(logits, labels), g["loss"], 1e-6)
    # with an ignored row the mean runs over the remaining rows only (models/loss.py:85-102)
    labels2 = labels.clone()
    labels2[1, 0] = -1.0
    x = logits.clone().requires_grad_(True)
    O.au_loss(x, labels2).backward()
    _close(O.au_loss_grad(logits, labels2), x.grad, 1e-12)
    assert float(x.grad[1].abs().max()) == 0.0


def test_fp32_oracle_close_to_fp64(golden_dir):
    """The fp32 instance (what bench.py times as the CPU baseline) agrees with the fp64 checker."""
    g = _load(golden_dir, "hot_T16.npz")
    seed, B, T = int(g["seed"]), int(g["batch"]), 16
    p = O.make_state_dict(seed, T, hot_path_only=True)
    out = O.hot_path_forward(*O.synth_hot_path_inputs(seed, B, T), p, T)
    _close(out["logits"], g["logits"], 1e-4)
    assert out["logits"].dtype == torch.float32


def test_training_oracle_matches_reference_gradients(golden_dir):
    """loss.backward() of the REFERENCE modules (grad_T16.npz) vs autograd through the restated forward: every small
    gradient tensor element-wise, every large one by its norm, plus d(loss)/d(fusion input)."""
    g = _load(golden_dir, "grad_T16.npz")
    T, B, seed = int(g["n_frames"]), int(g["batch"]), int(g["seed"])
    p = O.cast_params(O.make_state_dict(seed, T, hot_path_only=True), torch.float64)
    stage3, frame, audio = O.synth_hot_path_inputs(seed, B, T, torch.float64)
    loss, grads, gin, out = O.hot_path_grads(stage3, frame, audio, torch.from_numpy(g["labels"]).double(), p, T)
    _close(loss, g["loss"], 1e-6)
    _close(out["logits"], g["logits"], 2e-5)
    n_full = n_norm = 0
    for k, v in g.items():
        if k.startswith("g:"):
            _close(grads[k[2:]], v, 2e-6 * max(1.0, float(np.abs(v).max())), 1e-4)
            n_full += 1
        elif k.startswith("gnorm:"):
            assert abs(grads[k[6:]].norm().item() - float(v)) <= 1e-5 * max(float(v), 1e-6), k
            n_norm += 1
    assert n_full > 100 and n_norm > 150
    # the SFormer is not on the loss path of the isolated hot path unless the probe term is switched on
    assert gin["stage3"] is None or float(gin["stage3"].abs().max()) == 0.0
    _, _, gin2, _ = O.hot_path_grads(stage3[:32], frame, audio, torch.from_numpy(g["labels"]).double(), p, T, sformer_loss_weight=1.0)
    assert float(gin2["stage3"].abs().max()) > 0.0


def test_adam_rule_matches_torch_optim():
    """The restated update (train.py:334: torch.optim.Adam with coupled weight decay; AdamW variant) against torch."""
    torch.manual_seed(0)
    for decoupled in (False, True):
        p0 = torch.randn(257, dtype=torch.float64)
        q = torch.nn.Parameter(p0.clone())
        opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)([q], lr=5e-4, weight_decay=5e-5)
        p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
        for step in range(1, 6):
            gr = torch.randn(257, dtype=torch.float64)
            q.grad = gr.clone()
            opt.step()
            p, m, v = O.adam_update(p, gr, m, v, step, 5e-4, (0.9, 0.999), 1e-8, 5e-5, decoupled)
        _close(q, p, 1e-12)


def test_batchnorm_train_restatement():
    """au_former(batch_stats=True) and bn_running_update against nn.BatchNorm1d in train()."""
    torch.manual_seed(1)
    bn = torch.nn.BatchNorm1d(512).double().train()
    with torch.no_grad():
        bn.weight.normal_(); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 1.5)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    x = torch.randn(6, 512, dtype=torch.float64) * 2 + 1
    y = bn(x)
    mu, var = x.mean(0), x.var(0, unbiased=False)
    _close((x - mu) / torch.sqrt(var + 1e-5) * bn.weight + bn.bias, y.detach(), 1e-12)
    rm, rv = O.bn_running_update(x, rm0, rv0)
    _close(rm, bn.running_mean, 1e-12)
    _close(rv, bn.running_var, 1e-12)


def test_variant_instantiations_match_reference_golden(golden_dir):
    """TFormer at dim 1536 (models/tformer.py:301), the dim-512 spatial transformer of VGGFormer
    (models/vggformer.py:252-258) and VA_former (models/heads.py:341-372) — the oracle restatements against outputs of the reference
    classes themselves (tests/golden/make_golden_variants.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_variants", os.path.join(golden_dir, "make_golden_variants.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = dict(np.load(os.path.join(golden_dir, "variants.npz")))
    frames, fmap, emb = mg.variant_inputs()
    cls = O.tformer(frames, O.make_variant_params("tformer1536", mg.SEED), "", 16)
    assert np.abs(cls.numpy() - g["tformer1536_cls"]).max() < 2e-5
    s_out = O.sformer_tokens(fmap, O.make_variant_params("sformer512", mg.SEED), "")
    assert np.abs(s_out.numpy() - g["sformer512_out"]).max() < 2e-5
    va_out, va_tok = O.va_former(emb, O.make_variant_params("va_former", mg.SEED))
    assert np.abs(va_out.numpy() - g["va_out"]).max() < 1e-5 and np.abs(va_tok.numpy() - g["va_tokens"]).max() < 1e-5
