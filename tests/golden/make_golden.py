#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference in the build container.

Run from the repo root (only where /root/reference exists):

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so parity is pinned on
its own outputs.  Weights and inputs are *not* stored: they are regenerated bit-identically from
numpy PCG64 seeds by ``oracle.avformer_oracle.make_state_dict / synth_*`` (loaded into the
reference with ``strict=True`` — which also pins the 462-key state-dict contract).  Only the
reference's outputs are stored, in float32, small enough to commit.

Import shims: bypass models/__init__.py (needs timm; imports a broken
avformer), stub audio.AudioFTDNNModel (models/avformer.py:16), alias heads.former_AU_head to
tformer.tformer_AU_head (models/avformer.py:19,87), patch torch.cuda.current_device while
constructing AULoss on a GPU-less host (models/loss.py:73).  forward() itself ends in a hard-coded
.cuda() (models/avformer.py:102), so the sub-modules are called in the same order instead.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import avformer_oracle as O  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    sys.path.insert(0, REF)
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REF, "models")]
    sys.modules["models"] = pkg
    audio = importlib.import_module("models.audio")
    audio.AudioFTDNNModel = None
    heads = importlib.import_module("models.heads")
    tformer = importlib.import_module("models.tformer")
    heads.former_AU_head = tformer.tformer_AU_head
    vformer = importlib.import_module("models.vformer")
    cd = torch.cuda.current_device
    torch.cuda.current_device = lambda: "cpu"
    try:
        av = importlib.import_module("models.avformer")
    finally:
        pass
    return av, vformer, heads, tformer, cd


def build_reference(av, vformer, cd, n_frames, seed):
    torch.cuda.current_device = lambda: "cpu"
    m = av.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU")
    torch.cuda.current_device = cd
    if n_frames != 16:   # models/vformer.py:271 hard-wires 16; other clip lengths need the ctor argument
        m.video_model.video_model.t_former = vformer.TFormer(num_patches=n_frames)
    sd = O.make_state_dict(seed, n_frames)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    ref_sd = m.state_dict()
    assert list(sorted(ref_sd)) == list(sorted(sd)) and len(sd) == 462
    for k in sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
    return m.eval(), sd


def ref_forward(m, clip, audio):
    a = m.audio_model(audio)
    v = m.video_model(clip)
    return a, v, m.au_head(torch.cat([a, v], dim=2))


class _Id(torch.nn.Module):
    def forward(self, x):
        return x


def ref_sformer_tokens(m, stage3):
    """Run reference lines models/vformer.py:245-259 unmodified by replacing the conv stages of a
    shallow copy of the ResFormer with identities (the map is fed in as a [1,F,256,7,7] 'clip')."""
    import copy
    s = copy.copy(m.video_model.video_model.s_former)
    s._modules = dict(s._modules)
    for name in ("conv1", "bn1", "relu", "maxpool", "layer1", "layer2", "layer3", "layer4", "avgpool"):
        s._modules[name] = _Id()
    out = s(stage3[None])                      # flatten(x, 1) of [F,256,7,7]
    return out.reshape(stage3.shape)


def save(name, **arrs):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (np.asarray(v.detach().numpy() if torch.is_tensor(v) else v)) for k, v in arrs.items()})
    print(f"wrote {path}: {os.path.getsize(path) / 1e3:.1f} kB")


def main():
    torch.set_grad_enabled(False)
    torch.backends.mkldnn.enabled = True
    av, vformer, heads, tformer_mod, cd = load_reference()
    sys.path.insert(0, REF)
    accf1 = importlib.import_module("metrics.accf1")

    # ---- A. hot-path blocks on synthetic hot-path inputs, T = 16 / 8 / 32 ---------------------
    for T, B in ((16, 2), (8, 2), (32, 2)):
        seed = 100 + T
        m, sd = build_reference(av, vformer, cd, T, seed)
        stage3, frame, audio_feat = O.synth_hot_path_inputs(seed, B, T)
        s_out = ref_sformer_tokens(m, stage3)
        cls = m.video_model.video_model.t_former(frame)
        au_v, tok_v = m.video_model.au_head(cls)
        au_a, tok_a = m.audio_model.au_head(audio_feat)
        fused = torch.cat([tok_a, tok_v], dim=2)
        logits = m.au_head(fused)
        labels = O.synth_inputs(seed, B, T, image=8)[2]
        labels[0, 0] = -1.0 if T == 8 else labels[0, 0]          # exercise the ignore-row rule once
        loss = m.get_au_loss(torch.cat([logits, torch.zeros(B, 9)], 1), labels)
        # a single encoder layer + its pieces for kernel-level tests (first SFormer layer, first 4 frames)
        x0 = stage3[:4].reshape(4, 256, 49).permute(0, 2, 1) + m.video_model.video_model.s_former.pos_embedding
        layer = m.video_model.video_model.s_former.spatial_transformer.layers[0]
        attn_sub = layer[0](x0)
        save(f"hot_T{T}.npz", seed=seed, batch=B, n_frames=T,
             sformer_out_head=s_out[:6].float(), sformer_out_sum=s_out.double().sum(dim=(1, 2, 3)),
             sformer_out_abssum=s_out.double().abs().sum(dim=(1, 2, 3)),
             sformer_attn_sublayer=attn_sub.float(),
             tformer_cls=cls, video_tokens=tok_v, audio_tokens=tok_a, video_au_out=au_v, audio_au_out=au_a,
             logits=logits, labels=labels, loss=loss)

    # ---- B. whole model, config 1 (B=2, T=8) and the native T=16 shape -------------------------
    for T, B in ((8, 2), (16, 2)):
        seed = 200 + T
        m, sd = build_reference(av, vformer, cd, T, seed)
        clip, audio, labels = O.synth_inputs(seed, B, T)
        inter = {}
        vm = m.video_model.video_model
        hooks = [vm.s_former.layer3.register_forward_hook(lambda mod, i, o: inter.__setitem__("stage3", o)),
                 vm.s_former.spatial_transformer.register_forward_hook(lambda mod, i, o: inter.__setitem__("sf", o)),
                 vm.s_former.register_forward_hook(lambda mod, i, o: inter.__setitem__("frame_feat", o)),
                 vm.t_former.register_forward_hook(lambda mod, i, o: inter.__setitem__("tformer_cls", o)),
                 m.audio_model.audio_model.register_forward_hook(lambda mod, i, o: inter.__setitem__("audio_feat", o))]
        a, v, logits = ref_forward(m, clip, audio)
        for h in hooks:
            h.remove()
        out21 = torch.zeros(B, 21)
        out21[:, :12] = logits
        loss = m.get_au_loss(out21, labels)
        pred = np.round(torch.sigmoid(logits).numpy())            # train.py:155
        metric = accf1.MultiLabelAccF1(ignore_index=-1)
        metric.update(pred, labels.numpy())
        acc, f1 = metric.get()
        sf = inter["sf"].permute(0, 2, 1).reshape(inter["stage3"].shape)
        save(f"full_T{T}.npz", seed=seed, batch=B, n_frames=T,
             stage3_frame0=inter["stage3"][:1], sformer_out_frame0=sf[:1],
             stage3_sum=inter["stage3"].double().sum(dim=(1, 2, 3)), sformer_out_sum=sf.double().sum(dim=(1, 2, 3)),
             frame_feat=inter["frame_feat"], tformer_cls=inter["tformer_cls"], audio_feat=inter["audio_feat"],
             audio_tokens=a, video_tokens=v, logits=logits, labels=labels, loss=loss,
             decisions=pred.astype(np.int64), acc=acc, f1=f1)

    # ---- C. gradients of the AU loss through the fusion head (reference default trainable set) --
    torch.set_grad_enabled(True)
    T, B, seed = 16, 4, 316
    m, sd = build_reference(av, vformer, cd, T, seed)
    m.eval()                                                      # dropout off: bit-matching Philox is out of scope
    _, frame, audio_feat = O.synth_hot_path_inputs(seed, B, T)
    labels = O.synth_inputs(seed, B, T, image=8)[2]
    for p_ in m.parameters():
        p_.requires_grad_(True)
    cls = m.video_model.video_model.t_former(frame)
    _, tok_v = m.video_model.au_head(cls)
    _, tok_a = m.audio_model.au_head(audio_feat)
    fused = torch.cat([tok_a, tok_v], dim=2)
    fused.retain_grad()
    logits = m.au_head(fused)
    logits.retain_grad()
    out21 = torch.cat([logits, torch.zeros(B, 9)], 1)
    loss = m.get_au_loss(out21, labels)
    loss.backward()
    grads = {}
    for k, p_ in m.named_parameters():
        if p_.grad is not None and not O.is_backbone_key(k):
            g = p_.grad.double()
            grads["gnorm:" + k] = g.norm()
            if p_.numel() <= 4096:
                grads["g:" + k] = p_.grad
    save("grad_T16.npz", seed=seed, batch=B, n_frames=T, loss=loss.detach(), logits=logits.detach(),
         dlogits=logits.grad, dfused=fused.grad, labels=labels, **grads)


if __name__ == "__main__":
    main()
