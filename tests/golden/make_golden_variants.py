#!/usr/bin/env python
"""Golden vectors for the other instantiations of the encoder block, made by running the UNMODIFIED reference
classes in the build container (only where /root/reference exists):

    python tests/golden/make_golden_variants.py

* ``models.tformer.TFormer(dim=128*12)``     (models/tformer.py:271-294, instantiated at :301) on [3*16, 1536] frame features;
* VGGFormer's spatial-transformer region      (models/vggformer.py:252-258 construct it; its forward does exactly what
  models/vformer.py:245-259 does: reshape/permute, + pos, Transformer(512, 1, 8, 32, 512), permute back) on a [3, 512, 7, 7] map —
  the region is rebuilt here from ``models.heads.Transformer`` because the class itself needs a VGGFace2 trunk that is not in the repo;
* ``models.heads.VA_former()``                (models/heads.py:341-372) on [5, 512] embeddings.

Weights and inputs are regenerated bit-identically from numpy PCG64 seeds (oracle.make_variant_params) and loaded with strict=True,
which also pins the variants' state-dict names; only the reference's outputs are stored.
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
SEED = 31


def variant_inputs(seed=SEED):
    rng = np.random.default_rng([seed, 99])
    frames = torch.from_numpy(np.abs(rng.standard_normal((3 * 16, 1536))) * 1.2).float()
    fmap = torch.from_numpy(np.maximum(rng.standard_normal((3, 512, 7, 7)) * 1.7 + 0.6, 0.0)).float()
    emb = torch.from_numpy(np.abs(rng.standard_normal((5, 512)))).float()
    return frames, fmap, emb


def main():
    sys.path.insert(0, REF)
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REF, "models")]
    sys.modules["models"] = pkg
    heads = importlib.import_module("models.heads")
    tformer = importlib.import_module("models.tformer")
    frames, fmap, emb = variant_inputs()
    out = {}
    with torch.no_grad():
        t = tformer.TFormer(dim=128 * 12).eval()
        t.load_state_dict(O.make_variant_params("tformer1536", SEED), strict=True)
        out["tformer1536_cls"] = t(frames).numpy()

        p = O.make_variant_params("sformer512", SEED)
        tr = heads.Transformer(512, 1, 8, 32, 512, 0.0).eval()
        tr.load_state_dict({k[len("spatial_transformer."):]: v for k, v in p.items() if k.startswith("spatial_transformer.")}, strict=True)
        b, c, h, w = fmap.shape
        x = fmap.reshape((b, c, h * w)).permute(0, 2, 1)                      # models/vformer.py:247-249
        x = x + p["pos_embedding"][:, : h * w]                                 # :253
        x = tr(x)                                                              # :255
        out["sformer512_out"] = x.permute(0, 2, 1).reshape((b, c, h, w)).numpy()   # :257-259

        va = heads.VA_former().eval()
        va.load_state_dict(O.make_variant_params("va_former", SEED), strict=True)
        va_out, va_tok = va(emb)
        out["va_out"], out["va_tokens"] = va_out.numpy(), va_tok.numpy()
    np.savez_compressed(os.path.join(OUT, "variants.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
