"""the other model variants instantiate the same encoder block at other (N, D, I, M) — TFormer at dim 1536
(models/tformer.py:301), the dim-512 spatial transformer of VGGFormer (models/vggformer.py:252-258) and VA_former with two tokens
(models/heads.py:341-372).  Each is run through the CUDA path and compared with outputs of the REFERENCE classes (tests/golden/
variants.npz) and with the fp64 oracle: fp32 mode 1e-4 relative, bf16 mode within the budgets of the hot path's own stacks."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import avformer_b200 as A
from oracle import avformer_oracle as O

pytestmark = pytest.mark.gpu


def _golden(golden_dir):
    spec = importlib.util.spec_from_file_location("make_golden_variants", os.path.join(golden_dir, "make_golden_variants.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg, dict(np.load(os.path.join(golden_dir, "variants.npz")))


def _err(a, b):
    return (a.double().cpu() - torch.as_tensor(b).double()).abs().max().item()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_tformer_dim_1536(golden_dir, precision):
    mg, g = _golden(golden_dir)
    frames, _, _ = mg.variant_inputs()
    m = A.TFormer(num_patches=16, dim=128 * 12)
    m.load_state_dict(O.make_variant_params("tformer1536", mg.SEED), strict=True)
    m = m.cuda().eval()
    m.spatial_transformer.precision = precision
    with torch.no_grad():
        cls = m(frames.cuda())
    assert cls.shape == (3, 1536)
    ref = torch.from_numpy(g["tformer1536_cls"])
    tol = 1e-4 * max(1.0, ref.abs().max().item()) if precision == "fp32" else 0.15
    assert _err(cls, ref) < tol, _err(cls, ref)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sformer_block_dim_512(golden_dir, precision):
    mg, g = _golden(golden_dir)
    _, fmap, _ = mg.variant_inputs()
    m = A.SFormerBlock(num_patches=49, dim=512, depth=1, heads=8, mlp_dim=512, dim_head=32)
    m.load_state_dict(O.make_variant_params("sformer512", mg.SEED), strict=True)
    m = m.cuda().eval()
    m.spatial_transformer.precision = precision
    x = fmap.cuda() if precision == "fp32" else fmap.bfloat16().cuda()
    out = m(x)
    assert out.shape == fmap.shape
    if precision == "fp32":
        ref = torch.from_numpy(g["sformer512_out"])
        assert _err(out, ref) < 1e-4 * max(1.0, ref.abs().max().item())
    else:
        ref = O.sformer_tokens(fmap.bfloat16().double(), O.cast_params(O.make_variant_params("sformer512", mg.SEED), torch.float64), "")
        d = (out.double().cpu() - ref).abs()
        assert d.max().item() < 0.2 and d.mean().item() < 2e-2


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_va_former_two_tokens(golden_dir, precision):
    mg, g = _golden(golden_dir)
    _, _, emb = mg.variant_inputs()
    m = A.VA_former()
    m.load_state_dict(O.make_variant_params("va_former", mg.SEED), strict=True)
    m = m.cuda().eval()
    m.corr_transformer.precision = precision
    va_out, tok = m(emb.cuda())
    assert va_out.shape == (5, 2) and tok.shape == (5, 2, 128)
    tol_t, tol_o = (1e-4 * max(1.0, np.abs(g["va_tokens"]).max()), 1e-4 * max(1.0, np.abs(g["va_out"]).max())) if precision == "fp32" else (6e-2, 2e-2)
    assert _err(tok, g["va_tokens"]) < tol_t and _err(va_out, g["va_out"]) < tol_o
    with pytest.raises(RuntimeError):
        m.train()(emb.cuda())
