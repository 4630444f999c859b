"""CPU-side checks: the C-ABI library loads, exports every symbol include/avformer_b200.h declares,
refuses to compute without a device, and the drop-in modules honour the reference's state-dict contract."""
import ctypes
import os
import re

import pytest
import torch

import avformer_b200 as A
from oracle import avformer_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "avformer_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(avf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(A._lib.build())
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/avformer_b200.h but not exported"
    assert set(names) == set(A._lib.SIGNATURES), "ctypes signature table and header disagree"
    assert A._lib.lib().avf_abi_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_no_device_means_error_not_fallback():
    L = A._lib.lib()
    a, b, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    assert L.avf_device_info(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)) == -2          # AVF_ENODEVICE
    assert b"no CPU fallback" in L.avf_last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        A.functional.to_bf16(torch.zeros(8))
    m = A.Transformer(128, 1, 8, 32, 256)
    with pytest.raises(RuntimeError, match="no CPU fallback"), torch.no_grad():
        m(torch.zeros(1, 12, 128))


def test_launch_policy_switches_round_trip_without_a_device():
    """avf_set_sm_cap / avf_set_pdl_enabled / avf_set_fused_enabled are plain process-wide switches (no device needed): each call
    returns the previous value; the developer probes report 'unsupported' in a product build."""
    L = A._lib.lib()
    assert L.avf_set_sm_cap(116) == 0 and L.avf_set_sm_cap(-5) == 116 and L.avf_set_sm_cap(0) == 0
    old = L.avf_set_pdl_enabled(0)
    assert old in (0, 1) and L.avf_set_pdl_enabled(1) == 0 and L.avf_set_pdl_enabled(old) == 1
    prev = L.avf_set_fused_enabled(0)
    assert L.avf_set_fused_enabled(prev) == 0
    buf = (ctypes.c_uint64 * 64)()
    assert L.avf_debug_gemm_prof(buf) != 0 or torch.cuda.is_available()


def test_state_dict_contract_matches_reference_names():
    m = A.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU")
    spec = {k: s for k, s, _ in O.state_dict_spec(16)}
    sd = m.state_dict()
    assert set(sd) == set(spec) and len(sd) == 462
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(spec[k]), k
    res = m.load_state_dict(O.make_state_dict(5, 16), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert m.modes == ["clip", "audio_features"] and m.task == "AU"
    assert sum(p.numel() for p in m.parameters()) == 32_764_416
    # dropout placement of the reference: 0.2 in the audio AU_former and the fusion head, 0 elsewhere
    assert m.audio_model.au_head.corr_transformer.dropout == 0.2 and m.au_head.corr_transformer.dropout == 0.2
    assert m.video_model.au_head.corr_transformer.dropout == 0.0
    # other clip lengths need a TFormer(num_patches=T) (models/vformer.py:271)
    m.set_clip_length(8)
    assert m.state_dict()["video_model.video_model.t_former.pos_embedding"].shape == (1, 9, 512)
    m.load_state_dict(O.make_state_dict(5, 8), strict=True)


def test_constructor_signatures():
    import inspect
    sig = inspect.signature(A.TwoStreamAuralVisualFormer.__init__)
    assert list(sig.parameters)[1:] == ["modality", "video_pretrained", "audio_pretrained", "task"]
    assert sig.parameters["task"].default == "EX" and sig.parameters["modality"].default == "A;V;M"
    assert list(inspect.signature(A.Transformer.__init__).parameters)[1:] == ["dim", "depth", "heads", "dim_head", "mlp_dim", "dropout"]
    assert list(inspect.signature(A.TFormer.__init__).parameters)[1:] == ["num_patches", "dim", "depth", "heads", "mlp_dim", "dim_head", "dropout"]
    assert list(inspect.signature(A.AU_former.__init__).parameters)[1:] == ["input_dim", "emb_dim", "dropout"]
    assert list(inspect.signature(A.former_AU_head.__init__).parameters)[1:] == ["emb_dim", "dropout"]
    assert A.tformer_AU_head is A.former_AU_head
    vm = A.VideoModel()
    vm.config_modality("A;V;M")
    assert vm.num_channels == 4 and vm.s_former.conv1.in_channels == 4


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "multi-modal-multi-label-facial-action-unit-detection-with-transformer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read(), f"{f} mentions the oracle"


def test_graph_weight_signature_watches_the_hot_path_cheaply():
    """GraphedHotPath runs this check on the host in front of every replay (graphs.py): it has to notice in-place edits,
    load_state_dict and replaced Parameter objects of the hot path's modules, ignore the conv backbones (not read by the captured
    kernels), and stay far below the 1.08 ms of the GPU step it precedes — the full-model scan took 0.9 ms and made the replay loop
    host-bound."""
    import time
    from avformer_b200.graphs import hot_path_modules, weights_signature
    m = A.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU").eval()
    watch = hot_path_modules(m)
    vm = m.video_model.video_model
    sig = weights_signature(m, watch)
    assert weights_signature(m, watch) == sig
    with torch.no_grad():
        vm.s_former.conv1.weight.mul_(1.0)                 # conv trunk inside the SFormer wrapper: outside the hot path
    assert weights_signature(m, watch) == sig
    assert weights_signature(m) != weights_signature(m, watch)   # the full scan covers more tensors
    full = weights_signature(m)
    with torch.no_grad():
        vm.s_former.conv1.weight.mul_(1.0)
    assert weights_signature(m) != full                     # ... and sees the conv edit (InferenceEngine relies on it)
    changes = []
    with torch.no_grad():
        vm.t_former.pos_embedding.add_(0.0)                 # in-place edit
    changes.append(weights_signature(m, watch))
    m.au_head.load_state_dict(m.au_head.state_dict())       # copy_ into every parameter of the fusion head
    changes.append(weights_signature(m, watch))
    lin = m.video_model.au_head.AU_linear_p3
    lin.weight = torch.nn.Parameter(lin.weight.detach().clone())          # a replaced Parameter object
    changes.append(weights_signature(m, watch))
    m.audio_model.au_head.AU_BN1.running_mean.add_(0.0)     # a buffer
    changes.append(weights_signature(m, watch))
    assert len({sig, *changes}) == 5
    t0 = time.perf_counter()
    for _ in range(20):
        weights_signature(m, watch)
    per_call = (time.perf_counter() - t0) / 20
    assert per_call < 0.5e-3, f"{per_call * 1e6:.0f} us per call"
