"""Whole-model evaluation throughput: model(x) with the stock torch fp32 conv backbones vs InferenceEngine
(BN-folded channels_last convs + CUDA graph) with fp32 / bf16 backbones.  One JSON line per variant.

    python tools/full_model.py [clips] [frames]
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avformer_b200 as A

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
torch.manual_seed(0)
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda")
model = A.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU").set_clip_length(T).to(dev).eval().set_precision("bf16")
x = {"clip": torch.randn(B, 3, T, 112, 112, device=dev), "audio_features": torch.randn(B, 1, 64, 1001, device=dev)}
FLOP_PER_CLIP_16 = 21.12e9


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


with torch.no_grad():
    ref = model(x).clone()
    rows = [("model(x): torch fp32 NCHW backbones, eager", lambda: model(x), ref)]
    for name, dt in (("InferenceEngine fp32 backbones (folded BN, channels_last, CUDA graph)", torch.float32),
                     ("InferenceEngine bf16 backbones (folded BN, channels_last, CUDA graph)", torch.bfloat16)):
        eng = A.InferenceEngine(model, backbone_dtype=dt)
        rows.append((name, (lambda e=eng: e(x)), eng(x).clone()))
    for name, fn, out in rows:
        ms = timed(fn)
        print(json.dumps({"variant": name, "clips": B, "frames": T, "ms": ms, "clips_per_s": B / ms * 1e3,
                          "tflops_full_model": B * FLOP_PER_CLIP_16 * T / 16 / ms / 1e9, "max_abs_diff_vs_module": float((out - ref).abs().max())}), flush=True)
