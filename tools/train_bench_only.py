"""Time only the GraphedTrainStep of bench.py (A/B runs): python tools/train_bench_only.py"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import avformer_b200 as A
import bench
dev = torch.device("cuda")
torch.manual_seed(bench.SEED)
model = A.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU").to(dev).eval().set_precision("bf16")
args = types.SimpleNamespace(steps=30, warmup=3)
for _ in range(3):
    ms, launches = bench.train_step_ms(model, A, dev, 0, 1, args)
    print(f"train step {ms:.3f} ms ({launches} launches eager)", flush=True)
