"""A few eager forward steps of the hot path (bench.py workload: 512 clips x 16 frames) for ncu launch lists / --set full captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avformer_b200 as A

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, T = 512, 16
torch.manual_seed(2024)
dev = torch.device("cuda")
model = A.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU").to(dev).eval().set_precision("bf16")
g = torch.Generator().manual_seed(2024)
stage3 = torch.clamp(torch.randn(B * T, 256, 7, 7, generator=g) * 1.7 + 0.6, min=0).bfloat16().to(dev)
frame = (torch.randn(B * T, 512, generator=g).abs() * 1.2).bfloat16().to(dev)
audio = torch.randn(B, 512, generator=g).abs().to(dev)
with torch.no_grad():
    for i in range(steps):
        s_out, out21, dec = model.hot_path(stage3, frame, audio, want_decisions=True)
torch.cuda.synchronize()
print("logits", out21[0, :3].tolist(), "launches", A._lib.lib().avf_launch_count())
