"""Raw pinned H2D time of the bench inputs vs hot_path_from_host with different chunk counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avformer_b200 as A
B, T = 512, 16
torch.manual_seed(2024)
dev = torch.device("cuda")
m = A.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU").to(dev).eval().set_precision("bf16")
g = torch.Generator().manual_seed(2024)
stage3 = torch.clamp(torch.randn(B * T, 256, 7, 7, generator=g) * 1.7 + 0.6, min=0).bfloat16().pin_memory()
frame = (torch.randn(B * T, 512, generator=g).abs() * 1.2).bfloat16().pin_memory()
audio = torch.randn(B, 512, generator=g).abs().pin_memory()
out_h = torch.empty((B, 21)).pin_memory(); dec_h = torch.empty((B, 12), dtype=torch.int32).pin_memory()
d3 = torch.empty_like(stage3, device=dev); df = torch.empty_like(frame, device=dev)

def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

def raw():
    d3.copy_(stage3, non_blocking=True); df.copy_(frame, non_blocking=True)
print(f"raw H2D of stage3+frame ({(stage3.numel()*2+frame.numel()*2)/1e6:.0f} MB): {timed(raw):.3f} ms")
with torch.no_grad():
    for ch in (2, 4, 8, 16, 32):
        print(f"chunks={ch}: {timed(lambda: m.hot_path_from_host(stage3, frame, audio, out_h, dec_h, chunks=ch)):.3f} ms", flush=True)
