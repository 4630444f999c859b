"""Run only the fused SFormer kernel (for ncu): python tools/fused_bench.py [frames] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avformer_b200 as A
AF = A.functional
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.manual_seed(0)
m = A.video.ResFormer(A.video.BasicBlock, [2, 2, 2, 2]).cuda().eval()
m.spatial_transformer.precision = "bf16"
fm = (torch.clamp(torch.randn(frames, 256, 7, 7) * 1.7 + 0.6, min=0)).bfloat16().cuda()
with torch.no_grad():
    for _ in range(2):
        m.sformer(fm)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        m.sformer(fm)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
print(f"sformer fused {frames} frames: {ms*1e3:.1f} us  {frames*53838848/ms/1e9:.1f} TFLOP/s")
