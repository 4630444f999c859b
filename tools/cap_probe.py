"""Fused SFormer kernel time vs number of SMs used (avf_set_sm_cap): per-tile time that falls with fewer active SMs means the
kernel is contending for a shared resource (L2 -> SM weight traffic), not for anything inside the SM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avformer_b200 as A
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
caps = [int(v) for v in sys.argv[2:]] or [148, 111, 74, 37]
torch.manual_seed(0)
m = A.video.ResFormer(A.video.BasicBlock, [2, 2, 2, 2]).cuda().eval()
m.spatial_transformer.precision = "bf16"
fm = (torch.clamp(torch.randn(frames, 256, 7, 7) * 1.7 + 0.6, min=0)).bfloat16().cuda()
L = A._lib.lib()
with torch.no_grad():
    for cap in caps:
        L.avf_set_sm_cap(cap)
        for _ in range(2):
            m.sformer(fm)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5):
            m.sformer(fm)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 5 * 1e3
        tiles = (frames + 1) // 2
        per_cta = -(-tiles // cap)
        print(f"cap {cap:3d}: {us:8.1f} us  {us / per_cta:6.2f} us per tile per CTA  ({frames * 53838848 / us / 1e6:.1f} TFLOP/s)")
    L.avf_set_sm_cap(0)
