"""Phase breakdown of the fused SFormer kernel (library must be built with AVF_NVCC_EXTRA=-DAVF_FUSED_PROF)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avformer_b200 as A
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
torch.manual_seed(0)
m = A.video.ResFormer(A.video.BasicBlock, [2, 2, 2, 2]).cuda().eval()
m.spatial_transformer.precision = "bf16"
fm = (torch.clamp(torch.randn(frames, 256, 7, 7) * 1.7 + 0.6, min=0)).bfloat16().cuda()
L = A._lib.lib()
buf = (ctypes.c_uint64 * 64)()
with torch.no_grad():
    for _ in range(2):
        m.sformer(fm)
    rc = L.avf_debug_fused_prof(buf, 1)
    if rc != 0:
        print("library built without -DAVF_FUSED_PROF"); sys.exit(0)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); m.sformer(fm); e1.record(); torch.cuda.synchronize()
    L.avf_debug_fused_prof(buf, 0)
print(f"kernel {e0.elapsed_time(e1)*1e3:.1f} us for {frames} frames")
W = ["INPUT", "VEC", "LN1", "WAIT_D1", "E1", "WAIT_O", "E3", "WAIT_S", "E2", "WAIT_X1", "LN2", "WAIT_HACC", "GELU", "WAIT_X2", "OUTPUT"]
M = ["WAIT_A0", "QKV", "WAIT_STAGED", "S", "WAIT_P", "PV", "WAIT_OD7", "OUT", "WAIT_A0B", "FF1", "WAIT_H", "FF2", "RINGWAIT", "RINGWAIT_QKV", "RINGWAIT_OUT", "RINGWAIT_FF1", "RINGWAIT_FF2"]
tiles = max(1, buf[15])
print(f"CTA 0: {tiles} tiles; cycles per tile")
tw = sum(buf[i] for i in range(15)); tm = sum(buf[32 + i] for i in range(len(M)))
print("worker thread: total %.0f" % (tw / tiles))
for i, n in enumerate(W):
    print(f"   {n:10s} {buf[i] / tiles:9.0f}  {100 * buf[i] / tw:5.1f}%")
for i, n in enumerate(["E2_LD", "E2_EXP", "E2_XCH", "E2_ST(rest of E2 = arrive)"]):
    print(f"   (inside E2) {n:10s} {buf[16 + i] / tiles:9.0f}")
print("MMA thread: total %.0f" % (tm / tiles))
for i, n in enumerate(M):
    print(f"   {n:12s} {buf[32 + i] / tiles:9.0f}  {100 * buf[32 + i] / tm:5.1f}%")
