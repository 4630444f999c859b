"""Milestones of CTA 0 of one tcgen05 GEMM launch (library built with AVF_NVCC_EXTRA=-DAVF_GEMM_PROF):
python tools/gemm_phases.py M N K [bf16|fp32] [res]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avformer_b200 as A
AF = A.functional
m, n, k = (int(v) for v in sys.argv[1:4])
out = torch.bfloat16 if (len(sys.argv) > 4 and sys.argv[4] == "bf16") else torch.float32
res = len(sys.argv) > 5
a = torch.randn(m, k, device="cuda").bfloat16()
b = torch.randn(n, k, device="cuda").bfloat16()
bi = torch.randn(n, device="cuda")
r = torch.randn(m, n, device="cuda") if res else None
fl = 1 | (4 if res else 0)
L = A._lib.lib()
buf = (ctypes.c_uint64 * 16)()
names = ["entry", "prologue done", "predecessor complete (pdl_wait)", "first operand stage landed", "MMAs of tile 0 issued", "tile 1 issued", "tile 2 issued",
         "tile 3+ issued", "accumulator 0 complete", "accumulator 1 complete", "accumulator 2 complete", "accumulator 3+ complete", "epilogue warp done", "kernel end"]
for rep in range(3):
    for _ in range(3):
        AF.gemm(a, b, False, False, bias=bi, residual=r, flags=fl, out_dtype=out, precision="bf16")
    torch.cuda.synchronize()
    if L.avf_debug_gemm_prof(buf) != 0:
        print("library built without -DAVF_GEMM_PROF"); sys.exit(0)
    t0 = buf[0]
    print(f"--- gemm {m}x{n}x{k} run {rep}: milestones of CTA 0 in us after kernel entry")
    for i, nm in enumerate(names):
        if buf[i] >= t0 and buf[i] - t0 < 10**9:
            print(f"   {nm:34s} {(buf[i] - t0) / 1e3:8.2f}")
