"""One GEMM shape of the path, repeated (for ncu): python tools/gemm_one.py M N K [bf16|fp32] [res]"""
import sys, os
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
for _ in range(4):
    AF.gemm(a, b, False, False, bias=bi, residual=r, flags=1 | (4 if res else 0), out_dtype=out, precision="bf16")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(20):
    AF.gemm(a, b, False, False, bias=bi, residual=r, flags=1 | (4 if res else 0), out_dtype=out, precision="bf16")
e1.record(); torch.cuda.synchronize()
print(f"gemm {m}x{n}x{k}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call (eager enqueue included)")
