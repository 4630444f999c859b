"""Achieved tensor throughput of gemm_umma_kernel at the GEMM shapes of the path (forward, dgrad, wgrad) against the measured
dense-bf16 peak (MEASURED_PEAKS.json), timed with CUDA events; operands rotate through > L2-sized pools where they are large.

    python tools/gemm_roofline.py > profiles/rNN_gemm_roofline.md
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import avformer_b200 as A
AF = A.functional
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
PEAK = json.load(open(pk))["bf16_tflops"] if os.path.exists(pk) else 1590.0
dev = torch.device("cuda")


def timed(fn, reps=20):
    """20 launches captured in one CUDA graph: the kernel time, not the Python/ctypes enqueue cost (~25 us per call)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


rows = []
def run(name, m, n, k, ta=False, tb=False, out=torch.float32, bias=False, res=False, gelu=False):
    a = torch.randn((k, m) if ta else (m, k), device=dev).bfloat16()
    b = torch.randn((k, n) if tb else (n, k), device=dev).bfloat16()
    bi = torch.randn(n, device=dev) if bias else None
    r = torch.randn(m, n, device=dev) if res else None
    flags = (1 if bias else 0) | (2 if gelu else 0) | (4 if res else 0)
    ms = timed(lambda: AF.gemm(a, b, ta, tb, bias=bi, residual=r, flags=flags, out_dtype=out, precision="bf16"))
    tf = 2.0 * m * n * k / ms / 1e9
    rows.append((name, m, n, k, "TN" if ta else ("NN" if tb else "NT"), "bf16" if out == torch.bfloat16 else "fp32", ms * 1e3, tf, tf / PEAK))

R = 512 * 17           # TFormer rows of the bench workload
run("TFormer to_qkv", R, 1536, 512, out=torch.bfloat16)
run("TFormer to_out (+bias +residual)", R, 512, 512, bias=True, res=True)
run("TFormer net.0 (+bias, GELU)", R, 1024, 512, out=torch.bfloat16, bias=True, gelu=True)
run("TFormer net.3 (+bias +residual)", R, 512, 1024, bias=True, res=True)
Ra = 512 * 12
run("AU_former front (12 stacked Linear(512,128))", 512, 1536, 512, bias=True)
run("AU_former to_qkv", Ra, 768, 128, out=torch.bfloat16)
run("AU_former to_out", Ra, 128, 256, bias=True, res=True)
run("AU_former net.0", Ra, 256, 128, out=torch.bfloat16, bias=True, gelu=True)
run("AU_former net.3", Ra, 128, 256, bias=True, res=True)
Rs = 1024 * 49         # SFormer rows in the training step (64 clips x 16 frames)
run("train: SFormer to_qkv fwd", Rs, 768, 256, out=torch.bfloat16)
run("train: SFormer dgrad net.3 (dY W)", Rs, 512, 256, tb=True, out=torch.bfloat16)
run("train: SFormer dgrad to_qkv", Rs, 256, 768, tb=True)
run("train: SFormer wgrad to_qkv (dY^T X, split-K)", 768, 256, Rs, ta=True, tb=True)
run("train: SFormer wgrad net.0", 512, 256, Rs, ta=True, tb=True)
run("train: TFormer wgrad net.0 (64 clips)", 1024, 512, 64 * 17, ta=True, tb=True)
run("reference point: square 8192^3", 8192, 8192, 8192, out=torch.bfloat16)

print("# gemm_umma_kernel (tcgen05, persistent, TMA ring): achieved TFLOP/s at the shapes of the path\n")
print(f"Peak = {PEAK:.1f} TFLOP/s (MEASURED_PEAKS.json, cuBLAS bf16 8192^3 burst).  CUDA events around one graph replay of 20 launches (kernel time incl. launch gaps, no host enqueue cost), inputs resident.\n")
print("| GEMM | M | N | K | form | out | us | TFLOP/s | fraction of peak |")
print("|---|---|---|---|---|---|---|---|---|")
for name, m, n, k, form, o, us, tf, fr in rows:
    print(f"| {name} | {m} | {n} | {k} | {form} | {o} | {us:.1f} | {tf:.0f} | {fr:.2f} |")
