"""Achieved HBM bandwidth of the memory-bound kernels of the path (north_star: "achieved HBM GB/s against peak for the elementwise
and norm kernels") at the bench / training sizes, timed with CUDA events, against MEASURED_PEAKS.json.  Buffers are larger than the
126 MB L2 or rotated so that every launch streams from HBM.

    python tools/membound_roofline.py > profiles/rNN_membound_roofline.md
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avformer_b200 as A
AF = A.functional

peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
PEAK = peaks["hbm_gbs"]
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()                                  # evict the 126 MB L2 between launches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


rows = []
def report(name, ref, nbytes, fn):
    ms = timed(fn)
    gbs = nbytes / ms / 1e6
    rows.append((name, ref, nbytes / 1e6, ms * 1e3, gbs, gbs / PEAK))

R, D = 8192 * 49, 256                                  # SFormer token matrix of the bench workload
x = torch.randn(R, D, device=dev); g = torch.randn(D, device=dev); b = torch.randn(D, device=dev)
report("layernorm_kernel (fp32 -> bf16), 401k x 256", "heads.py:181", R * D * 6, lambda: AF.layernorm_fwd(x, g, b, "bf16"))
Rt, Dt = 512 * 17, 512
xt = torch.randn(Rt, Dt, device=dev); gt = torch.randn(Dt, device=dev)
report("layernorm_kernel (fp32 -> bf16), 8704 x 512 (TFormer)", "heads.py:181", Rt * Dt * 6, lambda: AF.layernorm_fwd(xt, gt, gt, "bf16"))
fm = torch.randn(8192, 256, 7, 7, device=dev).bfloat16(); pos = torch.randn(49, 256, device=dev)
report("sformer_pack_kernel (NCHW bf16 -> tokens fp32 + pos), 8192 frames", "vformer.py:247-253", fm.numel() * 6, lambda: AF.sformer_tokens_pack(fm, pos))
tokx = torch.randn(8192 * 49, 256, device=dev)
report("sformer_unpack_kernel (tokens fp32 -> NCHW bf16), 8192 frames", "vformer.py:257-259", fm.numel() * 6, lambda: AF.sformer_tokens_unpack(tokx, (8192, 256, 7, 7), torch.bfloat16))
Rb = 1024 * 49                                          # training: 64 clips x 16 frames
xb = torch.randn(Rb, D, device=dev); dyn = torch.randn(Rb, D, device=dev); dres = torch.randn(Rb, D, device=dev)
report("layernorm_bwd_kernel (+ bf16 copy, d-gamma/beta/bias), 50k x 256", "autograd of heads.py:169-185", Rb * D * (4 + 4 + 4 + 4 + 2),
       lambda: AF.layernorm_bwd_(xb, g, dyn, dres, want_bf16=True))
hb = torch.randn(Rb, 512, device=dev).bfloat16()
report("colsum (bf16 [50k x 512] -> fp32 [512])", "bias gradient, heads.py:192", Rb * 512 * 2, lambda: AF.colsum(hb))
n = 40_000_000
p = torch.randn(n, device=dev); gr = torch.randn(n, device=dev); m_ = torch.zeros(n, device=dev); v_ = torch.zeros(n, device=dev)
report("adam_kernel, 40 M parameters (flat bucket)", "train.py:334", n * 4 * 7, lambda: AF.adam_step_(p, gr, m_, v_, 1, 5e-4, 0.9, 0.999, 1e-8, 5e-5))
w = torch.randn(64 << 20, device=dev)
report("cast_f32_bf16_kernel, 64 M elements", "weight packing", w.numel() * 6, lambda: AF.to_bf16(w))
fr = torch.randn(512 * 16, 512, device=dev).bfloat16(); cls = torch.randn(512, device=dev); pe = torch.randn(17, 512, device=dev)
report("tformer_embed_kernel, 512 clips x 16 frames", "vformer.py:283-286", fr.numel() * 2 + 512 * 17 * 512 * 4, lambda: AF.tformer_embed(fr, cls, pe, 16))

print("# Memory-bound kernels: achieved HBM bandwidth (CUDA events, L2 flushed between launches)\n")
print(f"Peak = {PEAK:.1f} GB/s (MEASURED_PEAKS.json: device-to-device copy, read + write bytes).  Bytes = algorithmic bytes in + out.\n")
print("| kernel | reference | MB / launch | us | GB/s | fraction of peak |")
print("|---|---|---|---|---|---|")
for name, ref, mb, us, gbs, frac in rows:
    print(f"| {name} | `{ref}` | {mb:.1f} | {us:.1f} | {gbs:.0f} | {frac:.2f} |")
