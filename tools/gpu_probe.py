#!/usr/bin/env python
"""Bring-up probe for the GPU box: runs each kernel family against the CPU oracle / torch and PRINTS the
errors instead of asserting, one family per process (a trapped kernel poisons its CUDA context).

    python tools/gpu_probe.py all            # spawns one subprocess per family with a timeout
    python tools/gpu_probe.py gemm           # a single family in this process
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FAMILIES = ["rowops", "gemm_f32", "gemm", "attention", "stack_fp32", "stack_bf16", "model"]


def _err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    d = (a - b).abs()
    return f"max_abs={d.max().item():.3e} mean_abs={d.mean().item():.3e} ref_absmax={b.abs().max().item():.3e}"


def fam_rowops():
    import torch
    import avformer_b200 as A
    from oracle import avformer_oracle as O
    AF = A.functional
    print("device:", AF.device_info())
    torch.manual_seed(0)
    for D in (128, 256, 512):
        x = torch.randn(1000, D, device="cuda") * 2 + 0.5
        g, b = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
        ref = O.layer_norm(x.double().cpu(), g.double().cpu(), b.double().cpu())
        print(f"layernorm D={D} fp32:", _err(AF.layernorm_fwd(x, g, b, "fp32"), ref))
        print(f"layernorm D={D} bf16:", _err(AF.layernorm_fwd(x, g, b, "bf16").float(), ref))
    fm = torch.randn(5, 256, 7, 7, device="cuda")
    pos = torch.randn(49, 256, device="cuda")
    x = torch.empty(5 * 49, 256, device="cuda")
    AF.check(A._lib.lib().avf_sformer_tokens_pack(0, AF._ptr(fm), AF._ptr(pos), AF._ptr(x), 5, 256, 49, AF._stream()))
    ref = fm.reshape(5, 256, 49).permute(0, 2, 1) + pos
    print("pack:", _err(x.view(5, 49, 256), ref))
    back = torch.empty_like(fm)
    AF.check(A._lib.lib().avf_sformer_tokens_unpack(0, AF._ptr(x), AF._ptr(back), 5, 256, 49, AF._stream()))
    print("unpack:", _err(back, (fm.reshape(5, 256, 49) + pos.t()).reshape(5, 256, 7, 7)))
    fr = torch.randn(3 * 16, 512, device="cuda")
    cls, p2 = torch.randn(512, device="cuda"), torch.randn(17, 512, device="cuda")
    t = AF.tformer_embed(fr, cls, p2, 16)
    ref = torch.cat([cls.expand(3, 1, 512), fr.view(3, 16, 512)], 1) + p2
    print("embed:", _err(t.view(3, 17, 512), ref))
    print("cls_extract:", _err(AF.tformer_cls_extract(t, 3, 17), ref[:, 0]))
    logits = torch.randn(37, 21, device="cuda") * 2
    y = (torch.rand(37, 12, device="cuda") < 0.3).float()
    y[3, 0] = -1
    pw = torch.tensor(O.AU_POS_WEIGHT, device="cuda")
    loss, nv, grad = AF.au_bce_loss(logits, y, pw, want_grad=True)
    print("bce loss:", _err(loss, O.au_loss(logits[:, :12].double().cpu(), y.double().cpu())), "n_valid", nv.item())
    print("bce grad:", _err(grad, O.au_loss_grad(logits[:, :12].double().cpu(), y.double().cpu())))
    xt = torch.randn(37 * 12, 256, device="cuda")
    w = torch.randn(12, 256, device="cuda")
    out, dec = AF.au_logits(xt, w, 37, want_decisions=True)
    ref = (xt.view(37, 12, 256) * w).sum(-1)
    print("au_logits:", _err(out[:, :12], ref), "pad", out[:, 12:].abs().max().item(), "dec ok", bool((dec.bool() == (ref > 0)).all()))


def fam_gemm_f32():
    import torch
    import avformer_b200 as A
    AF = A.functional
    torch.manual_seed(1)
    for (m, n, k) in ((200, 768, 256), (131, 128, 512), (64, 1536, 512)):
        a, w = torch.randn(m, k, device="cuda"), torch.randn(n, k, device="cuda") / k ** 0.5
        b, r = torch.randn(n, device="cuda"), torch.randn(m, n, device="cuda")
        ref = a.double() @ w.double().t()
        print(f"gemm_f32 {m}x{n}x{k} plain:", _err(AF.linear_fwd(a, w, precision="fp32"), ref))
        ref2 = ref + b.double() + r.double()
        print(f"gemm_f32 {m}x{n}x{k} bias+res:", _err(AF.linear_fwd(a, w, b, r, precision="fp32"), ref2))
        g = torch.nn.functional.gelu(ref + b.double(), approximate="tanh")
        print(f"gemm_f32 {m}x{n}x{k} bias+gelu:", _err(AF.linear_fwd(a, w, b, gelu=True, precision="fp32"), g))


def fam_gemm():
    import torch
    import avformer_b200 as A
    AF = A.functional
    torch.manual_seed(2)
    # first: identity-like structured test to expose layout / descriptor mistakes
    for (m, n, k) in ((128, 128, 64), (128, 128, 256), (256, 256, 128), (300, 768, 256), (8704, 1536, 512), (1000, 64, 1024), (77, 128, 512)):
        a = (torch.randn(m, k, device="cuda")).bfloat16()
        w = (torch.randn(n, k, device="cuda") / k ** 0.5).bfloat16()
        b, r = torch.randn(n, device="cuda"), torch.randn(m, n, device="cuda")
        ref = a.double() @ w.double().t()
        t0 = time.time()
        c = AF.linear_fwd(a, w, precision="bf16")
        torch.cuda.synchronize()
        print(f"gemm_umma {m}x{n}x{k} plain f32out:", _err(c, ref), f"({(time.time() - t0) * 1e3:.1f} ms)")
        c = AF.linear_fwd(a, w, out_dtype=torch.bfloat16, precision="bf16")
        print(f"gemm_umma {m}x{n}x{k} plain bf16out:", _err(c.float(), ref))
        c = AF.linear_fwd(a, w, b, r, precision="bf16")
        print(f"gemm_umma {m}x{n}x{k} bias+res:", _err(c, ref + b.double() + r.double()))
        c = AF.linear_fwd(a, w, b, gelu=True, out_dtype=torch.bfloat16, precision="bf16")
        print(f"gemm_umma {m}x{n}x{k} bias+gelu bf16out:", _err(c.float(), torch.nn.functional.gelu(ref + b.double(), approximate="tanh")))
    # timings (CUDA events, 10 launches after 2 warm-ups)
    def timeit(fn, reps=10):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    for (m, n, k, od, gelu, res) in ((8704, 1536, 512, torch.bfloat16, False, False), (8704, 512, 512, torch.float32, False, True),
                                     (8704, 1024, 512, torch.bfloat16, True, False), (8704, 512, 1024, torch.float32, False, True),
                                     (401408, 768, 256, torch.bfloat16, False, False), (401408, 256, 256, torch.float32, False, True),
                                     (401408, 512, 256, torch.bfloat16, True, False), (401408, 256, 512, torch.float32, False, True),
                                     (6144, 768, 128, torch.bfloat16, False, False), (6144, 128, 256, torch.float32, False, True)):
        a = torch.randn(m, k, device="cuda").bfloat16()
        w = torch.randn(n, k, device="cuda").bfloat16()
        b = torch.randn(n, device="cuda")
        r = torch.randn(m, n, device="cuda") if res else None
        ms = timeit(lambda: AF.linear_fwd(a, w, b, r, gelu=gelu, out_dtype=od, precision="bf16"))
        ms2 = timeit(lambda: torch.matmul(a, w.t()))
        print(f"gemm_umma {m}x{n}x{k} out={str(od)[6:]} gelu={gelu} res={res}: {ms * 1e3:8.1f} us {2 * m * n * k / ms / 1e9:7.1f} TFLOP/s | cuBLAS plain: {ms2 * 1e3:8.1f} us")
    for (ns, nt, h, dh) in ((8192, 49, 8, 32), (512, 17, 8, 64), (512, 12, 8, 32)):
        qkv = torch.randn(ns * nt, 3 * h * dh, device="cuda").bfloat16()
        ms = timeit(lambda: AF.attention_fwd(qkv, ns, nt, h, dh))
        print(f"attention_mma seq={ns} tok={nt} dh={dh}: {ms * 1e3:8.1f} us  {4 * ns * h * nt * nt * dh / ms / 1e9:7.1f} TFLOP/s")


def fam_attention():
    import torch
    import avformer_b200 as A
    AF = A.functional
    torch.manual_seed(3)
    for (ns, nt, h, dh) in ((5, 49, 8, 32), (7, 17, 8, 64), (9, 12, 8, 32), (3, 33, 8, 64), (4, 9, 8, 64)):
        qkv = torch.randn(ns * nt, 3 * h * dh, device="cuda")
        q, k, v = (t.reshape(ns, nt, h, dh).permute(0, 2, 1, 3).double() for t in qkv.split(h * dh, -1))
        ref = (torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, -1) @ v).permute(0, 2, 1, 3).reshape(ns * nt, h * dh)
        print(f"attention fp32 seq={ns} tok={nt} dh={dh}:", _err(AF.attention_fwd(qkv, ns, nt, h, dh), ref))
        qb = qkv.bfloat16()
        q, k, v = (t.reshape(ns, nt, h, dh).permute(0, 2, 1, 3).double() for t in qb.split(h * dh, -1))
        ref = (torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, -1) @ v).permute(0, 2, 1, 3).reshape(ns * nt, h * dh)
        print(f"attention bf16 seq={ns} tok={nt} dh={dh}:", _err(AF.attention_fwd(qb, ns, nt, h, dh).float(), ref))


def _stack(precision):
    import torch
    import avformer_b200 as A
    from oracle import avformer_oracle as O
    T, B, seed = 16, 2, 116
    p64 = O.cast_params(O.make_state_dict(seed, T, hot_path_only=True), torch.float64)
    stage3, frame, audio = O.synth_hot_path_inputs(seed, B, T, torch.float64)
    ref = O.hot_path_forward(stage3, frame, audio, p64, T)
    m = A.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU")
    m.load_state_dict(O.make_state_dict(seed, T), strict=True)
    m = m.cuda().eval().set_precision(precision)
    with torch.no_grad():
        vm = m.video_model.video_model
        s_out = vm.s_former.sformer(stage3.float().cuda())
        print(f"[{precision}] sformer_out:", _err(s_out, ref["sformer_out"]))
        cls = vm.t_former(frame.float().cuda())
        print(f"[{precision}] tformer_cls:", _err(cls, ref["tformer_cls"]))
        _, vt = m.video_model.au_head(cls)
        print(f"[{precision}] video_tokens (from own cls):", _err(vt, ref["video_tokens"]))
        _, at = m.audio_model.au_head(audio.float().cuda())
        print(f"[{precision}] audio_tokens:", _err(at, ref["audio_tokens"]))
        logits = m.au_head(torch.cat([at, vt], 2))
        print(f"[{precision}] logits:", _err(logits, ref["logits"]))
        print(f"[{precision}] decision flips:", int(((logits.cpu() > 0) != (ref["logits"] > 0)).sum()), "of", logits.numel())


def fam_stack_fp32():
    _stack("fp32")


def fam_stack_bf16():
    _stack("bf16")


def fam_model():
    import numpy as np
    import torch
    import avformer_b200 as A
    from oracle import avformer_oracle as O
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for T in (8, 16):
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"full_T{T}.npz")))
        seed, B = int(g["seed"]), int(g["batch"])
        m = A.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU").set_clip_length(T)
        m.load_state_dict(O.make_state_dict(seed, T), strict=True)
        m = m.cuda().eval()
        clip, audio, labels = O.synth_inputs(seed, B, T)
        for prec in ("fp32", "bf16"):
            m.set_precision(prec)
            with torch.no_grad():
                out = m({"clip": clip.cuda(), "audio_features": audio.cuda(), "AU": labels.cuda()})
                loss = m.get_au_loss(out, labels.cuda())
            print(f"model T={T} [{prec}] logits vs reference golden:", _err(out[:, :12], torch.from_numpy(g["logits"])),
                  "pad", out[:, 12:].abs().max().item(), "loss", loss.item(), "ref", float(g["loss"]),
                  "flips", int(((out[:, :12].cpu() > 0).numpy() != (g["logits"] > 0)).sum()))


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":
        rc = 0
        for f in FAMILIES:
            print(f"===== {f} =====", flush=True)
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), f], timeout=240)
                print(f"----- {f}: exit {r.returncode}", flush=True)
                rc |= r.returncode != 0
            except subprocess.TimeoutExpired:
                print(f"----- {f}: TIMEOUT", flush=True)
                rc = 1
        sys.exit(rc)
    globals()["fam_" + which]()


if __name__ == "__main__":
    main()
