"""GPU probe for the fused encoder kernel: SFormer (NCHW bf16 io) and 12-token stacks vs the fp64 oracle and vs the unfused path."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avformer_b200 as A
from oracle import avformer_oracle as O
AF = A.functional


def err(a, b):
    d = (a.double().cpu() - b.double().cpu()).abs()
    return f"max_abs={d.max().item():.3e} mean_abs={d.mean().item():.3e} ref_absmax={b.double().abs().max().item():.3e}"


def sformer_case(n_frames, seed=3):
    torch.manual_seed(seed)
    m = A.video.ResFormer(A.video.BasicBlock, [2, 2, 2, 2]).cuda().eval()
    m.spatial_transformer.precision = "bf16"
    fm = (torch.clamp(torch.randn(n_frames, 256, 7, 7) * 1.7 + 0.6, min=0)).bfloat16().cuda()
    with torch.no_grad():
        AF.set_fused_enabled(False)
        ref_unfused = m.sformer(fm).float()
        AF.set_fused_enabled(True)
        out = m.sformer(fm).float()
        torch.cuda.synchronize()
        # fp64 oracle of the same region
        p = {k: v.detach().double().cpu() for k, v in m.state_dict().items()}
        yr = O.sformer_tokens(fm.double().cpu(), p, "")
    print(f"sformer F={n_frames}: fused vs unfused {err(out, ref_unfused)}")
    if True:
        print(f"sformer F={n_frames}: fused vs fp64   {err(out, yr)}")
        print(f"sformer F={n_frames}: unfused vs fp64 {err(ref_unfused, yr)}")
    return out


def head_case(n_clips, seed=5):
    torch.manual_seed(seed)
    h = A.heads.former_AU_head(emb_dim=256, dropout=0.2).cuda().eval()
    h.corr_transformer.precision = "bf16"
    x = torch.randn(n_clips, 12, 256, device="cuda") * 2
    with torch.no_grad():
        AF.set_fused_enabled(False)
        a = h(x)
        AF.set_fused_enabled(True)
        b = h(x)
        torch.cuda.synchronize()
    print(f"fusion head B={n_clips}: fused vs unfused {err(b, a)}")


if __name__ == "__main__":
    for f in (2, 1, 7, 300):
        sformer_case(f)
    for b in (1, 10, 33, 512):
        head_case(b)
    # timing
    torch.manual_seed(0)
    m = A.video.ResFormer(A.video.BasicBlock, [2, 2, 2, 2]).cuda().eval()
    m.spatial_transformer.precision = "bf16"
    fm = (torch.clamp(torch.randn(8192, 256, 7, 7) * 1.7 + 0.6, min=0)).bfloat16().cuda()
    with torch.no_grad():
        for en in (False, True):
            AF.set_fused_enabled(en)
            for _ in range(3):
                m.sformer(fm)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(10):
                m.sformer(fm)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"sformer 8192 frames fused={en}: {ms*1e3:.1f} us  {8192*53838848/ms/1e9:.1f} TFLOP/s")
