"""A/B of the SM partition of the graphed hot path: python tools/split_probe.py [clips] "116,32" "108,40" ...
Times GraphedHotPath (512 clips x 16 frames by default) with the SFormer after the chain (no split) and next to it on
disjoint SM sets, and checks that the results are bit-identical."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avformer_b200 as A

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
splits = [None] + [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]]
T = 16
torch.manual_seed(1234)
model = A.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU").cuda().eval().set_precision("bf16")
g = torch.Generator(device="cpu").manual_seed(7)
stage3 = torch.clamp(torch.randn(B * T, 256, 7, 7, generator=g) * 1.7 + 0.6, min=0).bfloat16().cuda()
frame = (torch.randn(B * T, 512, generator=g).abs() * 1.2).bfloat16().cuda()
audio = torch.randn(B, 512, generator=g).abs().cuda()
ref = None
with torch.no_grad():
    for sp in splits:
        gr = A.GraphedHotPath(model, stage3, frame, audio, sm_split=sp)
        for _ in range(3):
            out = gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(50):
            out = gr.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        res = [t.clone() for t in out]
        same = True if ref is None else all(torch.equal(a, b) for a, b in zip(ref, res))
        if ref is None:
            ref = res
        print(f"split {sp}: {ms:.3f} ms/step  {B / ms * 1e3:.0f} clips/s  identical to unsplit: {same}", flush=True)
