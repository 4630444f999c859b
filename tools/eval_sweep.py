"""BASELINE config 5: Aff-Wild2-shaped synthetic evaluation sweep — 32-frame clips, batch 1..1024, hot-path clips/s per batch
size (device-resident inputs, CUDA-graph replay and eager launches), one JSON line per batch size.

    python tools/eval_sweep.py [max_batch] > profiles/rNN_eval_sweep_T32.jsonl
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avformer_b200 as A

T = 32
max_b = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(2024)
dev = torch.device("cuda")
model = A.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU").set_clip_length(T).to(dev).eval().set_precision("bf16")
flop_per_clip = T * 53_838_848 + 421_926_912 + 2 * 11_308_032 + 28_760_064


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


B = 1
with torch.no_grad():
    while B <= max_b:
        g = torch.Generator().manual_seed(B)
        stage3 = torch.clamp(torch.randn(B * T, 256, 7, 7, generator=g) * 1.7 + 0.6, min=0).bfloat16().to(dev)
        frame = (torch.randn(B * T, 512, generator=g).abs() * 1.2).bfloat16().to(dev)
        audio = torch.randn(B, 512, generator=g).abs().to(dev)
        reps = 50 if B <= 64 else 10
        ms_eager = timed(lambda: model.hot_path(stage3, frame, audio, want_decisions=True), reps)
        graphed = A.GraphedHotPath(model, stage3, frame, audio)
        ms_graph = timed(graphed.replay, reps)
        print(json.dumps({"config": "eval sweep, T=32, hot path", "batch": B, "ms_graph": ms_graph, "ms_eager": ms_eager,
                          "clips_per_s_graph": B / ms_graph * 1e3, "clips_per_s_eager": B / ms_eager * 1e3,
                          "tflops_graph": B * flop_per_clip / ms_graph / 1e9}), flush=True)
        del graphed
        B *= 2
