"""A few eager training steps of the hot path (BASELINE config 4 shape: 64 clips x 16 frames) for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avformer_b200 as A

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, T = 64, 16
torch.manual_seed(0)
dev = torch.device("cuda")
model = A.TwoStreamAuralVisualFormer(video_pretrained=False, audio_pretrained=False, task="AU").to(dev).set_precision("bf16").train()
g = torch.Generator().manual_seed(1)
stage3 = torch.clamp(torch.randn(B * T, 256, 7, 7, generator=g) * 1.7 + 0.6, min=0).bfloat16().to(dev).requires_grad_(True)
frame = (torch.randn(B * T, 512, generator=g).abs() * 1.2).to(dev).requires_grad_(True)
audio = torch.randn(B, 512, generator=g).abs().to(dev).requires_grad_(True)
labels = (torch.rand(B, 12, generator=g) < 0.3).float().to(dev)
probe = torch.randn(B * T, 256, 7, 7, generator=g).bfloat16().to(dev) * 1e-3
hot = [p for k, p in model.named_parameters() if ".resnet." not in k and "s_former.conv1" not in k and "s_former.bn1" not in k and "s_former.layer" not in k]
opt = A.FusedAdam(hot, lr=5e-4, weight_decay=5e-5)
for i in range(steps):
    opt.zero_grad()
    s_out, out21 = model.hot_path_train(stage3, frame, audio)
    loss = model.get_au_loss(out21, labels)
    torch.autograd.backward([loss, s_out], [None, probe])
    opt.step()
torch.cuda.synchronize()
print("loss", loss.item(), "launches", A._lib.lib().avf_launch_count())
