"""train.py-equivalent runner: the reference's call pattern — build the model by name, Adam(lr, weight_decay)
over model.parameters(), per-step ``zero_grad / model(x) / get_au_loss / backward / step`` (train.py:206-236), per-epoch
``evaluate`` with MultiLabelAccF1 (train.py:106-169) and ``latest.pth`` / ``best.pth`` checkpoints whose keys are the
reference's own (train.py:97,247; loadable with ``load_state_dict(strict=True)`` on either side) — on a synthetic
Aff-Wild2-shaped feed (dataloader/aff2compdataset.py:114-247 needs LMDB/JPEG/torchaudio data that is not available
offline).  One process per GPU under torchrun: clips are sharded, FusedAdam all-reduces the flat gradient bucket.

    python -m torch.distributed.run --nproc-per-node N -m avformer_b200.runner --epochs 2 --batch 16 --frames 8
"""
from __future__ import annotations

import argparse
import os
import time
from typing import Dict, Iterator

import torch
import torch.distributed as dist

from . import dp
from .avformer import TwoStreamAuralVisualFormer
from .metrics import MultiLabelAccF1
from .optim import FusedAdam


class SyntheticAff2(torch.utils.data.IterableDataset):
    """Batches with the keys and shapes the reference's loader yields (train.py:207-218): 'clip' [B,3,T,112,112],
    'audio_features' [B,1,64,1001], 'AU' [B,12] in {0,1} (a row's first label -1 = unlabelled frame), 'EX', 'VA', 'Index'.
    The AU labels are a fixed random linear function of clip / audio statistics so that there is something to learn."""

    def __init__(self, n_batches: int, batch: int, frames: int, seed: int = 0, image: int = 112, unlabeled: float = 0.05):
        self.n_batches, self.batch, self.frames, self.seed, self.image, self.unlabeled = n_batches, batch, frames, seed, image, unlabeled

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        g = torch.Generator().manual_seed(self.seed)
        w = torch.randn(12, 4, generator=g)
        for i in range(self.n_batches):
            clip = torch.randn(self.batch, 3, self.frames, self.image, self.image, generator=g)
            audio = torch.randn(self.batch, 1, 64, 1001, generator=g)
            z = torch.randn(self.batch, 4, generator=g)
            clip = clip + z[:, :3, None, None, None]                      # the latent shifts the colour channels ...
            audio = audio + z[:, 3, None, None, None]                     # ... and the spectrogram level
            au = ((z @ w.t()) > 0.5).float()
            au[torch.rand(self.batch, generator=g) < self.unlabeled, 0] = -1.0
            yield {"clip": clip, "audio_features": audio, "AU": au, "EX": torch.zeros(self.batch), "VA": torch.zeros(self.batch, 2),
                   "Index": torch.arange(i * self.batch, (i + 1) * self.batch)}


@torch.no_grad()
def evaluate(model, batches, device, rank=0, world=1):
    """train.py:106-169 for task 'AU': eval mode, loss + MultiLabelAccF1(ignore_index=-1), score = 0.5 f1 + 0.5 acc."""
    model.eval()
    metric = MultiLabelAccF1(ignore_index=-1)
    total, n = 0.0, 0
    for data in batches:
        data = dp.shard_batch(data, rank, world)
        x = {k: v.to(device, non_blocking=True) for k, v in data.items()}
        result = model(x)
        labels = x["AU"].float()
        if bool((labels[:, 0] != -1).any()):
            total += model.get_au_loss(result, labels).item()
            n += 1
        metric.update_from_logits(result, labels)
    acc, f1 = metric.get()                       # counters are summed over the ranks inside get()
    model.train()
    return {"AU:acc": acc, "f1": f1, "score": 0.5 * f1 + 0.5 * acc, "loss": total / max(n, 1)}


def save_checkpoint(state, filepath, filename):
    os.makedirs(filepath, exist_ok=True)
    torch.save(state, os.path.join(filepath, filename))


def train(args) -> Dict[str, float]:
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    torch.manual_seed(args.seed)                                  # identical initial weights on every rank
    model = TwoStreamAuralVisualFormer(modality="A;V;M", video_pretrained=False, audio_pretrained=False, task="AU")
    model.video_model.video_model.config_modality("A;V")          # synthetic clips carry no mask channel
    model.set_clip_length(args.frames).to(device).set_precision(args.precision)
    if args.resume:
        model.load_state_dict(torch.load(args.resume, map_location="cpu"), strict=True)
    if args.freeze_backbones:                                     # the reference default once pretrained sub-models are loaded (models/avformer.py:78-85)
        for p in list(model.video_model.parameters()) + list(model.audio_model.parameters()):
            p.requires_grad = False
    model.train()
    optimizer = FusedAdam(model.parameters(), lr=args.learning_rate, weight_decay=args.weight_decay)       # train.py:334
    best, history = -1.0, []
    for epoch in range(args.epochs):
        t0, seen, loss_sum, steps = time.time(), 0, 0.0, 0
        for data in SyntheticAff2(args.steps_per_epoch, args.batch, args.frames, seed=args.seed + epoch, image=args.image):
            data = dp.shard_batch(data, rank, world)
            x = {k: v.to(device, non_blocking=True) for k, v in data.items()}        # train.py:216-218 copies every key
            optimizer.zero_grad()
            result = model(x)
            loss = model.get_au_loss(result, x["AU"].float())
            loss.backward()
            optimizer.step()
            loss_sum += loss.item()
            steps += 1
            seen += args.batch
        if rank == 0:
            save_checkpoint(model.state_dict(), args.checkpoint_path, "latest.pth")                       # train.py:247
        scores = evaluate(model, SyntheticAff2(args.eval_steps, args.batch, args.frames, seed=10_000 + args.seed, image=args.image),
                          device, rank, world)
        scores.update(epoch=epoch, train_loss=loss_sum / max(steps, 1), clips_per_s=seen / (time.time() - t0))
        history.append(scores)
        if rank == 0:
            print({k: (round(v, 5) if isinstance(v, float) else v) for k, v in scores.items()}, flush=True)
            if scores["score"] > best:
                best = scores["score"]
                save_checkpoint(model.state_dict(), args.checkpoint_path, "best.pth")
    return {"best": best, "history": history}


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--steps-per-epoch", type=int, default=8)
    ap.add_argument("--eval-steps", type=int, default=2)
    ap.add_argument("--batch", type=int, default=16, help="global batch (sharded over the ranks)")
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--image", type=int, default=112)
    ap.add_argument("--learning-rate", type=float, default=5e-4)
    ap.add_argument("--weight-decay", type=float, default=5e-5)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--freeze-backbones", action="store_true")
    ap.add_argument("--checkpoint-path", default="checkpoints")
    ap.add_argument("--resume", default=None)
    ap.add_argument("--seed", type=int, default=0)
    return ap.parse_args(argv)


def main(argv=None):
    out = train(parse_args(argv))
    if dist.is_initialized():
        dist.destroy_process_group()
    return out


if __name__ == "__main__":
    main()
