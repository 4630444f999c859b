"""Data-parallel plumbing on CPU: 2 ranks over gloo (the GPU box runs the same code over NCCL).  Covers SURVEY.md §8(e):
contiguous clip shards (ragged allowed), the evaluation-time logit gather, the gradient sum-all-reduce + mean, and that a
sharded evaluation reproduces the unsharded metric bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import avformer_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import avformer_b200 as A          # imports without a GPU; only the dp helpers are exercised here
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        n_total = 7                                                     # ragged: rank 0 takes 4 clips, rank 1 takes 3
        g = torch.Generator().manual_seed(5)
        full = {"clip": torch.randn(n_total, 3, 2, 4, 4, generator=g), "audio_features": torch.randn(n_total, 1, 4, 9, generator=g),
                "AU": (torch.rand(n_total, 12, generator=g) < 0.3).float()}
        shard = A.dp.shard_batch(full, rank, world)
        lo, hi = A.dp.shard_bounds(n_total, rank, world)
        assert shard["clip"].shape[0] == hi - lo and torch.equal(shard["clip"], full["clip"][lo:hi])
        # "logits" of a shard: a deterministic function of the clips, so the gathered result is checkable
        local = torch.zeros(hi - lo, 21)
        local[:, :12] = shard["clip"].flatten(1)[:, :12] * 3.0 - 0.5
        gathered = A.dp.gather_logits(local, n_total)
        expect = torch.zeros(n_total, 21)
        expect[:, :12] = full["clip"].flatten(1)[:, :12] * 3.0 - 0.5
        assert torch.equal(gathered, expect)
        # sharded evaluation == unsharded evaluation (decisions, accuracy, F1 of metrics/accf1.py)
        dec = O.decisions(gathered[:, :12])
        acc, f1, _ = O.multilabel_acc_f1(full["AU"].numpy(), dec, ignore_index=-1)
        acc0, f10, _ = O.multilabel_acc_f1(full["AU"].numpy(), O.decisions(expect[:, :12]), ignore_index=-1)
        assert acc == acc0 and f1 == f10
        # gradient bucket: sum over ranks / world == gradient of the mean loss over equal-sized shards
        bucket = torch.arange(10, dtype=torch.float32) * (rank + 1)
        A.dp.allreduce_mean_(bucket)
        assert torch.allclose(bucket, torch.arange(10, dtype=torch.float32) * 1.5)
        # pipelined gather (bench.py at N > 1): three batches in flight over two buffer pairs, results valid after wait() / two submits
        pipe = A.dp.PipelinedLogitGather()
        outs = []
        for step in range(3):
            mine = torch.full((4, 21), float(10 * step + rank))
            outs.append((step, pipe.submit(mine)))
            if step == 2:                                               # buffer pair 0 is being reused: its first result was waited for
                pass
        pipe.wait()
        for step, got in outs[1:]:                                      # (outs[0] shares its buffers with step 2)
            for r in range(world):
                assert torch.equal(got[4 * r: 4 * r + 4], torch.full((4, 21), float(10 * step + r)))
        # segment-wise gradient reduction overlapped with backward (optim.FusedAdam at N > 1): three segments of a flat bucket,
        # parameters reported stack by stack in backward order; the last segment is never reported and is reduced by finish()
        prm = [torch.nn.Parameter(torch.zeros(n)) for n in (3, 5, 2, 4, 6)]
        seg_of = {id(prm[0]): 0, id(prm[1]): 0, id(prm[2]): 1, id(prm[3]): 1, id(prm[4]): 2}
        flat = torch.arange(24, dtype=torch.float32) * (rank + 1)                   # offsets 0,3 | 8,10 | 16 (8-aligned like FusedAdam)
        red = A.dp.SegmentReducer(flat, bounds=[(0, 8), (8, 16), (16, 24)], counts=[2, 2, 1], seg_of=seg_of)
        red.arm()
        red.ready([prm[3]])
        assert red.launch_order == []                                               # segment 1 still waits for prm[2]
        red.ready([prm[0], prm[1]])
        assert red.launch_order == [0]
        red.ready([prm[2]])
        assert red.launch_order == [0, 1]
        red.ready([prm[2]])                                                         # a second report of a launched segment is ignored
        red.finish()
        assert red.launch_order == [0, 1, 2]
        assert torch.equal(flat, torch.arange(24, dtype=torch.float32) * 3.0)       # (1 + 2) x: every element summed exactly once
        red.disarm()
        red.ready([prm[4]])                                                         # not armed: nothing happens
        assert torch.equal(flat, torch.arange(24, dtype=torch.float32) * 3.0)
        # dropout seeds: identical torch seeds on every rank must still give rank-dependent mask seeds (encoder._rank_salt)
        from avformer_b200.encoder import _rank_salt
        salts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(salts, torch.tensor([_rank_salt()], dtype=torch.int64))
        assert int(salts[0]) == 0 and len({int(v) for v in salts}) == world
        tr = A.Transformer(128, 1, 8, 32, 256, dropout=0.2).train()
        torch.manual_seed(1234)
        _, seed, _ = tr.dropout_state()
        seeds = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(seeds, torch.tensor([seed], dtype=torch.int64))
        assert len({int(v) for v in seeds}) == world
        np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([1]))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_gather_and_allreduce(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0.npy") and os.path.exists(tmp_path / "ok1.npy")


def test_shard_bounds_cover_everything():
    import avformer_b200 as A
    for n in (1, 7, 64, 512, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [A.dp.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
