"""Time the gradient all-reduce of the hot path's flat bucket (10.4 M floats, 41.7 MB) alone: the one-kernel peer-memory version
(avf_grad_allreduce) against dist.all_reduce (NCCL), CUDA events, max over ranks.  Run under torchrun:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/allreduce_probe.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avformer_b200 as A  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 41658368 // 4
    ar = A.dp.PeerAllReduce(n)
    flat = torch.zeros(n, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(rank)
    src = torch.randn(n, device="cuda", generator=g)

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    ms_peer = timed(lambda: ar.reduce_())
    ms_nccl = timed(lambda: dist.all_reduce(flat))
    ar.grad.copy_(src)
    flat.copy_(src)
    ar.reduce_()
    dist.all_reduce(flat)
    ar.check()
    rel = ((ar.grad - flat).abs().max() / flat.abs().max()).item()
    if rank == 0:
        moved = 2 * (world - 1) / world * n * 4
        print(json.dumps({"probe": "grad_allreduce", "n_gpus": world, "bytes": n * 4, "peer_us": ms_peer * 1e3, "nccl_us": ms_nccl * 1e3,
                          "peer_nvlink_GBps_per_gpu_each_way": (world - 1) / world * n * 4 / (ms_peer * 1e-3) / 1e9,
                          "bus_bytes_per_gpu": moved, "max_rel_diff_vs_nccl": rel}))
    del ar
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
