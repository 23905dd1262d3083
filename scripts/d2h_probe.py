#!/usr/bin/env python
"""What the box can move between HBM and pinned host memory when N ranks copy at once (the ceiling of bench.py's
e2e leg at N GPUs): contiguous 1 GiB cudaMemcpyAsync device->host (and host->device), one process per GPU, all
ranks at the same time.  Prints per-rank and aggregate GB/s.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/d2h_probe.py
"""
import os
import time

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << 30
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    out = {}
    for name, dst, src in (("d2h", h, d), ("h2d", d, h)):
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        reps = 8
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gbs = reps * n / dt * 1e-9
        t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        out[name] = (gbs, float(t.item()))
    if rank == 0:
        print("ranks %d: d2h %.1f GB/s on rank 0, %.1f GB/s aggregate; h2d %.1f GB/s on rank 0, %.1f GB/s aggregate"
              % (world, out["d2h"][0], out["d2h"][1], out["h2d"][0], out["h2d"][1]), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
