#!/usr/bin/env python
"""Bare host->device copy bandwidth per GPU at N ranks (is the end-to-end scaling limit the host, the PCIe fabric, or us?).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/h2d_probe.py

Every rank copies one bench step's worth of frames (16 x 3840x2160x3 = 398 MB) from its own pinned buffer to its own GPU, 10 times,
all ranks at once (barrier before and after); rank 0 prints one JSON line with the per-rank and aggregate GB/s.  Nothing of the
library is involved: this is the ceiling the e2e leg of bench.py can reach with BGR24 host frames.
"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = 16 * 2160 * 3840 * 3
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.random_(0, 255)
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([nbytes * reps / dt / 1e9], device="cuda")
    if world > 1:
        all_ = [torch.zeros_like(gbs) for _ in range(world)]
        dist.all_gather(all_, gbs)
        vals = [float(v) for v in all_]
    else:
        vals = [float(gbs)]
    if rank == 0:
        print(json.dumps(dict(probe="bare pinned H2D, 398 MB x 10 per rank, all ranks concurrently", n_gpus=world, per_gpu_gbs=[round(v, 2) for v in vals],
                              aggregate_gbs=round(sum(vals), 2), frames_per_s_ceiling_bgr24=round(sum(vals) * 1e9 / (2160 * 3840 * 3), 1),
                              host_cpus=os.cpu_count())))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
