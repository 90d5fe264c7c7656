#!/usr/bin/env python
"""Device->host copy bandwidth with 1, 2, 4, ... ranks copying at the same time (one process per GPU, torchrun).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
      tools/microbench/d2h_scaling.py [--mb 128] [--reps 40] [--no-bind]

Answers one question for bench.py's end-to-end legs: is the per-GPU rate of the output copies at N = 8 (11 - 13 GB/s per
GPU against 47 GB/s alone) a property of the box (aggregate host-side limit) or of this repo's pipeline?  Every rank
copies `--mb` MB from its GPU into its own page-locked buffer `--reps` times, timed with CUDA events; in phase m only
ranks < m copy.  With / without binding the process to the GPU's NUMA node first (bind_to_gpu_numa).  Rank 0 prints one
JSON line per phase.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=128)
    ap.add_argument("--reps", type=int, default=40)
    ap.add_argument("--no-bind", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from clearwater_riverine_b200.backend import bind_to_gpu_numa
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    cpus = None if args.no_bind else bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.mb << 20
    dev = torch.empty(n, dtype=torch.uint8, device="cuda").fill_(1)
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.fill_(0)                                    # first touch here, after the binding
    stream = torch.cuda.Stream()
    m = 1
    phases = []
    while m <= world:
        phases.append(m)
        m *= 2
    for direction in ("d2h", "h2d"):
        for m in phases:
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            gbs = 0.0
            if rank < m:
                with torch.cuda.stream(stream):
                    for _ in range(3):
                        (host if direction == "d2h" else dev).copy_(dev if direction == "d2h" else host, non_blocking=True)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(args.reps):
                        (host if direction == "d2h" else dev).copy_(dev if direction == "d2h" else host, non_blocking=True)
                    e1.record()
                stream.synchronize()
                gbs = n * args.reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
            t = torch.tensor([gbs], device="cuda")
            if world > 1:
                all_t = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(all_t, t)
                rates = [float(x.item()) for x in all_t]
            else:
                rates = [gbs]
            if rank == 0:
                act = rates[:m]
                print(json.dumps({"direction": direction, "ranks_copying": m, "mb": args.mb, "numa_bound": cpus is not None,
                                  "gbs_per_rank": [round(r, 1) for r in act], "gbs_total": round(sum(act), 1)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
