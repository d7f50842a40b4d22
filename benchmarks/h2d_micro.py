#!/usr/bin/env python
"""The host link alone: every rank copies the same pinned buffer host -> device (and back) at the same time.
Explains bench.py's `e2e` curve: that number is 98 % host->device copy of the 3.2 GB of logits per step.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/h2d_micro.py

One cudaMemcpyAsync per transfer (torch Tensor.copy_ from pinned memory), CUDA events, barrier on both sides,
slowest rank reported; also prints each rank's CPU affinity and the GPU's NUMA node if sysfs exposes it."""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from activezero_b200 import dist_util  # noqa: E402


def main():
    rank, world, local = dist_util.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_util.init_from_env("nccl", dev)
    nbytes = 1 << 30
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    devbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    res = {}
    for name, (dst, src) in {"h2d": (devbuf, host), "d2h": (host, devbuf)}.items():
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist_util.barrier()
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        mine = 5 * nbytes / (a.elapsed_time(b) * 1e-3) / 1e9
        (slow,) = dist_util.max_over_ranks([-mine], dev)
        (tot,) = dist_util.sum_over_ranks([mine], dev)
        res[name] = {"slowest_rank_GBps": -slow, "sum_over_ranks_GBps": tot}
    try:
        aff = sorted(os.sched_getaffinity(0))
        aff_s = f"{aff[0]}-{aff[-1]} ({len(aff)} cpus)"
    except Exception:
        aff_s = "n/a"
    numa = "n/a"
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id  # torch >= 2.3
    except Exception:
        bus = None
    if rank == 0:
        print(json.dumps({"metric": "pinned host<->device copy, all ranks at once", "n_gpus": world, "bytes_per_copy": nbytes,
                          **res, "rank0_cpu_affinity": aff_s, "host_cpus": os.cpu_count()}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
