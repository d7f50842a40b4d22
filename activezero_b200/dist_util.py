"""One-process-per-GPU plumbing shared by ``bench.py`` and the training harness.

The hot path shards by stereo pair with NO data-path collective:
every op is independent per pair, and the reference computes the masked-mean
denominator of the reprojection loss per rank (``utils/reprojection.py:118``; DDP
averages gradients afterwards).  ``torch.distributed`` is used only for
rendezvous, barriers and reducing timings (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_from_env(backend: str, device=None):
    """Join the job torchrun describes (RANK/WORLD_SIZE/MASTER_*); no-op for a single process."""
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device} if (device is not None and backend == "nccl") else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def shard_pairs(total_pairs: int, rank: int, world: int):
    """Contiguous, even split of ``total_pairs`` stereo pairs over ``world`` ranks
    (what ``DistributedSampler`` gives the reference trainer, ``train.py:444-449``):
    returns (first_pair, n_pairs) of this rank; the first ``total % world`` ranks
    take one extra pair."""
    if not (0 <= rank < world) or total_pairs < 0:
        raise ValueError("bad rank/world/total")
    base, extra = divmod(total_pairs, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(values, device="cpu"):
    """Element-wise MAX of a list of floats over all ranks (timings are reported as the
    slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def sum_over_ranks(values, device="cpu"):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t]


def aggregate_throughput(units_this_rank: float, elapsed_ms_this_rank: float, device="cpu") -> float:
    """Whole-job throughput = units all ranks processed / slowest rank's time."""
    (total,) = sum_over_ranks([units_this_rank], device)
    (ms,) = max_over_ranks([elapsed_ms_this_rank], device)
    return total / (ms * 1e-3)
