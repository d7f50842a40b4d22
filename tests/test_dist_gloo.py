"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the pair sharding and
the max-over-ranks / whole-job throughput reductions bench.py uses with NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from activezero_b200 import dist_util


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w = dist_util.init_from_env("gloo")
    assert (r, w) == (rank, world) and dist.is_initialized()
    first, count = dist_util.shard_pairs(17, rank, world)
    # each rank "processes" its pairs: a per-pair checksum, summed across ranks
    mine = float(sum(p * p for p in range(first, first + count)))
    (total,) = dist_util.sum_over_ranks([mine])
    ms = 10.0 * (rank + 1)  # rank 1 is the slow one
    (slowest,) = dist_util.max_over_ranks([ms])
    thr = dist_util.aggregate_throughput(count, ms)
    dist_util.barrier()
    out[rank] = (first, count, total, slowest, thr)
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_reductions():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][:2] == (0, 9) and out[1][:2] == (9, 8)  # contiguous, disjoint, covering 17 pairs
    want = float(sum(p * p for p in range(17)))
    for r in range(world):
        assert out[r][2] == want          # no pair lost or duplicated
        assert out[r][3] == 20.0          # max over ranks
        assert out[r][4] == pytest.approx(17 / 20e-3)  # all units / slowest time


@pytest.mark.parametrize("total,world", [(8, 1), (8, 8), (7, 4), (3, 8), (0, 2), (64, 3)])
def test_shard_pairs_partition(total, world):
    seen = []
    for r in range(world):
        first, count = dist_util.shard_pairs(total, r, world)
        seen += list(range(first, first + count))
        assert count in (total // world, total // world + 1)
    assert seen == list(range(total))
    with pytest.raises(ValueError):
        dist_util.shard_pairs(4, 2, 2)


def test_single_process_passthrough():
    assert dist_util.max_over_ranks([3.0, 1.0]) == [3.0, 1.0]
    assert dist_util.aggregate_throughput(8, 2.0) == pytest.approx(4000.0)
    assert torch.distributed.is_available()
