"""World-size-2 gloo test of the N>1 host path: env sharding + the per-rollout all-gather of episode statistics."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rl_rocket_b200.distributed import gather_stats, shard
    off, n = shard(1000, world, rank)
    local = torch.tensor([n * 10, n * 17, rank + 1, rank, 2 * rank, 0, (rank + 1) * 70, int(-12.5 * (rank + 1) * 1048576)], dtype=torch.int64)
    per_rank, agg = gather_stats(local)
    q.put((rank, off, n, per_rank.tolist(), agg))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stat_gather():
    world, port = 2, 29533
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, off0, n0, pr0, agg0), (r1, off1, n1, pr1, agg1) = res
    assert (off0, n0, off1, n1) == (0, 500, 500, 500)
    assert pr0 == pr1 and len(pr0) == 2                 # every rank sees every rank's record
    assert agg0 == agg1
    assert agg0["steps"] == 10000 and agg0["done"] == 3 and agg0["len_sum"] == 210
    assert agg0["ret_sum"] == pytest.approx(-37.5) and agg0["mean_return"] == pytest.approx(-12.5)
