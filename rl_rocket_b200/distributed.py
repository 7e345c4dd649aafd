"""Multi-GPU plumbing: environments are independent, so a rank just owns a contiguous slice of the global env index
space (no data-path collective).  The only exchange is the per-rollout all-gather of each rank's episode-statistics
record (returns / done counts; 8 x int64 = 64 B per rank) -- NCCL over NVLink on GPUs, gloo in the CPU tests."""
import torch
import torch.distributed as dist

from ._native import NSTATS, STAT_NAMES


def shard(total_envs, world_size, rank):
    """Contiguous slice [offset, offset + n) of the global env indices owned by `rank` (remainder spread over the first ranks)."""
    if total_envs < 0 or world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad shard request")
    base, rem = divmod(total_envs, world_size)
    n = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, n


def gather_stats(local_stats, group=None):
    """all_gather of the per-rank statistics vectors -> (per_rank [W, NSTATS] tensor, aggregate dict)."""
    if local_stats.numel() != NSTATS or local_stats.dtype != torch.int64:
        raise ValueError("local_stats must be the int64[NSTATS] device accumulator")
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        out = torch.zeros((world, NSTATS), dtype=torch.int64, device=local_stats.device)
        dist.all_gather_into_tensor(out, local_stats.reshape(1, NSTATS).contiguous(), group=group)
    else:
        out = local_stats.reshape(1, NSTATS).clone()
    return out, aggregate(out)


def aggregate(per_rank):
    tot = per_rank.sum(0).tolist()
    d = dict(zip(STAT_NAMES, tot))
    d["ret_sum"] = d.pop("ret_sum_fx") / 1048576.0
    d["mean_return"] = d["ret_sum"] / d["done"] if d["done"] else float("nan")
    d["mean_length"] = d["len_sum"] / d["done"] if d["done"] else float("nan")
    return d
