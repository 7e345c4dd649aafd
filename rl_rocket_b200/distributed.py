"""Multi-GPU plumbing: environments are independent, so a rank just owns a contiguous slice of the global env index
space (no data-path collective).  The only exchange is the per-rollout all-gather of each rank's episode-statistics
record (returns / done counts; 8 x int64 = 64 B per rank): `StatsExchange` does it with ONE small kernel per rank that stores
the record straight into the peers' buffers over NVLink (C ABI rlr_exchange_*, CUDA-IPC mapped peer memory);
`gather_stats` is the plain torch.distributed collective (NCCL on GPUs, gloo in the CPU tests) and the fallback."""
import ctypes as C

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


class StatsExchange:
    """Per-rollout all-gather of the int64[NSTATS] statistics records of the ranks of one node through NVLink peer memory.

        ex = StatsExchange(device)               # once, after init_process_group (collective: exchanges the IPC handles)
        table = ex.all_gather(batch.stats)       # [world, NSTATS] int64 device tensor, enqueued on the current stream, no host sync

    Falls back to `dist.all_gather_into_tensor` when peer memory cannot be mapped (`ex.peer_memory` is then False)."""

    def __init__(self, device, group=None):
        from . import _native as N
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.out = torch.zeros((self.world, NSTATS), dtype=torch.int64, device=self.device)
        self.lib = N.load()
        self.handle = C.c_void_p()
        self.peer_memory = False
        ok = 1
        with torch.cuda.device(self.device):
            if self.lib.rlr_exchange_create(C.byref(self.handle), self.rank, self.world) != 0:
                ok, self.handle = 0, C.c_void_p()
            mine = (C.c_ubyte * 64)()
            if ok and self.lib.rlr_exchange_handle(self.handle, mine) != 0:
                ok = 0
            if self.world > 1:
                got = [None] * self.world
                dist.all_gather_object(got, (ok, bytes(mine)), group=group)
                ok = int(all(g[0] for g in got))
                if ok:
                    blob = b"".join(g[1] for g in got)
                    ok = int(self.lib.rlr_exchange_connect(self.handle, blob) == 0)
                flags = [None] * self.world
                dist.all_gather_object(flags, ok, group=group)       # all ranks take the same path
                ok = int(all(flags))
        self.peer_memory = bool(ok)
        self.error = None if ok else self.lib.rlr_last_error().decode()

    def all_gather(self, local_stats):
        if local_stats.numel() != NSTATS or local_stats.dtype != torch.int64 or local_stats.device != self.device:
            raise ValueError("local_stats must be the int64[NSTATS] accumulator on the exchange's device")
        if not self.peer_memory:
            if self.world > 1:
                dist.all_gather_into_tensor(self.out, local_stats.reshape(1, NSTATS).contiguous(), group=self.group)
            else:
                self.out.copy_(local_stats.reshape(1, NSTATS))
            return self.out
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        prev = torch.cuda.current_device()
        if prev != idx:
            torch.cuda.set_device(idx)
        rc = self.lib.rlr_exchange_allgather(self.handle, local_stats.data_ptr(), self.out.data_ptr(), torch.cuda.current_stream(idx).cuda_stream)
        if prev != idx:
            torch.cuda.set_device(prev)
        if rc:
            raise RuntimeError("rlr_exchange_allgather: " + self.lib.rlr_last_error().decode())
        return self.out

    def check(self):
        """Host-synchronising: raises if a peer's record ever failed to arrive."""
        if self.peer_memory and self.lib.rlr_exchange_status(self.handle) != 0:
            raise RuntimeError(self.lib.rlr_last_error().decode())

    def close(self):
        if self.handle:
            self.lib.rlr_exchange_destroy(self.handle)
            self.handle = C.c_void_p()


def aggregate(per_rank):
    tot = per_rank.sum(0).tolist()
    d = dict(zip(STAT_NAMES, tot))
    d["ret_sum"] = d.pop("ret_sum_fx") / 1048576.0
    d["mean_return"] = d["ret_sum"] / d["done"] if d["done"] else float("nan")
    d["mean_length"] = d["len_sum"] / d["done"] if d["done"] else float("nan")
    return d
