#!/usr/bin/env python
"""bench.py -- env-steps/sec of the 6DOF environment step on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # ours (CUDA path through the C ABI)
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's Python/scipy CPU path (oracle port)

A "step" is one pass of the hot path over one batch: ONE launch (rlr_step_chained; rlr_step with --serialized) advancing every
environment of the rank by one agent step (action mapping -> adaptive RK45 solve with ground event -> reward -> TimeLimit -> auto-reset).
N = 1 runs BASELINE config 3 (6DOF `env_config`, 65 536 envs); N > 1 keeps 65 536 envs per GPU (weak scaling; envs are
keyed by their global index, no data-path collective) and adds the per-rollout all-gather of the episode statistics
to the timed region.  Prints ONE JSON line on rank 0.

Timing: the K timed steps run back to back inside ONE bracket of CUDA events (barrier + synchronize on both sides).  So that every
launch finds its inputs in HBM, not in L2, the steps rotate over a RING of independent 65 536-env batches whose combined working
set is >= 2 x the L2 capacity ("inputs larger than L2"; the ring size is in `config.l2`).  Round 1 / early round 2 timed every
launch with its own event pair behind an L2-flush kernel; an event pair around an EMPTY kernel reads 5 - 6 us on this platform
(scripts/launch_overhead.cu), so that figure charged ~15 % of timestamping overhead to the kernel -- it is still reported, as
`variants.single_launch_event_pair_l2_flushed`.

Launch ordering: the workload is a synthetic random-action rollout -- every action exists before the rollout starts -- so the steps are
enqueued as an open-loop chain (rlr_step_chained): block b of a launch waits for block b of the previous launch on the same batch (the
only dependency a step has on its predecessor) instead of for the predecessor's whole grid, and the last third of a launch, when only
the warps with a third RK45 attempt / a ground event / resets are still running, overlaps the body of the next launch.  Same results
(tests/test_gpu_parity.py::test_chained_steps_equal_plain_steps: bit for bit in the exact build, flags / counters identical and the state
within 1e-15 in the fast build, whose chained kernel is a separate compilation).  `variants.serialized_launches_ring` is the same ring with
plain rlr_step launches (grid-wide ordering: what a policy network between two steps forces); `--serialized` makes that the headline.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 65536
# untimed steps queued right before the timed ones (after the barrier + synchronize): a chain of overlapping launches needs some steps
# from an idle GPU until its blocks have drifted apart to their steady state (20 timed steps read 19.7 / 18.9 / 18.7 us after 4 / 16 / 32
# lead-in steps; behind the device-side sleep below: 19.4 after 16, 18.9 - 19.0 after 48)
LEAD_IN = int(os.environ.get("RLR_BENCH_LEAD_IN", "48"))
GATE_CYCLES = int(os.environ.get("RLR_BENCH_GATE_CYCLES", "4000000"))     # device-side sleep the timed sequence is enqueued behind (~2 ms)
PREP_STEPS = 150            # untimed state preparation: desynchronise episodes so resets/events occur at their steady rate
FLOP_STEP, FLOP_ATTEMPT = 540, 1530      # algorithmic fp64 flop per 6DOF env step: 540 + 1530*k
BYTES_STEP = 320                          # algorithmic HBM bytes per env step when state round-trips
TRAFFIC_BYTES_PER_LAUNCH = 21.95e6        # measured DRAM read + write per step_kernel<Sys6> launch at 65 536 envs, in the ring (profiles/r02g)


def env_config():
    from rl_rocket_b200.config import env_config_6dof
    return env_config_6dof()


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 3 + j and r[3 + j].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows)}


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path (Python + scipy.solve_ivp), as restated in oracle/scipy_port.py
    (bit-identical to the live reference on the golden vectors), on every host core.  One bench "step" = a bounded sample:
    each of the P worker processes advances its environment for `--ref-seconds` of wall time."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import scipy_port as SP
    cores = os.cpu_count()
    vals = []
    farm = SP.CpuFarm(6, env_config(), 800, procs=cores)
    # bounded: the whole --steps K --warmup W run stays under ~2 minutes of sampling whatever K and W are
    per_step = max(0.25, min(args.ref_seconds, 110.0 / (args.warmup + args.steps)))
    for i in range(args.warmup + args.steps):
        r = farm.sample(per_step)
        if i >= args.warmup:
            vals.append(r)
    farm.close()
    steps = sum(v["steps"] for v in vals)
    wall = sum(v["wall_s"] for v in vals)
    value = steps / wall
    sample = f"{cores} single-env worker processes x {per_step:.2f} s of random-action 6DOF env_config steps per bench step"
    line = {"impl": "reference", "metric": "env_steps_per_sec_6dof", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "6DOF env_config (configuration_file.py:4-34), random U(-1,1) f32 actions, TimeLimit 800, auto-reset",
                       "envs": cores},
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from rl_rocket_b200.config import EnvSpec
    from rl_rocket_b200.engine import RocketBatch, fma_peak
    from rl_rocket_b200.vec_env import RocketVecEnv

    n = args.envs_per_gpu
    K, W = args.steps, args.warmup
    # The actions of a synthetic random-action / fixed-policy rollout exist before the rollout starts, so consecutive steps are enqueued as
    # an open-loop chain (rlr_step_chained: launch t+1 is ordered after launch t block by block, not grid by grid; same results as
    # rlr_step, tests/test_gpu_parity.py::test_chained_steps_equal_plain_steps).  --serialized times plain rlr_step launches instead (what
    # a policy network between two steps forces); that figure is always reported as variants.serialized_launches_ring.
    CHAINED = not args.serialized
    cfg = env_config()
    spec = EnvSpec(6, max_episode_steps=800, **cfg)
    # ring of independent batches: combined working set >= 2 x L2, so a launch never finds its state / outputs in L2 (--ring 1: one batch)
    l2_bytes = torch.cuda.get_device_properties(dev).L2_cache_size
    R = args.ring if args.ring > 0 else max(2, -(-2 * l2_bytes // (BYTES_STEP * n)))
    shared_stats = torch.zeros(8, dtype=torch.int64, device=dev)
    ring = [RocketBatch(spec, n, device=dev, seed=42 + 1000 * j, env_offset=rank * n, math_mode=args.math, stats=shared_stats) for j in range(R)]
    env = ring[0]
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    for b in ring:
        b.reset()
    chunk = torch.empty((PREP_STEPS, n, 3), device=dev)
    chunk.uniform_(-1, 1, generator=gen)
    actions = torch.empty((W + K, n, 3), device=dev)
    actions.uniform_(-1, 1, generator=gen)
    # L2 flush between timed steps: READ a 256 MiB buffer (clean lines; a write-flush would leave 126 MB of dirty lines whose
    # write-back lands inside the next timed kernel)
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    flush_sink = torch.zeros(1, dtype=torch.float32, device=dev)
    stat_gather = torch.zeros((world, 8), dtype=torch.int64, device=dev) if world > 1 else None
    exchange = None
    if world > 1:
        from rl_rocket_b200.distributed import StatsExchange
        exchange = StatsExchange(dev)            # NVLink peer-memory exchange (rlr_exchange_*); NCCL all-gather if it cannot be mapped

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    time.sleep(0.25)           # (every rank pauses alike: nvidia-smi needs a moment to start sampling on rank 0)
    # state preparation and warm-up run AFTER the sampler's start-up pause, so that the GPU is busy (clocks up) right before the timed region
    for t in range(PREP_STEPS):
        for b in ring:
            b.step_tensor(chunk[t], chained=CHAINED)
    for t in range(max(W, R)):
        ring[t % R].step_tensor(actions[t % W], chained=CHAINED)
    if world > 1:        # warm-up of the exchange too (first launch: module load; the NCCL fallback sets up its channels: milliseconds)
        for _ in range(3):
            exchange.all_gather(env.stats)
            dist.all_gather_into_tensor(stat_gather, env.stats.unsqueeze(0))
    env.read_stats(reset=True)
    ev0, ev_mid, ev1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    barrier()
    # The whole sequence -- rendezvous, lead-in, the K timed steps, the exchange -- is enqueued behind a ~2 ms device-side sleep, so the GPU
    # runs it from a full launch queue: a host hiccup of a few tens of microseconds (one rank of eight descheduled for 50 us was enough)
    # would otherwise starve the stream inside a region that lasts 0.4 ms at K = 20 and be booked as step or exchange time.
    if hasattr(torch.cuda, "_sleep") and GATE_CYCLES > 0:
        torch.cuda._sleep(GATE_CYCLES)
    if world > 1:
        # device-side rendezvous right before the timed region: the host barrier releases the ranks tens of microseconds apart, and a
        # rank that starts early would otherwise book that skew as time spent waiting in the rollout's exchange
        exchange.all_gather(env.stats)
    t_wall0 = time.perf_counter()
    for t in range(LEAD_IN):              # untimed lead-in on the same stream: the timed region starts on a busy GPU, as every step of a
        ring[(t - LEAD_IN) % R].step_tensor(actions[t % W], chained=CHAINED)    # rollout but the first does (ev0's timestamp = the last lead-in step's end)
    launches0 = sum(b.launches for b in ring)
    ev0.record()
    for t in range(K):
        ring[t % R].step_tensor(actions[W + t], chained=CHAINED)
    ev_mid.record()
    # per-rollout exchange of returns / done counts (all-gather of the 8-slot statistics record), inside the timed region; the skew the
    # ranks accumulate over their K steps is booked to it (it waits for the slowest rank)
    if world > 1:
        stat_gather = exchange.all_gather(env.stats)
    ev1.record()
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    step_ms = [ev0.elapsed_time(ev_mid)]
    gather_ms = ev_mid.elapsed_time(ev1) if world > 1 else 0.0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = sum(b.launches for b in ring) - launches0
    t_ms = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        mine = torch.tensor([sum(step_ms), gather_ms, wall_ms], dtype=torch.float64, device=dev)
        allr = torch.zeros((world, 3), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine.unsqueeze(0))
        per_rank = {"steps_ms_sum": [round(v, 4) for v in allr[:, 0].tolist()], "exchange_ms": [round(v, 4) for v in allr[:, 1].tolist()],
                    "host_wall_ms": [round(v, 3) for v in allr[:, 2].tolist()]}
    max_ms = float(t_ms.item())
    st = env.read_stats()
    if world > 1:
        tot = stat_gather.sum(0).cpu().tolist()
        st_all = {"steps": tot[0], "attempts": tot[1], "done": tot[2]}
    else:
        st_all = st
    value = n * world * K / (max_ms * 1e-3)

    # ---- end to end through the public VecEnv API with HOST buffers (pinned): H2D of the actions + D2H of obs / reward / done inside
    # the timed region, every step, on EVERY rank; whole-job value = all ranks' env-steps / the slowest rank's wall time.
    # Headline: RocketVecEnv.step_host_async (C ABI rlr_pipe_submit / rlr_pipe_wait, two slots): the host hands over step t + 1's
    # actions while step t's results are still crossing PCIe, and reads every step's result (done count) before reusing its slot.
    venv = RocketVecEnv(n, dof=6, max_episode_steps=800, device=dev, math_mode=args.math, env_offset=rank * n, **cfg)
    venv.reset_tensor()
    for t in range(PREP_STEPS):
        venv.step_tensor(chunk[t])
    Ke = min(K, 200)
    h_actions = [actions[W + t].cpu().pin_memory() for t in range(Ke)]
    for t in range(3):
        venv.step_host(h_actions[t % Ke])
    barrier()
    t0 = time.perf_counter()
    for t in range(Ke):
        venv.step_host(h_actions[t])
    torch.cuda.synchronize(dev)
    sync_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    D = max(2, args.e2e_depth)
    pend = [venv.step_host_async(h_actions[t % Ke], depth=D) for t in range(D)]   # warm-up of the pipeline (stream / event creation)
    for h in pend:
        h.wait()
    barrier()
    done_seen = 0
    t0 = time.perf_counter()
    pend = []
    for t in range(Ke):
        pend.append(venv.step_host_async(h_actions[t]))
        if len(pend) == D:
            o_, r_, d_ = pend.pop(0).wait()
            done_seen += int(d_[0])                    # the host reads the step's result
    while pend:
        o_, r_, d_ = pend.pop(0).wait()
        done_seen += int(d_[0])
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    # PCIe ceiling of this box, measured the same way on every rank at once: the step's D2H bytes and H2D bytes as plain
    # cudaMemcpyAsync on two streams, back to back, no kernel
    d_pack, h_pack = venv.batch.out_pack, venv._h_pack
    d_act = torch.empty_like(h_actions[0], device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    reps = 100
    for rep in range(2):
        barrier()
        t0 = time.perf_counter()
        for t in range(reps):
            with torch.cuda.stream(s_out):
                h_pack.copy_(d_pack, non_blocking=True)
            with torch.cuda.stream(s_in):
                d_act.copy_(h_actions[t % Ke], non_blocking=True)
        s_in.synchronize(); s_out.synchronize()
        pcie_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        for x in (sync_s, e2e_s, pcie_s):
            dist.all_reduce(x, op=dist.ReduceOp.MAX)
    e2e = n * world * Ke / float(e2e_s.item())
    e2e_sync = n * world * Ke / float(sync_s.item())
    pcie_ceiling = n * world * reps / float(pcie_s.item())
    d2h_gbs = venv.d2h_bytes_per_step() * world * reps / float(pcie_s.item()) / 1e9
    e2e_d = {"value": e2e, "unit": "env-steps/s", "h2d_bytes_per_step": venv.h2d_bytes_per_step() * world,
             "d2h_bytes_per_step": venv.d2h_bytes_per_step() * world, "steps": Ke,
             "api": f"RocketVecEnv.step_host_async (rlr_pipe_submit / rlr_pipe_wait, depth {D})",
             "synchronous_step_host": e2e_sync,
             "pcie_ceiling": {"value": pcie_ceiling, "unit": "env-steps/s", "d2h_GBps_all_ranks": d2h_gbs,
                              "how": "the same bytes per step as plain cudaMemcpyAsync D2H + H2D on two streams, all ranks at once, no kernel"},
             "frac_of_pcie_ceiling": e2e / pcie_ceiling,
             "note": "every rank: pinned host actions -> H2D copy -> step kernel -> D2H copy of obs|reward|done into pinned host memory, `depth` slots in "
                     "flight, the host reads each step's result; max wall time over ranks.  synchronous_step_host = one blocking call per step "
                     "(mapped pinned memory written by the kernel)"}

    # ---- BASELINE config 4 for the record: 2^20 environments in total, sharded over the ranks by global index (2^20 / world per GPU),
    # one launch per step, back to back, device-timed, max over ranks, the statistics all-gather included
    del venv
    n4 = (1 << 20) // world
    K4 = min(K, 100)
    env4 = RocketBatch(spec, n4, device=dev, seed=42, env_offset=rank * n4, math_mode=args.math)
    env4.reset()
    a4 = torch.empty((8, n4, 3), device=dev)
    a4.uniform_(-1, 1, generator=gen)
    for t in range(PREP_STEPS):
        env4.step_tensor(a4[t % 8])
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for t in range(K4):
        env4.step_tensor(a4[t % 8])
    if world > 1:
        exchange.all_gather(env4.stats)
    c1.record()
    c1.synchronize()
    t4 = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t4, op=dist.ReduceOp.MAX)
    config4 = n4 * world * K4 / (float(t4.item()) * 1e-3)
    del env4, a4

    line = None
    if rank == 0:
        # ---- the round-1 method for the record: every launch behind an L2 read-flush, timed by its own event pair
        Kf = min(K, 100)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Kf)]
        for rep in range(2):
            for t in range(Kf):
                torch.sum(flush, dim=(0,), keepdim=True, out=flush_sink)
                evs[t][0].record()
                env.step_tensor(actions[W + t % K])
                evs[t][1].record()
            torch.cuda.synchronize(dev)
        single_flushed = n * Kf / (sum(a_.elapsed_time(b_) for a_, b_ in evs) * 1e-3)
        # ---- back-to-back on ONE batch (state L2-resident) and fused-rollout variants, device-timed, for the record
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def bracket(step_fn, steps):
            for t in range(LEAD_IN):
                step_fn(t)
            e0.record()
            for t in range(steps):
                step_fn(t)
            e1.record()
            e1.synchronize()
            return n * steps / (e0.elapsed_time(e1) * 1e-3)

        torch.cuda.synchronize(dev)
        b2b = bracket(lambda t: env.step_tensor(actions[W + t % K]), K)
        b2b_chained = bracket(lambda t: env.step_tensor(actions[W + t % K], chained=True), K)
        ring_serialized = bracket(lambda t: ring[t % R].step_tensor(actions[W + t % K]), K)
        ring_chained = bracket(lambda t: ring[t % R].step_tensor(actions[W + t % K], chained=True), K)
        # two independent groups of environments stepped alternately on two streams: the tail of one launch overlaps the body of the next
        # (what the single-stream figures cannot do: griddepcontrol.wait orders consecutive launches of a stream)
        s2 = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
        torch.cuda.synchronize(dev)
        e0.record()
        for s_ in s2:
            s_.wait_event(e0)
        for t in range(K):
            with torch.cuda.stream(s2[t % 2]):
                ring[t % R].step_tensor(actions[W + t])
        for s_ in s2:
            torch.cuda.current_stream(dev).wait_stream(s_)
        e1.record()
        e1.synchronize()
        two_streams = n * K / (e0.elapsed_time(e1) * 1e-3)
        Tr = min(K, 200)
        env.rollout(actions[W:W + 3])            # first launch of the rollout kernel (module load) outside the timing
        e0.record()
        env.rollout(actions[W:W + Tr])
        e1.record()
        e1.synchronize()
        fused = n * Tr / (e0.elapsed_time(e1) * 1e-3)
        # ---- RLR_F32 precision (binary32 state, tolerance 1e-4, tests/test_gpu_f32.py) of the same workload, for the record
        env32 = RocketBatch(spec, n, device=dev, seed=42, env_offset=rank * n, math_mode="fast", precision="f32")
        env32.reset()
        for t in range(PREP_STEPS):
            env32.step_tensor(chunk[t])
        torch.cuda.synchronize(dev)
        e0.record()
        for t in range(K):
            env32.step_tensor(actions[W + t])
        e1.record()
        e1.synchronize()
        f32_b2b = n * K / (e0.elapsed_time(e1) * 1e-3)
        env32.rollout(actions[W:W + 3])
        e0.record()
        env32.rollout(actions[W:W + Tr])
        e1.record()
        e1.synchronize()
        f32_fused = n * Tr / (e0.elapsed_time(e1) * 1e-3)
        del env32
        peak = fma_peak("f64", device=dev)
        # ---- the rest of the measurement contract.  Fixed policies of test_6DOF_environment.py:19-20: timed like `value`.
        # BASELINE config 2 (3DOF, 4 096 envs) and config 1 (1DOF): batches far smaller than L2, so a ring cannot keep them out of it --
        # one launch per event pair behind an L2 read-flush (the figure includes the ~5 us an event pair around an empty kernel reads)
        def flushed_us(batch, act_fn, steps):
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for t in range(steps):
                torch.sum(flush, dim=(0,), keepdim=True, out=flush_sink)
                evs[t][0].record()
                batch.step_tensor(act_fn(t))
                evs[t][1].record()
            torch.cuda.synchronize(dev)
            return sum(a_.elapsed_time(b_) for a_, b_ in evs) * 1e3 / steps

        Kv = min(K, 100)
        extra = {}
        for name, act in (("fixed_policy_null_action_[0,0,-1]", (0.0, 0.0, -1.0)), ("fixed_policy_[1,1,-0.5]", (1.0, 1.0, -0.5))):
            # timed like `value`: Kv steps in one bracket over a ring of R batches
            a_fix = torch.tensor(act, dtype=torch.float32, device=dev).repeat(n, 1).contiguous()
            shared_p = torch.zeros(8, dtype=torch.int64, device=dev)
            ringp = [RocketBatch(spec, n, device=dev, seed=42 + 1000 * j, env_offset=rank * n, math_mode=args.math, stats=shared_p) for j in range(R)]
            for b in ringp:
                b.reset()
            for t in range(PREP_STEPS):
                for b in ringp:
                    b.step_tensor(a_fix)
            ringp[0].read_stats(reset=True)
            torch.cuda.synchronize(dev)
            e0.record()
            for t in range(Kv):
                ringp[t % R].step_tensor(a_fix, chained=CHAINED)
            e1.record()
            e1.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / Kv
            sp = ringp[0].read_stats()
            fl_ = FLOP_STEP * sp["steps"] + FLOP_ATTEMPT * sp["attempts"]
            extra[name] = {"env_steps_per_s": n / us * 1e6, "us_per_step": us, "attempts_per_step": sp["attempts"] / max(1, sp["steps"]),
                           "roofline_frac": fl_ / (us * 1e-6 * Kv) / 1e12 / peak}
            del ringp
        from rl_rocket_b200.config import env_config_1dof, env_config_3dof
        for name, dof_, cfg_, n_, tl_ in (("config2_3dof_4096_envs", 3, env_config_3dof(), 4096, 3000), ("3dof_65536_envs", 3, env_config_3dof(), 65536, 3000),
                                          ("config1_1dof_single_env", 1, env_config_1dof(), 1, 3000), ("1dof_4096_envs", 1, env_config_1dof(), 4096, 3000)):
            sp_ = EnvSpec(dof_, max_episode_steps=tl_, **cfg_)
            e_ = RocketBatch(sp_, n_, device=dev, seed=42, math_mode=args.math)
            e_.reset()
            a_ = torch.empty((16, n_, sp_.nact), device=dev).uniform_(-1, 1, generator=gen)
            for t in range(PREP_STEPS):
                e_.step_tensor(a_[t % 16])
            us_flushed = flushed_us(e_, lambda t: a_[t % 16], Kv)
            e_.read_stats(reset=True)
            torch.cuda.synchronize(dev)
            e0.record()
            for t in range(Kv):
                e_.step_tensor(a_[t % 16])
            e1.record()
            e1.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / Kv
            sp = e_.read_stats()
            # 3DOF flop model: 250 + 590 k per env step (trig calls counted as one flop each)
            fl_ = 250 * sp["steps"] + 590 * sp["attempts"]
            extra[name] = {"env_steps_per_s": n_ / us * 1e6, "us_per_step": us, "attempts_per_step": sp["attempts"] / max(1, sp["steps"]),
                           "roofline_frac": fl_ / (us * 1e-6 * Kv) / 1e12 / peak, "us_per_step_single_launch_event_pair_l2_flushed": us_flushed,
                           "timing": "Kv launches back to back in one event bracket (the batch is L2-resident: it is far smaller than L2)"}
            del e_
        # BASELINE config 5: SB3-style rollout collection through RocketVecEnv.step() (numpy in / numpy out + infos), wall clock
        for n_ in (4096, 65536):
            v_ = RocketVecEnv(n_, dof=6, max_episode_steps=800, device=dev, math_mode=args.math, **cfg)
            v_.reset()
            a_np = [actions[W + t, :n_].cpu().numpy() for t in range(8)]
            for t in range(10):
                v_.step(a_np[t % 8])
            Ks = 30
            t0 = time.perf_counter()
            for t in range(Ks):
                v_.step(a_np[t % 8])
            extra[f"config5_sb3_vecenv_step_{n_}_envs"] = {"env_steps_per_s": n_ * Ks / (time.perf_counter() - t0),
                                                            "note": "step_async + step_wait: numpy actions in, numpy obs / rewards / dones + info dicts out"}
            del v_
        peak32 = fma_peak("f32", device=dev)
        f32_flops_per_step = (FLOP_STEP + FLOP_ATTEMPT * st["attempts"] / max(1, st["steps"]))
        extra["f32_state_roofline"] = {"peak_fp32_fma_tflops": peak32, "back_to_back_frac": f32_b2b * f32_flops_per_step / 1e12 / peak32,
                                       "fused_rollout_frac": f32_fused * f32_flops_per_step / 1e12 / peak32,
                                       "note": "same flop model (540 + 1530 k) against the FFMA peak measured by rlr_fma_peak('f32')"}
        # ---- roofline of the dominant kernel (step_kernel<Sys6>): FP64 ALU bound, peak measured in this run
        # (the statistics record also counted the untimed lead-in steps: scale to the n * K timed env steps)
        flops = (FLOP_STEP * st["steps"] + FLOP_ATTEMPT * st["attempts"]) * (n * K / max(1, st["steps"]))
        kern_s = sum(step_ms) * 1e-3
        achieved = flops / kern_s / 1e12
        hbm_gbs = BYTES_STEP * n * K / kern_s / 1e9
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:  # noqa: BLE001
            hbm_peak = 6650.0
        roofline = {"bound": "fp64_alu", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": TRAFFIC_BYTES_PER_LAUNCH if n == ENVS_PER_GPU else None,
                    "traffic_source": "dram__bytes_read.sum (9.45 MB) + dram__bytes_write.sum (12.5 MB), mean of 39 consecutive step_kernel<Sys6> "
                                      "launches of the 13-batch ring, ncu --cache-control none, one pass (profiles/r02g_ring_dram_per_launch.csv); "
                                      "algorithmic bytes per launch = 320 B x 65536 = 21.0 MB",
                    "kernel": "step_kernel<Sys6>", "peak_source": "rlr_fma_peak DFMA microbench in this run "
                    "(MEASURED_PEAKS.json has no FP64 figure)", "flop_model": "540 + 1530*attempts per env step; attempts counted by the kernel",
                    "attempts_per_step": st["attempts"] / max(1, st["steps"]),
                    "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak, "bytes_per_env_step": BYTES_STEP}}
        # ---- CPU baseline: the reference's Python/scipy path (oracle port) on this box's host cores, bounded sample
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import scipy_port as SP
            r = SP.cpu_throughput(6, cfg, 800, seconds=args.cpu_seconds, warm=1.0)
            cpu = {"value": r["value"], "unit": "env-steps/s", "cores": r["cores"], "kind": "port",
                   "sample": f"{r['cores']} single-env processes x {args.cpu_seconds:.0f} s, random actions, auto-reset ({r['steps']} steps)"}
        line = {"metric": "env_steps_per_sec_6dof", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": max_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "BASELINE config 3: 6DOF env_config (configuration_file.py:4-34), %d envs per GPU, random U(-1,1) "
                           "f32 actions, TimeLimit 800, auto-reset, one %s launch per step" % (n, "rlr_step_chained (open-loop chain: consecutive "
                           "launches ordered block by block, so a launch's tail overlaps the next launch's body; same results as rlr_step)"
                           if CHAINED else "rlr_step"), "envs_per_gpu": n, "launch_ordering": "chained" if CHAINED else "serialized",
                           "total_envs": n * world, "math_mode": args.math,
                           "l2": ("inputs larger than L2: the K timed steps rotate over a ring of %d independent %d-env batches (%.0f MB of state + outputs "
                                  "touched per pass over the ring vs %.0f MB of L2), back to back inside ONE event bracket" % (R, n, R * BYTES_STEP * n / 1e6, l2_bytes / 1e6))
                                 if R > 1 else "one batch, back to back (state L2-resident)",
                           "ring_batches": R,
                           "state_prep": f"{PREP_STEPS} untimed random-action steps so episodes are desynchronised",
                           "collective": (("all-gather of the 64-byte episode-statistics record per rank, timed: " +
                                           ("one kernel per rank storing into the peers' buffers over NVLink (rlr_exchange_allgather)"
                                            if exchange.peer_memory else "NCCL all_gather (peer memory could not be mapped: %s)" % exchange.error))
                                          if world > 1 else None)},
                "clocks": clocks, "e2e": e2e_d, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "episodes_finished": st_all["done"], "wall_ms_timed_region": wall_ms,
                "variants": {"single_launch_event_pair_l2_flushed": single_flushed, "serialized_launches_ring": ring_serialized, "chained_launches_ring": ring_chained,
                             "back_to_back_one_batch_l2_warm": b2b, "back_to_back_one_batch_chained": b2b_chained, "ring_on_two_streams_alternating": two_streams, "fused_rollout_kernel": fused, "f32_state_back_to_back": f32_b2b,
                             "f32_state_fused_rollout": f32_fused, "unit": "env-steps/s (rank 0 only)",
                             "config4_2p20_envs_sharded_all_ranks": config4, **extra},
                "gather_ms": gather_ms, "steps_ms_sum": sum(step_ms), "per_rank": per_rank}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


_JSON_FD = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else written to fd 1 during the run (NCCL's version banner,
    library chatter) was redirected to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--math", default="fast", choices=["fast", "exact"])
    ap.add_argument("--serialized", action="store_true", help="time plain rlr_step launches (grid-wide ordering) instead of the open-loop chain")
    ap.add_argument("--ring", type=int, default=0, help="independent batches the timed steps rotate over (0: enough for 2 x L2; 1: one batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--ref-seconds", type=float, default=2.0)
    ap.add_argument("--e2e-depth", type=int, default=3)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun, one rank per GPU
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                   "--master-port", "29511"] + sys.argv
            sys.exit(subprocess.call(cmd))
        global _JSON_FD
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)             # keep stdout clean for the JSON line (NCCL prints its version banner to fd 1)
        run_ours(args)


if __name__ == "__main__":
    main()
