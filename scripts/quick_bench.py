"""Ad-hoc timing probe (development aid; bench.py is the contract benchmark)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_rocket_b200.config import EnvSpec
from rl_rocket_b200.engine import RocketBatch
from rl_rocket_b200.config import env_config_6dof, env_config_3dof

def timeit(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n

tag = os.environ.get("RLR_LIB", "default").split("/")[-1]
sizes = [int(x) for x in os.environ.get("QB_SIZES", "65536,131072").split(",")]
modes = os.environ.get("QB_MODES", "fast,exact").split(",")
prec = os.environ.get("QB_PREC", "f64")
res = {}
for N in sizes:
    for mode in modes:
        spec = EnvSpec(6, max_episode_steps=800, **env_config_6dof())
        env = RocketBatch(spec, N, seed=1, math_mode=mode, precision=prec)
        env.reset()
        T = 300
        a = torch.rand((T, N, 3), device="cuda") * 2 - 1
        for t in range(150): env.step_tensor(a[t])     # warm-up: desynchronise episodes
        env.read_stats(reset=True)
        it = iter(range(150, T))
        CH = os.environ.get("QB_CHAINED", "0") == "1"      # open-loop chaining (rlr_step_chained)
        ms = timeit(lambda: env.step_tensor(a[next(it)], chained=CH), 150)
        st = env.read_stats(reset=True)
        flops = 540 * st["steps"] + 1530 * st["attempts"]
        if os.environ.get("QB_FLUSH", "1") == "1":       # what bench.py times: one launch per event pair, L2 evicted in between
            flush = torch.zeros(64 * 1024 * 1024, device="cuda"); sink = torch.zeros(1, device="cuda")
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
            for rep in range(2):
                for t in range(100):
                    torch.sum(flush, dim=(0,), keepdim=True, out=sink)
                    ev[t][0].record(); env.step_tensor(a[50 + t]); ev[t][1].record()
                torch.cuda.synchronize()
            fl = sorted(x.elapsed_time(y) for x, y in ev)
            print(f"[{tag}] {prec} step {mode} N={N} flushed, per-launch event pairs: mean {sum(fl)/len(fl)*1e3:.1f} us  median {fl[50]*1e3:.1f} us  min {fl[0]*1e3:.1f} us", flush=True)
            env.read_stats(reset=True)
        print(f"[{tag}] {prec} step {mode} N={N}: {ms*1e3:.1f} us/step  {N/ms*1e3:.3e} env-steps/s  att/step {st['attempts']/st['steps']:.3f}  alg {flops/150/ms/1e9:.2f} Tflop/s", flush=True)
        if os.environ.get("QB_ROTATE"):                  # R batches stepped round-robin back to back: inputs colder than L2, code warm
            R = int(os.environ["QB_ROTATE"])
            envs = [env] + [RocketBatch(spec, N, seed=2 + r, math_mode=mode, precision=prec) for r in range(R - 1)]
            for e in envs[1:]:
                e.reset()
                for t in range(150): e.step_tensor(a[t])
            cnt = [0]
            def rot():
                envs[cnt[0] % R].step_tensor(a[150 + cnt[0] % 150], chained=CH); cnt[0] += 1
            for _ in range(2 * R): rot()
            msr = timeit(rot, 40 * R)
            print(f"[{tag}] {prec} step {mode} N={N} rotating over {R} batches, back to back: {msr*1e3:.1f} us/step", flush=True)
            del envs
        env.rollout(a[:20]); torch.cuda.synchronize()      # first launch: lazy module load + local-memory sizing
        ms2 = timeit(lambda: env.rollout(a[:100]), 3) / 100
        print(f"[{tag}] {prec} rollout {mode} N={N}: {ms2*1e3:.1f} us/step  {N/ms2*1e3:.3e} env-steps/s", flush=True)
        res[f"{mode}_{N}"] = dict(step_us=ms * 1e3, rollout_us=ms2 * 1e3)
if os.environ.get("QB_3DOF"):
    spec = EnvSpec(3, max_episode_steps=3000, **env_config_3dof())
    for N in (4096, 65536):
        env = RocketBatch(spec, N, seed=1); env.reset()
        a = torch.rand((200, N, 2), device="cuda") * 2 - 1
        for t in range(100): env.step_tensor(a[t])
        it = iter(range(100, 200))
        ms = timeit(lambda: env.step_tensor(a[next(it)]), 100)
        print(f"[{tag}] 3dof step N={N}: {ms*1e3:.1f} us/step {N/ms*1e3:.3e} env-steps/s", flush=True)
json.dump(res, open(f"gpurun_out/quick_bench_{tag}.json", "w"), indent=1)
