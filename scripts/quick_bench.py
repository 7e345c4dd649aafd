"""Ad-hoc timing probe (development aid; bench.py is the contract benchmark)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_rocket_b200.config import EnvSpec
from rl_rocket_b200.engine import RocketBatch
from tests import goldutil as G

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
        spec = EnvSpec(6, max_episode_steps=800, **G.cfg6())
        env = RocketBatch(spec, N, seed=1, math_mode=mode, precision=prec)
        env.reset()
        T = 300
        a = torch.rand((T, N, 3), device="cuda") * 2 - 1
        for t in range(150): env.step_tensor(a[t])     # warm-up: desynchronise episodes
        env.read_stats(reset=True)
        it = iter(range(150, T))
        ms = timeit(lambda: env.step_tensor(a[next(it)]), 150)
        st = env.read_stats(reset=True)
        flops = 540 * st["steps"] + 1530 * st["attempts"]
        print(f"[{tag}] {prec} step {mode} N={N}: {ms*1e3:.1f} us/step  {N/ms*1e3:.3e} env-steps/s  att/step {st['attempts']/st['steps']:.3f}  alg {flops/150/ms/1e9:.2f} Tflop/s", flush=True)
        env.rollout(a[:20]); torch.cuda.synchronize()      # first launch: lazy module load + local-memory sizing
        ms2 = timeit(lambda: env.rollout(a[:100]), 3) / 100
        print(f"[{tag}] {prec} rollout {mode} N={N}: {ms2*1e3:.1f} us/step  {N/ms2*1e3:.3e} env-steps/s", flush=True)
        res[f"{mode}_{N}"] = dict(step_us=ms * 1e3, rollout_us=ms2 * 1e3)
if os.environ.get("QB_3DOF"):
    spec = EnvSpec(3, max_episode_steps=3000, **G.cfg3())
    for N in (4096, 65536):
        env = RocketBatch(spec, N, seed=1); env.reset()
        a = torch.rand((200, N, 2), device="cuda") * 2 - 1
        for t in range(100): env.step_tensor(a[t])
        it = iter(range(100, 200))
        ms = timeit(lambda: env.step_tensor(a[next(it)]), 100)
        print(f"[{tag}] 3dof step N={N}: {ms*1e3:.1f} us/step {N/ms*1e3:.3e} env-steps/s", flush=True)
json.dump(res, open(f"gpurun_out/quick_bench_{tag}.json", "w"), indent=1)
