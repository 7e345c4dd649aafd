"""Ad-hoc timing probe (development aid; bench.py is the contract benchmark)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_rocket_b200.config import EnvSpec
from rl_rocket_b200.engine import RocketBatch, fma_peak
from tests import goldutil as G

def timeit(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n

res = {}
res["fp64_tflops"] = fma_peak("f64"); res["fp32_tflops"] = fma_peak("f32")
print(res, flush=True)
for N in (65536, 131072, 262144):
    for mode in ("exact", "fast"):
        spec = EnvSpec(6, max_episode_steps=800, **G.cfg6())
        env = RocketBatch(spec, N, seed=1, math_mode=mode)
        env.reset()
        T = 200
        a = torch.rand((T, N, 3), device="cuda") * 2 - 1
        for t in range(100): env.step_tensor(a[t])     # warm-up: desynchronise episodes
        env.read_stats(reset=True)
        it = iter(range(100, T))
        ms = timeit(lambda: env.step_tensor(a[next(it)]), 100)
        st = env.read_stats(reset=True)
        flops = 540 * st["steps"] + 1530 * st["attempts"]
        res[f"step_{mode}_{N}"] = dict(ms=ms, steps_per_s=N / ms * 1e3, att_per_step=st["attempts"] / st["steps"], tflops=flops / 100 / ms / 1e9)
        print(f"step {mode} N={N}: {ms:.4f} ms/step  {N/ms*1e3:.3e} env-steps/s  att/step {st['attempts']/st['steps']:.3f}  alg {flops/100/ms/1e9:.2f} Tflop/s", flush=True)
        ms = timeit(lambda: env.rollout(a[:100]), 3) / 100
        res[f"rollout_{mode}_{N}"] = dict(ms=ms, steps_per_s=N / ms * 1e3)
        print(f"rollout {mode} N={N}: {ms:.4f} ms/step  {N/ms*1e3:.3e} env-steps/s", flush=True)
spec = EnvSpec(3, max_episode_steps=3000, **G.cfg3())
for N in (4096, 65536):
    env = RocketBatch(spec, N, seed=1); env.reset()
    a = torch.rand((200, N, 2), device="cuda") * 2 - 1
    for t in range(100): env.step_tensor(a[t])
    it = iter(range(100, 200))
    ms = timeit(lambda: env.step_tensor(a[next(it)]), 100)
    print(f"3dof step N={N}: {ms:.4f} ms/step {N/ms*1e3:.3e} env-steps/s", flush=True)
    res[f"step3_{N}"] = dict(ms=ms, steps_per_s=N / ms * 1e3)
json.dump(res, open("gpurun_out/quick_bench.json", "w"), indent=1)
