"""Development aid: per-phase warp cycles of step_kernel<Sys6> from a -DRLR_PHASE_TIMING=1 tuning build.
usage: RLR_LIB=rl_rocket_b200/_C/librlrocket_PT.so python scripts/phase_probe.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_rocket_b200 import _native as N
from rl_rocket_b200.config import EnvSpec
from rl_rocket_b200.engine import RocketBatch
from rl_rocket_b200.config import env_config_6dof

NAMES = ["0 loads", "1 ctx+K0+init-step", "2 RK attempts", "3 event path", "4 normalise/obs/reward", "5 output stores", "6 auto-reset",
         "7 obs/state stores+stats", "8 IC sampling (warp-cooperative)", "9 terminal obs / fin stores"]
lib = N.load()
n = int(os.environ.get("QB_N", "65536"))
spec = EnvSpec(6, max_episode_steps=int(os.environ.get("PP_TL", "800")), **env_config_6dof())
env = RocketBatch(spec, n, seed=1, math_mode="fast"); env.reset()
a = torch.rand((250, n, 3), device="cuda") * 2 - 1
for t in range(150): env.step_tensor(a[t])
buf = (C.c_ulonglong * 48)()
lib.rlr_debug_phases(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(150, 250): env.step_tensor(a[t])
e1.record(); e1.synchronize()
lib.rlr_debug_phases(buf)
v = list(buf); warps = n // 32 * 100
tag = os.environ.get("RLR_LIB", "default").split("/")[-1]
print(f"[{tag}] {e0.elapsed_time(e1)*10:.1f} us/step (instrumented); cycles per warp-step by phase: mean over all warps | mean over the warps that ran it | max")
tot = 0
for k, nm in enumerate(NAMES):
    s, m, c = v[k], v[16 + k], v[32 + k]
    tot += s / warps
    print(f"  {nm:28s} {s/warps:9.0f} | {s/max(c,1):9.0f} ({c/warps*100:5.1f}% of warps) | {m:8d}")
print(f"  total mean {tot:.0f} cycles per warp-step")
