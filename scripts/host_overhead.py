"""Development probe: host-side cost of one RocketBatch.step_tensor call (Python + ctypes + launch), plain and chained: the time the
host needs to ENQUEUE 200 steps of a 65 536-env batch (the launch queue does not fill), against the device time of the same steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_rocket_b200.config import EnvSpec, env_config_6dof
from rl_rocket_b200.engine import RocketBatch
spec = EnvSpec(6, max_episode_steps=800, **env_config_6dof())
n = 65536
env = RocketBatch(spec, n, seed=1); env.reset()
a = torch.rand((n, 3), device="cuda") * 2 - 1
for chained in (False, True, False, True):
    for _ in range(200): env.step_tensor(a, chained=chained)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): env.step_tensor(a, chained=chained)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"chained={chained}: enqueue {1e6*(t1-t0)/200:.1f} us/call, device {1e6*(t2-t0)/200:.1f} us/step")
