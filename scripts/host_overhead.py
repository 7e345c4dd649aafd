"""Development probe: host-side cost of one RocketBatch.step_tensor call (Python + ctypes + launch), measured with a kernel that is
shorter than the call (32 envs) so that the loop is host-bound."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_rocket_b200.config import EnvSpec
from rl_rocket_b200.engine import RocketBatch
from tests import goldutil as G
spec = EnvSpec(6, max_episode_steps=800, **G.cfg6())
for n in (32, 65536):
    env = RocketBatch(spec, n, seed=1); env.reset()
    a = torch.rand((n, 3), device="cuda") * 2 - 1
    for _ in range(200): env.step_tensor(a)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3000): env.step_tensor(a)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"n={n}: enqueue {1e6*(t1-t0)/3000:.1f} us/call, total {1e6*(t2-t0)/3000:.1f} us/step")
