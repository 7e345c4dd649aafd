import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_rocket_b200.vec_env import RocketVecEnv
from tests import goldutil as G
n = 65536
venv = RocketVecEnv(n, dof=6, max_episode_steps=800, **G.cfg6())
venv.reset_tensor()
a = torch.rand((300, n, 3), device="cuda") * 2 - 1
for t in range(150): venv.step_tensor(a[t])
h_a = [a[150 + t].cpu().pin_memory() for t in range(100)]
for zc in (True, False):
    for t in range(5): venv.step_host(h_a[t], zero_copy=zc)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(100): venv.step_host(h_a[t], zero_copy=zc)
    torch.cuda.synchronize(); print(os.environ.get("RLR_LIB","")[-8:], "zero_copy" if zc else "copy", f"{(time.perf_counter()-t0)/100*1e6:.1f} us/step")
