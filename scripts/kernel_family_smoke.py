"""Small run of every kernel family (written for compute-sanitizer, which this GPU pool does not allow; useful as a quick soak): f64 step (device and mapped-host outputs), fused
rollout with observation buffers, f32 step / rollout, 3DOF step, resets and events included (short TimeLimit, low start)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_rocket_b200.config import EnvSpec
from rl_rocket_b200.engine import RocketBatch
from rl_rocket_b200.vec_env import RocketVecEnv
from tests import goldutil as G
N, T = 1000, 40
cfg = G.cfg6()
cfg["IC"] = [60, 20, 20, -30, 0, 0, 1, 0, 0, 0, 0, 0, 0, 45e3]          # low start: ground events within a few steps
a = torch.rand((T, N, 3), device="cuda") * 2 - 1
for prec in ("f64", "f32"):
    spec = EnvSpec(6, max_episode_steps=15, **cfg)
    env = RocketBatch(spec, N, seed=3, precision=prec, debug_outputs=True); env.reset()
    for t in range(T): env.step_tensor(a[t])
    obs_buf = torch.zeros((T, N, 14), device="cuda"); rb = torch.zeros((T, N), device="cuda"); db = torch.zeros((T, N), dtype=torch.uint8, device="cuda")
    env.rollout(a, obs_buf, rb, db)
    v = RocketVecEnv(N, dof=6, max_episode_steps=15, precision=prec, **cfg); v.reset()
    for t in range(10): v.step_host(a[t].cpu().pin_memory())
    print(prec, env.read_stats())
spec3 = EnvSpec(3, max_episode_steps=20, **G.cfg3())
e3 = RocketBatch(spec3, N, seed=1); e3.reset()
a3 = torch.rand((T, N, 2), device="cuda") * 2 - 1
for t in range(T): e3.step_tensor(a3[t])
e3.rollout(a3)
torch.cuda.synchronize(); print("3dof", e3.read_stats())
