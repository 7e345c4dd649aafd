"""Soak: binary32-state build against the f64 fast build at full size; environments are dropped at their first done / truncated
disagreement (from then on they are different episodes).  Reports the disagreement rate per env step and the worst deviation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_rocket_b200.config import EnvSpec
from rl_rocket_b200.engine import RocketBatch
from tests import goldutil as G
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
N = 65536
spec = EnvSpec(6, max_episode_steps=800, **G.cfg6())
f64 = RocketBatch(spec, N, seed=23, math_mode="fast", debug_outputs=True)
f32 = RocketBatch(spec, N, seed=23, math_mode="fast", precision="f32", debug_outputs=True)
f64.reset(); f32.reset()
g = torch.Generator(device="cuda").manual_seed(11)
norm = torch.as_tensor(spec.state_normalizer, device="cuda")[:, None]
alive = torch.ones(N, dtype=torch.bool, device="cuda")
worst = torch.zeros((), dtype=torch.float64, device="cuda")
alive_steps = torch.zeros((), dtype=torch.int64, device="cuda")
natt_diff = torch.zeros((), dtype=torch.int64, device="cuda")
for t in range(T):
    a = torch.rand((N, 3), device="cuda", generator=g) * 2 - 1
    f64.step_tensor(a); f32.step_tensor(a)
    alive_steps += alive.sum()
    natt_diff += ((f64.natt != f32.natt) & alive).sum()
    alive &= (f64.done == f32.done) & (f64.truncated == f32.truncated)
    dev = ((f64.y_pre_reset - f32.y_pre_reset.double()).abs() / norm).amax(0)
    worst = torch.maximum(worst, torch.where(alive, dev, torch.zeros_like(dev)).max())
d = int((~alive).sum())
print(f"f32 vs f64: {int(alive_steps):.3e} compared env steps, {d} environments ever disagreed on a flag ({d/max(1,int(alive_steps)):.2e} per env step), "
      f"{int(natt_diff)} attempt-count differences ({int(natt_diff)/max(1,int(alive_steps)):.2e} per env step), worst normalised deviation {float(worst):.3e}")
