"""Soak: fast and exact builds stepped side by side at full size; counts done / truncated / attempt-count disagreements and the worst
state deviation among environments that never disagreed.  usage: python scripts/soak_fast_vs_exact.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_rocket_b200.config import EnvSpec
from rl_rocket_b200.engine import RocketBatch
from tests import goldutil as G
T = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
dof = int(os.environ.get("DOF", "6"))
N = 65536
cfg = G.cfg6() if dof == 6 else G.cfg3()
spec = EnvSpec(dof, max_episode_steps=800 if dof == 6 else 3000, **cfg)
fast = RocketBatch(spec, N, seed=23, math_mode="fast", debug_outputs=True)
exact = RocketBatch(spec, N, seed=23, math_mode="exact", debug_outputs=True)
fast.reset(); exact.reset()
g = torch.Generator(device="cuda").manual_seed(11)
norm = torch.as_tensor(spec.state_normalizer, device="cuda")[:, None]
alive = torch.ones(N, dtype=torch.bool, device="cuda")
worst = torch.zeros((), dtype=torch.float64, device="cuda")
natt_diff = torch.zeros((), dtype=torch.int64, device="cuda")
t0 = time.time()
for t in range(T):
    a = torch.rand((N, spec.nact), device="cuda", generator=g) * 2 - 1
    fast.step_tensor(a); exact.step_tensor(a)
    alive &= (fast.done == exact.done) & (fast.truncated == exact.truncated)
    natt_diff += ((fast.natt != exact.natt) & alive).sum()
    dev = ((fast.y_pre_reset - exact.y_pre_reset).abs() / norm).amax(0)
    worst = torch.maximum(worst, torch.where(alive, dev, torch.zeros_like(dev)).max())
    if (t + 1) % 5000 == 0:
        print(f"step {t+1}: dropped {int((~alive).sum())}, attempt diffs {int(natt_diff)}, worst {float(worst):.3e}, {time.time()-t0:.0f} s", flush=True)
st = exact.read_stats()
print(f"{dof}DOF {N} envs x {T} steps = {N*T:.3e} env-steps, {st['done']} episodes ({st['landed']} landed): dropped {int((~alive).sum())}, "
      f"attempt-count differences {int(natt_diff)}, worst normalised deviation {float(worst):.3e}")
