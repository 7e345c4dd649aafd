import sys, os
sys.path.insert(0, os.getcwd())
import torch
from rl_rocket_b200.config import EnvSpec, env_config_6dof
from rl_rocket_b200.engine import RocketBatch
N, T = 20000, 300
spec = EnvSpec(6, max_episode_steps=60, **env_config_6dof())
g = torch.Generator(device="cuda").manual_seed(9)
acts = torch.rand((T, N, 3), device="cuda", generator=g) * 2 - 1
for mode in ("exact", "fast"):
    A = RocketBatch(spec, N, seed=5, math_mode=mode, debug_outputs=True); B = RocketBatch(spec, N, seed=5, math_mode=mode, debug_outputs=True)
    A.reset(); B.reset()
    norm = torch.as_tensor(spec.state_normalizer, device="cuda")[:, None]
    worst = 0.0; first = None; flagdiff = 0
    for t in range(T):
        A.step_tensor(acts[t], chained=True); B.step_tensor(acts[t])
        d = ((A.y - B.y).abs() / norm).max().item()
        flagdiff += int((A.done != B.done).sum()) + int((A.natt != B.natt).sum())
        if d > 0 and first is None: first = t
        worst = max(worst, d)
    print(mode, "worst normalised |chained - plain|", worst, "first differing step", first, "flag/attempt differences", flagdiff, "stats equal", A.read_stats() == B.read_stats())
