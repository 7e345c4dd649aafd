import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_rocket_b200.config import EnvSpec
from rl_rocket_b200.engine import RocketBatch
from tests import goldutil as G
spec = EnvSpec(3, max_episode_steps=3000, **G.cfg3())
N = int(os.environ.get("N3", "65536"))
env = RocketBatch(spec, N, seed=1); env.reset()
a = torch.rand((200, N, 2), device="cuda") * 2 - 1
for t in range(200): env.step_tensor(a[t])
torch.cuda.synchronize()
print(env.read_stats())
