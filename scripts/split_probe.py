"""Development aid: does splitting one step into sub-range launches on several streams (phase de-synchronisation) help?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_rocket_b200.config import EnvSpec, env_config_6dof
from rl_rocket_b200.engine import RocketBatch

n = 65536
spec = EnvSpec(6, max_episode_steps=800, **env_config_6dof())
env = RocketBatch(spec, n, seed=1, math_mode="fast"); env.reset()
a = torch.rand((300, n, 3), device="cuda") * 2 - 1
for t in range(150): env.step_tensor(a[t])
flush = torch.zeros(64 * 1024 * 1024, device="cuda"); sink = torch.zeros(1, device="cuda")
main = torch.cuda.current_stream()
side = [torch.cuda.Stream() for _ in range(7)]

def step_split(t, parts, streams, flushed):
    bounds = [round(n * x) // 64 * 64 for x in parts] + [n]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if flushed: torch.sum(flush, dim=(0,), keepdim=True, out=sink)
    ev0.record()
    if streams > 1:
        for s in side[:streams - 1]: s.wait_stream(main)
    for j in range(len(bounds) - 1):
        st = main if (streams == 1 or j % streams == 0) else side[j % streams - 1]
        env.step_raw(a[t].data_ptr(), env._ot, stream=st.cuda_stream, begin=bounds[j], count=bounds[j + 1] - bounds[j])
    if streams > 1:
        for s in side[:streams - 1]: main.wait_stream(s)
    ev1.record()
    return ev0, ev1

for name, parts, streams in (("single", [0.0], 1), ("2 halves, 2 streams", [0.0, 0.5], 2), ("2 halves, 1 stream (PDL chain)", [0.0, 0.5], 1),
                             ("60/40, 2 streams", [0.0, 0.6], 2), ("4 quarters, 4 streams", [0.0, 0.25, 0.5, 0.75], 4),
                             ("4 quarters, 2 streams", [0.0, 0.25, 0.5, 0.75], 2), ("8 parts, 8 streams", [i / 8 for i in range(8)], 8)):
    for flushed in (True, False):
        for rep in range(2):
            evs = [step_split(150 + (t % 100), parts, streams, flushed) for t in range(100)]
            torch.cuda.synchronize()
        ts = sorted(x.elapsed_time(y) for x, y in evs)
        print(f"{name:34s} {'flushed' if flushed else 'warm   '}: mean {sum(ts)/len(ts)*1e3:6.1f} us  median {ts[50]*1e3:6.1f}  min {ts[0]*1e3:6.1f}", flush=True)
