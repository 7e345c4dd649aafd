"""Development probe: end-to-end step with the kernel writing obs|reward|done straight into mapped pinned host memory (zero-copy)
versus the packed D2H copy of RocketVecEnv.step_host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_rocket_b200 import _native as N

from rl_rocket_b200.engine import _ptr
from rl_rocket_b200.vec_env import RocketVecEnv
from tests import goldutil as G

n = 65536
cfg = G.cfg6()
venv = RocketVecEnv(n, dof=6, max_episode_steps=800, **cfg)
venv.reset_tensor()
a = torch.rand((300, n, 3), device="cuda") * 2 - 1
for t in range(150): venv.step_tensor(a[t])
h_a = [a[150 + t].cpu().pin_memory() for t in range(100)]
def run(fn, reps=100):
    for t in range(5): fn(h_a[t])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(reps): fn(h_a[t])
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e6
print(f"step_host (packed D2H): {run(venv.step_host):.1f} us/step")
# zero-copy outputs
b = venv.batch
no = 14
h_pack = torch.zeros(n * no * 4 + n * 4 + n, dtype=torch.uint8).pin_memory()
h_obs = h_pack[:n * no * 4].view(torch.float32).view(n, no); h_rew = h_pack[n * no * 4:n * no * 4 + n * 4].view(torch.float32); h_done = h_pack[n * no * 4 + n * 4:]
ot = N.RlrStepOut(_ptr(h_obs), _ptr(h_rew), None, _ptr(h_done), _ptr(b.truncated), _ptr(b.terminal_obs), _ptr(b.fin_ret), _ptr(b.fin_len),
                  _ptr(b.flags), None, None, None, None, _ptr(b.stats))
d_act = torch.zeros((n, 3), device="cuda")
stream = torch.cuda.current_stream().cuda_stream
def zc(act_host):
    d_act.copy_(act_host, non_blocking=True)
    N.check(b.lib.rlr_step(b._params_ref, N.F64, b.math_mode, b._st, d_act.data_ptr(), 0, None, ot, n, 0, b.seed, stream))
    torch.cuda.current_stream().synchronize()
print(f"zero-copy outputs:      {run(zc):.1f} us/step")
def zc2(act_host):
    N.check(b.lib.rlr_step(b._params_ref, N.F64, b.math_mode, b._st, act_host.data_ptr(), 0, None, ot, n, 0, b.seed, stream))
    torch.cuda.current_stream().synchronize()
print(f"zero-copy in + out:     {run(zc2):.1f} us/step")
ot2 = N.RlrStepOut(_ptr(h_obs), _ptr(b.reward), None, _ptr(b.done), _ptr(b.truncated), _ptr(b.terminal_obs), _ptr(b.fin_ret), _ptr(b.fin_len),
                   _ptr(b.flags), None, None, None, None, _ptr(b.stats))
def zc3(act_host):
    N.check(b.lib.rlr_step(b._params_ref, N.F64, b.math_mode, b._st, act_host.data_ptr(), 0, None, ot2, n, 0, b.seed, stream))
    torch.cuda.current_stream().synchronize()
print(f"zero-copy in + obs only (reward/done stay on device): {run(zc3):.1f} us/step")
ot3 = N.RlrStepOut(_ptr(b.obs), _ptr(b.reward), None, _ptr(b.done), _ptr(b.truncated), _ptr(b.terminal_obs), _ptr(b.fin_ret), _ptr(b.fin_len),
                   _ptr(b.flags), None, None, None, None, _ptr(b.stats))
def zc4(act_host):
    N.check(b.lib.rlr_step(b._params_ref, N.F64, b.math_mode, b._st, act_host.data_ptr(), 0, None, ot3, n, 0, b.seed, stream))
    torch.cuda.current_stream().synchronize()
print(f"zero-copy in only (all outputs on device): {run(zc4):.1f} us/step")
def dev_only(act_host):
    N.check(b.lib.rlr_step(b._params_ref, N.F64, b.math_mode, b._st, d_act.data_ptr(), 0, None, ot3, n, 0, b.seed, stream))
    torch.cuda.current_stream().synchronize()
print(f"device only + sync: {run(dev_only):.1f} us/step")
print("obs finite:", bool(torch.isfinite(h_obs).all()), "done sum", int(h_done.sum()))
