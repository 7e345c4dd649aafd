"""Development probe: RocketVecEnv.step_host per rank under torchrun, zero-copy vs copy path (PCIe contention between GPUs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from rl_rocket_b200.vec_env import RocketVecEnv
from tests import goldutil as G
n = 65536
venv = RocketVecEnv(n, dof=6, max_episode_steps=800, device=dev, env_offset=rank * n, **G.cfg6())
venv.reset_tensor()
a = torch.rand((250, n, 3), device=dev) * 2 - 1
for t in range(150): venv.step_tensor(a[t])
h_a = [a[150 + t].cpu().pin_memory() for t in range(100)]
for zc in (True, False, True, False):
    for t in range(5): venv.step_host(h_a[t], zero_copy=zc)
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(100): venv.step_host(h_a[t], zero_copy=zc)
    torch.cuda.synchronize(); dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"world {world} {'zero_copy' if zc else 'copy'}: {dt.item()/100*1e6:.1f} us/step (max over ranks) -> {n*world*100/dt.item():.3e} env-steps/s", flush=True)
dist.destroy_process_group()
