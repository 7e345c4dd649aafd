"""GPU tests of the RLR_F32 precision (binary32 state) of the 6DOF path against the fp64 oracle.

Tolerance (BASELINE.json north_star): state within 1e-4 on the normaliser scale after 500 steps; done flags may differ from the
fp64 reference only where the fp64 margin to a threshold is below that tolerance -- such environments are counted, bounded and
dropped from the comparison from their first disagreement on (their trajectories are no longer the same episode)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from tests import goldutil as G

pytestmark = pytest.mark.gpu

F32_STATE_TOL = 1e-4
F32_REWARD_TOL = 1e-3


def _pair(N, tl, seed):
    from oracle import c_oracle as O
    from rl_rocket_b200.config import EnvSpec
    from rl_rocket_b200.engine import RocketBatch
    cfg = G.cfg6()
    spec = EnvSpec(6, max_episode_steps=tl, **cfg)
    env = RocketBatch(spec, N, seed=seed, math_mode="fast", precision="f32", debug_outputs=True)
    vec = O.VecOracle(O.make_params(6, max_episode_steps=tl, **cfg), N, seed=seed)
    return spec, env, vec


def test_f32_reset_is_exact():
    spec, env, vec = _pair(4096, 800, 3)
    assert np.array_equal(env.reset().cpu().numpy(), vec.reset())
    assert env.y.dtype == torch.float32
    assert np.array_equal(env.y.cpu().numpy().astype(np.float64), vec.y)          # ICs are f32 values


def test_f32_config3_one_step_65536_teacher_forced():
    """Config 3 at full size, one-step comparisons from the oracle's state (rounded to f32 on both sides)."""
    N = 65536
    spec, env, vec = _pair(N, 800, 2024)
    env.reset(); vec.reset()
    rng = np.random.RandomState(5)
    norm = spec.state_normalizer[:, None]
    worst, natt_diff = 0.0, 0
    for t in range(3):
        a = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
        env.step_tensor(torch.as_tensor(a).cuda())
        out = vec.step(a)
        same = env.done.cpu().numpy() == out["done"]
        assert (~same).sum() <= 2, "done flags differ on more than a threshold-margin handful of environments"
        dev = (np.abs(env.y_pre_reset.cpu().numpy().astype(np.float64) - out["y_pre_reset"]) / norm)[:, same]
        worst = max(worst, float(dev.max()))
        assert worst <= 1e-5, f"one-step f32 deviation {worst:.3e}"
        r = env.reward64.cpu().numpy()
        assert (np.abs(r - out["reward"])[same] <= F32_REWARD_TOL * np.maximum(1, np.abs(out["reward"][same]))).all()
        natt_diff += int((env.natt.cpu().numpy() != out["natt"]).sum())
        # teacher forcing: both sides continue from the f32-rounded oracle state
        y32 = vec.y.astype(np.float32)
        vec.y[:] = y32.astype(np.float64)
        env.y.copy_(torch.as_tensor(y32).cuda())
        env.k.copy_(torch.as_tensor(vec.k).cuda()); env.episode.copy_(torch.as_tensor(vec.episode).cuda())
        env.v0.copy_(torch.as_tensor(vec.v0).cuda()); env.ep_ret.copy_(torch.as_tensor(vec.ep_ret).cuda())
    assert natt_diff <= 3 * N * 0.002, f"{natt_diff} RK45 attempt counts differ"
    print(f"f32 one-step worst normalised deviation {worst:.3e}, attempt-count differences {natt_diff}")


def test_f32_free_running_500_steps():
    """2048 envs x 500 steps with auto-reset: every environment is compared until it first disagrees with the fp64 oracle on a done
    flag (a threshold crossing inside the f32 margin); the number of such environments is bounded."""
    N, T = 2048, 500
    spec, env, vec = _pair(N, 800, 99)
    env.reset(); vec.reset()
    rng = np.random.RandomState(6)
    norm = spec.state_normalizer[:, None]
    alive = np.ones(N, bool)
    worst, episodes = 0.0, 0
    for t in range(T):
        a = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
        env.step_tensor(torch.as_tensor(a).cuda())
        out = vec.step(a)
        alive &= env.done.cpu().numpy() == out["done"]
        dev = np.abs(env.y_pre_reset.cpu().numpy().astype(np.float64) - out["y_pre_reset"]) / norm
        worst = max(worst, float(dev[:, alive].max()))
        episodes += int(out["done"][alive].sum())
    dropped = int((~alive).sum())
    print(f"f32 free-running: worst normalised deviation {worst:.3e} over {int(alive.sum())} environments, {dropped} dropped, {episodes} episodes")
    assert worst <= F32_STATE_TOL
    assert dropped <= 0.02 * N, f"{dropped} environments crossed a threshold differently"
    assert episodes > N


def test_f32_rollout_kernel_equals_stepwise():
    from rl_rocket_b200.config import EnvSpec
    from rl_rocket_b200.engine import RocketBatch
    N, T = 1000, 60                      # ragged size: tail block partially filled
    spec = EnvSpec(6, max_episode_steps=40, **G.cfg6())
    a = torch.rand((T, N, 3), device="cuda") * 2 - 1
    e1 = RocketBatch(spec, N, seed=4, precision="f32"); e1.reset()
    e2 = RocketBatch(spec, N, seed=4, precision="f32"); e2.reset()
    obs_buf = torch.zeros((T, N, 14), device="cuda"); rew_buf = torch.zeros((T, N), device="cuda")
    done_buf = torch.zeros((T, N), dtype=torch.uint8, device="cuda")
    e1.rollout(a, obs_buf, rew_buf, done_buf)
    for t in range(T):
        o, r, d = e2.step_tensor(a[t])
        assert torch.equal(o, obs_buf[t]) and torch.equal(r, rew_buf[t]) and torch.equal(d, done_buf[t])
    assert torch.equal(e1.y, e2.y) and e1.read_stats() == e2.read_stats()
