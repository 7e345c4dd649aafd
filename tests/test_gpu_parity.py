"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors of the live reference and
against the C oracle on seeded inputs.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from tests import goldutil as G

pytestmark = pytest.mark.gpu


def _engine():
    from rl_rocket_b200.config import EnvSpec
    from rl_rocket_b200.engine import RocketBatch
    return EnvSpec, RocketBatch


def _spec(name, **over):
    EnvSpec, _ = _engine()
    dof, cfgf, tl, mode, akind = G.TRAJ[name]
    kw = dict(max_episode_steps=tl, reward_mode=mode, obs_mode=G.OBS_MODE.get(name, 0), aero_mode=G.AERO_MODE.get(name, 0), **cfgf())
    if akind == "discrete":
        kw["discrete_table"] = G.DISCRETE_TABLE
    kw.update(over)
    return EnvSpec(dof, **kw), akind


@pytest.mark.parametrize("math_mode", ["exact", "fast"])
@pytest.mark.parametrize("name", sorted(G.TRAJ))
def test_golden_trajectories(name, math_mode):
    """Every trajectory the live reference produced (tests/golden), replayed on the GPU with the same ICs and actions."""
    _, RocketBatch = _engine()
    spec, akind = _spec(name)
    g = G.load(name)
    T, N = g["done"].shape
    env = RocketBatch(spec, N, seed=int(g["seed"]), math_mode=math_mode, debug_outputs=True)
    ics = torch.as_tensor(g["ics"]).cuda() if "ics" in g.files else None      # crafted cases: every episode of env i restarts from ics[i]
    obs0 = env.reset(ic=ics).cpu().numpy()
    assert np.abs(obs0.astype(float) - g["obs0"]).max() <= G.OBS_TOL
    norm = spec.state_normalizer
    worst = 0.0
    for t in range(T):
        a = torch.as_tensor(g["actions"][t]).cuda()
        obs, rew, done = env.step_tensor(a, ic_inject=ics)
        gs = {k: g[k][t] for k in ("done", "truncated", "nfev", "y", "reward", "obs")}
        d, _ = G.compare_step(f"{name}/{math_mode}", t, norm, gs, env.y_pre_reset.cpu().numpy(), obs.cpu().numpy(),
                              env.reward64.cpu().numpy(), done.cpu().numpy(), env.truncated.cpu().numpy(), env.natt.cpu().numpy())
        dn = g["done"][t].astype(bool)
        assert np.abs(env.terminal_obs.cpu().numpy()[dn].astype(float) - g["terminal_obs"][t][dn]).max(initial=0) <= G.OBS_TOL
        worst = max(worst, d)
        if "rew_goal" in g.files:      # the success branch (rocket_env.py:1040-1061 / :449-476): landed flag bit-exact
            assert np.array_equal(((env.flags.cpu().numpy() >> 1) & 1).astype(bool), g["rew_goal"][t] != 0), f"{name} step {t}: landed flags differ"
    print(f"{name} [{math_mode}] worst normalised state deviation {worst:.3e}")


def _oracle_pair(dof, cfg, N, tl, seed, math_mode, mode=0):
    from oracle import c_oracle as O
    EnvSpec, RocketBatch = _engine()
    spec = EnvSpec(dof, max_episode_steps=tl, reward_mode=mode, **cfg)
    p = O.make_params(dof, max_episode_steps=tl, reward_mode=mode, **cfg)
    env = RocketBatch(spec, N, seed=seed, math_mode=math_mode, debug_outputs=True)
    vec = O.VecOracle(p, N, seed=seed)
    return spec, env, vec


def _compare_with_oracle(tag, spec, env, out, state_tol=G.STATE_TOL):
    norm = spec.state_normalizer[:, None]
    assert np.array_equal(env.done.cpu().numpy(), out["done"]), f"{tag}: done flags differ"
    assert np.array_equal(env.truncated.cpu().numpy(), out["truncated"]), f"{tag}: truncated flags differ"
    assert np.array_equal(env.natt.cpu().numpy(), out["natt"]), f"{tag}: RK45 attempt counts differ"
    dev = np.abs(env.y_pre_reset.cpu().numpy() - out["y_pre_reset"]) / norm
    assert dev.max() <= state_tol, f"{tag}: state deviates {dev.max():.3e}"
    r = env.reward64.cpu().numpy()
    assert (np.abs(r - out["reward"]) <= G.REWARD_TOL * np.maximum(1, np.abs(out["reward"]))).all(), f"{tag}: reward deviates"
    assert np.abs(env.obs.cpu().numpy().astype(float) - out["obs"]).max() <= G.OBS_TOL, f"{tag}: obs deviates"
    return float(dev.max())


@pytest.mark.parametrize("math_mode", ["exact", "fast"])
def test_config3_6dof_65536_teacher_forced(math_mode):
    """BASELINE config 3 at full size: 65 536 envs, Philox ICs, random actions; every env compared with the oracle for 3 steps."""
    N = 65536
    spec, env, vec = _oracle_pair(6, G.cfg6(), N, 800, 2024, math_mode)
    assert np.array_equal(env.reset().cpu().numpy(), vec.reset())        # reset is pure f32/f64 IEEE arithmetic: bit-exact
    rng = np.random.RandomState(5)
    worst = 0.0
    for t in range(3):
        a = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
        env.step_tensor(torch.as_tensor(a).cuda())
        out = vec.step(a)
        worst = max(worst, _compare_with_oracle(f"cfg3 step {t}", spec, env, out))
        # teacher forcing: continue both sides from the oracle's state so that every step is a fresh one-step comparison
        env.y.copy_(torch.as_tensor(vec.y).cuda())
    print(f"config 3 one-step [{math_mode}] worst normalised deviation {worst:.3e}")


@pytest.mark.parametrize("math_mode", ["exact", "fast"])
def test_config3_6dof_subset_500_steps(math_mode):
    """Free-running parity (auto-reset included) on 2048 envs x 500 steps of config 3."""
    N, T = 2048, 500
    spec, env, vec = _oracle_pair(6, G.cfg6(), N, 800, 99, math_mode)
    env.reset(); vec.reset()
    rng = np.random.RandomState(6)
    worst, episodes = 0.0, 0
    for t in range(T):
        a = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
        env.step_tensor(torch.as_tensor(a).cuda())
        out = vec.step(a)
        worst = max(worst, _compare_with_oracle(f"cfg3 free t={t}", spec, env, out))
        episodes += int(out["done"].sum())
    assert episodes > N            # every env finished several episodes
    st = env.read_stats()
    assert st["steps"] == N * T and st["done"] == episodes
    print(f"config 3 free-running [{math_mode}] worst {worst:.3e}, {episodes} episodes")


@pytest.mark.parametrize("math_mode", ["exact", "fast"])
def test_config2_3dof_4096_500_steps(math_mode):
    """BASELINE config 2: 3DOF planar env (main.py config), 4096 envs, 500 random-action steps."""
    N, T = 4096, 500
    spec, env, vec = _oracle_pair(3, G.cfg3(), N, 3000, 7, math_mode)
    env.reset(); vec.reset()
    rng = np.random.RandomState(8)
    worst, episodes = 0.0, 0
    for t in range(T):
        a = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        env.step_tensor(torch.as_tensor(a).cuda())
        out = vec.step(a)
        worst = max(worst, _compare_with_oracle(f"cfg2 t={t}", spec, env, out))
        episodes += int(out["done"].sum())
    assert episodes > 0
    print(f"config 2 [{math_mode}] worst {worst:.3e}, {episodes} episodes")


def test_rollout_kernel_equals_stepwise():
    """The fused T-step rollout launch is bit-identical to T single-step launches (exact math: the two kernels execute the
    same IEEE operation sequence; in fast mode the compiler may contract FMAs differently in the two kernels)."""
    EnvSpec, RocketBatch = _engine()
    N, T = 4096, 200
    spec = EnvSpec(6, max_episode_steps=800, **G.cfg6())
    a = torch.rand((T, N, 3), device="cuda") * 2 - 1
    e1 = RocketBatch(spec, N, seed=3, math_mode="exact")
    e2 = RocketBatch(spec, N, seed=3, math_mode="exact")
    e1.reset(); e2.reset()
    obs_buf = torch.zeros((T, N, 14), device="cuda")
    rew_buf = torch.zeros((T, N), device="cuda")
    done_buf = torch.zeros((T, N), dtype=torch.uint8, device="cuda")
    e2.rollout(a, obs_buf, rew_buf, done_buf)
    for t in range(T):
        obs, rew, done = e1.step_tensor(a[t])
        assert torch.equal(obs, obs_buf[t]) and torch.equal(rew, rew_buf[t]) and torch.equal(done, done_buf[t])
    assert torch.equal(e1.y, e2.y) and torch.equal(e1.k, e2.k) and torch.equal(e1.episode, e2.episode)
    assert e1.read_stats() == e2.read_stats()


def test_rollout_kernel_fast_matches_stepwise_within_tolerance():
    EnvSpec, RocketBatch = _engine()
    N, T = 4096, 200
    spec = EnvSpec(6, max_episode_steps=800, **G.cfg6())
    a = torch.rand((T, N, 3), device="cuda") * 2 - 1
    e1 = RocketBatch(spec, N, seed=3, math_mode="fast")
    e2 = RocketBatch(spec, N, seed=3, math_mode="fast")
    e1.reset(); e2.reset()
    done_buf = torch.zeros((T, N), dtype=torch.uint8, device="cuda")
    rew_buf = torch.zeros((T, N), device="cuda")
    e2.rollout(a, None, rew_buf, done_buf)
    for t in range(T):
        obs, rew, done = e1.step_tensor(a[t])
        assert torch.equal(done, done_buf[t])
        assert (rew - rew_buf[t]).abs().max().item() <= 1e-5
    norm = torch.as_tensor(spec.state_normalizer, device="cuda")[:, None]
    assert ((e1.y - e2.y).abs() / norm).max().item() <= G.STATE_TOL
    assert torch.equal(e1.k, e2.k) and torch.equal(e1.episode, e2.episode)


def test_sharding_is_invariant():
    """Trajectories are keyed by the GLOBAL env index: two half-size shards reproduce the full batch bit for bit."""
    EnvSpec, RocketBatch = _engine()
    N, T = 2048, 150
    spec = EnvSpec(6, max_episode_steps=800, **G.cfg6())
    a = torch.rand((T, N, 3), device="cuda") * 2 - 1
    full = RocketBatch(spec, N, seed=11)
    lo = RocketBatch(spec, N // 2, seed=11, env_offset=0)
    hi = RocketBatch(spec, N // 2, seed=11, env_offset=N // 2)
    for e in (full, lo, hi):
        e.reset()
    for t in range(T):
        full.step_tensor(a[t]); lo.step_tensor(a[t, : N // 2].contiguous()); hi.step_tensor(a[t, N // 2:].contiguous())
    assert torch.equal(full.y[:, : N // 2], lo.y) and torch.equal(full.y[:, N // 2:], hi.y)
    assert torch.equal(full.obs[N // 2:], hi.obs)
    s, a_, b = full.read_stats(), lo.read_stats(), hi.read_stats()
    assert all(s[k] == a_[k] + b[k] for k in ("steps", "attempts", "done", "landed", "viol", "trunc", "len_sum"))


def test_full_size_invariants():
    """Size-independent properties at 65 536 envs x 100 steps: unit quaternions, nfev in {8,14,20,...}, finite obs, counters."""
    EnvSpec, RocketBatch = _engine()
    N, T = 65536, 100
    spec = EnvSpec(6, max_episode_steps=800, **G.cfg6())
    env = RocketBatch(spec, N, seed=1, debug_outputs=True)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    steps = torch.zeros(N, dtype=torch.int64, device="cuda")
    for t in range(T):
        a = torch.rand((N, 3), device="cuda", generator=g) * 2 - 1
        obs, rew, done = env.step_tensor(a)
        q = env.y[6:10]
        assert (q.pow(2).sum(0).sqrt() - 1).abs().max().item() < 1e-6        # f32-normalised ICs right after a reset
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
        assert int(env.natt.min()) >= 1
        steps += 1
        # k counts steps since the last reset
        assert torch.equal(env.k.long() * (1 - done.long()), steps * (1 - done.long()))
        steps = steps * (1 - done.long())
    st = env.read_stats()
    assert st["steps"] == N * T and st["done"] > 0 and st["attempts"] >= st["steps"]
    assert st["len_sum"] > 0 and st["viol"] + st["trunc"] <= st["done"]


def test_empty_and_ragged_sizes():
    EnvSpec, RocketBatch = _engine()
    spec = EnvSpec(6, max_episode_steps=800, **G.cfg6())
    for n in (1, 31, 33, 129):
        env = RocketBatch(spec, n, seed=2)
        ref = RocketBatch(spec, 256, seed=2)
        env.reset(); ref.reset()
        a = torch.rand((256, 3), device="cuda") * 2 - 1
        for _ in range(5):
            env.step_tensor(a[:n].contiguous()); ref.step_tensor(a)
        assert torch.equal(env.y, ref.y[:, :n])
    env0 = RocketBatch(spec, 0, seed=2)
    env0.reset()
    env0.step_tensor(torch.zeros((0, 3), device="cuda"))


def test_config4_one_million_envs_sharded():
    """BASELINE config 4 (2^20 envs = 8 x 131 072): the eight shards, stepped separately with their global env offsets, reproduce the
    single full batch bit for bit (reset RNG and inputs keyed by the global index), and the shard statistics add up."""
    EnvSpec, RocketBatch = _engine()
    N, W, T = 1 << 20, 8, 3
    spec = EnvSpec(6, max_episode_steps=800, **G.cfg6())
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.rand((T, N, 3), device="cuda", generator=g) * 2 - 1
    full = RocketBatch(spec, N, seed=21)
    full.reset()
    for t in range(T):
        full.step_tensor(a[t])
    tot = {k: 0 for k in ("steps", "attempts", "done")}
    n = N // W
    for r in range(W):
        sh = RocketBatch(spec, n, seed=21, env_offset=r * n)
        sh.reset()
        for t in range(T):
            sh.step_tensor(a[t, r * n:(r + 1) * n].contiguous())
        assert torch.equal(sh.y, full.y[:, r * n:(r + 1) * n]) and torch.equal(sh.obs, full.obs[r * n:(r + 1) * n])
        st = sh.read_stats()
        for k in tot:
            tot[k] += st[k]
        del sh
    fs = full.read_stats()
    assert all(fs[k] == tot[k] for k in tot) and fs["steps"] == N * T


def test_non_finite_state_ends_the_episode_instead_of_hanging():
    """A NaN / Inf state (e.g. a user-injected initial condition) must not loop forever in the step-size controller: the solve reports
    failure, which ends the episode exactly like scipy's status -1 does in the reference (rocket_env.py:702)."""
    EnvSpec, RocketBatch = _engine()
    spec = EnvSpec(6, max_episode_steps=800, **G.cfg6())
    for mode in ("exact", "fast"):
        env = RocketBatch(spec, 64, seed=1, math_mode=mode)
        env.reset()
        env.y[3, 5] = float("nan")
        env.y[0, 9] = float("inf")
        obs, rew, done = env.step_tensor(torch.zeros((64, 3), device="cuda"))
        torch.cuda.synchronize()
        d = done.cpu().numpy().astype(bool)
        assert d[5] and d[9] and d.sum() == 2
        assert torch.isfinite(env.y).all()          # both were reset


def test_fast_and_exact_builds_agree_on_every_flag_at_full_size():
    """65 536 envs x 300 free-running steps (2e7 env steps, ~2.7e5 episodes): the throughput build (FMA contraction, hardware
    reciprocals, Newton event root, algebraic rearrangements) takes the same done / truncated decisions and RK45 attempt counts as the
    op-for-op IEEE build, and their states stay within the parity tolerance.  The north star asks for bit-exact done flags and step counts:
    over these 65 536 x 300 = 2.0e7 env steps (seeded, deterministic) NOT ONE environment may disagree on a flag or an attempt count."""
    EnvSpec, RocketBatch = _engine()
    N, T = 65536, 300
    spec = EnvSpec(6, max_episode_steps=800, **G.cfg6())
    fast = RocketBatch(spec, N, seed=17, math_mode="fast", debug_outputs=True)
    exact = RocketBatch(spec, N, seed=17, math_mode="exact", debug_outputs=True)
    fast.reset(); exact.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    norm = torch.as_tensor(spec.state_normalizer, device="cuda")[:, None]
    alive = torch.ones(N, dtype=torch.bool, device="cuda")
    worst = torch.zeros((), dtype=torch.float64, device="cuda")
    natt_diff = torch.zeros((), dtype=torch.int64, device="cuda")
    for t in range(T):
        a = torch.rand((N, 3), device="cuda", generator=g) * 2 - 1
        fast.step_tensor(a); exact.step_tensor(a)
        alive &= (fast.done == exact.done) & (fast.truncated == exact.truncated)
        natt_diff += ((fast.natt != exact.natt) & alive).sum()
        dev = ((fast.y_pre_reset - exact.y_pre_reset).abs() / norm).amax(0)
        worst = torch.maximum(worst, torch.where(alive, dev, torch.zeros_like(dev)).max())
    dropped = int((~alive).sum())
    print(f"fast vs exact at full size: worst {float(worst):.3e}, {dropped} envs dropped, {int(natt_diff)} attempt-count differences, "
          f"{exact.read_stats()['done']} episodes")
    assert dropped == 0 and int(natt_diff) == 0
    assert float(worst) <= G.STATE_TOL


@pytest.mark.parametrize("math_mode", ["exact", "fast"])
def test_3dof_65536_envs_one_step_teacher_forced(math_mode):
    """The 3DOF path at 65 536 envs (batches above 56 832 envs run the 128-register instantiation): every env against the oracle."""
    N = 65536
    spec, env, vec = _oracle_pair(3, G.cfg3(), N, 3000, 31, math_mode)
    assert np.array_equal(env.reset().cpu().numpy(), vec.reset())
    rng = np.random.RandomState(9)
    worst = 0.0
    for t in range(3):
        a = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        env.step_tensor(torch.as_tensor(a).cuda())
        out = vec.step(a)
        worst = max(worst, _compare_with_oracle(f"3dof 65536 step {t}", spec, env, out))
        env.y.copy_(torch.as_tensor(vec.y).cuda())
    print(f"3DOF 65536 one-step [{math_mode}] worst normalised deviation {worst:.3e}")


@pytest.mark.parametrize("math_mode", ["exact", "fast"])
@pytest.mark.parametrize("dof", [6, 3])
def test_min_thrust_extension_matches_the_oracle(dof, math_mode):
    """`min_thrust` (extension: the reference's minThrust TODO, rocket_env.py:405, :980): CUDA against the C oracle, both builds."""
    from oracle import c_oracle as O
    EnvSpec, RocketBatch = _engine()
    cfg, tl, na = (G.cfg6(), 800, 3) if dof == 6 else (G.cfg3(), 3000, 2)
    N, T, floor = 2048, 40, 0.3 * 981e3
    spec = EnvSpec(dof, max_episode_steps=tl, min_thrust=floor, **cfg)
    p = O.make_params(dof, max_episode_steps=tl, min_thrust=floor, **cfg)
    env = RocketBatch(spec, N, seed=8, math_mode=math_mode, debug_outputs=True)
    vec = O.VecOracle(p, N, seed=8)
    assert np.array_equal(env.reset().cpu().numpy(), vec.reset())
    rng = np.random.RandomState(2)
    for t in range(T):
        a = rng.uniform(-1, 1, (N, na)).astype(np.float32)
        env.step_tensor(torch.as_tensor(a).cuda())
        _compare_with_oracle(f"min_thrust {dof}dof step {t}", spec, env, vec.step(a))
    # the mass flow is thrust / (Isp g0) >= floor / (Isp g0): after T steps every environment has burnt at least that much
    burnt = np.float64(cfg["IC"][-1]) + np.float64(cfg["ICRange"][-1]) / 2 - env.y[-1].cpu().numpy()
    alive = env.k.cpu().numpy() == T
    assert (burnt[alive] >= T * spec.timestep * floor / (360 * 9.81) * 0.999).all()


@pytest.mark.parametrize("math_mode", ["exact", "fast"])
def test_cooperative_auto_reset_equals_the_reset_kernel(math_mode):
    """The step kernel's auto-reset (computed and stored cooperatively by the warp: lane b of a group of eight owns components 2b, 2b+1
    of a finished environment's initial condition) against the reset kernel (one thread per environment, scalar sampler): with
    TimeLimit 1 every environment finishes at every step -- 32 resets per warp, eight passes of four -- and must restart bit-identical
    to an explicit reset of the same episode index.  1000 environments: the last warp is partial."""
    EnvSpec, RocketBatch = _engine()
    N = 1000
    a = torch.rand((N, 3), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) * 2 - 1
    A = RocketBatch(EnvSpec(6, max_episode_steps=1, **G.cfg6()), N, seed=7, math_mode=math_mode)
    B = RocketBatch(EnvSpec(6, max_episode_steps=800, **G.cfg6()), N, seed=7, math_mode=math_mode)
    assert torch.equal(A.reset(), B.reset())
    for t in range(3):
        obs, rew, done = A.step_tensor(a)
        assert bool(done.all()) and bool(A.truncated.bool().logical_or(A.flags.bool()).all())
        obs_b = B.reset()                      # the reset kernel, next episode index
        assert torch.equal(A.y, B.y) and torch.equal(A.v0, B.v0) and torch.equal(A.episode, B.episode)
        assert int(A.k.abs().max()) == 0 and float(A.ep_ret.abs().max()) == 0.0
        assert torch.equal(obs, obs_b)
    # injected initial conditions take the same cooperative path
    ic = torch.as_tensor(np.tile(np.float32(G.cfg6()["IC"]), (N, 1)) + np.random.RandomState(1).uniform(-0.05, 0.05, (N, 14)).astype(np.float32))
    obs, rew, done = A.step_tensor(a, ic_inject=ic)
    C = RocketBatch(EnvSpec(6, max_episode_steps=800, **G.cfg6()), N, seed=7, math_mode=math_mode)
    obs_c = C.reset(ic=ic)
    assert torch.equal(A.y, C.y) and torch.equal(A.v0, C.v0) and torch.equal(obs, obs_c)


@pytest.mark.parametrize("math_mode", ["exact", "fast"])
def test_aero_mode_extension_matches_the_oracle(math_mode):
    """`aero_mode = 1` (extension, SURVEY.md 8(f) N4: the angle-of-attack line of Simulator3DOF.RHS, simulator.py:103-104, enabled):
    CUDA (`Sys3Aero`, generic solver in both builds) against the C oracle, 2 048 envs x 300 free-running steps with auto-reset; the
    golden files of the live reference's RHS are replayed by test_golden_trajectories."""
    from oracle import c_oracle as O
    EnvSpec, RocketBatch = _engine()
    cfg, N, T = G.cfg3_default(), 2048, 300
    spec = EnvSpec(3, max_episode_steps=2000, aero_mode=1, **cfg)
    p = O.make_params(3, max_episode_steps=2000, aero_mode=1, **cfg)
    env = RocketBatch(spec, N, seed=11, math_mode=math_mode, debug_outputs=True)
    vec = O.VecOracle(p, N, seed=11)
    assert np.abs(env.reset().cpu().numpy().astype(float) - vec.reset(None)).max() <= G.OBS_TOL
    rng = np.random.RandomState(5)
    worst, finished = 0.0, 0
    for t in range(T):
        a = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        env.step_tensor(torch.as_tensor(a).cuda())
        out = vec.step(a)
        worst = max(worst, _compare_with_oracle(f"aero 3dof step {t}", spec, env, out))
        finished += int(out["done"].sum())
    assert finished > N // 2          # touchdowns / bound violations happened: events and resets are covered
    print(f"aero_mode [{math_mode}] worst normalised state deviation {worst:.3e}, {finished} finished episodes")
    # the fused rollout takes the same path
    env2 = RocketBatch(spec, 512, seed=3, math_mode=math_mode); env3 = RocketBatch(spec, 512, seed=3, math_mode=math_mode)
    env2.reset(); env3.reset()
    acts = torch.rand((20, 512, 2), device="cuda") * 2 - 1
    env2.rollout(acts)
    for t in range(20):
        env3.step_tensor(acts[t])
    dev = ((env2.y - env3.y).abs() / torch.as_tensor(spec.state_normalizer, device="cuda")[:, None]).max().item()
    assert dev <= (0.0 if math_mode == "exact" else 1e-9)


@pytest.mark.parametrize("math_mode", ["exact", "fast"])
def test_chained_steps_equal_plain_steps(math_mode):
    """Open-loop chaining (rlr_step_chained: block b of a launch orders itself after block b of the previous chained launch on the
    same state, not after the whole predecessor grid) changes when blocks run, not what they compute: 300 chained steps of one batch
    -- with resets, ground events and a chain broken by a plain step, a masked reset and a rollout in the middle -- and a ring of three
    batches stepped in turn against plain steps.  Exact build: bit-identical state, outputs and statistics.  Fast build: the chained
    and the plain kernel are two compilations of the same source and ptxas contracts multiply-adds per kernel, so the state may differ
    in the last bits (measured 1e-15 of the normaliser); done / truncated flags, episode counters, step counters and the statistics'
    counts must still be identical."""
    EnvSpec, RocketBatch = _engine()
    N, T = 20000, 300                       # 313 blocks, the last one partial
    spec = EnvSpec(6, max_episode_steps=60, **G.cfg6())
    norm = torch.as_tensor(spec.state_normalizer, device="cuda")[:, None]
    tol = 0.0 if math_mode == "exact" else 1e-12

    def same(a, b, t):
        assert torch.equal(a.done, b.done) and torch.equal(a.truncated, b.truncated) and torch.equal(a.flags, b.flags), f"step {t}: flags"
        assert torch.equal(a.k, b.k) and torch.equal(a.episode, b.episode), f"step {t}: counters"
        assert ((a.y - b.y).abs() / norm).max().item() <= tol, f"step {t}: state"
        assert (a.obs - b.obs).abs().max().item() <= (0.0 if tol == 0.0 else 2e-7), f"step {t}: obs"
        # (the reward is evaluated on the float32 copy of the state, rocket_env.py:695: a last-bit difference of the f64 state can flip
        # a float32 rounding, which moves a reward term by ~1e-8 and the running return with it)
        assert (a.reward64 - b.reward64).abs().max().item() <= tol * 1e6 and (a.ep_ret - b.ep_ret).abs().max().item() <= tol * 1e8, f"step {t}: reward"

    def same_stats(a, b):
        sa, sb = a.read_stats(), b.read_stats()
        assert {k: v for k, v in sa.items() if k != "ret_sum"} == {k: v for k, v in sb.items() if k != "ret_sum"}
        assert abs(sa["ret_sum"] - sb["ret_sum"]) <= (0.0 if tol == 0.0 else 1e-3)
        return sa

    g = torch.Generator(device="cuda").manual_seed(9)
    acts = torch.rand((T, N, 3), device="cuda", generator=g) * 2 - 1
    A = RocketBatch(spec, N, seed=5, math_mode=math_mode)
    B = RocketBatch(spec, N, seed=5, math_mode=math_mode)
    A.reset(); B.reset()
    for t in range(T):
        if t == 100:                        # a plain step, a masked reset and a fused rollout break the chain and restart it
            A.step_tensor(acts[t]); B.step_tensor(acts[t])
            m = (torch.arange(N, device="cuda") % 7 == 0)
            A.reset(mask=m); B.reset(mask=m)
            A.rollout(acts[t:t + 3]); B.rollout(acts[t:t + 3])
            continue
        A.step_tensor(acts[t], chained=True)
        B.step_tensor(acts[t])
        if t % 25 == 0 or t == T - 1:
            same(A, B, t)
    assert same_stats(A, B)["done"] > N
    # three batches stepped in turn on one stream (what bench.py's ring does), chained vs plain
    ring_c = [RocketBatch(spec, 4096, seed=20 + j, math_mode=math_mode) for j in range(3)]
    ring_p = [RocketBatch(spec, 4096, seed=20 + j, math_mode=math_mode) for j in range(3)]
    for b in ring_c + ring_p:
        b.reset()
    norm = norm[:, :1]
    acts4 = acts[:90, :4096].contiguous()           # (the actions of a chained step must exist before the step is enqueued)
    torch.cuda.synchronize()
    for t in range(90):
        ring_c[t % 3].step_tensor(acts4[t], chained=True)
    for t in range(90):
        ring_p[t % 3].step_tensor(acts4[t])
    for c, p_ in zip(ring_c, ring_p):
        same(c, p_, 90)
        same_stats(c, p_)
    with pytest.raises(ValueError):
        RocketBatch(EnvSpec(3, **G.cfg3()), 64).step_tensor(torch.zeros((64, 2), device="cuda"), chained=True)
