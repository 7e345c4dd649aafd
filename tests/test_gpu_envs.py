"""GPU tests of the reference-facing surface: gym single-env facades, the SB3-style VecEnv (infos / auto-reset / Monitor stats),
wrappers, and an SB3 `collect_rollouts`-shaped harness (BASELINE config 5).  Golden data = the live reference's own outputs."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from tests import goldutil as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["null", "nonnull", "mixed"])
def test_rocket6dof_facade_replays_kat2(tag):
    """test_6DOF_environment.py:12-34 scenario through the gym facade (list / ndarray actions, reward as float, info dict)."""
    from rl_rocket_b200.envs import Rocket6DOF
    g = G.load("kat_6dof.npz")
    cfg = G.cfg6()
    cfg["ICRange"] = [0] * 14
    env = Rocket6DOF(**cfg)
    assert env.observation_space.shape == (14,) and env.action_space.shape == (3,)
    obs = env.reset()
    assert obs.dtype == np.float32 and np.array_equal(obs, g[f"kat2_{tag}_obs"][0])
    act = g[f"kat2_{tag}_action"]
    act = act.tolist() if tag == "null" else act              # the reference script passes a python list
    n, done, R = 0, False, 0.0
    while not done:
        obs, r, done, info = env.step(act)
        assert isinstance(r, float) and isinstance(done, bool) and obs.dtype == np.float32
        assert np.abs((env.state - g[f"kat2_{tag}_y"][n]) / env.state_normalizer).max() < 1e-12
        assert abs(r - g[f"kat2_{tag}_reward"][n]) < 1e-11
        assert done == bool(g[f"kat2_{tag}_done"][n]) and info["bounds_violation"] == bool(g[f"kat2_{tag}_viol"][n])
        assert np.abs(obs.astype(float) - g[f"kat2_{tag}_obs"][n + 1]).max() < G.OBS_TOL
        assert np.abs(np.array(list(info["rewards_dict"].values())) - g[f"kat2_{tag}_terms"][n]).max() < 1e-11
        R += r
        n += 1
    assert n == len(g[f"kat2_{tag}_reward"]) and len(env.SIM.states) == n + 1
    assert abs(R - g[f"kat2_{tag}_reward"].sum()) < 1e-9
    assert env.used_mass() >= 0 and env.states_to_dataframe().shape == (n + 1, 14)


def test_rocket3dof_facade_replays_kat3():
    from rl_rocket_b200.envs import Rocket
    g = G.load("kat_3dof.npz")
    cfg = G.cfg3()
    cfg["ICRange"] = [0] * 7
    env = Rocket(**cfg)
    obs = env.reset()
    assert obs.dtype == np.float64 and np.abs(obs - g["kat3_mixed_obs"][0]).max() < 1e-15
    n, done = 0, False
    while not done:
        obs, r, done, info = env.step(g["kat3_mixed_action"])
        assert np.abs((env.y - g["kat3_mixed_y"][n]) / env.state_normalizer).max() < 1e-11
        assert abs(r - g["kat3_mixed_reward"][n]) < 1e-10 and done == bool(g["kat3_mixed_done"][n])
        assert np.abs(obs - g["kat3_mixed_obs"][n + 1]).max() < 1e-6
        n += 1
    assert n == 285


def test_wrappers_on_the_facades():
    from rl_rocket_b200.envs import Rocket, Rocket1DOF
    from rl_rocket_b200 import wrappers as W
    env = W.TimeLimit(W.DiscreteActions3DOF(Rocket(**G.cfg3())), max_episode_steps=5)
    env.reset()
    for k in range(5):
        obs, r, done, info = env.step(2)
    assert done and info["TimeLimit.truncated"] is True and obs.shape == (7,)
    e1 = Rocket1DOF(timestep=0.05, reward_coeff=G.cfg3()["reward_coeff"])
    o = e1.reset()
    assert o.shape == (2,) and e1.action_space.shape == (1,)
    o, r, d, info = e1.step(np.float32([0.3]))
    assert o.shape == (2,) and abs(e1.y[0]) < 1e-6 and abs(e1.y[3]) < 1e-5        # gimbal 0: stays on the vertical (theta = f32(pi/2))
    a = W.RewardAnnealing(Rocket(**G.cfg3()))
    a.reset()
    _, r, _, info = a.step(np.float32([0.1, -0.5]))
    assert set(info["rewards_dict"]) == {"attitude_constraint", "rew_goal", "thrust_penalty"} and r == pytest.approx(-0.004 * 0.5)


def test_vecenv_matches_golden_dummyvecenv_semantics():
    """RocketVecEnv.step == DummyVecEnv([Monitor(TimeLimit(env))]).step on the golden trajectory: obs after auto-reset,
    f32 rewards, dones, terminal_observation, TimeLimit.truncated and Monitor's episode r / l."""
    from rl_rocket_b200.vec_env import RocketVecEnv
    for name, tl in (("traj_6dof.npz", 800), ("traj_6dof_timelimit.npz", 25)):
        g = G.load(name)
        T, N = g["done"].shape
        venv = RocketVecEnv(N, dof=6, max_episode_steps=tl, math_mode="exact", **G.cfg6())
        venv.seed(int(g["seed"]))
        obs = venv.reset()
        assert obs.shape == (N, 14) and obs.dtype == np.float32 and np.abs(obs.astype(float) - g["obs0"]).max() <= G.OBS_TOL
        ep_r, ep_l = np.zeros(N), np.zeros(N, int)
        n_eps = 0
        for t in range(T):
            obs, rew, done, infos = venv.step(g["actions"][t])
            assert rew.dtype == np.float32 and done.dtype == bool and len(infos) == N
            assert np.array_equal(done, g["done"][t].astype(bool))
            assert np.abs(obs.astype(float) - g["obs"][t]).max() <= G.OBS_TOL
            assert np.abs(rew - g["reward"][t].astype(np.float32)).max() <= 1e-5
            ep_r += g["reward"][t]
            ep_l += 1
            for i in range(N):
                if done[i]:
                    inf = infos[i]
                    assert np.abs(inf["terminal_observation"].astype(float) - g["terminal_obs"][t, i]).max() <= G.OBS_TOL
                    assert inf["TimeLimit.truncated"] == bool(g["truncated"][t, i])
                    assert inf["episode"]["l"] == ep_l[i] and abs(inf["episode"]["r"] - ep_r[i]) < 1e-5
                    assert inf["bounds_violation"] == bool(g["viol"][t, i])
                    ep_r[i], ep_l[i] = 0.0, 0
                    n_eps += 1
                else:
                    assert infos[i] == {}
        st = venv.episode_stats()
        assert st["done"] == n_eps == int(g["done"].sum()) and st["steps"] == T * N


def test_vecenv_discrete_and_annealing_modes():
    from rl_rocket_b200.vec_env import RocketVecEnv
    g = G.load("traj_3dof_discrete.npz")
    T, N = g["done"].shape
    venv = RocketVecEnv(N, dof=3, max_episode_steps=3000, discrete_actions=True, math_mode="exact", **G.cfg3())
    venv.seed(int(g["seed"]))
    venv.reset()
    assert venv.action_space.n == 4
    for t in range(T):
        obs, rew, done, infos = venv.step(g["actions"][t])
        assert np.array_equal(done, g["done"][t].astype(bool)) and np.abs(rew - g["reward"][t].astype(np.float32)).max() <= 1e-4
    g = G.load("traj_6dof_annealing.npz")
    T, N = g["done"].shape
    venv = RocketVecEnv(N, dof=6, max_episode_steps=800, math_mode="exact", **G.cfg6())
    venv.set_reward_annealing(True)              # second training phase of main_6DOF.py:95-103
    venv.seed(int(g["seed"]))
    venv.reset()
    for t in range(T):
        obs, rew, done, infos = venv.step(g["actions"][t])
        assert np.array_equal(done, g["done"][t].astype(bool)) and np.abs(rew - g["reward"][t].astype(np.float32)).max() <= 1e-6


def test_config5_sb3_style_rollout_collection():
    """BASELINE config 5: the calling pattern of SB3's OnPolicyAlgorithm.collect_rollouts (policy forward on the GPU, clip to
    the Box, env.step, bootstrap on TimeLimit.truncated via terminal_observation, episode info buffer), on 6DOF with the
    vtarg-shaped reward.  Checked against the oracle stepping the same actions."""
    from oracle import c_oracle as O
    from rl_rocket_b200.vec_env import RocketVecEnv
    N, n_steps, tl = 256, 128, 40
    cfg = G.cfg6()
    venv = RocketVecEnv(N, dof=6, max_episode_steps=tl, math_mode="exact", **cfg)
    venv.seed(77)
    vec = O.VecOracle(O.make_params(6, max_episode_steps=tl, **cfg), N, seed=77)
    torch.manual_seed(0)
    policy = torch.nn.Sequential(torch.nn.Linear(14, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                                 torch.nn.Linear(64, 3)).cuda()
    value = torch.nn.Sequential(torch.nn.Linear(14, 64), torch.nn.Tanh(), torch.nn.Linear(64, 1)).cuda()
    obs = venv.reset()
    assert np.array_equal(obs, vec.reset())
    buf_obs = np.zeros((n_steps, N, 14), np.float32)
    buf_rew = np.zeros((n_steps, N), np.float32)
    ep_info, n_boot = [], 0
    for t in range(n_steps):
        with torch.no_grad():
            mean = policy(torch.as_tensor(obs).cuda())
            actions = (mean + 0.3 * torch.randn_like(mean)).cpu().numpy()
        clipped = np.clip(actions, venv.action_space.low, venv.action_space.high)
        new_obs, rewards, dones, infos = venv.step(clipped)
        out = vec.step(clipped.astype(np.float32))
        assert np.array_equal(dones, out["done"].astype(bool)) and np.abs(new_obs - out["obs"]).max() <= G.OBS_TOL
        assert np.abs(rewards - out["reward"].astype(np.float32)).max() <= 1e-5
        for i in np.nonzero(dones)[0]:
            ep_info.append(infos[i]["episode"])
            if infos[i].get("TimeLimit.truncated", False):
                with torch.no_grad():
                    rewards[i] += 0.99 * value(torch.as_tensor(infos[i]["terminal_observation"]).cuda()).item()
                n_boot += 1
            assert infos[i]["episode"]["l"] == int(out["fin_len"][i]) and abs(infos[i]["episode"]["r"] - out["fin_ret"][i]) < 1e-5
        buf_obs[t], buf_rew[t] = obs, rewards
        obs = new_obs
    assert len(ep_info) >= N * (n_steps // tl) - N and n_boot > 0
    assert np.isfinite(buf_obs).all() and np.isfinite(buf_rew).all()


def test_trajectory_trace_matches_oracle_history():
    """device-side trace of selected environments = what the reference keeps in SIM.states / actions / times and
    vtarg_history, here checked against the oracle's per-step states and vtarg for the traced columns, across auto-resets."""
    from oracle import c_oracle as O
    from rl_rocket_b200.vec_env import RocketVecEnv
    N, T, tl = 128, 90, 30
    cfg = G.cfg6()
    idx = [3, 64, 127]
    venv = RocketVecEnv(N, dof=6, max_episode_steps=tl, math_mode="exact", trace_indices=idx, trace_capacity=256, **cfg)
    venv.seed(11)
    p = O.make_params(6, max_episode_steps=tl, **cfg)
    vec = O.VecOracle(p, N, seed=11)
    venv.reset(); vec.reset()
    ic0 = vec.y[:, idx].T.copy()
    rng = np.random.RandomState(2)
    hist_y, hist_a = [], []
    for t in range(T):
        a = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
        venv.step_tensor(torch.as_tensor(a).cuda())
        out = vec.step(a)
        hist_y.append(out["y_pre_reset"][:, idx].T.copy()); hist_a.append(a[idx])
    h = venv.trace.to_host()
    assert h["states"].shape == (T + 1, 3, 14)
    assert np.array_equal(h["states"][0], ic0)
    norm = venv.spec.state_normalizer
    assert (np.abs(h["states"][1:] - np.array(hist_y)) / norm).max() <= G.STATE_TOL
    assert np.array_equal(h["actions"][1:], np.array(hist_a))
    # the last (running) episode of every traced env starts at an initial condition (k == 0) and has consecutive times
    for j in range(3):
        s, a, vt, times = venv.trace.last_episode(j)
        assert times[0] == 0.0 and np.allclose(np.diff(times), cfg["timestep"])
        assert len(s) == len(a) == len(vt) == len(times) and len(s) <= tl + 1
        assert np.array_equal(s[0].astype(np.float32).astype(np.float64), s[0])          # an IC: f32-valued state
        assert abs(venv.trace.used_mass(j) - (s[0, -1] - s[-1, -1])) == 0
        assert list(venv.trace.vtarg_to_dataframe(j).columns) == ["v_x", "v_y", "v_z"]
    # vtarg of the traced rows against the scalar oracle's _compute_vtarg on the same f32 states
    env = O.ScalarEnv(p)
    s_last, a_last, vt_last, _ = venv.trace.last_episode(0)
    if len(s_last) > 1:
        env.reset(s_last[0].astype(np.float32))
        for r in range(1, len(s_last)):
            _, _, _, info = env.step(a_last[r])
            assert np.abs(info["vtarg"] - vt_last[r, :3]).max() <= 1e-9 * max(1.0, np.abs(info["vtarg"]).max())
            assert abs(info["t_go"] - vt_last[r, 3]) <= 1e-9 * max(1.0, abs(info["t_go"]))


def test_facade_vtarg_history_matches_oracle():
    """Rocket6DOF.vtarg_history / vtarg_to_dataframe (rocket_env.py:1012, :1026-1029) against the oracle on the KAT-2 replay."""
    from oracle import c_oracle as O
    from rl_rocket_b200.envs import Rocket6DOF
    cfg = G.cfg6()
    cfg["IC"] = [500, 100, 100, -50, 0, 0, 1, 0, 0, 0, 0, 0, 0, 45e3]
    cfg["ICRange"] = [0] * 14
    env = Rocket6DOF(**cfg)
    ref = O.ScalarEnv(O.make_params(6, **cfg))
    env.reset(); ref.reset(np.float32(cfg["IC"]))
    for t in range(40):
        a = np.float32([0.3, -0.2, 0.5])
        _, _, done, _ = env.step(a)
        _, _, rdone, info = ref.step(a)
        assert done == rdone
        assert np.abs(env.vtarg_history[-1] - info["vtarg"]).max() <= 1e-9 * max(1.0, np.abs(info["vtarg"]).max())
        if done:
            break
    df = env.vtarg_to_dataframe()
    assert df.shape == (len(env.vtarg_history), 3)


@pytest.mark.parametrize("N", [4096, 1000])          # 1000: ragged -- the last warp takes the per-lane store path
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_step_host_zero_copy_equals_copy_path(precision, N):
    """RocketVecEnv.step_host: the mapped-host-memory path (kernel reads actions from / writes obs|reward|done to pinned host memory)
    returns exactly what the H2D + packed D2H path returns."""
    from rl_rocket_b200.vec_env import RocketVecEnv
    T = 40
    cfg = G.cfg6()
    v1 = RocketVecEnv(N, dof=6, max_episode_steps=30, precision=precision, **cfg)
    v2 = RocketVecEnv(N, dof=6, max_episode_steps=30, precision=precision, **cfg)
    v1.reset(); v2.reset()
    rng = np.random.RandomState(0)
    for t in range(T):
        a = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
        o1, r1, d1 = v1.step_host(torch.as_tensor(a).pin_memory(), zero_copy=True)
        o2, r2, d2 = v2.step_host(a, zero_copy=False)
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)
    assert int(d1.sum()) >= 0 and v1.episode_stats() == v2.episode_stats()
    # numpy (unpinned) actions go through the pinned staging buffer
    o1, r1, d1 = v1.step_host(a)
    o2, r2, d2 = v2.step_host(a, zero_copy=False)
    assert torch.equal(o1, o2)


def test_step_is_cuda_graph_capturable():
    """rlr_step only enqueues work on the caller's stream (no allocation, no synchronisation), so a sequence of steps can be captured
    into a CUDA graph and replayed; the replay reproduces the eager trajectory bit for bit."""
    from rl_rocket_b200.config import EnvSpec
    from rl_rocket_b200.engine import RocketBatch
    N, T = 4096, 8
    spec = EnvSpec(6, max_episode_steps=800, **G.cfg6())
    a = torch.rand((T, N, 3), device="cuda") * 2 - 1
    eager = RocketBatch(spec, N, seed=9); eager.reset()
    graphed = RocketBatch(spec, N, seed=9); graphed.reset()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):                       # warm-up on the capture stream (launch geometry / attribute queries happen here)
        graphed.step_tensor(a[0])
    side.synchronize()
    graphed.reseed(9)                                   # a second reset() alone would move on to the next episode's initial conditions
    graphed.reset()
    obs_hist = torch.zeros((T, N, 14), device="cuda")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for t in range(T):
            o, r, d = graphed.step_tensor(a[t])
            obs_hist[t].copy_(o)
    g.replay()
    torch.cuda.synchronize()
    for t in range(T):
        o, r, d = eager.step_tensor(a[t])
        assert torch.equal(o, obs_hist[t])
    assert torch.equal(eager.y, graphed.y) and torch.equal(eager.k, graphed.k)
    g.replay()                                           # a second replay continues the trajectories
    torch.cuda.synchronize()
    for t in range(T):
        eager.step_tensor(a[t])
    assert torch.equal(eager.y, graphed.y)


@pytest.mark.parametrize("depth", [2, 3])
def test_step_host_async_pipeline_equals_synchronous_path(depth):
    """RocketVecEnv.step_host_async (rlr_pipe_*: H2D / kernel / D2H on three streams, `depth` slots) returns, step for step, what the
    synchronous step_host returns -- including when the host runs `depth` submits ahead of the results it collects."""
    from rl_rocket_b200.vec_env import RocketVecEnv
    N, T = 5000, 60
    cfg = G.cfg6()
    v1 = RocketVecEnv(N, dof=6, max_episode_steps=25, **cfg)
    v2 = RocketVecEnv(N, dof=6, max_episode_steps=25, **cfg)
    v1.reset(); v2.reset()
    rng = np.random.RandomState(1)
    acts = [torch.as_tensor(rng.uniform(-1, 1, (N, 3)).astype(np.float32)).pin_memory() for _ in range(T)]
    ref = []
    for t in range(T):
        o, r, d = v2.step_host(acts[t], zero_copy=False)
        ref.append((o.clone(), r.clone(), d.clone()))
    pending = []
    got = []
    for t in range(T):
        pending.append(v1.step_host_async(acts[t], depth=depth))
        if len(pending) == depth:                       # collect the oldest before its slot is resubmitted
            got.append(tuple(x.clone() for x in pending.pop(0).wait()))
    while pending:
        got.append(tuple(x.clone() for x in pending.pop(0).wait()))
    assert len(got) == T
    for t in range(T):
        assert torch.equal(got[t][0], ref[t][0]) and torch.equal(got[t][1], ref[t][1]) and torch.equal(got[t][2], ref[t][2]), t
    assert v1.episode_stats() == v2.episode_stats() and v1.episode_stats()["done"] > 0
    with pytest.raises(ValueError):
        v1.step_host_async(torch.zeros((N, 3)))         # not pinned


def test_vecenv_reset_draws_fresh_initial_conditions_and_seed_reproduces():
    """Every reset() samples new initial conditions (init_space.sample(), rocket_env.py:142, :672); seed(s) + reset() reproduces."""
    from rl_rocket_b200.vec_env import RocketVecEnv
    v = RocketVecEnv(256, dof=6, max_episode_steps=800, **G.cfg6())
    v.seed(7)
    o1 = v.reset()
    o2 = v.reset()
    assert not np.array_equal(o1, o2)
    v.step(np.zeros((256, 3), np.float32))
    o3 = v.reset()
    assert not np.array_equal(o3, o1) and not np.array_equal(o3, o2)
    v.seed(7)
    assert np.array_equal(v.reset(), o1) and np.array_equal(v.reset(), o2)
    # infos of non-terminal environments: a fresh dict per step (writing into one must not leak into later steps)
    _, _, done, infos = v.step(np.zeros((256, 3), np.float32))
    infos[0]["mine"] = 1
    _, _, done, infos2 = v.step(np.zeros((256, 3), np.float32))
    assert "mine" not in infos2[0]


def test_vecenv_gaudet_observation_matches_golden():
    """RocketVecEnv(gaudet_obs=True) (RLR_OBS_GAUDET) against the corrected GaudetStateObs wrapper over the live reference."""
    from rl_rocket_b200.vec_env import RocketVecEnv
    for name in ("traj_3dof_gaudet.npz", "traj_3dof_default_gaudet.npz"):
        dof, cfgf, tl, mode, akind = G.TRAJ[name]
        g = G.load(name)
        T, N = g["done"].shape
        v = RocketVecEnv(N, dof=3, max_episode_steps=tl, math_mode="exact", gaudet_obs=True, **cfgf())
        v.seed(int(g["seed"]))
        assert v.observation_space.shape == (4,)
        obs = v.reset()
        assert obs.shape == (N, 4) and np.abs(obs.astype(float) - g["obs0"]).max() <= 1e-6 * np.maximum(1, np.abs(g["obs0"])).max()
        for t in range(T):
            obs, rew, done, infos = v.step(g["actions"][t])
            assert np.array_equal(done, g["done"][t].astype(bool))
            tol = 1e-6 * np.maximum(1.0, np.abs(g["obs"][t]))
            assert (np.abs(obs.astype(float) - g["obs"][t]) <= tol).all(), (name, t)
            for i in np.nonzero(done)[0]:
                assert (np.abs(infos[i]["terminal_observation"].astype(float) - g["terminal_obs"][t][i]) <= 1e-6 * np.maximum(1, np.abs(g["terminal_obs"][t][i]))).all()


def test_discrete_actions_with_reward_annealing_matches_golden():
    """DiscreteActions3DOF(RewardAnnealing(Rocket)) (main.py:139-146): the annealed thrust penalty sees the table row."""
    from rl_rocket_b200.vec_env import RocketVecEnv
    g = G.load("traj_3dof_discrete_annealing.npz")
    T, N = g["done"].shape
    v = RocketVecEnv(N, dof=3, max_episode_steps=3000, discrete_actions=True, reward_annealing=True, math_mode="exact", **G.cfg3())
    v.seed(int(g["seed"]))
    v.reset()
    for t in range(T):
        obs, rew, done, infos = v.step(g["actions"][t])
        assert np.array_equal(done, g["done"][t].astype(bool))
        assert np.abs(rew.astype(float) - g["reward"][t]).max() <= 1e-6 * max(1.0, np.abs(g["reward"][t]).max())
    assert len({float(x) for x in np.round(g["reward"], 9).ravel()}) >= 2      # throttle -1 rows (0) and throttle +1 rows (-2 xi) both occur


def test_3dof_vtarg_output_matches_oracle():
    """SURVEY A14 on the GPU: the step path's vtarg output of the 3DOF kernel against the C oracle's or_env_step on the same pre-step
    states (KAT-4, tests/test_oracle_golden.py, pins the oracle's _compute_vtarg to the reference's test_vtarg.py grid)."""
    import ctypes as C
    from oracle import c_oracle as O
    from rl_rocket_b200.config import EnvSpec
    from rl_rocket_b200.engine import RocketBatch
    N, T = 512, 40
    lib = O.lib()
    for cfg, tl in ((G.cfg3(), 3000), (G.cfg3_default(), 2000)):
        for mm in ("exact", "fast"):
            spec = EnvSpec(3, max_episode_steps=tl, **cfg)
            p = O.make_params(3, max_episode_steps=tl, **cfg)
            env = RocketBatch(spec, N, seed=3, math_mode=mm, debug_outputs=True)
            env.reset()
            rng = np.random.RandomState(5)
            for t in range(T):
                a = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
                y0, k0, v0 = env.y.cpu().numpy().T.copy(), env.k.cpu().numpy(), env.v0.cpu().numpy()
                env.step_tensor(torch.as_tensor(a).cuda())
                vt = env.vtarg.cpu().numpy()
                ref = np.zeros((N, 4))
                obs, r, fl = np.zeros(14, np.float32), C.c_double(0), np.zeros(6, np.int32)
                for i in range(N):
                    yy = y0[i].copy()
                    lib.or_env_step(C.byref(p), O._p(yy, C.c_double), C.c_int(int(k0[i])), C.c_float(float(v0[i])), O._p(a[i], C.c_float), C.c_int(0),
                                    O._p(obs, C.c_float), C.byref(r), None, O._p(fl, C.c_int), O._p(ref[i], C.c_double))
                scale = np.maximum(1.0, np.abs(ref.T))
                assert (np.abs(vt - ref.T) <= 1e-9 * scale).all(), (mm, t, np.abs(vt - ref.T).max())


def _exchange_worker(rank, world, port, q):
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch, torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from rl_rocket_b200.distributed import StatsExchange, aggregate
    ex = StatsExchange(torch.device("cuda", rank))
    rows = []
    for it in range(50):                         # many epochs: the parity double-buffering and the flags must hold up
        local = torch.tensor([100 * rank + it, it, rank + 1, rank, 2 * rank, 0, (rank + 1) * 70, -(rank + 1) * it], dtype=torch.int64, device=f"cuda:{rank}")
        rows.append(ex.all_gather(local).cpu().tolist())
    ex.check()
    q.put((rank, ex.peer_memory, ex.error, rows))
    dist.barrier()
    ex.close()
    dist.destroy_process_group()


def test_stats_exchange_single_rank_and_two_gpus():
    """rlr_exchange_allgather: world 1 on this GPU; world 2 over NVLink peer memory when the box has two GPUs."""
    from rl_rocket_b200.distributed import StatsExchange
    ex = StatsExchange(torch.device("cuda:0"))
    assert ex.peer_memory, ex.error
    for it in range(5):
        local = torch.arange(8, dtype=torch.int64, device="cuda") * (it + 1)
        assert ex.all_gather(local).cpu().tolist() == [[(it + 1) * j for j in range(8)]]
    ex.check(); ex.close()
    if torch.cuda.device_count() < 2:
        return
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, 29547, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1], (res[0][2], res[1][2])          # peer memory was mapped on both ranks
    for it in range(50):
        want = [[100 * r + it, it, r + 1, r, 2 * r, 0, (r + 1) * 70, -(r + 1) * it] for r in range(2)]
        assert res[0][3][it] == want and res[1][3][it] == want, it
