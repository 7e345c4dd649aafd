"""Generate tests/golden/*.npz from the LIVE reference (dev container only).

    python -m oracle.make_golden

Every array written here is produced by the reference's own unmodified sources
(/root/reference, loaded by oracle/live_reference.py).  Initial conditions are injected
(`init_space.sample` is replaced by a feeder of pre-drawn f32 ICs) because gym's RNG is not
available; the ICs themselves come from the oracle's Philox sampler so that the CUDA path
can regenerate the same ones on the device.
"""
import ast
import os
import sys

import numpy as np

from . import live_reference as L
from . import c_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class LiveVec:
    """DummyVecEnv(TimeLimit(env)) over live reference envs with injected ICs and nfev capture."""

    def __init__(self, make_env, N, params, seed, max_episode_steps, obs_fn=None, act_fn=None, ics=None):
        ns = L.load()
        import my_environment.utils.simulator as simmod
        self.simmod = simmod
        self._orig = simmod.solve_ivp
        self.envs = [make_env() for _ in range(N)]
        self.N, self.p, self.seed, self.T = N, params, seed, max_episode_steps
        self.k = np.zeros(N, np.int32)
        self.episode = np.zeros(N, np.int32)
        self.obs_fn = obs_fn or (lambda o: o)
        self.act_fn = act_fn or (lambda a: a)
        self.last_nfev = 0
        self.ics = None if ics is None else np.asarray(ics, dtype=np.float32)      # fixed per-env ICs (every episode restarts from them)

    def _solve(self, *a, **k):
        sol = self._orig(*a, **k)
        self.last_nfev = sol.nfev
        return sol

    def _reset(self, i):
        ic = O.sample_ic(self.p, self.seed, i, int(self.episode[i])) if self.ics is None else self.ics[i].copy()
        e = self.envs[i].unwrapped
        e.init_space.sample = lambda ic=ic: ic.copy()
        self.k[i] = 0
        return ic, self.obs_fn(np.asarray(self.envs[i].reset())), e

    def reset(self):
        out = [self._reset(i) for i in range(self.N)]
        return np.stack([o[0] for o in out]), np.stack([o[1] for o in out]).astype(np.float32)

    def step(self, actions):
        self.simmod.solve_ivp = self._solve
        try:
            n = O.ndim(self.p)
            res = dict(obs=[], reward=[], done=[], truncated=[], y=[], nfev=[], ic=[], terminal_obs=[], viol=[], rew_goal=[])
            for i, env in enumerate(self.envs):
                o, r, d, info = env.step(self.act_fn(actions[i]))
                e = env.unwrapped
                ystate = np.array(e.SIM.state, dtype=np.float64)
                self.k[i] += 1
                tr = False
                if self.T and self.k[i] >= self.T:
                    tr = not d
                    d = True
                o = self.obs_fn(np.asarray(o))
                res["terminal_obs"].append(o.astype(np.float32) if d else np.zeros_like(o, dtype=np.float32))
                ic = np.zeros(n, np.float32)
                if d:
                    self.episode[i] += 1
                    ic, o, _ = self._reset(i)
                res["obs"].append(o.astype(np.float32)); res["reward"].append(float(r)); res["done"].append(d)
                res["truncated"].append(tr); res["y"].append(ystate); res["nfev"].append(self.last_nfev)
                res["ic"].append(ic); res["viol"].append(bool(info["bounds_violation"]))
                res["rew_goal"].append(float(info["rewards_dict"]["rew_goal"]))
            return {k: np.asarray(v) for k, v in res.items()}
        finally:
            self.simmod.solve_ivp = self._orig


def rollout(vec, actions):
    ic0, obs0 = vec.reset()
    steps = [vec.step(a) for a in actions]
    out = {k: np.stack([s[k] for s in steps]) for k in steps[0]}
    out["ic0"], out["obs0"], out["actions"] = ic0, obs0, actions
    return out


def save(name, **arrs):
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name)
    np.savez_compressed(path, **arrs)
    print("wrote", path, os.path.getsize(path), "bytes")


def kat_6dof(ns):
    # KAT-1: test_6DOF_simulator.py
    s = ns.Simulator6DOF([100, 100, 100, 0, 3, 4, 1, 1, 0, 0, 0, 0, 0, 50e3])
    st, status, t = s.step([0, 0, 1])
    out = dict(kat1_state=np.asarray(st, dtype=np.float64), kat1_status=status, kat1_t=t)
    # KAT-2: env_config with zero ICRange (test_6DOF_reward.py:10-11), fixed actions (test_6DOF_environment.py:19-20)
    cfg = L.env_config_6dof()
    cfg["ICRange"] = [0] * 14
    for tag, act in (("null", np.float32([0, 0, -1])), ("nonnull", np.float32([1, 1, -0.5])), ("mixed", np.float32([0.3, -0.2, 0.5]))):
        env = ns.Rocket6DOF(**cfg)
        obs = [env.reset()]
        ys, rs, ds, vs, terms, vt = [], [], [], [], [], []
        done = False
        while not done:
            o, r, done, info = env.step(act)
            obs.append(o); ys.append(np.array(env.state)); rs.append(r); ds.append(done); vs.append(info["bounds_violation"])
            terms.append([float(v) for v in info["rewards_dict"].values()]); vt.append(np.array(env.vtarg_history[-1]))
        out.update({f"kat2_{tag}_action": act, f"kat2_{tag}_obs": np.array(obs), f"kat2_{tag}_y": np.array(ys),
                    f"kat2_{tag}_reward": np.array(rs), f"kat2_{tag}_done": np.array(ds), f"kat2_{tag}_viol": np.array(vs),
                    f"kat2_{tag}_terms": np.array(terms), f"kat2_{tag}_vtarg": np.array(vt)})
    save("kat_6dof.npz", **out)


def kat_3dof(ns):
    cfg = L.env_config_3dof()
    cfg["ICRange"] = [0] * 7
    out = {}
    for tag, act in (("null", np.float32([0, -1])), ("mixed", np.float32([0.5, 0.2]))):
        env = ns.Rocket(**cfg)
        obs = [env.reset()]
        ys, rs, ds, terms = [], [], [], []
        done = False
        while not done:
            o, r, done, info = env.step(act)
            obs.append(o); ys.append(np.array(env.y)); rs.append(r); ds.append(done)
            terms.append([float(v) for v in info["rewards_dict"].values()])
        out.update({f"kat3_{tag}_action": act, f"kat3_{tag}_obs": np.array(obs), f"kat3_{tag}_y": np.array(ys),
                    f"kat3_{tag}_reward": np.array(rs), f"kat3_{tag}_done": np.array(ds), f"kat3_{tag}_terms": np.array(terms)})
    save("kat_3dof.npz", **out)


def kat_vtarg():
    # KAT-4: the stand-alone _compute_vtarg of test_vtarg.py:47-70, executed from its own source text
    src = open(os.path.join(L.REFERENCE_ROOT, "test_vtarg.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "_compute_vtarg"][0]
    g = {"np": np}
    exec(compile(ast.Module([fn], []), "test_vtarg.py", "exec"), g)
    ic = [-1600, 2000, np.pi * 3 / 4, 180, -90, 0, 50e3]
    pos = np.linspace(start=ic[0:2], stop=[0, 0], num=200)
    vel = np.linspace(start=ic[3:5], stop=[0, 0], num=200)
    with np.errstate(all="ignore"):
        res = [g["_compute_vtarg"](r, v, IC=ic) for r, v in zip(pos, vel)]
    save("kat_vtarg.npz", r=pos, v=vel, vtarg=np.array([x[0] for x in res]), tgo=np.array([x[1] for x in res]))


def traj_6dof(ns, N=16, T=320, seed=1234):
    cfg = L.env_config_6dof()
    p = O.make_params(6, max_episode_steps=800, **cfg)
    vec = LiveVec(lambda: ns.Rocket6DOF(**cfg), N, p, seed, 800)
    actions = np.random.RandomState(7).uniform(-1, 1, (T, N, 3)).astype(np.float32)
    save("traj_6dof.npz", seed=seed, **rollout(vec, actions))
    # short TimeLimit to exercise truncation, constant hover-ish thrust so that episodes survive
    p2 = O.make_params(6, max_episode_steps=25, **cfg)
    vec = LiveVec(lambda: ns.Rocket6DOF(**cfg), 4, p2, seed + 1, 25)
    a2 = np.random.RandomState(8).uniform(-0.2, 0.2, (60, 4, 3)).astype(np.float32)
    save("traj_6dof_timelimit.npz", seed=seed + 1, **rollout(vec, a2))
    # RewardAnnealing wrapper (wrappers.py:68-86)
    if ns.wrappers is not None:
        vec = LiveVec(lambda: ns.wrappers.RewardAnnealing(ns.Rocket6DOF(**cfg)), 4, p, seed + 2, 800)
        a3 = np.random.RandomState(9).uniform(-1, 1, (150, 4, 3)).astype(np.float32)
        save("traj_6dof_annealing.npz", seed=seed + 2, **rollout(vec, a3))


def traj_3dof(ns, N=16, T=400, seed=4321):
    cfg = L.env_config_3dof()
    p = O.make_params(3, max_episode_steps=3000, **cfg)
    vec = LiveVec(lambda: ns.Rocket(**cfg), N, p, seed, 3000)
    actions = np.random.RandomState(17).uniform(-1, 1, (T, N, 2)).astype(np.float32)
    save("traj_3dof.npz", seed=seed, **rollout(vec, actions))
    # default-ctor 3DOF (dt 0.1, low altitude: episodes end on the ground -> event path + landing checks)
    dcfg = dict(IC=[100, 500, np.pi / 2, -10, -50, 0, 50e3], ICRange=[10, 50, 0.1, 1, 10, 0.1, 1e3], timestep=0.1, seed=42,
                reward_coeff={"alfa": -0.01, "beta": -1e-8, "eta": 2, "gamma": -10, "delta": -5, "kappa": 10, "waypoint": 50, "landing_radius": 30})
    pd_ = O.make_params(3, max_episode_steps=2000, **dcfg)
    vec = LiveVec(lambda: ns.Rocket(**dcfg), 8, pd_, seed + 1, 2000)
    a = np.random.RandomState(18).uniform(-1, 1, (300, 8, 2)).astype(np.float32)
    save("traj_3dof_default.npz", seed=seed + 1, **rollout(vec, a))
    if ns.wrappers is not None:
        # DiscreteActions3DOF (wrappers.py:24-35): integer actions through the live wrapper
        vec = LiveVec(lambda: ns.wrappers.DiscreteActions3DOF(ns.Rocket(**cfg)), 4, p, seed + 2, 3000, act_fn=lambda a: int(a))
        ai = np.random.RandomState(19).randint(0, 4, (480, 4)).astype(np.int64)
        save("traj_3dof_discrete.npz", seed=seed + 2, **rollout(vec, ai))
        vec = LiveVec(lambda: ns.wrappers.RewardAnnealing(ns.Rocket(**cfg)), 4, p, seed + 3, 3000)
        a3 = np.random.RandomState(20).uniform(-1, 1, (480, 4, 2)).astype(np.float32)
        save("traj_3dof_annealing.npz", seed=seed + 3, **rollout(vec, a3))
    # config 1: 1DOF toy problem (README.md:4-8; defined in SURVEY.md 8(d)): 3DOF env, gimbal 0, obs (z, vz)
    c1 = dict(IC=[0, 2000, np.pi / 2, 0, -90, 0, 50e3], ICRange=[0, 50, 0, 0, 0, 0, 1e3], timestep=0.05, seed=42,
              reward_coeff=cfg["reward_coeff"])
    p1 = O.make_params(1, max_episode_steps=3000, **c1)
    vec = LiveVec(lambda: ns.Rocket(**c1), 1, p1, 42, 3000, obs_fn=lambda o: np.asarray(o)[[1, 4]],
                  act_fn=lambda a: np.float32([0.0, a[0]]))
    a1 = np.random.RandomState(0).uniform(-1, 1, (500, 1, 1)).astype(np.float32)
    save("traj_1dof.npz", seed=42, **rollout(vec, a1))


# ---------------------------------------------------------------------------------------------- round 2: the unpinned branches
def _bisect_f32(pred, lo, hi):
    """Adjacent float32 values (a, b), a < b, with pred(a) != pred(b), by bisection between lo and hi (pred differs there)."""
    lo, hi = np.float32(lo), np.float32(hi)
    plo = pred(lo)
    assert plo != pred(hi)
    while np.nextafter(lo, np.float32(np.inf)) < hi:
        mid = np.float32((np.float64(lo) + np.float64(hi)) / 2)
        if pred(mid) == plo:
            lo = mid
        else:
            hi = mid
    return lo, hi


def _oracle_episode(p, ic, act):
    """(landed, steps) of one episode under a constant normalised action, on the C oracle (used to place the threshold cases)."""
    e = O.ScalarEnv(p)
    e.reset(np.float32(ic))
    for n in range(1, 2000):
        _, _, done, info = e.step(np.float32(act))
        if done:
            return info["landed"], n
    raise RuntimeError("episode did not end")


LAND6_LIMITS = {"waypoint": 50, "landing_radius": 30, "maximum_velocity": 10, "landing_attitude_limit": [0.1, 0.1, 0.1],
                "omega_lim": [0.2, 0.2, 0.2]}


def landing_cases_6dof():
    """Crafted touchdown episodes of the 6DOF env (rocket_env.py:1040-1061): each row = (IC f32[14], constant action).  Threshold
    cases are adjacent float32 initial conditions found by bisection on the C oracle; the live reference then decides."""
    cfg = L.env_config_6dof()
    hover = [0.0, 0.0, -0.1]
    base = [3, 0, 0, -1.5, 0, 0, 1, 0, 0, 0, 0, 0, 0, 45e3]

    def ic(**kw):
        v = list(base)
        for k, x in kw.items():
            v[int(k[1:])] = x
        return v
    cases = [
        ("soft", ic(), [0.01, -0.02, -0.1]),                                  # landed
        ("slow_descent", ic(i0=30, i1=5, i2=5, i3=-9), hover),                # landed, |v| = 8.8
        ("too_fast", ic(i0=30, i1=5, i2=5, i3=-12), hover),                   # |v| = 11.9 -> crash
        ("radius_in", ic(i1=29.9), hover), ("radius_out", ic(i1=30.1), hover),
        ("radius_3d_out", ic(i1=22, i2=22), hover),                           # |y|, |z| < 30 but the 3-D norm is 31.1 (rocket_env.py:1042)
        ("omega_any", ic(i0=0.5, i3=-2, i10=0.05, i11=0.3, i12=0.3), hover),  # two rates above 0.2, one below: np.any -> landed
        ("omega_none", ic(i0=0.5, i3=-2, i10=0.3, i11=0.3, i12=0.3), hover),  # all three above -> not landed
        ("free_fall", ic(i0=2, i3=0), [0, 0, -1]),                            # engine off
        ("zero_mass", [500, 100, 100, -50, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0.0], [0.2, 0.1, 0.3]),   # infinite dv/dt: solve_ivp status -1 (rocket_env.py:702)
    ]
    # speed threshold (maximum_velocity = 10): adjacent float32 descent rates on either side
    p = O.make_params(6, max_episode_steps=800, **dict(cfg, ICRange=[0] * 14))
    a, b = _bisect_f32(lambda vx: _oracle_episode(p, ic(i0=30, i3=float(vx)), hover)[0], -12.0, -9.0)
    cases += [("speed_just_out", ic(i0=30, i3=float(a)), hover), ("speed_just_in", ic(i0=30, i3=float(b)), hover)]
    a, b = _bisect_f32(lambda yy: _oracle_episode(p, ic(i1=float(yy)), hover)[0], 29.9, 30.1)
    cases += [("radius_just_in", ic(i1=float(a)), hover), ("radius_just_out", ic(i1=float(b)), hover)]
    return cases


def landing_cases_6dof_attitude():
    """Same with landing_attitude_limit = [0.1, 0.1, 0.1] rad (the env_config limits [10 deg, 10 deg, 2 pi] can never fail because
    |angle| <= pi): tilted touchdowns that separate np.any from np.all (rocket_env.py:1051-1054)."""
    hover = [0.0, 0.0, -0.1]

    def tilt(ax, ang):
        q = [np.cos(ang / 2), 0, 0, 0]
        q[1 + ax] = np.sin(ang / 2)
        return [0.5, 0, 0, -2] + [0, 0] + q + [0, 0, 0, 45e3]

    def tilt2(a1, a2, a3):
        from scipy.spatial.transform import Rotation
        x, y, z, w = Rotation.from_euler("zyx", [a1, a2, a3]).as_quat()
        return [0.5, 0, 0, -2, 0, 0, w, x, y, z, 0, 0, 0, 45e3]
    return [("upright", tilt(0, 0.0), hover), ("one_axis", tilt(1, 0.3), hover), ("two_axes", tilt2(0.3, 0.3, 0.0), hover),
            ("all_axes", tilt2(0.3, 0.3, 0.3), hover), ("all_axes_small", tilt2(0.05, 0.05, 0.05), hover)]


def landing_cases_3dof():
    """Crafted 3DOF touchdowns (rocket_env.py:449-476; `all` of the five conditions): IC = [x, z, theta, vx, vz, omega, m]."""
    cfg = L.env_config_3dof()
    hover = [0.0, 0.0]
    base = [0, 3, np.pi / 2, 0, -1.5, 0, 50e3]

    def ic(**kw):
        v = list(base)
        for k, x in kw.items():
            v[int(k[1:])] = x
        return v
    cases = [("soft", ic(), hover), ("too_fast", ic(i1=20, i4=-20), hover), ("radius_in", ic(i0=29.9), hover), ("radius_out", ic(i0=30.1), hover),
             ("tilted", ic(i2=np.pi / 2 + 0.25), hover), ("tilted_ok", ic(i2=np.pi / 2 + 0.15), hover), ("spinning", ic(i1=0.5, i4=-2, i5=0.25), hover),
             ("spinning_ok", ic(i1=0.5, i4=-2, i5=0.15), hover), ("free_fall", ic(i1=2, i4=0), [0, -1]),
             ("zero_mass", [0, 500, np.pi / 2, 0, -50, 0, 0.0], [0.3, 0.2])]
    p = O.make_params(3, max_episode_steps=3000, **dict(cfg, ICRange=[0] * 7))
    a, b = _bisect_f32(lambda vz: _oracle_episode(p, ic(i1=40, i4=float(vz)), hover)[0], -20.0, -10.0)
    cases += [("speed_just_out", ic(i1=40, i4=float(a)), hover), ("speed_just_in", ic(i1=40, i4=float(b)), hover)]
    return cases


def traj_cases(ns, name, dof, cfg, cases, T, tl):
    """One env per crafted case, constant action, ICs injected (every episode restarts from the same IC), T steps."""
    n = 14 if dof == 6 else 7
    cfg0 = dict(cfg, ICRange=[0] * n)
    p = O.make_params(dof, max_episode_steps=tl, **cfg0)
    ics = np.float32([c[1] for c in cases])
    cls = ns.Rocket6DOF if dof == 6 else ns.Rocket
    vec = LiveVec(lambda: cls(**cfg0), len(cases), p, 0, tl, ics=ics)
    actions = np.repeat(np.float32([c[2] for c in cases])[None], T, axis=0)
    with np.errstate(all="ignore"):
        out = rollout(vec, actions)
    ends = [(c[0], int(out["done"][:, i].argmax()) + 1) for i, c in enumerate(cases)]
    print(name, "first episode ends:", ends)
    save(name, seed=0, ics=ics, case_names=np.array([c[0] for c in cases]), **out)
    return out


def traj_landing(ns):
    cfg6 = L.env_config_6dof()
    traj_cases(ns, "traj_6dof_landing.npz", 6, cfg6, landing_cases_6dof(), 40, 800)
    cfg6a = dict(cfg6, landing_params=LAND6_LIMITS)
    traj_cases(ns, "traj_6dof_landing_attitude.npz", 6, cfg6a, landing_cases_6dof_attitude(), 8, 800)
    traj_cases(ns, "traj_3dof_landing.npz", 3, L.env_config_3dof(), landing_cases_3dof(), 60, 3000)
    # random low-altitude episodes: Philox ICs in a box that straddles the speed / radius / spin-rate limits
    cfgr = dict(cfg6, IC=[4, 0, 0, -6, 0, 0, 1, 0, 0, 0, 0, 0, 0, 45e3], ICRange=[4, 70, 70, 12, 4, 4, 0.2, 0.2, 0.2, 0.2, 0.8, 0.8, 0.8, 1e3])
    p = O.make_params(6, max_episode_steps=800, **cfgr)
    vec = LiveVec(lambda: ns.Rocket6DOF(**cfgr), 32, p, 77, 800)
    rng = np.random.RandomState(21)
    a = np.stack([rng.uniform(-0.3, 0.3, (90, 32)), rng.uniform(-0.3, 0.3, (90, 32)), rng.uniform(-0.4, 0.1, (90, 32))], -1).astype(np.float32)
    save("traj_6dof_landing_random.npz", seed=77, **rollout(vec, a))


def traj_unclipped(ns, N=8, T=200, seed=555):
    """Actions outside the Box: the environments never clip (rocket_env.py:969-981), so gimbal angles beyond pi/4 reach numpy's
    float32 sin / cos range reduction (simulator.py:351-357)."""
    cfg = L.env_config_6dof()
    p = O.make_params(6, max_episode_steps=800, **cfg)
    vec = LiveVec(lambda: ns.Rocket6DOF(**cfg), N, p, seed, 800)
    rng = np.random.RandomState(31)
    a = np.stack([rng.uniform(-3, 3, (T, N)), rng.uniform(-3, 3, (T, N)), rng.uniform(-1, 1, (T, N))], -1).astype(np.float32)
    a[::7, :, 0] *= 40          # a few very large gimbal commands (several turns)
    save("traj_6dof_unclipped.npz", seed=seed, **rollout(vec, a))
    cfg3 = L.env_config_3dof()
    p3 = O.make_params(3, max_episode_steps=3000, **cfg3)
    vec = LiveVec(lambda: ns.Rocket(**cfg3), 4, p3, seed + 1, 3000)
    a3 = np.stack([rng.uniform(-3, 3, (300, 4)), rng.uniform(-1, 1, (300, 4))], -1).astype(np.float32)
    save("traj_3dof_unclipped.npz", seed=seed + 1, **rollout(vec, a3))


def traj_discrete_annealing(ns, seed=4400):
    """DiscreteActions3DOF(RewardAnnealing(Rocket)) -- the composition main.py:139-146 sketches: the annealed thrust penalty sees the
    table row, -xi * (row[1] + 1) in float64 (wrappers.py:34, :80)."""
    cfg = L.env_config_3dof()
    p = O.make_params(3, max_episode_steps=3000, reward_mode=1, **cfg)
    vec = LiveVec(lambda: ns.wrappers.DiscreteActions3DOF(ns.wrappers.RewardAnnealing(ns.Rocket(**cfg))), 4, p, seed, 3000,
                  act_fn=lambda a: int(a))
    ai = np.random.RandomState(23).randint(0, 4, (480, 4)).astype(np.int64)
    save("traj_3dof_discrete_annealing.npz", seed=seed, **rollout(vec, ai))


def gaudet_fixed(ns):
    """GaudetStateObs (wrappers.py:51-66) with its two defects repaired, over the LIVE reference env: the reference version calls
    `self.env.unwrapped.compute_vtarg` (no such method: it is `_compute_vtarg`, rocket_env.py:219) and feeds it the normalised
    observation although the method works on the physical state.  This subclass keeps everything else."""
    base = ns.wrappers.GaudetStateObs

    class GaudetStateObsFixed(base):
        def observation(self, observation):
            e = self.env.unwrapped
            s = np.asarray(e.y).astype(np.float32)                 # the state Rocket.step hands to _compute_vtarg (rocket_env.py:155)
            v_targ, t_go = e._compute_vtarg(s[0:2], s[3:5])
            return np.float32([s[3] - v_targ[0], s[4] - v_targ[1], t_go, s[1]])
    return GaudetStateObsFixed


def traj_gaudet(ns, seed=6100):
    cfg = L.env_config_3dof()
    p = O.make_params(3, max_episode_steps=3000, obs_mode=1, **cfg)
    W = gaudet_fixed(ns)
    vec = LiveVec(lambda: W(ns.Rocket(**cfg)), 4, p, seed, 3000)
    a = np.random.RandomState(41).uniform(-1, 1, (400, 4, 2)).astype(np.float32)
    save("traj_3dof_gaudet.npz", seed=seed, **rollout(vec, a))
    # low-altitude default-constructor env: touchdowns, so the waypoint branch of _compute_vtarg and terminal observations are covered
    dcfg = dict(IC=[100, 500, np.pi / 2, -10, -50, 0, 50e3], ICRange=[10, 50, 0.1, 1, 10, 0.1, 1e3], timestep=0.1, seed=42,
                reward_coeff={"alfa": -0.01, "beta": -1e-8, "eta": 2, "gamma": -10, "delta": -5, "kappa": 10, "waypoint": 50, "landing_radius": 30})
    pd_ = O.make_params(3, max_episode_steps=2000, obs_mode=1, **dcfg)
    vec = LiveVec(lambda: W(ns.Rocket(**dcfg)), 4, pd_, seed + 1, 2000)
    a = np.random.RandomState(42).uniform(-1, 1, (200, 4, 2)).astype(np.float32)
    save("traj_3dof_default_gaudet.npz", seed=seed + 1, **rollout(vec, a))


def aero_fixed(ns):
    """Simulator3DOF with the angle-of-attack line of its RHS enabled (EXTENSION N4; OrParams.aero_mode = 1).  The reference keeps
    `alfa = self._computeAoA(y)` commented out (simulator.py:103-104) and the call could not run as written: it passes the height
    instead of the state, and `_computeAoA` ends in `self._normalize`, which does not exist (simulator.py:132-148).  This subclass
    compiles the reference's OWN RHS source with that one line enabled and the state passed, and supplies the missing `_normalize` as
    the wrap to (-pi, pi] built on the class's `_wrapTo2Pi`; everything downstream (Cn, N, the N terms of ax / ay / dom) is the
    reference's code."""
    import inspect
    import textwrap
    import my_environment.utils.simulator as simmod
    src = textwrap.dedent(inspect.getsource(ns.Simulator3DOF.RHS))
    off, on = "alfa = 0\n", "alfa = self._computeAoA(state)\n"
    assert src.count(off) == 1 and "#alfa = self._computeAoA(y)" in src, "the reference's RHS is not the one this hook was written against"
    scope = dict(vars(simmod))
    exec(compile(src.replace(off, on), "<Simulator3DOF.RHS with the angle-of-attack line enabled>", "exec"), scope)

    class Simulator3DOFAero(ns.Simulator3DOF):
        RHS = scope["RHS"]

        def _normalize(self, angle):
            a = self._wrapTo2Pi(angle)
            return a - 2. * np.pi if a > np.pi else a
    return Simulator3DOFAero


def traj_3dof_aero(ns, seed=8100):
    """3DOF trajectories with the angle-of-attack normal force on: the live Rocket env with its simulator class swapped for aero_fixed."""
    import my_environment.envs.rocket_env as envmod
    orig = envmod.Simulator3DOF
    envmod.Simulator3DOF = aero_fixed(ns)
    try:
        cfg = L.env_config_3dof()
        p = O.make_params(3, max_episode_steps=3000, aero_mode=1, **cfg)
        vec = LiveVec(lambda: ns.Rocket(**cfg), 4, p, seed, 3000)
        a = np.random.RandomState(61).uniform(-1, 1, (400, 4, 2)).astype(np.float32)
        save("traj_3dof_aero.npz", seed=seed, **rollout(vec, a))
        # low-altitude default-constructor env: touchdowns (ground events) and resets under the aerodynamic torque
        dcfg = dict(IC=[100, 500, np.pi / 2, -10, -50, 0, 50e3], ICRange=[10, 50, 0.1, 1, 10, 0.1, 1e3], timestep=0.1, seed=42,
                    reward_coeff={"alfa": -0.01, "beta": -1e-8, "eta": 2, "gamma": -10, "delta": -5, "kappa": 10, "waypoint": 50, "landing_radius": 30})
        pd_ = O.make_params(3, max_episode_steps=2000, aero_mode=1, **dcfg)
        vec = LiveVec(lambda: ns.Rocket(**dcfg), 4, pd_, seed + 1, 2000)
        a = np.random.RandomState(62).uniform(-1, 1, (200, 4, 2)).astype(np.float32)
        save("traj_3dof_default_aero.npz", seed=seed + 1, **rollout(vec, a))
    finally:
        envmod.Simulator3DOF = orig


def traj_3dof_timelimit(ns, seed=7300):
    """3DOF TimeLimit truncation (gym.wrappers.TimeLimit over Rocket, main.py:65): 30-step limit, gentle actions, so every episode is cut."""
    cfg = L.env_config_3dof()
    p = O.make_params(3, max_episode_steps=30, **cfg)
    vec = LiveVec(lambda: ns.Rocket(**cfg), 4, p, seed, 30)
    a = np.random.RandomState(51).uniform(-0.3, 0.3, (100, 4, 2)).astype(np.float32)
    save("traj_3dof_timelimit.npz", seed=seed, **rollout(vec, a))


def main():
    if not L.available():
        sys.exit("live reference not present; golden files can only be generated in the dev container")
    ns = L.load()
    kat_6dof(ns)
    kat_3dof(ns)
    kat_vtarg()
    traj_6dof(ns)
    traj_3dof(ns)
    traj_landing(ns)
    traj_unclipped(ns)
    traj_discrete_annealing(ns)
    traj_gaudet(ns)
    traj_3dof_timelimit(ns)
    traj_3dof_aero(ns)


if __name__ == "__main__":
    main()
