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

    def __init__(self, make_env, N, params, seed, max_episode_steps, obs_fn=None, act_fn=None):
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

    def _solve(self, *a, **k):
        sol = self._orig(*a, **k)
        self.last_nfev = sol.nfev
        return sol

    def _reset(self, i):
        ic = O.sample_ic(self.p, self.seed, i, int(self.episode[i]))
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
            res = dict(obs=[], reward=[], done=[], truncated=[], y=[], nfev=[], ic=[], terminal_obs=[], viol=[])
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


def main():
    if not L.available():
        sys.exit("live reference not present; golden files can only be generated in the dev container")
    ns = L.load()
    kat_6dof(ns)
    kat_3dof(ns)
    kat_vtarg()
    traj_6dof(ns)
    traj_3dof(ns)


if __name__ == "__main__":
    main()
