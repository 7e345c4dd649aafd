"""Python/scipy port of the reference's environment step -- CPU BASELINE + second checker (TEST INFRASTRUCTURE ONLY).

/root/reference cannot travel to the GPU box, so the "reference's Python/scipy CPU path" that bench.py times next to
the GPU number is this port: the same library calls in the same places (one `scipy.integrate.solve_ivp` call with a
terminal ground event per env step, `scipy.spatial.transform.Rotation` for the attitude matrix and Euler angles,
`np.cross` / small `ndarray.dot` products in the right-hand side), so its cost per step is the reference's cost per
step.  It is validated against the golden vectors of the live reference in tests/test_scipy_port.py.

What it follows:
  sim6_rhs / sim6_advance   my_environment/utils/simulator.py:227-378
  sim3_rhs / sim3_advance   my_environment/utils/simulator.py:55-130, :150-163
  PortEnv6.step / .reset    my_environment/envs/rocket_env.py:665-719, :825-859, :963-1061
  PortEnv3.step / .reset    my_environment/envs/rocket_env.py:137-247, :395-476
  normalizers / bounds      oracle/params.py (rocket_env.py:69-100, :575-620)
"""
from math import fmod

import numpy as np
from scipy.integrate import solve_ivp
from scipy.spatial.transform import Rotation

from .params import derived_3dof, derived_6dof

G0 = 9.81
ISP = 360
J6 = np.diag([75350.25, 6037675.13, 6037675.13])
J6_INV = np.linalg.inv(J6)
ARM_T = [-15, 0, 0]
ARM_CP = [5, 0, 0]
TWO_PI = 2.0 * np.pi


# ------------------------------------------------------------------------------------------ 6DOF simulator
def _body_to_inertial(q):
    return Rotation.from_quat([q[1], q[2], q[3], q[0]]).as_matrix()


def _thrust_body(u):
    dy, dz, thrust = u[0], u[1], u[2]
    m = np.array([[np.cos(dy) * np.cos(dz), -np.sin(dy), -np.cos(dy) * np.sin(dz)],
                  [np.sin(dy) * np.cos(dz), np.cos(dy), -np.sin(dy) * np.sin(dz)],
                  [np.sin(dz), 0, np.cos(dz)]])
    return m @ [thrust, 0.0, 0.0]


def sim6_rhs(t, y, u):
    v, q, w, mass = y[3:6], y[6:10], y[10:13], y[13]
    aero = np.array([0, 0, 0])
    force = _body_to_inertial(q).dot(_thrust_body(u) + aero)
    dv = 1 / mass * force + [-G0, 0, 0]
    wx, wy, wz = w
    big_omega = np.array([[0, -wx, -wy, -wz], [wx, 0, wz, -wy], [wy, -wz, 0, wx], [wz, wy, -wx, 0]])
    torque = np.cross(ARM_T, _thrust_body(u)) + np.cross(ARM_CP, aero)
    dq = 0.5 * big_omega.dot(q)
    dw = J6_INV.dot(torque - np.cross(w, np.dot(J6, w)))
    dm = -u[2] / (G0 * ISP)
    return np.concatenate([v, dv, dq, dw, [dm]])


def sim6_advance(y, t, dt, u):
    def ground(t_, y_):
        return y_[0]
    ground.terminal = True
    sol = solve_ivp(fun=lambda t_, y_: sim6_rhs(t_, y_, u), t_span=[t, t + dt], y0=y, events=ground)
    y = np.array([row[-1] for row in sol.y])
    y[6:10] = y[6:10] / np.linalg.norm(y[6:10])
    return y, sol.status, round(t + dt, 3), sol.nfev


# ------------------------------------------------------------------------------------------ 3DOF simulator
def sim3_rhs(t, y, u):
    x, z, phi, vx, vz, om, mass = y
    delta, thrust = u[0], u[1]
    rho, cd, sref, inertia = 1.225, 0.3, 10.5, 6.04e6
    cn = 1 * 0
    q_dyn = 0.5 * rho * (vx ** 2 + vz ** 2)
    drag = cd * q_dyn * sref
    normal = cn * q_dyn * sref
    ax = (thrust * np.cos(delta + phi) - normal * np.sin(phi) - drag * np.cos(phi)) / mass
    az = (thrust * np.sin(delta + phi) + normal * np.cos(phi) - drag * np.cos(phi)) / mass - G0
    dom = (normal * (10 - 20) - thrust * np.sin(delta) * (40 - 10)) / inertia
    dm = -thrust / (ISP * G0)
    return np.array([vx, vz, om, ax, az, dom, dm])


def sim3_advance(y, t, dt, u):
    def ground(t_, y_):
        return y_[1]
    ground.terminal = True
    sol = solve_ivp(fun=lambda t_, y_: sim3_rhs(t_, y_, u), t_span=[t, t + dt], y0=y, events=ground)
    y = np.array([row[-1] for row in sol.y])
    y[2] = fmod(fmod(y[2], TWO_PI) + TWO_PI, TWO_PI)
    return y, sol.status, round(t + dt, 3), sol.nfev


# ------------------------------------------------------------------------------------------ environments
def _vtarg(r_hat, v_hat, tau, v0):
    t_go = np.linalg.norm(r_hat) / np.linalg.norm(v_hat)
    return -v0 * (np.array(r_hat) / max(1e-3, np.linalg.norm(r_hat))) * (1 - np.exp(-t_go / tau)), t_go


class PortEnv6:
    """Rocket6DOF without rendering/history; reset(ic) takes the f32 initial condition explicitly."""

    def __init__(self, IC, ICRange, timestep, reward_coeff, trajectory_limits, landing_params, seed=42):
        d = derived_6dof(IC, ICRange)
        self.normalizer, self.low, self.high = d["normalizer"], d["low"], d["high"]
        self.pos_low, self.pos_high = d["pos_low"], d["pos_high"]
        self.max_gimbal, self.max_thrust = np.deg2rad(20), 981e3
        self.dt, self.rc = timestep, reward_coeff
        self.att_limit = trajectory_limits["attitude_limit"]
        self.lp = landing_params
        self.omega_lim = np.array([0.2, 0.2, 0.2])
        self.rng = np.random.RandomState(seed)

    def sample_ic(self):
        return self.rng.uniform(self.low, self.high).astype(np.float32)

    def reset(self, ic=None):
        ic = self.sample_ic() if ic is None else np.array(ic, dtype=np.float32)
        ic[6:10] = ic[6:10] / np.linalg.norm(ic[6:10])
        self.ic, self.y, self.t = ic, ic, 0
        return (self.y / self.normalizer).astype("float32")

    def step(self, a):
        u = np.float32([a[0] * self.max_gimbal, a[1] * self.max_gimbal, (a[2] + 1) / 2.0 * self.max_thrust])
        self.y, status, self.t, self.nfev = sim6_advance(self.y, self.t, self.dt, u)
        s = self.y.astype(np.float32)
        rot = Rotation.from_quat(np.roll(s[6:10], -1))
        viol = not bool(np.can_cast(np.float32, np.float32) and np.all(np.float32(s[0:3]) >= self.pos_low)
                        and np.all(np.float32(s[0:3]) <= self.pos_high))
        done = bool(status) or viol
        r, v = s[0:3], s[3:6]
        v0 = np.linalg.norm(self.ic[3:6])
        if r[0] > self.lp["waypoint"]:
            vt, _ = _vtarg(r - [self.lp["waypoint"], 0, 0], v - [-2, 0, 0], 20, v0)
        else:
            vt, _ = _vtarg([r[0] + 1, 0, 0], v - [-1, 0, 0], 100, v0)
        euler = rot.as_euler("zyx")
        landed = all([s[0] <= 1e-3, np.linalg.norm(v) < self.lp["maximum_velocity"], np.linalg.norm(r) < self.lp["landing_radius"],
                      np.any(abs(euler) < self.lp["landing_attitude_limit"]), np.any(abs(s[10:13]) < self.omega_lim)])
        terms = {"velocity_tracking": self.rc["alfa"] * np.linalg.norm(v - vt), "thrust_penalty": self.rc["beta"] * u[2],
                 "eta": self.rc["eta"], "attitude_constraint": self.rc["gamma"] * np.any(np.abs(euler) > self.att_limit),
                 "rew_goal": self.rc["kappa"] * landed}
        reward = sum(terms.values())
        if viol:
            reward += -50
        return (self.y / self.normalizer).astype("float32"), reward, done, {"bounds_violation": viol, "rewards_dict": terms}


class PortEnv3:
    """Rocket (3DOF) without rendering/history."""

    def __init__(self, IC, ICRange, timestep, reward_coeff, seed=42):
        d = derived_3dof(IC, ICRange)
        self.normalizer, self.low, self.high = d["normalizer"], d["low"], d["high"]
        self.x_left, self.x_right, self.y_up = d["x_left"], d["x_right"], d["y_up"]
        self.max_gimbal, self.max_thrust = np.deg2rad(20), 981e3
        self.dt, self.rc = timestep, reward_coeff
        self.rng = np.random.RandomState(seed)

    def sample_ic(self):
        return self.rng.uniform(self.low, self.high).astype(np.float32)

    def reset(self, ic=None):
        ic = self.sample_ic() if ic is None else np.array(ic, dtype=np.float32)
        self.ic, self.y, self.t = ic, ic, 0
        return self.y.astype(np.float32) / self.normalizer

    def step(self, a):
        u = np.float32([a[0] * self.max_gimbal, (a[1] + 1) / 2.0 * self.max_thrust])
        self.y, status, self.t, self.nfev = sim3_advance(self.y, self.t, self.dt, u)
        s = self.y.astype(np.float32)
        viol = bool(s[0] <= self.x_left or s[0] >= self.x_right or s[1] >= self.y_up)
        done = bool(status) or viol
        r, v = s[0:2], s[3:5]
        zeta = s[2] - np.pi / 2
        v0 = np.linalg.norm(self.ic[3:5])
        if r[1] > self.rc["waypoint"]:
            vt, _ = _vtarg(r - [0, self.rc["waypoint"]], v - [0, -2], 20, v0)
        else:
            vt, _ = _vtarg([0, r[1]], v - [0, -1], 100, v0)
        landed = all([s[1] <= 1e-3, np.linalg.norm(v) < 15, np.linalg.norm(r) < self.rc["landing_radius"], abs(zeta) < 0.2,
                      abs(s[5]) < 0.2])
        terms = {"velocity_tracking": self.rc["alfa"] * np.linalg.norm(v - vt), "thrust_penalty": self.rc["beta"] * u[1],
                 "eta": self.rc["eta"], "attitude_constraint": self.rc["gamma"] * float(abs(zeta) > 2 * np.pi),
                 "attitude_hint": self.rc["delta"] * np.maximum(0, abs(zeta) - np.pi / 2), "rew_goal": self.rc["kappa"] * landed}
        reward = sum(terms.values())
        if viol:
            reward += -50
        return s / self.normalizer, reward, done, {"bounds_violation": viol, "rewards_dict": terms}


# ------------------------------------------------------------------------------------------ throughput harness
_W = {}


def _worker_init(dof, cfg, cap, base_seed):
    """Pool initializer: one persistent single-env worker per process (BLAS threads pinned to 1)."""
    import os
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    seed = base_seed + os.getpid() % 100000
    if dof == 6:
        env = PortEnv6(cfg["IC"], cfg["ICRange"], cfg["timestep"], cfg["reward_coeff"], cfg["trajectory_limits"], cfg["landing_params"], seed)
        na = 3
    else:
        env = PortEnv3(cfg["IC"], cfg["ICRange"], cfg["timestep"], cfg["reward_coeff"], seed)
        na = 2
    env.reset()
    _W.update(env=env, na=na, cap=cap, rng=np.random.RandomState(seed), k=0)


def _worker_run(seconds):
    """Random-action rollout with auto-reset for `seconds` of wall time; returns (steps, elapsed)."""
    import time
    env, na, cap, rng = _W["env"], _W["na"], _W["cap"], _W["rng"]
    k, steps = _W["k"], 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        _, _, done, _ = env.step(rng.uniform(-1, 1, na).astype(np.float32))
        k += 1
        steps += 1
        if done or k >= cap:
            env.reset()
            k = 0
    _W["k"] = k
    return steps, time.perf_counter() - t0


class CpuFarm:
    """`procs` persistent single-env worker processes (default: every host core) running the port."""

    def __init__(self, dof, cfg, cap, procs=None, seed=1000):
        import multiprocessing as mp
        import os
        self.procs = procs or os.cpu_count()
        self.pool = mp.get_context("spawn").Pool(self.procs, initializer=_worker_init, initargs=(dof, cfg, cap, seed))

    def sample(self, seconds):
        res = self.pool.map(_worker_run, [seconds] * self.procs, chunksize=1)
        steps = sum(r[0] for r in res)
        wall = max(r[1] for r in res)
        return dict(value=steps / wall, cores=self.procs, steps=steps, wall_s=wall)

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_throughput(dof, cfg, cap, seconds=8.0, warm=1.0, procs=None):
    """Aggregate env-steps/s of `procs` single-env worker processes (default: every host core)."""
    farm = CpuFarm(dof, cfg, cap, procs)
    try:
        farm.sample(warm)
        return farm.sample(seconds)
    finally:
        farm.close()
