"""Host-side constants of the reference environments -- ORACLE side (test infrastructure).

Restates, with the same numpy expressions (so the same float32/float64 promotion), what the
reference constructors compute once:
  Rocket6DOF.__init__  my_environment/envs/rocket_env.py:511-663
  Rocket.__init__      my_environment/envs/rocket_env.py:27-135
"""
import ctypes as C

import numpy as np

NMAX = 14


class OrParams(C.Structure):
    _fields_ = [
        ("dof", C.c_int), ("reward_mode", C.c_int), ("max_episode_steps", C.c_int), ("dt_ms", C.c_int),
        ("dt", C.c_double),
        ("normalizer", C.c_double * NMAX),
        ("ic_low", C.c_float * NMAX), ("ic_high", C.c_float * NMAX),
        ("pos_low", C.c_float * 3), ("pos_high", C.c_float * 3),
        ("x_left", C.c_double), ("x_right", C.c_double), ("y_up", C.c_double),
        ("alfa", C.c_double), ("beta", C.c_double), ("eta", C.c_double), ("gamma", C.c_double),
        ("delta", C.c_double), ("kappa", C.c_double), ("xi", C.c_double),
        ("waypoint", C.c_double), ("landing_radius", C.c_double), ("maximum_velocity", C.c_double),
        ("att_traj_limit", C.c_double * 3), ("att_land_limit", C.c_double * 3), ("omega_lim", C.c_double * 3),
        ("max_gimbal", C.c_double), ("max_thrust", C.c_float), ("obs_mode", C.c_int), ("min_thrust", C.c_float),
        ("aero_mode", C.c_int),
    ]


def _dt_ms(timestep):
    ms = round(timestep * 1000.0)
    if abs(ms - timestep * 1000.0) > 1e-9 or ms <= 0:
        raise ValueError("timestep must be a positive whole number of milliseconds (the reference rounds t to 3 decimals)")
    return int(ms)


def derived_6dof(IC, ICRange):
    """normalizer f64[14], bounds f32[3]x2, IC box f32[14]x2  (rocket_env.py:557-620)."""
    ICMean = np.float32(IC)
    ICR = np.float32(ICRange)
    low = ICMean - ICR / 2
    high = ICMean + ICR / 2
    max_gimbal = np.deg2rad(20)
    max_thrust = 981e3
    t_free_fall = (-ICMean[3] + np.sqrt(ICMean[3] ** 2 + 2 * 9.81 * ICMean[0])) / 9.81
    inertia = 6.04e6
    lever_arm = 15.0
    omega_max = max_thrust * np.sin(max_gimbal) * lever_arm / (inertia) * t_free_fall / 5.0
    v_max = 2 * 9.81 * t_free_fall
    normalizer = np.maximum(
        np.array([1.2 * abs(ICMean[0]), 1.5 * abs(ICMean[1]), 1.5 * abs(ICMean[2]), v_max, v_max, v_max,
                  1.1, 1.1, 1.1, 1.1, omega_max, omega_max, omega_max, ICMean[13] + ICR[13]]), 1)
    pb_high = 0.9 * np.maximum(normalizer[0:3], 100)
    pb_low = -0.9 * np.maximum(normalizer[1:3], 100)
    pb_low = np.insert(pb_low, 0, -30)
    return dict(normalizer=np.asarray(normalizer, dtype=np.float64), low=low.astype(np.float32), high=high.astype(np.float32),
                pos_low=pb_low.astype(np.float32), pos_high=pb_high.astype(np.float32),
                max_gimbal=float(max_gimbal), max_thrust=max_thrust)


def derived_3dof(IC, ICRange):
    """normalizer f64[7], x/y bounds, IC box  (rocket_env.py:51-100)."""
    ICMean = np.float32(IC)
    ICR = np.float32(ICRange)
    low = ICMean - ICR / 2
    high = ICMean + ICR / 2
    max_gimbal = np.deg2rad(20)
    max_thrust = 981e3
    t_free_fall = (-ICMean[4] + np.sqrt(ICMean[4] ** 2 + 2 * 9.81 * ICMean[1])) / 9.81
    inertia = 6.04e6
    lever_arm = 30.0
    normalizer = np.maximum(
        np.array([1.5 * abs(ICMean[0]), 1.5 * abs(ICMean[1]), 2 * np.pi, 2 * 9.81 * t_free_fall, 2 * 9.81 * t_free_fall,
                  max_thrust * np.sin(max_gimbal) * lever_arm / (inertia) * t_free_fall / 5.0, ICMean[6] + ICR[6]]), 1)
    x_right = 0.9 * np.maximum(normalizer[0], 100)
    y_up = 0.9 * np.maximum(normalizer[1], 100)
    return dict(normalizer=np.asarray(normalizer, dtype=np.float64), low=low.astype(np.float32), high=high.astype(np.float32),
                x_right=float(x_right), x_left=float(-x_right), y_up=float(y_up),
                max_gimbal=float(max_gimbal), max_thrust=max_thrust)


_DEF_RC6 = {"alfa": -0.01, "beta": -1e-8, "eta": 2, "gamma": -10, "delta": -5, "kappa": 10, "xi": 0.004}
_DEF_RC3 = {"alfa": -0.01, "beta": -1e-8, "eta": 2, "gamma": -10, "delta": -5, "kappa": 10, "waypoint": 50, "landing_radius": 30}
_DEF_TL = {"attitude_limit": [1.5, 1.5, 2 * np.pi]}
_DEF_LP = {"waypoint": 50, "landing_radius": 30, "maximum_velocity": 10,
           "landing_attitude_limit": [0.2, 0.2, 2 * np.pi], "omega_lim": [0.2, 0.2, 0.2]}


def make_params(dof, IC=None, ICRange=None, timestep=0.1, seed=42, reward_coeff=None, trajectory_limits=None,
                landing_params=None, max_episode_steps=0, reward_mode=0, obs_mode=0, min_thrust=0.0, aero_mode=0):
    p = OrParams()
    p.obs_mode = int(obs_mode)
    p.min_thrust = float(min_thrust)
    p.aero_mode = int(aero_mode)
    if aero_mode and dof == 6:
        raise ValueError("the angle-of-attack hook is defined for the 3DOF / 1DOF environment (the 6DOF aerodynamic force is a zero stub)")
    if obs_mode and dof != 3:
        raise ValueError("the Gaudet observation is defined for the 3DOF environment")
    p.dof = dof
    p.reward_mode = reward_mode
    p.max_episode_steps = int(max_episode_steps or 0)
    p.dt = float(timestep)
    p.dt_ms = _dt_ms(timestep)
    if dof == 6:
        IC = [500, 100, 100, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 50e3] if IC is None else IC
        ICRange = [50, 10, 10, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 1e3] if ICRange is None else ICRange
        rc = dict(_DEF_RC6) if reward_coeff is None else reward_coeff
        tl = _DEF_TL if trajectory_limits is None else trajectory_limits
        lp = _DEF_LP if landing_params is None else landing_params
        d = derived_6dof(IC, ICRange)
        n = 14
        p.pos_low[:] = d["pos_low"].tolist()
        p.pos_high[:] = d["pos_high"].tolist()
        p.waypoint = lp["waypoint"]
        p.landing_radius = lp["landing_radius"]
        p.maximum_velocity = lp["maximum_velocity"]
        p.att_traj_limit[:] = [float(v) for v in tl["attitude_limit"]]
        p.att_land_limit[:] = [float(v) for v in lp["landing_attitude_limit"]]
        p.omega_lim[:] = [0.2, 0.2, 0.2]          # rocket_env.py:656 ignores landing_params["omega_lim"]
    else:
        IC = [100, 500, np.pi / 2, -10, -50, 0, 50e3] if IC is None else IC
        ICRange = [10, 50, 0.1, 1, 10, 0.1, 1e3] if ICRange is None else ICRange
        rc = dict(_DEF_RC3) if reward_coeff is None else reward_coeff
        d = derived_3dof(IC, ICRange)
        n = 7
        p.x_left, p.x_right, p.y_up = d["x_left"], d["x_right"], d["y_up"]
        p.waypoint = rc["waypoint"]
        p.landing_radius = rc["landing_radius"]
    for i in range(n):
        p.normalizer[i] = float(d["normalizer"][i])
        p.ic_low[i] = float(d["low"][i])
        p.ic_high[i] = float(d["high"][i])
    for i in range(n, NMAX):
        p.normalizer[i] = 1.0
    p.alfa, p.beta, p.eta, p.gamma, p.kappa = rc["alfa"], rc["beta"], rc["eta"], rc["gamma"], rc["kappa"]
    p.delta = rc.get("delta", 0.0)
    p.xi = rc.get("xi", 0.01)
    p.max_gimbal = d["max_gimbal"]
    p.max_thrust = d["max_thrust"]
    return p
