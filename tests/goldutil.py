"""Shared helpers for the golden-vector tests (configs + comparison rules).

Tolerances: done / truncated / episode lengths / RK attempt counts bit-exact;
fp64 state |a-b|/normalizer <= 1e-9; reward |a-b| <= 1e-9*max(1,|b|); obs (f32) compared on the
normalised scale with an absolute 2e-7 (1 ulp of O(1) f32 values; components that are pure
rounding noise around zero cannot be compared in ulps).
"""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STATE_TOL = 1e-9
REWARD_TOL = 1e-9
OBS_TOL = 2e-7


def load(name):
    return np.load(os.path.join(GOLD, name))


def cfg6():
    pi = np.pi
    return {
        "timestep": 0.1, "seed": 42,
        "IC": [500, 100, 100, -50, 0, 0, 1, 0, 0, 0, 0, 0, 0, 45e3],
        "ICRange": [50, 10, 10, 10, 10, 10, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 1e3],
        "reward_coeff": {"alfa": -0.01, "beta": -1e-7, "delta": -5, "eta": 0.05, "gamma": -10, "kappa": 10, "xi": 0.004},
        "trajectory_limits": {"attitude_limit": [1.5, 1.5, 2 * pi]},
        "landing_params": {"waypoint": 50, "landing_radius": 30, "maximum_velocity": 10,
                           "landing_attitude_limit": [10 / 180 * pi, 10 / 180 * pi, 2 * pi], "omega_lim": [0.2, 0.2, 0.2]},
    }


def cfg3():
    pi = np.pi
    return {
        "IC": [-1600, 2000, pi * 3 / 4, 180, -90, 0, 50e3], "ICRange": [5, 50, 0, 0, 0, 0, 1e3], "timestep": 0.05, "seed": 42,
        "reward_coeff": {"alfa": -0.01, "beta": 0, "delta": -5, "eta": 0.2, "gamma": -10, "kappa": 10, "xi": 0.004,
                         "waypoint": 30, "landing_radius": 30},
    }


def cfg3_default():
    return dict(IC=[100, 500, np.pi / 2, -10, -50, 0, 50e3], ICRange=[10, 50, 0.1, 1, 10, 0.1, 1e3], timestep=0.1, seed=42,
                reward_coeff={"alfa": -0.01, "beta": -1e-8, "eta": 2, "gamma": -10, "delta": -5, "kappa": 10,
                              "waypoint": 50, "landing_radius": 30})


def cfg1():
    return dict(IC=[0, 2000, np.pi / 2, 0, -90, 0, 50e3], ICRange=[0, 50, 0, 0, 0, 0, 1e3], timestep=0.05, seed=42,
                reward_coeff=cfg3()["reward_coeff"])


DISCRETE_TABLE = [[0, -1], [-1, +1], [0, +1], [+1, +1]]


def discrete_to_physical(a_int, table=DISCRETE_TABLE):
    """What Rocket._denormalize_action (rocket_env.py:395-406) yields for DiscreteActions3DOF's int64 rows."""
    max_gimbal, max_thrust = np.deg2rad(20), 981e3
    rows = np.asarray(table)[np.asarray(a_int)]
    return np.stack([np.float32(rows[..., 0] * max_gimbal), np.float32((rows[..., 1] + 1) / 2.0 * max_thrust)], -1).astype(np.float32)


# name -> (dof, cfg factory, max_episode_steps, reward_mode, action kind)
TRAJ = {
    "traj_6dof.npz": (6, cfg6, 800, 0, "norm"),
    "traj_6dof_timelimit.npz": (6, cfg6, 25, 0, "norm"),
    "traj_6dof_annealing.npz": (6, cfg6, 800, 1, "norm"),
    "traj_3dof.npz": (3, cfg3, 3000, 0, "norm"),
    "traj_3dof_default.npz": (3, cfg3_default, 2000, 0, "norm"),
    "traj_3dof_discrete.npz": (3, cfg3, 3000, 0, "discrete"),
    "traj_3dof_annealing.npz": (3, cfg3, 3000, 1, "norm"),
    "traj_1dof.npz": (1, cfg1, 3000, 0, "norm"),
}


def compare_step(tag, t, normalizer, g, got_y, got_obs, got_reward, got_done, got_trunc, got_natt=None, state_tol=STATE_TOL,
                 reward_tol=REWARD_TOL, obs_tol=OBS_TOL):
    """g: dict of golden arrays for step t (already indexed)."""
    assert np.array_equal(np.asarray(got_done, bool), g["done"].astype(bool)), f"{tag} step {t}: done flags differ"
    assert np.array_equal(np.asarray(got_trunc, bool), g["truncated"].astype(bool)), f"{tag} step {t}: truncated flags differ"
    if got_natt is not None:
        assert np.array_equal(2 + 6 * np.asarray(got_natt), g["nfev"]), f"{tag} step {t}: RK45 attempt counts differ"
    dev = np.abs(np.asarray(got_y).T - g["y"]) / normalizer
    assert np.nanmax(dev) <= state_tol, f"{tag} step {t}: state deviates {np.nanmax(dev):.3e}"
    rdev = np.abs(np.asarray(got_reward) - g["reward"]) / np.maximum(1.0, np.abs(g["reward"]))
    assert rdev.max() <= reward_tol, f"{tag} step {t}: reward deviates {rdev.max():.3e}"
    odev = np.abs(np.asarray(got_obs, np.float64) - g["obs"].astype(np.float64))
    assert odev.max() <= obs_tol, f"{tag} step {t}: obs deviates {odev.max():.3e}"
    return float(np.nanmax(dev)), float(rdev.max())
