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
    from rl_rocket_b200.config import env_config_6dof
    return env_config_6dof()


def cfg3():
    from rl_rocket_b200.config import env_config_3dof
    return env_config_3dof()


def cfg3_default():
    return dict(IC=[100, 500, np.pi / 2, -10, -50, 0, 50e3], ICRange=[10, 50, 0.1, 1, 10, 0.1, 1e3], timestep=0.1, seed=42,
                reward_coeff={"alfa": -0.01, "beta": -1e-8, "eta": 2, "gamma": -10, "delta": -5, "kappa": 10,
                              "waypoint": 50, "landing_radius": 30})


def cfg1():
    from rl_rocket_b200.config import env_config_1dof
    return env_config_1dof()


def cfg6_landing_attitude():
    """env_config with a landing attitude limit that CAN fail (the configured [10 deg, 10 deg, 2 pi] cannot: |angle| <= pi)."""
    c = cfg6()
    c["landing_params"] = {"waypoint": 50, "landing_radius": 30, "maximum_velocity": 10, "landing_attitude_limit": [0.1, 0.1, 0.1],
                           "omega_lim": [0.2, 0.2, 0.2]}
    return c


def cfg6_landing_random():
    c = cfg6()
    c["IC"] = [4, 0, 0, -6, 0, 0, 1, 0, 0, 0, 0, 0, 0, 45e3]
    c["ICRange"] = [4, 70, 70, 12, 4, 4, 0.2, 0.2, 0.2, 0.2, 0.8, 0.8, 0.8, 1e3]
    return c


def _zero_range(cfgf):
    def f():
        c = cfgf()
        c["ICRange"] = [0] * len(c["IC"])
        return c
    return f


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
    # round 2: the branches round 1 left unpinned -- touchdowns with rew_goal != 0 (np.any vs np.all, each landing limit from both
    # sides, adjacent float32 ICs around the speed / radius thresholds), integrator failure (zero mass: solve_ivp status -1),
    # actions outside the Box (numpy's float32 sin / cos range reduction), DiscreteActions3DOF(RewardAnnealing(...)).
    # Files that carry an "ics" array restart every episode of env i from ics[i] (injected initial conditions).
    "traj_6dof_landing.npz": (6, _zero_range(cfg6), 800, 0, "norm"),
    "traj_6dof_landing_attitude.npz": (6, _zero_range(cfg6_landing_attitude), 800, 0, "norm"),
    "traj_6dof_landing_random.npz": (6, cfg6_landing_random, 800, 0, "norm"),
    "traj_3dof_landing.npz": (3, _zero_range(cfg3), 3000, 0, "norm"),
    "traj_6dof_unclipped.npz": (6, cfg6, 800, 0, "norm"),
    "traj_3dof_unclipped.npz": (3, cfg3, 3000, 0, "norm"),
    "traj_3dof_discrete_annealing.npz": (3, cfg3, 3000, 1, "discrete"),
    "traj_3dof_timelimit.npz": (3, cfg3, 30, 0, "norm"),
}
# observation mode per file (default: the normalised state): GaudetStateObs (wrappers.py:51-66) with its defects fixed, generated
# by oracle/make_golden.py::gaudet_fixed over the live reference environment
OBS_MODE = {"traj_3dof_gaudet.npz": 1, "traj_3dof_default_gaudet.npz": 1}
TRAJ["traj_3dof_gaudet.npz"] = (3, cfg3, 3000, 0, "norm")
TRAJ["traj_3dof_default_gaudet.npz"] = (3, cfg3_default, 2000, 0, "norm")


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
    gobs = g["obs"].astype(np.float64)
    odev = np.abs(np.asarray(got_obs, np.float64) - gobs) / np.maximum(1.0, np.abs(gobs))     # (1 f32 ulp, also for un-normalised rows)
    assert odev.max() <= obs_tol, f"{tag} step {t}: obs deviates {odev.max():.3e}"
    return float(np.nanmax(dev)), float(rdev.max())
