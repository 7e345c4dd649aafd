"""oracle/scipy_port.py (the Python/scipy CPU-baseline port) against the live reference's golden vectors."""
import numpy as np
import pytest

from oracle import scipy_port as SP
from tests import goldutil as G


@pytest.mark.parametrize("tag", ["null", "nonnull", "mixed"])
def test_port_6dof_kat2(tag):
    g = G.load("kat_6dof.npz")
    cfg = G.cfg6()
    cfg["ICRange"] = [0] * 14
    env = SP.PortEnv6(cfg["IC"], cfg["ICRange"], cfg["timestep"], cfg["reward_coeff"], cfg["trajectory_limits"], cfg["landing_params"])
    obs = env.reset(np.float32(cfg["IC"]))
    assert np.array_equal(obs, g[f"kat2_{tag}_obs"][0])
    act = g[f"kat2_{tag}_action"]
    n, done = 0, False
    while not done:
        obs, r, done, info = env.step(act)
        # same library calls in the same order as the reference: bit-identical
        assert np.array_equal(env.y, g[f"kat2_{tag}_y"][n])
        assert r == g[f"kat2_{tag}_reward"][n] and done == bool(g[f"kat2_{tag}_done"][n])
        assert np.array_equal(obs, g[f"kat2_{tag}_obs"][n + 1])
        n += 1
    assert n == len(g[f"kat2_{tag}_reward"])


@pytest.mark.parametrize("tag", ["null", "mixed"])
def test_port_3dof_kat3(tag):
    g = G.load("kat_3dof.npz")
    cfg = G.cfg3()
    cfg["ICRange"] = [0] * 7
    env = SP.PortEnv3(cfg["IC"], cfg["ICRange"], cfg["timestep"], cfg["reward_coeff"])
    env.reset(np.float32(cfg["IC"]))
    act = g[f"kat3_{tag}_action"]
    n, done = 0, False
    while not done:
        obs, r, done, info = env.step(act)
        assert np.array_equal(env.y, g[f"kat3_{tag}_y"][n])
        assert r == g[f"kat3_{tag}_reward"][n] and done == bool(g[f"kat3_{tag}_done"][n])
        assert np.array_equal(obs, g[f"kat3_{tag}_obs"][n + 1])
        n += 1
    assert n == len(g[f"kat3_{tag}_reward"])


def test_port_6dof_random_trajectory_env0():
    g = G.load("traj_6dof.npz")
    cfg = G.cfg6()
    env = SP.PortEnv6(cfg["IC"], cfg["ICRange"], cfg["timestep"], cfg["reward_coeff"], cfg["trajectory_limits"], cfg["landing_params"])
    env.reset(g["ic0"][0])
    for t in range(200):
        obs, r, done, info = env.step(g["actions"][t, 0])
        assert np.array_equal(env.y, g["y"][t, 0]) and r == g["reward"][t, 0] and done == bool(g["done"][t, 0])
        assert env.nfev == g["nfev"][t, 0]
        if done:
            env.reset(g["ic"][t, 0])
