import numpy as np


def check_env(env, warn=True, skip_render_check=True):
    """env_checker.py: the assertions of check_env that concern the reset / step surface (spaces, shapes, dtypes, types)."""
    from gym import spaces
    assert isinstance(env.observation_space, spaces.Box) and isinstance(env.action_space, (spaces.Box, spaces.Discrete))
    obs = env.reset()
    assert isinstance(obs, np.ndarray), "reset() must return a numpy array"
    assert obs.shape == env.observation_space.shape, (obs.shape, env.observation_space.shape)
    assert np.can_cast(obs.dtype, env.observation_space.dtype) or obs.dtype == np.float64
    assert np.all(obs >= env.observation_space.low - 1e-6) and np.all(obs <= env.observation_space.high + 1e-6)
    if isinstance(env.action_space, spaces.Box):
        assert np.all(env.action_space.low == -1) and np.all(env.action_space.high == 1), "action space must be normalised"
    a = env.action_space.sample()
    out = env.step(a)
    assert len(out) == 4
    o, r, d, info = out
    assert isinstance(o, np.ndarray) and o.shape == env.observation_space.shape
    assert isinstance(r, (float, int, np.floating)), type(r)
    assert isinstance(d, (bool, np.bool_)), type(d)
    assert isinstance(info, dict)
    if not skip_render_check:
        assert "render.modes" in env.metadata or "render_modes" in env.metadata
