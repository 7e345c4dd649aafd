"""CPU-side tests of the host logic: the C-ABI library loads and exports what include/rl_rocket_b200.h declares,
EnvSpec reproduces the reference constructors' constants, spaces / wrappers / sharding behave.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from rl_rocket_b200 import _native as N
from rl_rocket_b200 import spaces
from rl_rocket_b200.config import EnvSpec
from tests import goldutil as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "rl_rocket_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rlr_[a-z0-9_]+)\s*\(", hdr)))


def test_library_loads_and_exports_every_declared_symbol():
    from rl_rocket_b200.csrc import build
    build.build()
    lib = N.load()
    syms = _declared_symbols()
    assert set(syms) == set(N.EXPORTS), (syms, N.EXPORTS)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert lib.rlr_abi_version() == N.ABI_VERSION == 6
    assert lib.rlr_sizeof_params() == ctypes.sizeof(N.RlrParams)


def test_abi_argument_validation_without_gpu():
    lib = N.load()
    spec = EnvSpec(6, **G.cfg6())
    assert lib.rlr_ndim(ctypes.byref(spec.params)) == 14 and lib.rlr_nact(ctypes.byref(spec.params)) == 3
    spec3 = EnvSpec(1, **G.cfg1())
    assert lib.rlr_ndim(ctypes.byref(spec3.params)) == 7 and lib.rlr_nobs(ctypes.byref(spec3.params)) == 2
    st = N.RlrState(None, None, None, None, None)
    # empty batch is a no-op, NULL state is rejected before any CUDA call
    assert lib.rlr_reset(ctypes.byref(spec.params), N.F64, st, None, None, None, 0, 0, 0, None) == 0
    assert lib.rlr_reset(ctypes.byref(spec.params), N.F64, st, None, None, None, 4, 0, 0, None) == -1
    assert b"NULL" in lib.rlr_last_error()
    bad = EnvSpec(6, **G.cfg6()).params
    bad.dof = 5
    assert lib.rlr_reset(ctypes.byref(bad), N.F64, st, None, None, None, 4, 0, 0, None) == -1
    spec3d = EnvSpec(3, **G.cfg3())
    assert lib.rlr_reset(ctypes.byref(spec3d.params), N.F32, N.RlrState(1, 1, 1, 1, 1), None, None, None, 4, 0, 0, None) == -3      # f32 state: 6DOF only


def test_step_chained_argument_validation_without_gpu():
    lib = N.load()
    spec = EnvSpec(6, **G.cfg6())
    st, out = N.RlrState(1, 1, 1, 1, 1), N.RlrStepOut()
    out.obs, out.done = 1, 1
    P = ctypes.byref(spec.params)
    # block_seq, and a non-zero seq_publish different from seq_wait, are required
    assert lib.rlr_step_chained(P, N.F64, N.MATH_FAST, st, 1, N.ACT_NORMALIZED, None, out, 64, 0, 0, None, 0, 1, None) == -1
    assert lib.rlr_step_chained(P, N.F64, N.MATH_FAST, st, 1, N.ACT_NORMALIZED, None, out, 64, 0, 0, 1, 0, 0, None) == -1
    assert lib.rlr_step_chained(P, N.F64, N.MATH_FAST, st, 1, N.ACT_NORMALIZED, None, out, 64, 0, 0, 1, 3, 3, None) == -1
    assert b"seq_publish" in lib.rlr_last_error()
    # 6DOF, f64 state only; an empty batch is a no-op
    assert lib.rlr_step_chained(P, N.F32, N.MATH_FAST, st, 1, N.ACT_NORMALIZED, None, out, 64, 0, 0, 1, 0, 1, None) == -3
    assert lib.rlr_step_chained(ctypes.byref(EnvSpec(3, **G.cfg3()).params), N.F64, N.MATH_FAST, st, 1, N.ACT_NORMALIZED, None, out, 64, 0, 0, 1, 0, 1,
                                None) == -3
    assert lib.rlr_step_chained(P, N.F64, N.MATH_FAST, st, 1, N.ACT_NORMALIZED, None, out, 0, 0, 0, 1, 0, 1, None) == 0
    assert lib.rlr_step_chained(P, N.F64, N.MATH_FAST, N.RlrState(None, None, None, None, None), 1, N.ACT_NORMALIZED, None, out, 64, 0, 0, 1, 0, 1, None) == -1


def test_envspec_matches_reference_constants():
    # values the live reference computes for configuration_file.env_config
    s = EnvSpec(6, **G.cfg6())
    assert np.allclose(s.state_normalizer[[0, 1, 2, 6, 13]], [600, 150, 150, 1.1, 46000])
    assert abs(s.state_normalizer[3] - 321.900879) < 1e-5 and abs(s.state_normalizer[10] - 2.73418671) < 1e-7
    assert s.pos_low.tolist() == [-30, -135, -135] and s.pos_high.tolist() == [540, 135, 135]
    # main.py config
    s3 = EnvSpec(3, **G.cfg3())
    assert np.allclose(s3.state_normalizer[[0, 1, 6]], [2400, 3000, 51000]) and abs(s3.state_normalizer[2] - 2 * np.pi) < 1e-15
    assert abs(s3.x_bound_right - 2160) < 1e-9 and abs(s3.y_bound_up - 2700) < 1e-9
    # the golden obs0 of the live reference is IC / normalizer
    g = G.load("traj_6dof.npz")
    sel = [0, 1, 2, 3, 4, 5, 10, 11, 12, 13]          # (the quaternion is re-normalised at reset)
    assert np.abs((g["ic0"] / s.state_normalizer).astype(np.float32) - g["obs0"])[:, sel].max() < 2e-7


def test_envspec_product_and_oracle_params_agree():
    from oracle import c_oracle as O
    for dof, cfg in ((6, G.cfg6()), (3, G.cfg3()), (3, G.cfg3_default()), (1, G.cfg1())):
        s = EnvSpec(dof, max_episode_steps=123, reward_mode=1, **cfg)
        p = O.make_params(dof, max_episode_steps=123, reward_mode=1, **cfg)
        for f, _ in O.OrParams._fields_:
            a, b = getattr(s.params, f), getattr(p, f)
            a = list(a) if hasattr(a, "__len__") else a
            b = list(b) if hasattr(b, "__len__") else b
            assert a == b, f


def test_envspec_validation_and_discrete_table():
    with pytest.raises(ValueError):
        EnvSpec(6, timestep=0.0125, **{k: v for k, v in G.cfg6().items() if k != "timestep"})
    with pytest.raises(ValueError):
        EnvSpec(6, IC=[1, 2, 3])
    with pytest.raises(ValueError):
        EnvSpec(6, discrete_table=G.DISCRETE_TABLE, **G.cfg6())
    s = EnvSpec(3, discrete_table=G.DISCRETE_TABLE, **G.cfg3())
    assert s.params.n_discrete == 4
    assert np.array_equal(s.discrete_table, G.discrete_to_physical(np.arange(4)))
    s2 = s.with_(reward_mode=1)
    assert s2.params.reward_mode == 1 and s2.params.n_discrete == 4 and s.params.reward_mode == 0


def test_spaces_shim():
    b = spaces.Box(low=-1, high=1, shape=(14,))
    assert b.shape == (14,) and b.dtype == np.float32 and b.contains(np.zeros(14, np.float32))
    assert not b.contains(np.full(14, 2, np.float32)) and not b.contains(np.full(14, np.nan, np.float32))
    b.seed(3)
    x = b.sample()
    assert x.dtype == np.float32 and b.contains(x)
    d = spaces.Discrete(4)
    assert d.n == 4 and d.contains(3) and not d.contains(4)


def test_no_cpu_fallback_in_product(monkeypatch):
    """The product path must fail loudly without the CUDA extension / a GPU, and must never import the oracle."""
    import torch
    from rl_rocket_b200 import engine
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            engine.RocketBatch(EnvSpec(6, **G.cfg6()), 4)
    monkeypatch.setattr(N, "_lib", None)
    monkeypatch.setattr(N, "LIB_PATH", "/nonexistent/librlrocket.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        N.load()
    pkg = os.path.join(ROOT, "rl_rocket_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
                assert "oracle/" not in src and "rocket_oracle" not in src, f"{f} references the oracle"


def test_wrappers_logic_with_a_stub_env():
    from rl_rocket_b200 import wrappers as W

    class Stub(spaces.GymEnv):
        reward_coefficients = {"xi": 0.004}
        action_space = spaces.Box(-1, 1, (2,))
        observation_space = spaces.Box(-1, 1, (7,))

        def __init__(self):
            self.last = None

        def reset(self):
            return np.zeros(7)

        def step(self, a):
            self.last = np.asarray(a)
            return np.zeros(7), -1.0, False, {"rewards_dict": {"attitude_constraint": -10.0, "rew_goal": 10, "eta": 5}}

    d = W.DiscreteActions3DOF(Stub())
    d.step(1)
    assert d.env.last.tolist() == [-1, 1] and d.action_space.n == 4
    a = W.RewardAnnealing(Stub())
    _, r, _, info = a.step(np.float32([0.0, 0.5]))
    assert r == pytest.approx(-10.0 + 10 - 0.004 * 1.5) and set(info["rewards_dict"]) == {"attitude_constraint", "rew_goal", "thrust_penalty"}
    t = W.TimeLimit(Stub(), 3)
    t.reset()
    flags = [t.step([0, 0])[2:] for _ in range(3)]
    assert [f[0] for f in flags] == [False, False, True] and flags[-1][1]["TimeLimit.truncated"] is True


def test_shard_partition():
    from rl_rocket_b200.distributed import shard
    for total, world in ((1 << 20, 8), (65536, 2), (10, 4), (3, 8)):
        parts = [shard(total, world, r) for r in range(world)]
        assert parts[0][0] == 0 and sum(n for _, n in parts) == total
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
    assert shard(1 << 20, 8, 3) == (3 * 131072, 131072)
