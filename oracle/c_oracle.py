"""ctypes front-end of the C oracle (oracle/rocket_oracle.c) -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C

import numpy as np

from . import build_oracle
from .params import OrParams, make_params  # noqa: F401

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle.build())
        assert _lib.or_sizeof_params() == C.sizeof(OrParams), "OrParams layout mismatch"
        _lib.or_f32_norm.restype = C.c_float
        _lib.or_np_sinf.restype = C.c_float
        _lib.or_np_cosf.restype = C.c_float
        _lib.or_np_sinf.argtypes = [C.c_float]
        _lib.or_np_cosf.argtypes = [C.c_float]
        _lib.or_sim_step.restype = C.c_int
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def ndim(p):
    return 14 if p.dof == 6 else 7


def nact(p):
    return {6: 3, 3: 2, 1: 1}[p.dof]


def nobs(p):
    if p.dof == 3 and p.obs_mode == 1:
        return 4
    return 2 if p.dof == 1 else ndim(p)


def sample_ic(p, seed, env, episode):
    ic = np.zeros(ndim(p), np.float32)
    lib().or_sample_ic(C.byref(p), C.c_uint64(seed), C.c_int64(env), C.c_int32(episode), _p(ic, C.c_float))
    return ic


def sim_step(p, y, k, u):
    """Bare simulator step; returns (y_new f64[n], status, nfev, natt)."""
    y = np.array(y, dtype=np.float64)
    u = np.asarray(u, dtype=np.float32)
    nf, na = C.c_int(0), C.c_int(0)
    st = lib().or_sim_step(C.byref(p), _p(y, C.c_double), C.c_int(k), _p(u, C.c_float), C.byref(nf), C.byref(na))
    return y, st, nf.value, na.value


class ScalarEnv:
    """Single environment without TimeLimit (mirrors Rocket6DOF / Rocket reset()/step())."""

    def __init__(self, p):
        self.p = p
        self.n = ndim(p)
        self.y = np.zeros(self.n)
        self.k = 0
        self.v0 = C.c_float(0)

    def reset(self, ic):
        ic = np.asarray(ic, dtype=np.float32)
        obs = np.zeros(self.n, np.float32)
        lib().or_reset_from_ic(C.byref(self.p), _p(ic, C.c_float), _p(self.y, C.c_double), C.byref(self.v0), _p(obs, C.c_float))
        self.k = 0
        return obs

    def step(self, action, action_kind=0):
        a = np.asarray(action, dtype=np.float32)
        obs = np.zeros(self.n, np.float32)
        r = C.c_double(0)
        terms = np.zeros(6)
        flags = np.zeros(6, np.int32)
        vt = np.zeros(4)
        lib().or_env_step(C.byref(self.p), _p(self.y, C.c_double), C.c_int(self.k), self.v0, _p(a, C.c_float), C.c_int(action_kind),
                          _p(obs, C.c_float), C.byref(r), _p(terms, C.c_double), _p(flags, C.c_int32), _p(vt, C.c_double))
        self.k += 1
        info = dict(status=int(flags[0]), viol=bool(flags[1]), landed=bool(flags[2]), nfev=int(flags[4]), natt=int(flags[5]),
                    terms=terms, vtarg=vt[:3].copy(), t_go=float(vt[3]))
        return obs, r.value, bool(flags[3]), info


class VecOracle:
    """Batched oracle with TimeLimit + auto-reset; SoA state like the CUDA path."""

    def __init__(self, p, N, seed=42, env_offset=0):
        self.p, self.N, self.seed, self.env_offset = p, int(N), int(seed), int(env_offset)
        n = ndim(p)
        self.y = np.zeros((n, self.N))
        self.k = np.zeros(self.N, np.int32)
        self.episode = np.zeros(self.N, np.int32)
        self.v0 = np.zeros(self.N, np.float32)
        self.ep_ret = np.zeros(self.N)
        self._started = False

    def reset(self, ic=None):
        """First reset: ICs keyed by the current episode indices; later explicit resets advance them first (fresh ICs)."""
        if self._started:
            self.episode += 1
        self._started = True
        obs = np.zeros((self.N, nobs(self.p)), np.float32)
        if ic is not None:
            ic = np.ascontiguousarray(ic, dtype=np.float32)
        lib().or_vec_reset(C.byref(self.p), C.c_int64(self.N), C.c_int64(self.env_offset), C.c_uint64(self.seed),
                           _p(self.y, C.c_double), _p(self.k, C.c_int32), _p(self.episode, C.c_int32), _p(self.v0, C.c_float),
                           _p(self.ep_ret, C.c_double), _p(ic, C.c_float), _p(obs, C.c_float))
        return obs

    def step(self, actions, action_kind=0, ic_inject=None):
        N, p = self.N, self.p
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(N, nact(p))
        if ic_inject is not None:
            ic_inject = np.ascontiguousarray(ic_inject, dtype=np.float32)
        out = dict(
            obs=np.zeros((N, nobs(p)), np.float32), reward=np.zeros(N), done=np.zeros(N, np.uint8),
            truncated=np.zeros(N, np.uint8), terminal_obs=np.zeros((N, nobs(p)), np.float32),
            fin_ret=np.zeros(N), fin_len=np.zeros(N, np.int32), fin_landed=np.zeros(N, np.uint8),
            natt=np.zeros(N, np.int32), y_pre_reset=np.zeros((ndim(p), N)))
        lib().or_vec_step(C.byref(p), C.c_int64(N), C.c_int64(self.env_offset), C.c_uint64(self.seed),
                          _p(self.y, C.c_double), _p(self.k, C.c_int32), _p(self.episode, C.c_int32), _p(self.v0, C.c_float),
                          _p(self.ep_ret, C.c_double), _p(a, C.c_float), C.c_int(action_kind), _p(ic_inject, C.c_float),
                          _p(out["obs"], C.c_float), _p(out["reward"], C.c_double), _p(out["done"], C.c_uint8),
                          _p(out["truncated"], C.c_uint8), _p(out["terminal_obs"], C.c_float),
                          _p(out["fin_ret"], C.c_double), _p(out["fin_len"], C.c_int32), _p(out["fin_landed"], C.c_uint8),
                          _p(out["natt"], C.c_int32), _p(out["y_pre_reset"], C.c_double))
        return out


def f32_norm(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    return np.float32(lib().or_f32_norm(_p(x, C.c_float), C.c_int(x.size)))


def vtarg_simple(r, v, v0):
    r = np.asarray(r, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    vt = np.zeros(2)
    tgo = C.c_double(0)
    lib().or_vtarg_simple(_p(r, C.c_double), _p(v, C.c_double), C.c_double(v0), _p(vt, C.c_double), C.byref(tgo))
    return vt, tgo.value
