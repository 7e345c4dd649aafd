"""ctypes binding of the C ABI in include/rl_rocket_b200.h (rl_rocket_b200/_C/librlrocket.so).

There is NO CPU fallback: if the CUDA library is missing or does not load, importing the engine
raises.  Build it with `python -m rl_rocket_b200.csrc.build` (nvcc, sm_100a).
"""
import ctypes as C
import os

NMAX = 14
MAX_DISCRETE = 16
NSTATS = 8
STAT_NAMES = ("steps", "attempts", "done", "landed", "viol", "trunc", "len_sum", "ret_sum_fx")

ABI_VERSION = 6
FIN_HEADER = 5
ACT_NORMALIZED, ACT_PHYSICAL, ACT_DISCRETE = 0, 1, 2
MATH_EXACT, MATH_FAST = 0, 1
F64, F32 = 0, 1
OBS_STATE, OBS_GAUDET = 0, 1

LIB_PATH = os.environ.get("RLR_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_C", "librlrocket.so")   # RLR_LIB: tuning builds only


class RlrParams(C.Structure):
    _fields_ = [
        ("dof", C.c_int32), ("reward_mode", C.c_int32), ("max_episode_steps", C.c_int32), ("dt_ms", C.c_int32),
        ("dt", C.c_double),
        ("normalizer", C.c_double * NMAX), ("inv_normalizer", C.c_double * NMAX),
        ("ic_low", C.c_float * NMAX), ("ic_high", C.c_float * NMAX),
        ("pos_low", C.c_float * 3), ("pos_high", C.c_float * 3),
        ("x_left", C.c_double), ("x_right", C.c_double), ("y_up", C.c_double),
        ("alfa", C.c_double), ("beta", C.c_double), ("eta", C.c_double), ("gamma", C.c_double),
        ("delta", C.c_double), ("kappa", C.c_double), ("xi", C.c_double),
        ("waypoint", C.c_double), ("landing_radius", C.c_double), ("maximum_velocity", C.c_double),
        ("att_traj_limit", C.c_double * 3), ("att_land_limit", C.c_double * 3), ("omega_lim", C.c_double * 3),
        ("att_traj_c", C.c_double * 3), ("att_land_c", C.c_double * 3),
        ("max_gimbal", C.c_double), ("max_thrust", C.c_float),
        ("n_discrete", C.c_int32), ("discrete_table", C.c_float * (2 * MAX_DISCRETE)),
        ("discrete_thr", C.c_float * MAX_DISCRETE), ("obs_mode", C.c_int32), ("min_thrust", C.c_float),
        ("aero_mode", C.c_int32),
    ]


class RlrState(C.Structure):
    _fields_ = [("y", C.c_void_p), ("k", C.c_void_p), ("episode", C.c_void_p), ("v0", C.c_void_p), ("ep_ret", C.c_void_p)]


class RlrStepOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("obs", "reward", "reward64", "done", "truncated", "terminal_obs", "fin_ret", "fin_len",
                                           "flags", "natt", "y_pre_reset", "reward_terms", "vtarg", "stats", "fin_count", "fin_records")]


EXPORTS = ("rlr_abi_version", "rlr_last_error", "rlr_sizeof_params", "rlr_ndim", "rlr_nact", "rlr_nobs",
           "rlr_reset", "rlr_step", "rlr_step_range", "rlr_step_chained", "rlr_rollout", "rlr_fma_peak",
           "rlr_pipe_create", "rlr_pipe_destroy", "rlr_pipe_submit", "rlr_pipe_wait",
           "rlr_exchange_create", "rlr_exchange_handle", "rlr_exchange_connect", "rlr_exchange_allgather", "rlr_exchange_status",
           "rlr_exchange_destroy")

_lib = None


def load():
    """Load librlrocket.so (once).  Raises RuntimeError when it is absent -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build the CUDA extension with `python -m rl_rocket_b200.csrc.build` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise RuntimeError(f"{LIB_PATH} does not export {name}")
    lib.rlr_last_error.restype = C.c_char_p
    if lib.rlr_abi_version() != ABI_VERSION:
        raise RuntimeError("librlrocket.so ABI version mismatch")
    if lib.rlr_sizeof_params() != C.sizeof(RlrParams):
        raise RuntimeError("RlrParams layout mismatch between Python and librlrocket.so")
    P = C.POINTER(RlrParams)
    lib.rlr_ndim.argtypes = lib.rlr_nact.argtypes = lib.rlr_nobs.argtypes = [P]
    lib.rlr_reset.argtypes = [P, C.c_int, RlrState, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_void_p]
    lib.rlr_step.argtypes = [P, C.c_int, C.c_int, RlrState, C.c_void_p, C.c_int, C.c_void_p, RlrStepOut, C.c_int64, C.c_int64,
                             C.c_uint64, C.c_void_p]
    lib.rlr_step_range.argtypes = [P, C.c_int, C.c_int, RlrState, C.c_void_p, C.c_int, C.c_void_p, RlrStepOut, C.c_int64, C.c_int64, C.c_int64,
                                   C.c_int64, C.c_uint64, C.c_void_p]
    lib.rlr_step_chained.argtypes = [P, C.c_int, C.c_int, RlrState, C.c_void_p, C.c_int, C.c_void_p, RlrStepOut, C.c_int64, C.c_int64,
                                     C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.rlr_rollout.argtypes = [P, C.c_int, C.c_int, RlrState, C.c_void_p, C.c_int, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_void_p]
    lib.rlr_pipe_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.rlr_pipe_destroy.argtypes = [C.c_void_p]
    lib.rlr_pipe_submit.argtypes = [C.c_void_p, C.c_int, P, C.c_int, C.c_int, RlrState, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, RlrStepOut,
                                    C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64, C.c_int64, C.c_uint64, C.c_void_p]
    lib.rlr_pipe_wait.argtypes = [C.c_void_p, C.c_int]
    lib.rlr_exchange_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int]
    lib.rlr_exchange_handle.argtypes = [C.c_void_p, C.c_void_p]
    lib.rlr_exchange_connect.argtypes = [C.c_void_p, C.c_void_p]
    lib.rlr_exchange_allgather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.rlr_exchange_status.argtypes = [C.c_void_p]
    lib.rlr_exchange_destroy.argtypes = [C.c_void_p]
    lib.rlr_fma_peak.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError(f"librlrocket error {rc}: {load().rlr_last_error().decode()}")
