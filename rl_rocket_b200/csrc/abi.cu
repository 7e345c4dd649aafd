// abi.cu -- the extern "C" boundary declared in include/rl_rocket_b200.h.
// No torch types, no allocation, no synchronisation: raw device pointers in, kernels enqueued on the caller's stream.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "rl_rocket_b200.h"

#define RLR_DECLARE_LAUNCHERS(NS)                                                                                                          \
    namespace NS {                                                                                                                         \
    int launch_step(const RlrParams*, RlrState, const void*, int, const float*, RlrStepOut, int64_t, int64_t, int64_t, int64_t, uint64_t,   \
                    cudaStream_t, uint32_t*, uint32_t, uint32_t);                                                                           \
    int launch_reset(const RlrParams*, RlrState, const uint8_t*, const float*, float*, int64_t, int64_t, uint64_t, cudaStream_t);         \
    int launch_rollout(const RlrParams*, RlrState, const void*, int, int, float*, float*, uint8_t*, uint8_t*, float*, unsigned long long*, \
                       int64_t, int64_t, uint64_t, cudaStream_t);                                                                          \
    }
RLR_DECLARE_LAUNCHERS(rlr_exact)
RLR_DECLARE_LAUNCHERS(rlr_fast)

namespace rlr_fast {      // RLR_F32 precision (6DOF, fast math only)
int launch_step_f32(const RlrParams*, RlrState, const void*, int, const float*, RlrStepOut, int64_t, int64_t, int64_t, int64_t, uint64_t,
                    cudaStream_t);
int launch_reset_f32(const RlrParams*, RlrState, const uint8_t*, const float*, float*, int64_t, int64_t, uint64_t, cudaStream_t);
int launch_rollout_f32(const RlrParams*, RlrState, const void*, int, int, float*, float*, uint8_t*, uint8_t*, float*, unsigned long long*,
                       int64_t, int64_t, uint64_t, cudaStream_t);
}

namespace rlr_peak {
int launch_fma_peak(int precision, int blocks, int threads, int iters, double* sink, double* flops_out, cudaStream_t stream);
}

static thread_local char g_err[256] = "";

static int fail(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
static int cuda_status(int err) {
    if (err == 0) return RLR_OK;
    snprintf(g_err, sizeof(g_err), "CUDA error %d: %s", err, cudaGetErrorString((cudaError_t)err));
    return RLR_E_CUDA;
}
static int check_params(const RlrParams* p) {
    if (p == nullptr) return fail(RLR_E_INVALID, "params is NULL");
    if (p->dof != 6 && p->dof != 3 && p->dof != 1) return fail(RLR_E_INVALID, "dof must be 6, 3 or 1");
    if (p->dt_ms <= 0 || !(p->dt > 0)) return fail(RLR_E_INVALID, "timestep must be a positive whole number of milliseconds");
    if (p->n_discrete < 0 || p->n_discrete > RLR_MAX_DISCRETE) return fail(RLR_E_INVALID, "n_discrete out of range");
    if (p->obs_mode != RLR_OBS_STATE && p->obs_mode != RLR_OBS_GAUDET) return fail(RLR_E_INVALID, "bad obs_mode");
    if (!(p->min_thrust >= 0.0f) || p->min_thrust > p->max_thrust) return fail(RLR_E_INVALID, "min_thrust must lie in [0, max_thrust]");
    if (p->aero_mode != 0 && p->aero_mode != 1) return fail(RLR_E_INVALID, "bad aero_mode");
    if (p->aero_mode != 0 && p->dof == 6) return fail(RLR_E_UNSUPPORTED, "aero_mode is defined for the 3DOF / 1DOF environment (the 6DOF aerodynamic force of the reference is a zero stub, simulator.py:359-360)");
    if (p->obs_mode == RLR_OBS_GAUDET && p->dof != 3) return fail(RLR_E_UNSUPPORTED, "RLR_OBS_GAUDET is defined for the 3DOF environment (wrappers.py:52)");
    return RLR_OK;
}
static int check_precision(const RlrParams* p, int precision, int math_mode) {
    if (precision == RLR_F64) return RLR_OK;
    if (precision != RLR_F32) return fail(RLR_E_INVALID, "bad precision");
    if (p->dof != 6) return fail(RLR_E_UNSUPPORTED, "RLR_F32 state is implemented for the 6DOF environment only");
    if (math_mode != RLR_MATH_FAST) return fail(RLR_E_UNSUPPORTED, "RLR_F32 state needs RLR_MATH_FAST (the exact build is the op-for-op fp64 reference)");
    return RLR_OK;
}
static int check_state(const RlrState& st) {
    if (!st.y || !st.k || !st.episode || !st.v0 || !st.ep_ret) return fail(RLR_E_INVALID, "RlrState has a NULL pointer");
    return RLR_OK;
}

extern "C" int rlr_set_error_(int code, const char* msg) { return fail(code, msg); }      // shared with host_pipe.cu (not in the public header)

extern "C" {

int rlr_abi_version(void) { return RLR_ABI_VERSION; }
const char* rlr_last_error(void) { return g_err; }
int rlr_sizeof_params(void) { return (int)sizeof(RlrParams); }
int rlr_ndim(const RlrParams* p) { return p->dof == 6 ? 14 : 7; }
int rlr_nact(const RlrParams* p) { return p->dof == 6 ? 3 : (p->dof == 3 ? 2 : 1); }
int rlr_nobs(const RlrParams* p) { return p->dof == 6 ? 14 : (p->dof == 3 ? (p->obs_mode == RLR_OBS_GAUDET ? 4 : 7) : 2); }

int rlr_reset(const RlrParams* p, int precision, RlrState st, const uint8_t* mask, const float* ic_inject, float* obs, int64_t n,
              int64_t env_offset, uint64_t seed, void* stream) {
    int rc = check_params(p);
    if (rc) return rc;
    if (n < 0) return fail(RLR_E_INVALID, "n < 0");
    if (n == 0) return RLR_OK;            /* empty batch: nothing to do (zero-size buffers may be NULL) */
    if ((rc = check_state(st))) return rc;
    if ((rc = check_precision(p, precision, RLR_MATH_FAST))) return rc;
    if (precision == RLR_F32) return cuda_status(rlr_fast::launch_reset_f32(p, st, mask, ic_inject, obs, n, env_offset, seed, (cudaStream_t)stream));
    return cuda_status(rlr_exact::launch_reset(p, st, mask, ic_inject, obs, n, env_offset, seed, (cudaStream_t)stream));
}

int rlr_step_range(const RlrParams* p, int precision, int math_mode, RlrState st, const void* actions, int action_kind, const float* ic_inject,
                   RlrStepOut out, int64_t n, int64_t begin, int64_t count, int64_t env_offset, uint64_t seed, void* stream) {
    int rc = check_params(p);
    if (rc) return rc;
    if (n < 0 || begin < 0 || count < 0 || begin + count > n) return fail(RLR_E_INVALID, "need 0 <= begin, 0 <= count, begin + count <= n");
    if (count == 0) return RLR_OK;            /* empty batch / empty range: nothing to do (zero-size buffers may be NULL) */
    if ((rc = check_state(st))) return rc;
    if ((rc = check_precision(p, precision, math_mode))) return rc;
    if (!actions || !out.obs || !out.done) return fail(RLR_E_INVALID, "actions, out.obs and out.done are required");
    if ((out.fin_records == nullptr) != (out.fin_count == nullptr)) return fail(RLR_E_INVALID, "out.fin_count and out.fin_records go together");
    if (action_kind < 0 || action_kind > 2) return fail(RLR_E_INVALID, "bad action_kind");
    if (action_kind == RLR_ACT_DISCRETE && (p->dof != 3 || p->n_discrete <= 0))
        return fail(RLR_E_INVALID, "discrete actions need dof == 3 and a discrete_table");
    cudaStream_t s = (cudaStream_t)stream;
    if (precision == RLR_F32)
        return cuda_status(rlr_fast::launch_step_f32(p, st, actions, action_kind, ic_inject, out, n, begin, count, env_offset, seed, s));
    if (math_mode == RLR_MATH_EXACT)
        return cuda_status(rlr_exact::launch_step(p, st, actions, action_kind, ic_inject, out, n, begin, count, env_offset, seed, s, nullptr, 0u, 0u));
    if (math_mode == RLR_MATH_FAST)
        return cuda_status(rlr_fast::launch_step(p, st, actions, action_kind, ic_inject, out, n, begin, count, env_offset, seed, s, nullptr, 0u, 0u));
    return fail(RLR_E_INVALID, "bad math_mode");
}

int rlr_step_chained(const RlrParams* p, int precision, int math_mode, RlrState st, const void* actions, int action_kind, const float* ic_inject,
                     RlrStepOut out, int64_t n, int64_t env_offset, uint64_t seed, uint32_t* block_seq, uint32_t seq_wait, uint32_t seq_publish,
                     void* stream) {
    int rc = check_params(p);
    if (rc) return rc;
    if (n < 0) return fail(RLR_E_INVALID, "n < 0");
    if (p->dof != 6 || precision != RLR_F64) return fail(RLR_E_UNSUPPORTED, "rlr_step_chained is implemented for the 6DOF environment with RLR_F64 state");
    if (!block_seq || seq_publish == 0u || seq_publish == seq_wait) return fail(RLR_E_INVALID, "rlr_step_chained: block_seq, and a non-zero seq_publish different from seq_wait, are required");
    if (n == 0) return RLR_OK;
    if ((rc = check_state(st))) return rc;
    if (!actions || !out.obs || !out.done) return fail(RLR_E_INVALID, "actions, out.obs and out.done are required");
    if ((out.fin_records == nullptr) != (out.fin_count == nullptr)) return fail(RLR_E_INVALID, "out.fin_count and out.fin_records go together");
    if (action_kind != RLR_ACT_NORMALIZED && action_kind != RLR_ACT_PHYSICAL) return fail(RLR_E_INVALID, "bad action_kind");
    cudaStream_t s = (cudaStream_t)stream;
    if (math_mode == RLR_MATH_EXACT)
        return cuda_status(rlr_exact::launch_step(p, st, actions, action_kind, ic_inject, out, n, 0, n, env_offset, seed, s, block_seq, seq_wait, seq_publish));
    if (math_mode == RLR_MATH_FAST)
        return cuda_status(rlr_fast::launch_step(p, st, actions, action_kind, ic_inject, out, n, 0, n, env_offset, seed, s, block_seq, seq_wait, seq_publish));
    return fail(RLR_E_INVALID, "bad math_mode");
}

int rlr_step(const RlrParams* p, int precision, int math_mode, RlrState st, const void* actions, int action_kind, const float* ic_inject,
             RlrStepOut out, int64_t n, int64_t env_offset, uint64_t seed, void* stream) {
    if (n < 0) return fail(RLR_E_INVALID, "n < 0");
    return rlr_step_range(p, precision, math_mode, st, actions, action_kind, ic_inject, out, n, 0, n, env_offset, seed, stream);
}

int rlr_rollout(const RlrParams* p, int precision, int math_mode, RlrState st, const void* actions, int action_kind, int32_t T, float* obs_buf,
                float* reward_buf, uint8_t* done_buf, uint8_t* trunc_buf, float* last_obs, unsigned long long* stats, int64_t n,
                int64_t env_offset, uint64_t seed, void* stream) {
    int rc = check_params(p);
    if (rc) return rc;
    if (n < 0 || T < 0) return fail(RLR_E_INVALID, "n < 0 or T < 0");
    if (n == 0 || T == 0) return RLR_OK;
    if ((rc = check_state(st))) return rc;
    if ((rc = check_precision(p, precision, math_mode))) return rc;
    if (!actions) return fail(RLR_E_INVALID, "actions is NULL");
    if (action_kind < 0 || action_kind > 2) return fail(RLR_E_INVALID, "bad action_kind");
    if (action_kind == RLR_ACT_DISCRETE && (p->dof != 3 || p->n_discrete <= 0))
        return fail(RLR_E_INVALID, "discrete actions need dof == 3 and a discrete_table");
    cudaStream_t s = (cudaStream_t)stream;
    if (precision == RLR_F32)
        return cuda_status(rlr_fast::launch_rollout_f32(p, st, actions, action_kind, T, obs_buf, reward_buf, done_buf, trunc_buf, last_obs, stats, n,
                                                        env_offset, seed, s));
    if (math_mode == RLR_MATH_EXACT)
        return cuda_status(rlr_exact::launch_rollout(p, st, actions, action_kind, T, obs_buf, reward_buf, done_buf, trunc_buf, last_obs, stats, n,
                                                     env_offset, seed, s));
    if (math_mode == RLR_MATH_FAST)
        return cuda_status(rlr_fast::launch_rollout(p, st, actions, action_kind, T, obs_buf, reward_buf, done_buf, trunc_buf, last_obs, stats, n,
                                                    env_offset, seed, s));
    return fail(RLR_E_INVALID, "bad math_mode");
}

int rlr_fma_peak(int precision, int blocks, int threads, int iters, double* sink, double* flops_out, void* stream) {
    if (!sink || blocks <= 0 || threads <= 0 || threads > 1024 || iters <= 0) return fail(RLR_E_INVALID, "bad fma_peak arguments");
    return cuda_status(rlr_peak::launch_fma_peak(precision, blocks, threads, iters, sink, flops_out, (cudaStream_t)stream));
}

}  // extern "C"
