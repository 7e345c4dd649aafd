/*
 * rl_rocket_b200.h -- C ABI of the B200-native batched RL_rocket environment step.
 *
 * One shared library (rl_rocket_b200/_C/librlrocket.so, built by nvcc for sm_100a) exports the
 * entry points below.  They are what a binding of the reference's environment layer would call
 * instead of its Python/scipy path; INTEGRATION.md shows the ctypes stub.
 *
 * Reference interfaces replaced (paths relative to Tuxliri/RL_rocket):
 *   rlr_reset      Rocket6DOF.reset  my_environment/envs/rocket_env.py:665-688   (+ Simulator6DOF.__init__ utils/simulator.py:180-225)
 *                  Rocket.reset      my_environment/envs/rocket_env.py:137-148   (+ Simulator3DOF.__init__ utils/simulator.py:13-53)
 *   rlr_step       Rocket6DOF.step   my_environment/envs/rocket_env.py:690-719   -> Simulator6DOF.step utils/simulator.py:227-257
 *                  Rocket.step       my_environment/envs/rocket_env.py:150-175   -> Simulator3DOF.step utils/simulator.py:55-86
 *                  + gym TimeLimit.step / SB3 DummyVecEnv.step_wait semantics (main_6DOF.py:21-22,71; main.py:65-70,105)
 *                  + wrappers DiscreteActions3DOF / RewardAnnealing (my_environment/wrappers/wrappers.py:24-35,68-86)
 *   rlr_step_range the same step restricted to the environments [begin, begin + count) of the batch (sub-batch pipelining:
 *                  several ranges on different streams, so one range's PCIe transfers overlap another's integration)
 *   rlr_step_chained  rlr_step for open-loop action sources (the reference's own rollouts step fixed / sampled actions,
 *                  test_6DOF_environment.py:19-27): consecutive launches ordered block by block instead of grid by grid
 *   rlr_rollout    T consecutive rlr_step calls fused in one launch (state stays in registers)
 *   RlrStepOut.stats / fin_ret / fin_len   the per-episode bookkeeping SB3's Monitor does on the host (main.py:65-70),
 *                  accumulated on the device by the step / rollout kernels (there is no separate statistics entry point)
 *   RlrParams.obs_mode = RLR_OBS_GAUDET    the GaudetStateObs observation wrapper (my_environment/wrappers/wrappers.py:51-66)
 *
 * Conventions
 *   - Every function returns 0 on success or a negative RLR_E_* code; rlr_last_error() gives a
 *     thread-local message.  Numerical failure of the integrator is NOT an error: as in the
 *     reference (rocket_env.py:702, bool(status)) it terminates the episode.
 *   - All pointers are DEVICE pointers into caller-owned memory (PyTorch tensors' data_ptr());
 *     the library never allocates, frees or synchronises.  Work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - State is structure-of-arrays: y[c * n + i] for component c of environment i, n = number of
 *     environments in this call.  Observations / actions are row-major [n][dim] as SB3 expects.
 */
#ifndef RL_ROCKET_B200_H
#define RL_ROCKET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLR_ABI_VERSION 6
#define RLR_NMAX 14          /* state components: 14 (6DOF), 7 (3DOF / 1DOF) */
#define RLR_MAX_DISCRETE 16

enum { RLR_OK = 0, RLR_E_INVALID = -1, RLR_E_CUDA = -2, RLR_E_UNSUPPORTED = -3 };

/* action_kind */
enum {
    RLR_ACT_NORMALIZED = 0,  /* f32 [n][3|2|1] in [-1,1]; mapped as _denormalize_action (rocket_env.py:969-981 / :395-406) */
    RLR_ACT_PHYSICAL   = 1,  /* f32 [n][3|2]  already de-normalised (gimbal [rad] ..., thrust [N]) */
    RLR_ACT_DISCRETE   = 2   /* i64 [n] index into RlrParams.discrete_table (DiscreteActions3DOF, wrappers.py:24-35) */
};

/* math_mode */
enum {
    RLR_MATH_EXACT = 0,      /* IEEE op-for-op sequence of the reference (no FMA contraction, true divisions) */
    RLR_MATH_FAST  = 1       /* FMA contraction + reciprocal multiplies; deviation bounded in tests (<= 1e-9 normalised) */
};

/* obs_mode */
enum {
    RLR_OBS_STATE  = 0,      /* normalised state (rocket_env.py:963-964 / :209-210); 1DOF: (z, vz) */
    RLR_OBS_GAUDET = 1       /* 3DOF only: f32 [vx - vtarg_x, vz - vtarg_z, t_go, z] of the PHYSICAL state -- GaudetStateObs
                                (wrappers.py:51-66) with its two defects fixed: it calls the existing _compute_vtarg
                                (rocket_env.py:219-247) and feeds it the un-normalised r, v the method is written for */
};

/* precision of the integrated state */
enum { RLR_F64 = 0, RLR_F32 = 1 };

/* Host-side POD, copied by value into kernel parameters (< 1 KB). Filled by rl_rocket_b200/config.py. */
typedef struct RlrParams {
    int32_t dof;                 /* 6, 3 or 1 (1DOF toy problem of README.md:4-8: 3DOF with gimbal 0, obs (z, vz)) */
    int32_t reward_mode;         /* 0 shaped reward; 1 RewardAnnealing */
    int32_t max_episode_steps;   /* TimeLimit; <= 0 disables */
    int32_t dt_ms;               /* timestep in whole milliseconds: t_k = k*dt_ms/1000 (simulator.py:245 rounds t to 3 decimals) */
    double  dt;
    double  normalizer[RLR_NMAX];
    double  inv_normalizer[RLR_NMAX];              /* 1 / normalizer, used by RLR_MATH_FAST for the observation */
    float   ic_low[RLR_NMAX], ic_high[RLR_NMAX];   /* init_space Box (rocket_env.py:564-567 / :57-60) */
    float   pos_low[3], pos_high[3];               /* 6DOF position_bounds_space (rocket_env.py:615-620) */
    double  x_left, x_right, y_up;                 /* 3DOF bounds (rocket_env.py:97-100) */
    double  alfa, beta, eta, gamma, delta, kappa, xi;
    double  waypoint, landing_radius, maximum_velocity;
    double  att_traj_limit[3], att_land_limit[3], omega_lim[3];
    double  att_traj_c[3], att_land_c[3];          /* cos(limit) for the two atan2 angles, sin(limit) for the asin angle (RLR_MATH_FAST) */
    double  max_gimbal;
    float   max_thrust;
    int32_t n_discrete;
    float   discrete_table[2 * RLR_MAX_DISCRETE];  /* physical (gimbal, thrust) pairs */
    float   discrete_thr[RLR_MAX_DISCRETE];        /* the rows' normalised throttle (row[1], an integer in the reference's table): what
                                                      RewardAnnealing sees under DiscreteActions3DOF, -xi * (row[1] + 1) in float64 */
    int32_t obs_mode;                              /* RLR_OBS_* */
    float   min_thrust;                            /* lower bound applied to the de-normalised thrust [N] (the reference's open TODO at
                                                      rocket_env.py:405, :980); 0 = off = the reference's behaviour */
    int32_t aero_mode;                             /* 3DOF / 1DOF: 0 = the reference's behaviour (alfa = 0, simulator.py:103); 1 = the
                                                      angle-of-attack line the reference keeps commented out (simulator.py:104) enabled
                                                      -- alfa = wrap_(-pi,pi](phi - atan2(vz, vx)), Cn = Cnalfa alfa, N = Cn Q Sref, and the
                                                      N terms of the reference's own ax / ay / dom (simulator.py:107-127) */
} RlrParams;

/* Persistent per-environment state (device pointers, caller-owned). */
typedef struct RlrState {
    void*    y;          /* f64 (RLR_F64) or f32 (RLR_F32) [ndim][n] */
    int32_t* k;          /* [n] steps taken in the current episode */
    int32_t* episode;    /* [n] episode counter (keys the reset RNG) */
    float*   v0;         /* [n] |v| of the episode's initial condition (vtarg, rocket_env.py:989-991) */
    double*  ep_ret;     /* [n] running undiscounted return */
} RlrState;

/* Per-call outputs; any pointer may be NULL except obs, done. */
typedef struct RlrStepOut {
    float*   obs;            /* [n][nobs]  observation after the step (after auto-reset for finished envs) */
    float*   reward;         /* [n] f32 (what SB3 stores) */
    double*  reward64;       /* [n] f64 (what Rocket*.step returns) */
    uint8_t* done;           /* [n] */
    uint8_t* truncated;      /* [n] info["TimeLimit.truncated"] */
    float*   terminal_obs;   /* [n][nobs]  info["terminal_observation"], written for finished envs only */
    double*  fin_ret;        /* [n] info["episode"]["r"] for finished envs */
    int32_t* fin_len;        /* [n] info["episode"]["l"] for finished envs */
    uint8_t* flags;          /* [n] bit0 bounds_violation, bit1 landed, bit2 ground event, bit3 integrator failed */
    int32_t* natt;           /* [n] RK45 attempts spent in this step (nfev = 2 + 6*natt) */
    void*    y_pre_reset;    /* [ndim][n] state after the step, before auto-reset (same dtype as y) */
    double*  reward_terms;   /* [6][n] velocity_tracking, thrust_penalty, eta, attitude_constraint, attitude_hint, rew_goal */
    double*  vtarg;          /* [4][n] target velocity (3 components; 2 for 3DOF, third = 0) and t_go of _compute_vtarg
                                (rocket_env.py:986-1014 / :215-247) -- what the reference appends to vtarg_history */
    unsigned long long* stats; /* [RLR_NSTATS] device accumulators, atomically added (see below) */
    /* Compact queue of the episodes that finished in this step -- what Monitor / DummyVecEnv put into `infos` (main.py:65-70, SB3
     * dummy_vec_env.step_wait): the step appends one record per finished environment, so the host reads `fin_count` records instead
     * of gathering five [n]-sized arrays.  Record = RLR_FIN_HEADER + nobs doubles: [env index, episode return, episode length,
     * TimeLimit.truncated, flags (as above), terminal_observation[nobs]].  Record j of a step goes to slot (count_before + j) mod n: the
     * caller either zeroes *fin_count before the step or remembers the previous count; order within a step is arbitrary.
     * Both NULL = not recorded. */
    unsigned int* fin_count;   /* device counter */
    double*  fin_records;      /* [n][RLR_FIN_HEADER + nobs] */
} RlrStepOut;
#define RLR_FIN_HEADER 5

/* stats slots (all u64; returns / lengths are fixed point) */
enum {
    RLR_STAT_STEPS = 0,      /* env steps executed */
    RLR_STAT_ATTEMPTS = 1,   /* RK45 attempts (for the flop count: 540 + 1530*k per 6DOF step) */
    RLR_STAT_DONE = 2,       /* finished episodes */
    RLR_STAT_LANDED = 3,
    RLR_STAT_VIOL = 4,
    RLR_STAT_TRUNC = 5,
    RLR_STAT_LEN_SUM = 6,    /* sum of finished-episode lengths */
    RLR_STAT_RET_SUM = 7,    /* sum of finished-episode returns, as (int64)(ret * 2^20) two's complement */
    RLR_NSTATS = 8
};

int         rlr_abi_version(void);
const char* rlr_last_error(void);
int         rlr_sizeof_params(void);

/* Number of state / action / observation components for a params block. */
int rlr_ndim(const RlrParams* p);
int rlr_nact(const RlrParams* p);
int rlr_nobs(const RlrParams* p);

/* reset: (re)initialise environments.  mask NULL = all; ic_inject NULL = Philox IC keyed by
 * (seed, env_offset + i, episode[i]); otherwise f32 [n][ndim] initial conditions to use verbatim. */
int rlr_reset(const RlrParams* p, int precision, RlrState st, const uint8_t* mask, const float* ic_inject,
              float* obs, int64_t n, int64_t env_offset, uint64_t seed, void* stream);

/* one environment step for n environments, TimeLimit + auto-reset fused */
int rlr_step(const RlrParams* p, int precision, int math_mode, RlrState st, const void* actions, int action_kind,
             const float* ic_inject, RlrStepOut out, int64_t n, int64_t env_offset, uint64_t seed, void* stream);

/* rlr_step for the environments [begin, begin + count) only; every per-environment pointer (state, actions, outputs) is still
 * indexed by the absolute environment index and n stays the batch size (the structure-of-arrays stride).  Ranges are independent:
 * launch them on different streams to pipeline one range's host transfers behind another's integration. */
int rlr_step_range(const RlrParams* p, int precision, int math_mode, RlrState st, const void* actions, int action_kind,
                   const float* ic_inject, RlrStepOut out, int64_t n, int64_t begin, int64_t count, int64_t env_offset,
                   uint64_t seed, void* stream);

/* Open-loop chaining of consecutive steps (6DOF, RLR_F64): rlr_step, except that block b of the launch orders itself after block b
 * of the PREVIOUS chained launch on the same state only -- not after everything enqueued before it -- so the tail of one step (the
 * last third of a launch, when a few warps with a third RK45 attempt / a ground event / resets are still running) overlaps the
 * body of the next.  For action sources that do not depend on the previous step's results: random / scripted / repeated actions,
 * the reference's own rollouts (test_6DOF_environment.py:19-27 steps fixed actions), several environment groups stepped in turn.
 *   block_seq   device u32 [ceil(n / 64)], zero-initialised once, owned by the caller together with the state; one entry per block
 *   seq_wait    0: this launch waits for everything enqueued before it (start of a chain, or after ANY other operation touched the
 *               state or the step's outputs); otherwise the seq_publish of the previous chained launch on this state
 *   seq_publish non-zero, different from seq_wait: stored to block_seq[b] (release) once block b's results are in memory
 * The caller guarantees that `actions` / `ic_inject` are not produced by work still pending on the stream, and that nothing else on
 * the stream touches the state or `out` between two launches of one chain.  Results: bit-identical to rlr_step's with RLR_MATH_EXACT;
 * with RLR_MATH_FAST the chained kernel is a separate compilation (FMA contraction may differ in the last bit: <= 1e-15 normalised),
 * flags / counters / attempt counts identical. */
int rlr_step_chained(const RlrParams* p, int precision, int math_mode, RlrState st, const void* actions, int action_kind,
                     const float* ic_inject, RlrStepOut out, int64_t n, int64_t env_offset, uint64_t seed,
                     uint32_t* block_seq, uint32_t seq_wait, uint32_t seq_publish, void* stream);

/* T fused steps.  actions: [T][n][nact] (or i64 [T][n] for discrete).  Outputs, when non-NULL, are
 * [T][n][...] rollout buffers (obs/reward/done/truncated) -- the layout SB3's RolloutBuffer uses. */
int rlr_rollout(const RlrParams* p, int precision, int math_mode, RlrState st, const void* actions, int action_kind,
                int32_t T, float* obs_buf, float* reward_buf, uint8_t* done_buf, uint8_t* trunc_buf, float* last_obs,
                unsigned long long* stats, int64_t n, int64_t env_offset, uint64_t seed, void* stream);

/* Host-buffer pipeline: what DummyVecEnv.step_wait does on the host for one step (main_6DOF.py:71, main.py:105) -- take the actions
 * from host memory, step, hand obs / reward / done back in host memory -- as ONE call that does not block, with `depth` slots so that
 * the PCIe transfers of one step overlap the integration of the next (the step kernels of consecutive submits still run in order on
 * compute_stream: the state dependency between steps is kept on the device).
 *   rlr_pipe_submit(slot): H2D copy of h_actions (pinned) into d_actions on the pipe's copy stream; rlr_step on compute_stream with
 *       `out` (its obs / reward / done must point into d_pack: obs at 0, reward at n*nobs*4, done at n*nobs*4 + n*4); D2H copy of
 *       d_pack into h_pack (pinned) on the pipe's second copy stream.  The slot's buffers (h_actions, d_actions, d_pack, h_pack) are
 *       caller-owned and must stay untouched until rlr_pipe_wait(slot) returns; a slot is resubmitted only after it was waited for.
 *   rlr_pipe_wait(slot): blocks the host until the slot's result is in h_pack.
 * Create / destroy on the device the environments live on. */
typedef struct RlrHostPipe RlrHostPipe;
int rlr_pipe_create(RlrHostPipe** pipe, int depth);
int rlr_pipe_destroy(RlrHostPipe* pipe);
int rlr_pipe_submit(RlrHostPipe* pipe, int slot, const RlrParams* p, int precision, int math_mode, RlrState st, const void* h_actions,
                    void* d_actions, size_t action_bytes, int action_kind, RlrStepOut out, const void* d_pack, void* h_pack,
                    size_t pack_bytes, int64_t n, int64_t env_offset, uint64_t seed, void* compute_stream);
int rlr_pipe_wait(RlrHostPipe* pipe, int slot);

/* Per-rollout exchange of the episode-statistics records (returns / done counts: RLR_NSTATS u64 per rank) between the GPUs of ONE
 * node over NVLink peer memory -- the one exchange step of the multi-GPU path (environments shard by index; there is no data-path
 * collective).  Each rank creates an exchange on its device, the 64-byte CUDA-IPC handles are all-gathered once by the host side
 * (torch.distributed, any backend), and rlr_exchange_allgather then enqueues ONE single-block kernel that writes the rank's record
 * into every peer's table, flags it, waits for the peers' flags and copies the [world][RLR_NSTATS] table to `out` (device).
 * Every rank must call it the same number of times.  rlr_exchange_status (host-synchronising) reports a peer that never arrived. */
typedef struct RlrExchange RlrExchange;
int rlr_exchange_create(RlrExchange** ex, int rank, int world);
int rlr_exchange_handle(RlrExchange* ex, unsigned char* out64);
int rlr_exchange_connect(RlrExchange* ex, const unsigned char* handles /* [world][64] */);
int rlr_exchange_allgather(RlrExchange* ex, const unsigned long long* local_stats, unsigned long long* out, void* stream);
int rlr_exchange_status(RlrExchange* ex);
int rlr_exchange_destroy(RlrExchange* ex);

/* FP64 / FP32 FMA roofline calibration: launches a dependent-chain FMA kernel, returns flops issued. */
int rlr_fma_peak(int precision, int blocks, int threads, int iters, double* sink, double* flops_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RL_ROCKET_B200_H */
