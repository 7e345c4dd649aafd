// rocket_kernels.cuh -- sm_100a device code of the batched RL_rocket environment step.
//
// Included by kernels_exact.cu (nvcc -fmad=false, RLR_EXACT=1) and kernels_fast.cu (FMA
// contraction on, RLR_EXACT=0); each translation unit wraps it in its own namespace RLR_NS.
//
// One thread = one environment.  The state vector lives in registers from the load at the top
// of the kernel to the store at the bottom, across the whole adaptive RK45 solve (scipy
// solve_ivp semantics: initial-step heuristic, error-controlled sub-steps, FSAL, terminal
// ground event located with Brent's method on the quartic dense output), the reward, the
// TimeLimit and the auto-reset.
//
// Reference call path reproduced (Tuxliri/RL_rocket):
//   Rocket6DOF.step  my_environment/envs/rocket_env.py:690-719  -> Simulator6DOF.step utils/simulator.py:227-257
//   Rocket.step      my_environment/envs/rocket_env.py:150-175  -> Simulator3DOF.step utils/simulator.py:55-86
//   solve_ivp/RK45   scipy/integrate/_ivp/{ivp.py:645-700, rk.py:14-176,541-565,715-737, common.py:63-134}
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <mutex>

#include "rl_rocket_b200.h"

#ifndef RLR_REGS6
#define RLR_REGS6 128        // registers are granted per warp in units of 1024, i.e. 96 / 128 / 160 per thread: 128 lets 7 blocks
                             // x 64 threads (shared memory: 7 x 31 KB) be resident = 448 threads/SM, which holds all 65 536
                             // envs of BASELINE config 3 in ONE wave on 148 SMs (144 registers would drop to 6 blocks/SM)
#endif
// register cap of the step / rollout kernels: S::MAX_REGS registers per thread (see Sys6 / Sys3)
#define RLR_KERNEL_BOUNDS(S) __maxnreg__(S::MAX_REGS)
#ifndef RLR_NS
#error "define RLR_NS before including rocket_kernels.cuh"
#endif

namespace RLR_NS {

#define RLR_DEV __device__ __forceinline__

// Development aid (tuning builds only, -DRLR_PHASE_TIMING=1): per-warp clock64() deltas between phase marks, summed / maxed per phase.
#ifndef RLR_PHASE_TIMING
#define RLR_PHASE_TIMING 0
#endif
#ifndef RLR_WARP_TRACE
#define RLR_WARP_TRACE 0
#endif
#if RLR_PHASE_TIMING
__device__ unsigned long long g_phase_sum[16], g_phase_max[16], g_phase_cnt[16];
RLR_DEV void phase_mark(int k) {
    __shared__ long long tp[8];
    if ((threadIdx.x & 31) == 0) {
        const long long now = clock64();
        if (k >= 0) {
            const unsigned long long d = (unsigned long long)(now - tp[threadIdx.x >> 5]);
            atomicAdd(&g_phase_sum[k], d); atomicMax(&g_phase_max[k], d); atomicAdd(&g_phase_cnt[k], 1ull);
        }
        tp[threadIdx.x >> 5] = now;
    }
}
#define RLR_PHASE(k) phase_mark(k)
#elif defined(RLR_WARP_TRACE) && RLR_WARP_TRACE
#define RLR_PHASE(k) trace_mark(k)
RLR_DEV void trace_mark(int k);
#else
#define RLR_PHASE(k)
#endif
// Development aid (-DRLR_WARP_TRACE=1): one record per warp of the last step launch -- SM id, global-timer start / end, and what
// the warp carried (most RK45 attempts of a lane, ground event, finished lanes) -- to see which warps make the tail of the launch.
// tuning switches (A/B builds): 1 = stage 6 peeled and fused with the error / quadrature sums; 1 = dead scalars parked in shared memory
#ifndef RLR_FUSED_ERR
#define RLR_FUSED_ERR 1
#endif
#ifndef RLR_PARK
#define RLR_PARK 1
#endif
#ifndef RLR_PREFETCH
#define RLR_PREFETCH 1
#endif
#ifndef RLR_OPAQUE_STORE_INDEX
#define RLR_OPAQUE_STORE_INDEX 1
#endif
#ifndef RLR_WARP_TRACE
#define RLR_WARP_TRACE 0
#endif
#if RLR_WARP_TRACE
__device__ unsigned long long g_warp_trace[8192 * 16];      // per warp: sm, t0, t1, flags, then the (last) arrival time at phase marks 0..9
RLR_DEV unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
RLR_DEV unsigned smid_() { unsigned s; asm volatile("mov.u32 %0, %smid;" : "=r"(s)); return s; }
RLR_DEV void trace_mark(int k) {        // the lowest active lane of whichever part of the warp passes the mark; the last part wins (max)
    const unsigned am = __activemask();
    const unsigned long long w = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if ((int)(threadIdx.x & 31) == __ffs(am) - 1 && w < 8192) {
        const unsigned long long now = globaltimer_ns();
        if (k < 0) { g_warp_trace[w * 16 + 1] = now; for (int j = 4; j < 16; ++j) g_warp_trace[w * 16 + j] = 0ull; }
        else *(volatile unsigned long long*)&g_warp_trace[w * 16 + 4 + k] = now;      // absolute; the host subtracts the warp's start
    }
}
#endif

// Programmatic dependent launch (sm_90+): a step kernel lets the next kernel of the stream begin launching at once and itself waits
// for its predecessor's completion only right before its first global access, so consecutive step launches (action repeat, graph
// replays, random-action loops) overlap the launch latency and block placement of step t+1 with the tail of step t.
RLR_DEV void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
RLR_DEV void pdl_wait_for_predecessor() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Open-loop chaining (rlr_step_chained): block b of a launch orders itself after block b of the previous chained launch on the same
// state -- the only launch-to-launch dependency a step has, since block b works on the same environments in every launch -- through a
// per-block sequence word, instead of waiting for the predecessor's whole grid.  wait = 0: the ordinary grid-wide wait.
struct StepChain { uint32_t* block_seq; uint32_t wait, publish; };
RLR_DEV void chain_wait(const StepChain& ch) {
    if (ch.wait == 0u) { pdl_wait_for_predecessor(); return; }
    if (threadIdx.x == 0) {
        const uint32_t* f = ch.block_seq + blockIdx.x;
        uint32_t v;
        for (int spins = 0; spins < (1 << 22); ++spins) {         // bounded (~2 s): a broken chain must not hang the GPU
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v == ch.wait) break;
            __nanosleep(40);
        }
    }
    __syncthreads();
}
RLR_DEV void chain_publish(const StepChain& ch) {
    if (ch.publish == 0u) return;
    // The pattern of cooperative groups' grid.sync(): block barrier, then ONE thread fences and publishes.  The barrier orders every
    // thread's result stores before thread 0's fence + release store, and a release is cumulative over what the releasing thread has
    // synchronised with (PTX memory model, causality order), so the waiting block's acquire load sees all of them.  (A fence in every
    // thread instead kept each warp stalled ~1 us on its outstanding stores: membar 0.53 stall cycles per issue in ncu.)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ch.block_seq + blockIdx.x), "r"(ch.publish) : "memory");
    }
}

// ------------------------------------------------------------------ numpy float32 helpers
// numpy's SIMD float32 sin/cos for |x| < pi/4 (5-FMA Horner). Explicit fmaf:
// these FMAs are part of the reference's arithmetic, not a contraction choice.
RLR_DEV float np_sinf_small(float x) {
    float x2 = __fmul_rn(x, x), r;
    r = __fmaf_rn(0x1.7d3bbcp-19f, x2, -0x1.a06bbap-13f);
    r = __fmaf_rn(r, x2, 0x1.11119ap-07f);
    r = __fmaf_rn(r, x2, -0x1.555556p-03f);
    r = __fmaf_rn(r, x2, 0.0f);
    r = __fmaf_rn(r, x, x);
    return r;
}
RLR_DEV float np_cosf_small(float x) {
    float x2 = __fmul_rn(x, x), r;
    r = __fmaf_rn(0x1.98e616p-16f, x2, -0x1.6c06dcp-10f);
    r = __fmaf_rn(r, x2, 0x1.55553cp-05f);
    r = __fmaf_rn(r, x2, -0x1.0p-01f);
    r = __fmaf_rn(r, x2, 1.0f);
    return r;
}
// Any argument up to numpy's Cody-Waite limit (cos 71476.0625, sin 117435.992): quadrant k = rint(x 2/pi) by the 1.5 * 2^23
// magic-number add with a FUSED multiply-add, three-term f32 reduction, both polynomials' choice by the parity of k (k + 1 for
// cosine), negation by bit 1.  Bit-identical to np.sin / np.cos on every float32 with |x| <= 64 (checked exhaustively on the
// host, tests/test_oracle_golden.py, through the C restatement the tests use).  The environments never clip the action (rocket_env.py:969-981),
// so gimbal angles beyond pi/4 (|action| > 2.25) are legal inputs.  Beyond the limit numpy calls libm: correctly rounded double.
__device__ __noinline__ float np_sincosf_reduced(float x, bool want_cos) {
    const float lim = want_cos ? 71476.0625f : 117435.992f;
    if (!(fabsf(x) <= lim)) return want_cos ? (float)cos((double)x) : (float)sin((double)x);
    const float magic = 0x1.8p23f;
    const float k = __fsub_rn(__fmaf_rn(x, 0.636619772367581343075535053490057448f, magic), magic);
    float r = __fmaf_rn(k, -0x1.921fb0p+00f, x);
    r = __fmaf_rn(k, -0x1.5110b4p-22f, r);
    r = __fmaf_rn(k, -0x1.846988p-48f, r);
    int ik = (int)k;
    if (want_cos) ik += 1;
    float v = ((ik & 1) == 0) ? np_sinf_small(r) : np_cosf_small(r);
    if ((ik & 2) == 2) v = __fsub_rn(0.0f, v);
    return v;
}
// |x| < pi/4 (every action inside the Box): k = 0 and r = x, i.e. the polynomials themselves -- kept inline, the rest out of line
RLR_DEV float np_sinf(float x) { return fabsf(x) < 0.78539816f ? np_sinf_small(x) : np_sincosf_reduced(x, false); }
RLR_DEV float np_cosf(float x) { return fabsf(x) < 0.78539816f ? np_cosf_small(x) : np_sincosf_reduced(x, true); }

// f32 ops whose rounding is part of the reference's dtype flow: never contracted.
RLR_DEV float fmul(float a, float b) { return __fmul_rn(a, b); }
RLR_DEV float fadd(float a, float b) { return __fadd_rn(a, b); }
RLR_DEV float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// numpy float32 linalg.norm on 2..4 elements: double accumulator of f32-rounded products
RLR_DEV float f32_norm2(float a, float b) { return __fsqrt_rn((float)((double)fmul(a, a) + (double)fmul(b, b))); }
RLR_DEV float f32_norm3(float a, float b, float c) {
    return __fsqrt_rn((float)(((double)fmul(a, a) + (double)fmul(b, b)) + (double)fmul(c, c)));
}
RLR_DEV float f32_norm4(float a, float b, float c, float d) {
    return __fsqrt_rn((float)((((double)fmul(a, a) + (double)fmul(b, b)) + (double)fmul(c, c)) + (double)fmul(d, d)));
}
// 10 * |nextafter(t, +inf) - t| for t >= 0  (rk.py:120)
RLR_DEV double min_step_at(double t) { return 10 * fabs(__longlong_as_double(__double_as_longlong(t) + 1) - t); }
// n / 1000.0 correctly rounded for an integer 0 <= n < 2^32 without the IEEE division: q = RN(n * RN(1/1000)), the exact residual by
// one FMA, one correction (Markstein).  Bit-identical to the division for EVERY n < 2^32 (checked exhaustively on the host,
// scripts/div1000_check.c).  The division's out-of-line slow path was taken by every lane with k = 0 (0 / 1000: a third of the warps
// of a launch hold a freshly reset environment) and cost those warps ~100 serial instructions far away in the instruction cache.
RLR_DEV double div1000_(long long n) {
    const double a = (double)n;
    if (n < 0 || n >= (1ll << 32)) return a / 1000.0;
    const double q = __dmul_rn(a, 0.001);
    const double r = __fma_rn(-q, 1000.0, a);
    return __fma_rn(r, 0.001, q);
}
RLR_DEV double py_max(double a, double b) { return (b > a) ? b : a; }
RLR_DEV double py_min(double a, double b) { return (b < a) ? b : a; }

// ------------------------------------------------------------------ division / sqrt / pow primitives
// EXACT: IEEE division, sqrt and libdevice pow, i.e. the operations numpy / scipy perform.
// FAST : branch-free hardware seed (MUFU.RCP64H / MUFU.RSQ64H, or an f32 exp2/log2 seed for the fifth root) refined by
//        two Newton steps to ~1 ulp.  No slow-path calls: an IEEE `a / b` costs ~25 SASS instructions with a divergent
//        call, 157 such sites made the first kernel 139 KB and instruction-cache bound (profiles/r01a_generic_kernel.md).
#if RLR_EXACT
RLR_DEV double rcp_(double x) { return 1.0 / x; }
RLR_DEV double sqrt_(double x) { return sqrt(x); }
RLR_DEV double rsqrt_(double x) { return 1.0 / sqrt(x); }
RLR_DEV double pow_m02(double x) { return pow(x, -0.2); }
RLR_DEV double pow_p02(double x) { return pow(x, 1.0 / 5.0); }
#define RLR_DIV(a, b) ((a) / (b))
#else
// Measured on B200 (scripts/seed_accuracy.cu, 2^26 arguments over 2^+-20): the hardware seeds are good to 2^-20; seed + one Newton
// step 1e-12; rcp seed + two Newton steps reproduces the IEEE quotient; rsqrt seed + two Newton steps 2.7e-16; x^(-1/5) from the f32
// exp2/log2 seed + one Newton step 4e-13.  (A single cubic step r0 (1 + e + e^2) is also 1 ulp with one FMA less, but measured no faster.)
RLR_DEV double rcp_(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
// reciprocal of an error-control scale (atol + rtol |y|): hardware seed + ONE Newton step (1e-12).  These factors only
// weight the RMS norms that choose the step size, where a relative error of 1e-12 moves h by 1e-13 and the state by < 1e-15.
RLR_DEV double rcp1_(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
// max(|a|, |b|) with one compare and two selects (fmax() expands to a NaN-aware six-instruction sequence)
RLR_DEV double absmax_(double a, double b) { const double fa = fabs(a), fb = fabs(b); return (fa > fb) ? fa : fb; }
RLR_DEV double rsqrt_(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x * r, r, 1.0);
    r = fma(0.5 * r, e, r);
    e = fma(-x * r, r, 1.0);
    r = fma(0.5 * r, e, r);
    return r;
}
RLR_DEV double sqrt_(double x) { return (x == 0.0) ? 0.0 : x * rsqrt_(x); }      // NaN propagates
// Step-size-control grade (1e-12, like rcp1_): square root and x^(-1/5) with ONE Newton step.  They only see the RMS norms and the
// error ratio that choose h; the integrated state never passes through them.
RLR_DEV double sqrtc_(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x * r, r, 1.0);
    r = fma(0.5 * r, e, r);
    return (x == 0.0) ? 0.0 : x * r;
}
RLR_DEV double powc_m02(double x) {
    const float xf = fminf(fmaxf((float)x, 1e-30f), 1e30f);
    const double y = (double)__powf(xf, -0.2f);
    const double y2 = y * y;
    return y * fma(-0.2 * x, (y2 * y2) * y, 1.2);
}
RLR_DEV double powc_p02(double x) { const double y = powc_m02(x), y2 = y * y; return x * (y2 * y2); }
// x^(-1/5): f32 seed (rel. err ~1e-6), Newton y <- y*(1.2 - 0.2*x*y^5) twice (error 3e^2 per step).  Out-of-range
// or non-finite x yields garbage/NaN that the callers' max(0.2, .) / min(10, .) clamps resolve like the exact path.
RLR_DEV double pow_m02(double x) {
    const float xf = fminf(fmaxf((float)x, 1e-30f), 1e30f);
    double y = (double)__powf(xf, -0.2f);
    const double mx = -0.2 * x;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double y2 = y * y;
        const double y5 = (y2 * y2) * y;
        y = y * fma(mx, y5, 1.2);
    }
    return y;
}
RLR_DEV double pow_p02(double x) {          // x^(1/5) = x * (x^(-1/5))^4
    const double y = pow_m02(x), y2 = y * y;
    return x * (y2 * y2);
}
#define RLR_DIV(a, b) ((a) * rcp_(b))
#endif

#if !RLR_EXACT
// sin and cos of a moderate argument (|x| < ~1e5; the 3DOF pitch angle is wrapped to [0, 2 pi) every step) in ~35 FP64 instructions:
// three-term Cody-Waite reduction by pi/2 and the fdlibm kernel polynomials on [-pi/4, pi/4] (< 1 ulp each).  libdevice's sincos()
// carries the Payne-Hanek path for huge arguments and costs more than twice as much; the 3DOF right-hand side is dominated by it.
RLR_DEV void sincos_fast(double x, double* sn, double* cs) {
    const double n = rint(x * 6.36619772367581382433e-01);          // x * 2 / pi
    double r = fma(-n, 1.57079632673412561417e+00, x);               // pi/2 in three parts (fdlibm pio2_1, pio2_2, pio2_3)
    r = fma(-n, 6.07710050630396597660e-11, r);
    r = fma(-n, 2.02226624879595063154e-21, r);
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double s0 = fma(z * r, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double c0 = fma(z * z, pc, fma(z, -0.5, 1.0));
    const int q = (int)n & 3;
    const double sa = (q & 1) ? c0 : s0, ca = (q & 1) ? s0 : c0;
    *sn = (q & 2) ? -sa : sa;
    *cs = ((q + 1) & 2) ? -ca : ca;
}
#endif

// ------------------------------------------------------------------ Dormand-Prince tableau
#define A10 (1.0 / 5)
#define A20 (3.0 / 40)
#define A21 (9.0 / 40)
#define A30 (44.0 / 45)
#define A31 (-56.0 / 15)
#define A32 (32.0 / 9)
#define A40 (19372.0 / 6561)
#define A41 (-25360.0 / 2187)
#define A42 (64448.0 / 6561)
#define A43 (-212.0 / 729)
#define A50 (9017.0 / 3168)
#define A51 (-355.0 / 33)
#define A52 (46732.0 / 5247)
#define A53 (49.0 / 176)
#define A54 (-5103.0 / 18656)
#define B0 (35.0 / 384)
#define B2 (500.0 / 1113)
#define B3 (125.0 / 192)
#define B4 (-2187.0 / 6784)
#define B5 (11.0 / 84)
#define E0 (-71.0 / 57600)
#define E2 (71.0 / 16695)
#define E3 (-71.0 / 1920)
#define E4 (17253.0 / 339200)
#define E5 (-22.0 / 525)
#define E6 (1.0 / 40)

__constant__ double c_P[7][4] = {
    {1, -8048581381.0 / 2820520608, 8663915743.0 / 2820520608, -12715105075.0 / 11282082432},
    {0, 0, 0, 0},
    {0, 131558114200.0 / 32700410799, -68118460800.0 / 10900136933, 87487479700.0 / 32700410799},
    {0, -1754552775.0 / 470086768, 14199869525.0 / 1410260304, -10690763975.0 / 1880347072},
    {0, 127303824393.0 / 49829197408, -318862633887.0 / 49829197408, 701980252875.0 / 199316789632},
    {0, -282668133.0 / 205662961, 2019193451.0 / 616988883, -1453857185.0 / 822651844},
    {0, 40617522.0 / 29380423, -110615467.0 / 29380423, 69997945.0 / 29380423}};

// ------------------------------------------------------------------ dynamical systems
// 6DOF: y = [r(3), v(3), q0..q3, w(3), m]   (simulator.py:259-378)
struct Sys6 {
    static constexpr int N = 14, EV = 0, NA = 3, NOBS = 14, MAX_REGS = RLR_REGS6;
    static constexpr bool AERO = false;
    struct Ctx { double Tb0, Tb1, Tb2, tau1, tau2, dm; };

    RLR_DEV static void make_ctx(Ctx& c, const float* u) {
        float cy = np_cosf(u[0]), sy = np_sinf(u[0]), cz = np_cosf(u[1]), sz = np_sinf(u[1]);
        float cycz = fmul(cy, cz), sycz = fmul(sy, cz);
        double T = (double)u[2];
        c.Tb0 = (double)cycz * T;        // first column of _rot_mat_thrust_to_body times thrust (simulator.py:316-318)
        c.Tb1 = (double)sycz * T;
        c.Tb2 = (double)sz * T;
        c.tau1 = 15.0 * c.Tb2;           // r_T_B x T_b with r_T_B = [-15,0,0]  (simulator.py:373-378)
        c.tau2 = -15.0 * c.Tb1;
        c.dm = (double)fdiv(-u[2], 3531.6f);   // f32 division under NEP-50 (simulator.py:291-292)
    }

    RLR_DEV static void rhs(const Ctx& c, const double (&y)[N], double (&d)[N]) {
        const double q0 = y[6], q1 = y[7], q2 = y[8], q3 = y[9];
        const double wx = y[10], wy = y[11], wz = y[12];
        // Rotation.from_quat([q1,q2,q3,q0]) normalises; as_matrix() closed form
        const double n2 = q1 * q1 + q2 * q2 + q3 * q3 + q0 * q0;
#if RLR_EXACT
        const double nq = sqrt(n2);
        const double x = q1 / nq, yy = q2 / nq, z = q3 / nq, w = q0 / nq;
        const double inv_m = 1.0 / y[13];
#else
        const double rn = rsqrt(n2);
        const double x = q1 * rn, yy = q2 * rn, z = q3 * rn, w = q0 * rn;
        const double inv_m = 1.0 / y[13];
#endif
        const double x2 = x * x, y2 = yy * yy, z2 = z * z, w2 = w * w;
        const double xy = x * yy, zw = z * w, xz = x * z, yw = yy * w, yz = yy * z, xw = x * w;
        const double R00 = x2 - y2 - z2 + w2, R01 = 2 * (xy - zw), R02 = 2 * (xz + yw);
        const double R10 = 2 * (xy + zw), R11 = -x2 + y2 - z2 + w2, R12 = 2 * (yz - xw);
        const double R20 = 2 * (xz - yw), R21 = 2 * (yz + xw), R22 = -x2 - y2 + z2 + w2;
        const double F0 = R00 * c.Tb0 + R01 * c.Tb1 + R02 * c.Tb2;
        const double F1 = R10 * c.Tb0 + R11 * c.Tb1 + R12 * c.Tb2;
        const double F2 = R20 * c.Tb0 + R21 * c.Tb1 + R22 * c.Tb2;
        d[0] = y[3]; d[1] = y[4]; d[2] = y[5];
        d[3] = inv_m * F0 + (-9.81);
        d[4] = inv_m * F1;
        d[5] = inv_m * F2;
        // dq = 0.5 * OMEGA(w) . q with the un-normalised q (simulator.py:287, :362-370)
        d[6] = 0.5 * (((-wx) * q1 + (-wy) * q2) + (-wz) * q3);
        d[7] = 0.5 * ((wx * q0 + wz * q2) + (-wy) * q3);
        d[8] = 0.5 * ((wy * q0 + (-wz) * q1) + wx * q3);
        d[9] = 0.5 * ((wz * q0 + wy * q1) + (-wx) * q2);
        // dom = Jinv . (tau - w x (J w)),  J = diag(75350.25, 6037675.13, 6037675.13)
        const double J0 = 75350.25, J1 = 6037675.13, J2 = 6037675.13;
        const double Jw0 = J0 * wx, Jw1 = J1 * wy, Jw2 = J2 * wz;
        const double c0 = wy * Jw2 - wz * Jw1, c1 = wz * Jw0 - wx * Jw2, c2 = wx * Jw1 - wy * Jw0;
        d[10] = (1.0 / J0) * (0.0 - c0);
        d[11] = (1.0 / J1) * (c.tau1 - c1);
        d[12] = (1.0 / J2) * (c.tau2 - c2);
        d[13] = c.dm;
    }
};

// 3DOF: y = [x, z, theta, vx, vz, omega, m]   (simulator.py:88-130)
struct Sys3 {
    static constexpr int N = 7, EV = 1, NA = 2, NOBS = 7, MAX_REGS = 168;
    static constexpr bool AERO = false;
    struct Ctx { double delta, T, torque, dm, sd, cd; };

    RLR_DEV static void make_ctx(Ctx& c, const float* u) {
        c.delta = (double)u[0];
        c.T = (double)u[1];
#if !RLR_EXACT
        sincos(c.delta, &c.sd, &c.cd);      // fast build: sin / cos (delta + phi) by angle addition, one sincos(phi) per right-hand side
#endif
        float ts = fmul(fmul(u[1], np_sinf(u[0])), 30.0f);   // f32 product T*sin(delta)*(x_T - x_CG)  (simulator.py:124-125)
        c.torque = (double)ts;
        c.dm = (double)fdiv(-u[1], 3531.6f);                  // simulator.py:126
    }

    RLR_DEV static void rhs(const Ctx& c, const double (&y)[N], double (&d)[N]) {
        const double phi = y[2], vx = y[3], vz = y[4], mass = y[6];
        const double v2 = vx * vx + vz * vz;
        const double Q = (0.5 * 1.225) * v2;
        const double A = (0.3 * Q) * 10.5;
        double sdp, cdp, cp;
#if RLR_EXACT
        sincos(c.delta + phi, &sdp, &cdp);
        cp = cos(phi);
#else
        double sp;
        sincos_fast(phi, &sp, &cp);
        sdp = c.sd * cp + c.cd * sp;
        cdp = c.cd * cp - c.sd * sp;
#endif
        // NB: drag multiplies cos(phi) on BOTH axes -- reference behaviour (simulator.py:122-123)
        const double Acp = A * cp;
        d[0] = vx; d[1] = vz; d[2] = y[5];
        d[3] = RLR_DIV((c.T * cdp - 0.0) - Acp, mass);
        d[4] = RLR_DIV((c.T * sdp + 0.0) - Acp, mass) - 9.81;
        d[5] = (-0.0 - c.torque) / 6.04e6;
        d[6] = c.dm;
    }
};
// Same system under a 128-register cap: 4 blocks x 128 threads per SM instead of 3, i.e. one wave instead of 1.15 for 65 536
// environments (15.0 vs 16.7 us per step); small batches are latency-bound and prefer the spill-free 168-register build.
struct Sys3Dense : Sys3 { static constexpr int MAX_REGS = 128; };

// EXTENSION (RlrParams.aero_mode = 1): the same system with the angle-of-attack line the reference keeps
// commented out enabled (simulator.py:103-104) -- alfa = _computeAoA(state) = phi - arctan2(vz, vx) (0 for a zero velocity;
// simulator.py:132-148), folded to (-pi, pi] by the `_normalize` the reference forgot to define (here: its own _wrapTo2Pi,
// simulator.py:150-163, minus 2 pi above pi), Cn = Cnalfa alfa (Cnalfa = 1), N = Cn Q Sref, and then the reference's own
// ax / ay / dom with their N terms (simulator.py:119-125; x_CG - x_CP = -10).  d(omega)/dt now depends on the state, so this system
// always runs the generic solver (the closed forms of rk45_solve3 need a constant angular acceleration).
struct Sys3Aero : Sys3 {
    static constexpr bool AERO = true;
    RLR_DEV static void rhs(const Ctx& c, const double (&y)[N], double (&d)[N]) {
        const double phi = y[2], vx = y[3], vz = y[4], mass = y[6];
        const double v2 = vx * vx + vz * vz;
        const double Q = (0.5 * 1.225) * v2;
        const double A = (0.3 * Q) * 10.5;
        const double pi = 3.141592653589793, pi2 = 2.0 * 3.141592653589793;
        double alfa = 0.0;
        if (!(vx == 0.0 && vz == 0.0)) alfa = phi - atan2(vz, vx);
        alfa = fmod(fmod(alfa, pi2) + pi2, pi2);
        if (alfa > pi) alfa = alfa - pi2;
        const double Nf = ((1 * alfa) * Q) * 10.5;
        double sdp, cdp, sp, cp;
#if RLR_EXACT
        sincos(c.delta + phi, &sdp, &cdp);
        sincos(phi, &sp, &cp);
#else
        sincos_fast(phi, &sp, &cp);
        sdp = c.sd * cp + c.cd * sp;
        cdp = c.cd * cp - c.sd * sp;
#endif
        const double Acp = A * cp;      // the drag multiplies cos(phi) on both axes -- reference behaviour (simulator.py:122-123)
        d[0] = vx; d[1] = vz; d[2] = y[5];
        d[3] = RLR_DIV((c.T * cdp - Nf * sp) - Acp, mass);
        d[4] = RLR_DIV((c.T * sdp + Nf * cp) - Acp, mass) - 9.81;
        d[5] = (Nf * (double)(10 - 20) - c.torque) / 6.04e6;
        d[6] = c.dm;
    }
};


// ------------------------------------------------------------------ RK45 (scipy semantics)
template <int N>
RLR_DEV double rms_ratio(const double (&a)[N], const double (&s)[N]) {
    double acc = 0.0;
#pragma unroll
#if RLR_EXACT
    for (int i = 0; i < N; ++i) { double q = a[i] / s[i]; acc += q * q; }
    return sqrt(acc) / sqrt((double)N);
#else
    // step-size-control grade arithmetic (1e-12), see rcp1_ / sqrtc_
    for (int i = 0; i < N; ++i) { double q = a[i] * rcp1_(s[i]); acc += q * q; }
    return sqrtc_(acc) * (N == 7 ? 0.37796447300922722721 : 0.26726124191242438468);      // 1 / sqrt(N)
#endif
}

// dense-output polynomial row for component i: Q_i[c] = sum_j K[j][i] * P[j][c]
template <int N>
RLR_DEV void dense_row(const double (&K)[7][N], int i, double (&Q)[4]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 7; ++j) acc += K[j][i] * c_P[j][c];
        Q[c] = acc;
    }
}
RLR_DEV double dense_at(const double (&Q)[4], double y_old, double t_old, double h, double t) {
    double x = (t - t_old) / h;
    double p0 = x, p1 = p0 * x, p2 = p1 * x, p3 = p2 * x;
    double s = ((Q[0] * p0 + Q[1] * p1) + Q[2] * p2) + Q[3] * p3;
    return h * s + y_old;
}

// scipy.optimize.brentq (Zeros/brentq.c), xtol = rtol = 4*eps, maxiter = 100, on the event component
__device__ __noinline__ double brentq_event(double q0, double q1, double q2, double q3, double y_old, double t_old, double h, double xa, double xb) {
    const double Q[4] = {q0, q1, q2, q3};
    const double xtol = 4 * 2.220446049250313e-16, rtol = 4 * 2.220446049250313e-16;
    double xpre = xa, xcur = xb, xblk = 0, fblk = 0, spre = 0, scur = 0;
    double fpre = dense_at(Q, y_old, t_old, h, xpre), fcur = dense_at(Q, y_old, t_old, h, xcur);
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
#pragma unroll 1
    for (int it = 0; it < 100; ++it) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) { xblk = xpre; fblk = fpre; spre = scur = xcur - xpre; }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        double delta = (xtol + rtol * fabs(xcur)) / 2;
        double sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            double stry;
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);
            } else {
                double dpre = (fpre - fcur) / (xpre - xcur);
                double dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            if (2 * fabs(stry) < fmin(fabs(spre), 3 * fabs(sbis) - delta)) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else { spre = sbis; scur = sbis; }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = dense_at(Q, y_old, t_old, h, xcur);
    }
    return xcur;
}

// solve_ivp(RHS, [t0, t_bound], y, events = terminal height event) with RK45 defaults.
// Returns status: 0 reached t_bound, 1 ground event, -1 step size underflow.
template <class S>
RLR_DEV int rk45_solve(const typename S::Ctx& c, double t0, double t_bound, double (&y)[S::N], int& natt) {
    constexpr int N = S::N;
    constexpr double rtol = 1e-3, atol = 1e-6;
    double K[7][N];
    S::rhs(c, y, K[0]);
    natt = 0;
    // ---- select_initial_step (common.py:68-134)
    double h_abs;
    {
        const double interval = fabs(t_bound - t0);
        double scale[N], y1[N], f1[N];
#pragma unroll
        for (int i = 0; i < N; ++i) scale[i] = atol + fabs(y[i]) * rtol;
        const double d0 = rms_ratio<N>(y, scale);
        const double d1 = rms_ratio<N>(K[0], scale);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : RLR_DIV(0.01 * d0, d1);
        h0 = py_min(h0, interval);
#pragma unroll
        for (int i = 0; i < N; ++i) y1[i] = y[i] + h0 * K[0][i];
        S::rhs(c, y1, f1);
#pragma unroll
        for (int i = 0; i < N; ++i) f1[i] = f1[i] - K[0][i];
        const double d2 = RLR_DIV(rms_ratio<N>(f1, scale), h0);
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = py_max(1e-6, h0 * 1e-3);
#if RLR_EXACT
        else h1 = pow_p02(0.01 / py_max(d1, d2));
#else
        else h1 = powc_p02(0.01 * rcp_(py_max(d1, d2)));
#endif
        h_abs = py_min(py_min(100 * h0, h1), interval);
    }
    double t = t0;
    double min_step = min_step_at(t);
    if (h_abs < min_step) h_abs = min_step;
    bool rejected = false;
    int status;
#pragma unroll 1
    for (;;) {
        if (!(h_abs >= min_step)) { status = -1; break; }      // (negated form: a non-finite step size also ends the solve instead of looping forever)
        double t_new = t + h_abs;
        if (t_new - t_bound > 0) t_new = t_bound;
        const double h = t_new - t;
        h_abs = fabs(h);
        ++natt;
        double ys[N], y_new[N];
        // ---- rk_step (rk.py:14-71): dy = dot(K[:s].T, a[:s]) * h
#pragma unroll
        for (int i = 0; i < N; ++i) ys[i] = y[i] + (K[0][i] * A10) * h;
        S::rhs(c, ys, K[1]);
#pragma unroll
        for (int i = 0; i < N; ++i) ys[i] = y[i] + (K[0][i] * A20 + K[1][i] * A21) * h;
        S::rhs(c, ys, K[2]);
#pragma unroll
        for (int i = 0; i < N; ++i) ys[i] = y[i] + ((K[0][i] * A30 + K[1][i] * A31) + K[2][i] * A32) * h;
        S::rhs(c, ys, K[3]);
#pragma unroll
        for (int i = 0; i < N; ++i) ys[i] = y[i] + (((K[0][i] * A40 + K[1][i] * A41) + K[2][i] * A42) + K[3][i] * A43) * h;
        S::rhs(c, ys, K[4]);
#pragma unroll
        for (int i = 0; i < N; ++i)
            ys[i] = y[i] + ((((K[0][i] * A50 + K[1][i] * A51) + K[2][i] * A52) + K[3][i] * A53) + K[4][i] * A54) * h;
        S::rhs(c, ys, K[5]);
#pragma unroll
        for (int i = 0; i < N; ++i)
            y_new[i] = y[i] + h * ((((K[0][i] * B0 + K[2][i] * B2) + K[3][i] * B3) + K[4][i] * B4) + K[5][i] * B5);
        S::rhs(c, y_new, K[6]);
        // ---- error norm (rk.py:101-107,143-144)
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double num = (((((K[0][i] * E0 + K[2][i] * E2) + K[3][i] * E3) + K[4][i] * E4) + K[5][i] * E5) + K[6][i] * E6) * h;
#if RLR_EXACT
            const double e = num / (atol + fmax(fabs(y[i]), fabs(y_new[i])) * rtol);
#else
            const double e = num * rcp1_(fma(absmax_(y[i], y_new[i]), rtol, atol));
#endif
            acc += e * e;
        }
#if RLR_EXACT
        const double err = sqrt(acc) / sqrt((double)N);
#else
        const double err = sqrtc_(acc) * (N == 7 ? 0.37796447300922722721 : 0.26726124191242438468);
#endif
        if (err < 1) {
#if RLR_EXACT
            double factor = (err == 0) ? 10.0 : py_min(10.0, 0.9 * pow_m02(err));
#else
            double factor = (err == 0) ? 10.0 : py_min(10.0, 0.9 * powc_m02(err));
#endif
            if (rejected) factor = py_min(1.0, factor);
            h_abs *= factor;
            // ---- terminal event g(y) = y[EV]  (ivp.py:134-158, 678-699)
            const double g = y[S::EV], g_new = y_new[S::EV];
            if (((g <= 0) && (g_new >= 0)) || ((g >= 0) && (g_new <= 0))) {
                double Q[4];
                dense_row<N>(K, S::EV, Q);
                const double root = brentq_event(Q[0], Q[1], Q[2], Q[3], y[S::EV], t, h, t, t_new);
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    dense_row<N>(K, i, Q);
                    y[i] = dense_at(Q, y[i], t, h, root);
                }
                status = 1;
                break;
            }
#pragma unroll
            for (int i = 0; i < N; ++i) { y[i] = y_new[i]; K[0][i] = K[6][i]; }
            t = t_new;
            if (t - t_bound >= 0) { status = 0; break; }
            rejected = false;
            min_step = min_step_at(t);
            if (h_abs < min_step) h_abs = min_step;
        } else {
#if RLR_EXACT
            h_abs *= py_max(0.2, 0.9 * pow(err, -0.2));
#else
            h_abs *= py_max(0.2, 0.9 * powc_m02(err));
#endif
            rejected = true;
        }
    }
    return status;
}


// ------------------------------------------------------------------ 6DOF: shared-memory stage matrix, rolled stage loop
// Row s (1..6) of c_RKA: coefficients a_sj, j = 0..5; row 6 is B (y_new is "stage 6": FSAL).
__constant__ double c_RKA[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {A10, 0, 0, 0, 0, 0},
    {A20, A21, 0, 0, 0, 0},
    {A30, A31, A32, 0, 0, 0},
    {A40, A41, A42, A43, 0, 0},
    {A50, A51, A52, A53, A54, 0},
    {B0, 0, B2, B3, B4, B5}};
__constant__ double c_RKB[6] = {B0, 0, B2, B3, B4, B5};
__constant__ double c_RKE[7] = {E0, 0, E2, E3, E4, E5, E6};
__constant__ double c_RKC[7] = {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1, 1};      // nodes c_s = sum_j a_sj (rk.py:282)
#define RLR_INV_RS14 0.2672612419124243846      // 1 / sqrt(14)

// Right-hand side restricted to the components that need arithmetic: d = [dv(3), dq(4), dw(3)].
// dr is the stage velocity itself and dm is constant over the env step (simulator.py:280, :291-292).
RLR_DEV void rhs6r(const Sys6::Ctx& c, const double (&q)[4], const double (&wv)[3], double m, double (&d)[10]) {
    const double q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
    const double wx = wv[0], wy = wv[1], wz = wv[2];
    const double J0 = 75350.25, J1 = 6037675.13, J2 = 6037675.13;
#if RLR_EXACT
    const double n2 = q1 * q1 + q2 * q2 + q3 * q3 + q0 * q0;
    const double nq = sqrt(n2);
    const double x = q1 / nq, yy = q2 / nq, z = q3 / nq, w = q0 / nq;
    const double inv_m = rcp_(m);
    const double x2 = x * x, y2 = yy * yy, z2 = z * z, w2 = w * w;
    const double xy = x * yy, zw = z * w, xz = x * z, yw = yy * w, yz = yy * z, xw = x * w;
    const double R00 = x2 - y2 - z2 + w2, R01 = 2 * (xy - zw), R02 = 2 * (xz + yw);
    const double R10 = 2 * (xy + zw), R11 = -x2 + y2 - z2 + w2, R12 = 2 * (yz - xw);
    const double R20 = 2 * (xz - yw), R21 = 2 * (yz + xw), R22 = -x2 - y2 + z2 + w2;
    const double F0 = R00 * c.Tb0 + R01 * c.Tb1 + R02 * c.Tb2;
    const double F1 = R10 * c.Tb0 + R11 * c.Tb1 + R12 * c.Tb2;
    const double F2 = R20 * c.Tb0 + R21 * c.Tb1 + R22 * c.Tb2;
    d[0] = inv_m * F0 + (-9.81);
    d[1] = inv_m * F1;
    d[2] = inv_m * F2;
    d[3] = 0.5 * (((-wx) * q1 + (-wy) * q2) + (-wz) * q3);
    d[4] = 0.5 * ((wx * q0 + wz * q2) + (-wy) * q3);
    d[5] = 0.5 * ((wy * q0 + (-wz) * q1) + wx * q3);
    d[6] = 0.5 * ((wz * q0 + wy * q1) + (-wx) * q2);
    const double Jw0 = J0 * wx, Jw1 = J1 * wy, Jw2 = J2 * wz;
    const double c0 = wy * Jw2 - wz * Jw1, c1 = wz * Jw0 - wx * Jw2, c2 = wx * Jw1 - wy * Jw0;
    d[7] = (1.0 / J0) * (0.0 - c0);
    d[8] = (1.0 / J1) * (c.tau1 - c1);
    d[9] = (1.0 / J2) * (c.tau2 - c2);
#else
    // R(q/|q|) Tb = [(w^2 - u.u) Tb + 2 (u.Tb) u + 2 w (u x Tb)] / |q|^2  with q = (w, u): 23 operations instead of the 37 of the
    // explicit rotation matrix, one reciprocal (1 / (m |q|^2)), no rsqrt, and the products do not wait for it.
    const double w = q0, x = q1, yy = q2, z = q3;
    const double uu = (x * x + yy * yy) + z * z;
    const double inv_m = rcp_(m * (uu + w * w));
    const double a = w * w - uu;
    const double b = 2.0 * ((x * c.Tb0 + yy * c.Tb1) + z * c.Tb2);
    const double w2 = w + w;
    const double cx = yy * c.Tb2 - z * c.Tb1, cy = z * c.Tb0 - x * c.Tb2, cz = x * c.Tb1 - yy * c.Tb0;
    const double F0 = (a * c.Tb0 + b * x) + w2 * cx;
    const double F1 = (a * c.Tb1 + b * yy) + w2 * cy;
    const double F2 = (a * c.Tb2 + b * z) + w2 * cz;
    d[0] = inv_m * F0 + (-9.81);
    d[1] = inv_m * F1;
    d[2] = inv_m * F2;
    const double hx = 0.5 * wx, hy = 0.5 * wy, hz = 0.5 * wz;
    d[3] = ((-hx) * q1 + (-hy) * q2) + (-hz) * q3;
    d[4] = (hx * q0 + hz * q2) + (-hy) * q3;
    d[5] = (hy * q0 + (-hz) * q1) + hx * q3;
    d[6] = (hz * q0 + hy * q1) + (-hx) * q2;
    // w x (J w) with J = diag(J0, J1, J1): component 0 vanishes, the others are (J0 - J1) wx wz and (J1 - J0) wx wy
    d[7] = 0.0;
    d[8] = (1.0 / J1) * c.tau1 - (((J0 - J2) / J1) * wx) * wz;
    d[9] = (1.0 / J2) * c.tau2 - (((J1 - J0) / J2) * wx) * wy;
#endif
}

constexpr int BS6 = 64;                                   // threads per block of the 6DOF kernels (compile time: LDS offsets are immediates)
constexpr int KROW = 5 * BS6;                             // double2 slots between consecutive stages (10 doubles per thread and stage)
constexpr size_t SMEM6_K_BYTES = 6 * KROW * sizeof(double2);   // stage matrix: 30 KB per block
// ... followed by a 20-byte-per-thread parking area for the per-environment scalars that are dead during the solve (running return,
// step counter, episode counter, |v| of the initial condition).  At the 128-register cap ptxas spilled them to local memory, and with
// 210 KB of the SM's 256 KB configured as shared memory the 448 threads' stack frames do not fit what is left of L1: the reloads in
// the epilogue were L2 round trips (6.5 % of the step kernel's stall samples, profiles/r01e).  7 blocks x (32 000 + 1 024) B still fit an SM.
constexpr size_t SMEM6_BYTES = SMEM6_K_BYTES + BS6 * (sizeof(double) + 3 * sizeof(int32_t));
struct Park6 {
    volatile double* ep_ret; volatile int32_t* k; volatile int32_t* episode; volatile float* v0;
    RLR_DEV explicit Park6(double* sK) {
        char* base = reinterpret_cast<char*>(sK) + SMEM6_K_BYTES;
        ep_ret = reinterpret_cast<double*>(base) + threadIdx.x;
        k = reinterpret_cast<int32_t*>(base + BS6 * 8) + threadIdx.x;
        episode = reinterpret_cast<int32_t*>(base + BS6 * 12) + threadIdx.x;
        v0 = reinterpret_cast<float*>(base + BS6 * 16) + threadIdx.x;
    }
};

// Terminal ground event of the 6DOF solve (ivp.py:678-699): dense output Q = K^T P (rk.py:554-565, :715-737), root of the height
// polynomial, final state y(t_e).  buf = [y(14) | K6(10) | v_new(3) | dm, t_old, h, t_new]; y(t_e) is written to buf[0..13].
// K0..K5 come from shared memory.  P[.][0] = (1,0,..,0) and P[1][.] = 0 make those terms exact no-ops; the position rows of K are
// the stage velocities, recomputed from the dv rows.
#if RLR_EXACT
// Exact build: scipy's own evaluation order -- Q = K^T P per component (three 6-term chains), then the polynomial in x, and
// Brent's method for the root.
__device__ __noinline__ void event_path6(const double2* __restrict__ sKt, double* __restrict__ buf) {
#define KS(j, cc) (((cc) & 1) ? sKt[(j) * KROW + ((cc) >> 1) * BS6].y : sKt[(j) * KROW + ((cc) >> 1) * BS6].x)
    const double dm = buf[27], t = buf[28], h = buf[29], t_new = buf[30];
    double vs[5][3];                       // stage velocities 2..5 (K_1 has zero weight in P)
#pragma unroll
    for (int s = 2; s <= 5; ++s) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double a3 = KS(0, i) * c_RKA[s][0];
#pragma unroll
            for (int j = 1; j < s; ++j) a3 += KS(j, i) * c_RKA[s][j];
            vs[s - 1][i] = buf[3 + i] + a3 * h;
        }
    }
#define DENSE_Q(k0, k2, k3, k4, k5, k6, cc) \
    ((((((k0) * c_P[0][cc] + (k2) * c_P[2][cc]) + (k3) * c_P[3][cc]) + (k4) * c_P[4][cc]) + (k5) * c_P[5][cc]) + (k6) * c_P[6][cc])
    const double vx0 = buf[3], vx6 = buf[24];
    const double hq0 = vx0, hq1 = DENSE_Q(vx0, vs[1][0], vs[2][0], vs[3][0], vs[4][0], vx6, 1),
                 hq2 = DENSE_Q(vx0, vs[1][0], vs[2][0], vs[3][0], vs[4][0], vx6, 2),
                 hq3 = DENSE_Q(vx0, vs[1][0], vs[2][0], vs[3][0], vs[4][0], vx6, 3);
    const double root = brentq_event(hq0, hq1, hq2, hq3, buf[0], t, h, t, t_new);
    const double xx = RLR_DIV(root - t, h);
    const double p1 = xx, p2 = p1 * xx, p3 = p2 * xx, p4 = p3 * xx;
#define DENSE_Y(yold, k0, k2, k3, k4, k5, k6)                                                                \
    (h * ((((k0) * p1 + DENSE_Q(k0, k2, k3, k4, k5, k6, 1) * p2) + DENSE_Q(k0, k2, k3, k4, k5, k6, 2) * p3) + \
          DENSE_Q(k0, k2, k3, k4, k5, k6, 3) * p4) + (yold))
#pragma unroll
    for (int i = 0; i < 3; ++i) buf[i] = DENSE_Y(buf[i], buf[3 + i], vs[1][i], vs[2][i], vs[3][i], vs[4][i], buf[24 + i]);   // uses the OLD v
#pragma unroll
    for (int cc = 0; cc < 10; ++cc) buf[3 + cc] = DENSE_Y(buf[3 + cc], KS(0, cc), KS(2, cc), KS(3, cc), KS(4, cc), KS(5, cc), buf[14 + cc]);
    buf[13] = DENSE_Y(buf[13], dm, dm, dm, dm, dm, dm);
#undef DENSE_Y
#undef DENSE_Q
#undef KS
}
#else
// Fast build: the same dense output with the sums regrouped so that the per-component work is ONE 6-term chain instead of three:
//   y_i(t_e) = y_i + h sum_j w_j K_j[i],   w_j = sum_c P[j][c] x^(c+1)   (x = (t_e - t) / h; w_1 = 0),
// and for the position rows (K_j[r] = stage velocity v + h sum_l a_jl K_l[dv])
//   r_i(t_e) = r_i + h ((sum_j w_j) v_i + h sum_l W_l K_l[dv_i]),   W_l = sum_{j > l} w_j a_jl   (row 6 of the tableau = B -> v_new),
// so no stage velocity is rebuilt except the four height-rate values the root finder needs.  The root of the height quartic
//   g(x) = r0 + h x (q0 + x (q1 + x (q2 + x q3)))  in [0, 1]
// comes from Newton's method started at the secant point (quadratic convergence: at most 5 steps, stopped when the update is below
// one ulp of x); anything unexpected (non-finite, left the bracket, residual not at rounding level) falls back to brentq_event, so
// the result is always a root to ~4 ulp -- the tolerance scipy's brentq (xtol = rtol = 4 eps) guarantees.
// No local-memory argument buffer (its loads were L2 round trips on the critical chain of the last warps of a launch: ground event
// -> reward -> auto-reset): six scalars come in registers, and the results go back through the lane's own stage-matrix column, which
// is dead once the episode has ended --
//   row 0, pair p:  D[2p], D[2p+1]   D_c = sum_{j<=5} w_j K_j[c]  (c = dv, dq, dw components; the caller adds w_6 K6[c] from registers)
//   row 1:          (PW_0, PW_1), (PW_2, w_6), (S, x)            PW_i = sum_l W_l K_l[dv_i],  S = sum_j w_j
__device__ __noinline__ void event_path6(double2* sKt, double g0, double vx0, double vx6, double h, double t, double t_new) {
#define KS2(j, p) sKt[(j) * KROW + (p) * BS6]
    double xx;
    {
        // height-rate stage values 2..5 (component 0 of the dv rows) and the height polynomial (component 0 of Q = K^T P)
        double kx[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) kx[j] = KS2(j, 0).x;
        double vsx[6];
#pragma unroll
        for (int s = 2; s <= 5; ++s) {
            double a3 = kx[0] * c_RKA[s][0];
#pragma unroll
            for (int j = 1; j < s; ++j) a3 = fma(kx[j], c_RKA[s][j], a3);
            vsx[s] = fma(a3, h, vx0);
        }
        double hq[4];
        hq[0] = vx0;
#pragma unroll
        for (int cc = 1; cc < 4; ++cc)
            hq[cc] = ((((vx0 * c_P[0][cc] + vsx[2] * c_P[2][cc]) + vsx[3] * c_P[3][cc]) + vsx[4] * c_P[4][cc]) + vsx[5] * c_P[5][cc]) + vx6 * c_P[6][cc];
        const double g1 = h * (((hq[0] + hq[1]) + hq[2]) + hq[3]) + g0;
        double x = RLR_DIV(g0, g0 - g1);
        const double d3 = 4.0 * hq[3], d2 = 3.0 * hq[2], d1 = 2.0 * hq[1];
#pragma unroll 1
        for (int it = 0; it < 5; ++it) {
            const double s = fma(fma(fma(hq[3], x, hq[2]), x, hq[1]), x, hq[0]);
            const double ds = fma(fma(fma(d3, x, d2), x, d1), x, hq[0]);
            const double gx = fma(h * x, s, g0);
            const double dx = RLR_DIV(gx, h * ds);
            x = x - dx;
            if (fabs(dx) <= 1.2e-16 * fabs(x)) break;
        }
        const double s = fma(fma(fma(hq[3], x, hq[2]), x, hq[1]), x, hq[0]);
        const double gx = fma(h * x, s, g0);
        const double tol = 64 * 2.220446049250313e-16 * (fabs(g0) + fabs(g1));
        if (x >= 0.0 && x <= 1.0 && fabs(gx) <= tol) xx = x;
        else xx = RLR_DIV(brentq_event(hq[0], hq[1], hq[2], hq[3], g0, t, h, t, t_new) - t, h);
    }
    // w_j(x), j = 0, 2..6 (row 1 of P is zero)
    double w[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) w[j] = (j == 1) ? 0.0 : xx * fma(fma(fma(c_P[j][3], xx, c_P[j][2]), xx, c_P[j][1]), xx, c_P[j][0]);
    const double S = ((((w[0] + w[2]) + w[3]) + w[4]) + w[5]) + w[6];
    // positions: PW_i = sum_l W_l K_l[dv_i], the rows streamed from shared memory
    double PW[3];
    {
        double W[6];
#pragma unroll
        for (int l = 0; l < 6; ++l) {
            double a = w[6] * c_RKA[6][l];
#pragma unroll
            for (int j = 5; j > l; --j) { if (j != 1) a = fma(w[j], c_RKA[j][l], a); }
            W[l] = a;
        }
        PW[0] = PW[1] = PW[2] = 0.0;
#pragma unroll
        for (int l = 0; l < 6; ++l) {
            const double2 a = KS2(l, 0);
            const double b = KS2(l, 1).x;
            PW[0] = fma(a.x, W[l], PW[0]); PW[1] = fma(a.y, W[l], PW[1]); PW[2] = fma(b, W[l], PW[2]);
        }
    }
    // dv, dq, dw: one 5-term chain per component over K0, K2..K5; pair p of row 0 is dead once D of pair p is formed
#pragma unroll
    for (int p = 0; p < 5; ++p) {
        const double2 k0 = KS2(0, p), k2 = KS2(2, p), k3 = KS2(3, p), k4 = KS2(4, p), k5 = KS2(5, p);
        const double dx_ = (((k0.x * w[0] + k2.x * w[2]) + k3.x * w[3]) + k4.x * w[4]) + k5.x * w[5];
        const double dy_ = (p == 3) ? 0.0 : (((k0.y * w[0] + k2.y * w[2]) + k3.y * w[3]) + k4.y * w[4]) + k5.y * w[5];      // d(omega_x)/dt == 0
        KS2(0, p) = make_double2(dx_, dy_);
    }
    KS2(1, 0) = make_double2(PW[0], PW[1]);
    KS2(1, 1) = make_double2(PW[2], w[6]);
    KS2(1, 2) = make_double2(S, xx);
#undef KS2
}
#endif

// solve_ivp for the 6DOF system.  y = [r(3), v(3), q(4), w(3), m] in registers; the stage derivatives K0..K5 (10 doubles
// each: dv, dq, dw -- dr is the stage velocity, dm a per-step constant) live in shared memory at
// sK[(stage*10 + c) * BS6 + tid]: consecutive threads -> consecutive 8-byte words (conflict-free), a fixed stride between
// stages so that the stage loop walks one pointer.  K0 is f(y) (FSAL), K6 = f(y_new) stays in registers until accepted.
RLR_DEV int rk45_solve6(const Sys6::Ctx& c, double t0, double t_bound, double (&y)[14], int& natt, double2* __restrict__ sKt) {
    // sKt = (double2*)sK + tid; slot (stage, pair p) holds components 2p, 2p+1 -> 128-bit shared-memory accesses
    constexpr double rtol = 1e-3, atol = 1e-6;
#if !RLR_EXACT
    // Zero mass: dv/dt is infinite.  scipy then gets h0 = 0.01 d0 / inf = 0, clamps it to min_step, rejects the one attempt (its error
    // norm is NaN) and gives up with status -1 (rk.py:111-176; pinned by the zero_mass cases of tests/golden/traj_*_landing.npz).
    // The exact build walks through this with IEEE infinities; the Newton reciprocals of the fast build would turn the infinity into
    // a NaN at once, so the outcome is stated directly.
    if (y[13] == 0.0) { natt = 1; return -1; }
#endif
#define KS2(j, p) sKt[(j) * KROW + (p) * BS6]
#define KS(j, cc) (((cc) & 1) ? KS2(j, (cc) >> 1).y : KS2(j, (cc) >> 1).x)
    double kk[10];
    {
        const double q[4] = {y[6], y[7], y[8], y[9]};
        const double wv[3] = {y[10], y[11], y[12]};
        rhs6r(c, q, wv, y[13], kk);
#pragma unroll
        for (int p = 0; p < 5; ++p) KS2(0, p) = make_double2(kk[2 * p], kk[2 * p + 1]);
    }
    natt = 0;
    // ---- select_initial_step (common.py:68-134) on the full 14-vector f0 = [v, dv, dq, dw, dm]
    double h_abs;
    {
        const double interval = fabs(t_bound - t0);
        double isc[14];
#pragma unroll
        for (int i = 0; i < 14; ++i) {
            const double sc = atol + fabs(y[i]) * rtol;
#if RLR_EXACT
            isc[i] = sc;
#else
            isc[i] = rcp1_(sc);
#endif
        }
#if RLR_EXACT
#define SCALED(a, i) ((a) / isc[i])
#else
#define SCALED(a, i) ((a) * isc[i])
#endif
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int i = 0; i < 14; ++i) { const double u_ = SCALED(y[i], i); a0 += u_ * u_; }
#pragma unroll
        for (int i = 0; i < 3; ++i) { const double u_ = SCALED(y[3 + i], i); a1 += u_ * u_; }
#if RLR_EXACT
#define RLR_LIVE(cc) true
#else
#define RLR_LIVE(cc) ((cc) != 7)            // d(omega_x)/dt == 0 exactly in the fast build (rhs6r)
#endif
#pragma unroll
        for (int i = 0; i < 10; ++i) { if (RLR_LIVE(i)) { const double u_ = SCALED(kk[i], 3 + i); a1 += u_ * u_; } }
        { const double u_ = SCALED(c.dm, 13); a1 += u_ * u_; }
#if RLR_EXACT
        const double rs14 = sqrt(14.0);
        const double d0 = RLR_DIV(sqrt_(a0), rs14), d1 = RLR_DIV(sqrt_(a1), rs14);
#else
        const double d0 = sqrtc_(a0) * RLR_INV_RS14, d1 = sqrtc_(a1) * RLR_INV_RS14;
#endif
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : RLR_DIV(0.01 * d0, d1);
        h0 = py_min(h0, interval);
        double v1[3], q1[4], w1[3], f1[10];
#pragma unroll
        for (int i = 0; i < 3; ++i) v1[i] = y[3 + i] + h0 * kk[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) q1[i] = y[6 + i] + h0 * kk[3 + i];
#pragma unroll
        for (int i = 0; i < 3; ++i) w1[i] = RLR_LIVE(7 + i) ? y[10 + i] + h0 * kk[7 + i] : y[10 + i];
        const double m1 = y[13] + h0 * c.dm;
        rhs6r(c, q1, w1, m1, f1);
        double a2 = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) { const double u_ = SCALED(v1[i] - y[3 + i], i); a2 += u_ * u_; }
#pragma unroll
        for (int i = 0; i < 10; ++i) { if (RLR_LIVE(i)) { const double u_ = SCALED(f1[i] - kk[i], 3 + i); a2 += u_ * u_; } }
#if RLR_EXACT
        { const double u_ = SCALED(c.dm - c.dm, 13); a2 += u_ * u_; }
        const double d2 = RLR_DIV(RLR_DIV(sqrt_(a2), rs14), h0);
#else
        const double d2 = RLR_DIV(sqrtc_(a2) * RLR_INV_RS14, h0);
#endif
#undef RLR_LIVE
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = py_max(1e-6, h0 * 1e-3);
#if RLR_EXACT
        else h1 = pow_p02(RLR_DIV(0.01, py_max(d1, d2)));
#else
        else h1 = powc_p02(RLR_DIV(0.01, py_max(d1, d2)));
#endif
        h_abs = py_min(py_min(100 * h0, h1), interval);
#undef SCALED
    }
    double t = t0;
    double min_step = min_step_at(t);
    if (h_abs < min_step) h_abs = min_step;
    bool rejected = false;
    int status;
    RLR_PHASE(1);                       // 1: make_ctx, K0, select_initial_step
#pragma unroll 1
    for (;;) {
        if (!(h_abs >= min_step)) { status = -1; break; }      // (negated form: a non-finite step size also ends the solve instead of looping forever)
        double t_new = t + h_abs;
        if (t_new - t_bound > 0) t_new = t_bound;
        const double h = t_new - t;
        h_abs = fabs(h);
        ++natt;
        double sv[3], sq[4], sw[3], sm = y[13];
#if RLR_EXACT
        double aB[3], aE[3];           // position rows of K are the stage velocities: accumulated on the fly (fast build: see the error norm)
#pragma unroll
        for (int i = 0; i < 3; ++i) { aB[i] = y[3 + i] * B0; aE[i] = y[3 + i] * E0; }
#endif
#if RLR_EXACT || !RLR_FUSED_ERR
        // ---- rk_step (rk.py:14-71); stage 6 is y_new (row 6 of c_RKA = B).
        // Fast build: d(omega_x)/dt is identically zero (J_y == J_z, see rhs6r), so component 7 of the stage matrix is skipped and
        // omega_x keeps its value; dm is constant, so the mass stage value is m + (c_s h) dm.
#if RLR_EXACT
#define RLR_LIVE(cc) true
#else
#define RLR_LIVE(cc) ((cc) != 7)
#endif
#pragma unroll 1
        for (int s = 1; s <= 6; ++s) {
            double acc[10];
            double a = c_RKA[s][0];
#pragma unroll
            for (int p = 0; p < 5; ++p) {
                const double2 k2 = KS2(0, p);
                acc[2 * p] = k2.x * a;
                if (RLR_LIVE(2 * p + 1)) acc[2 * p + 1] = k2.y * a;
            }
#if RLR_EXACT
            double accm = c.dm * a;
#endif
#if RLR_EXACT
            const double2* kp = sKt + KROW;
#pragma unroll 1
            for (int j = 1; j < s; ++j, kp += KROW) {
                a = c_RKA[s][j];
#pragma unroll
                for (int p = 0; p < 5; ++p) {
                    const double2 k2 = kp[p * BS6];
                    acc[2 * p] += k2.x * a;
                    acc[2 * p + 1] += k2.y * a;
                }
                accm += c.dm * a;
            }
#else
            // rows 1 .. s-1: unrolled with warp-uniform guards, so that every shared-memory and constant-bank address is base +
            // immediate (the rolled form stalls on the write-after-read hazards of its index registers: 5 % of a lone warp's time)
#pragma unroll
            for (int j = 1; j <= 5; ++j) {
                if (j < s) {
                    const double aj = c_RKA[s][j];
#pragma unroll
                    for (int p = 0; p < 5; ++p) {
                        const double2 k2 = KS2(j, p);
                        acc[2 * p] += k2.x * aj;
                        if (RLR_LIVE(2 * p + 1)) acc[2 * p + 1] += k2.y * aj;
                    }
                }
            }
#endif
#pragma unroll
            for (int i = 0; i < 3; ++i) sv[i] = y[3 + i] + acc[i] * h;
#pragma unroll
            for (int i = 0; i < 4; ++i) sq[i] = y[6 + i] + acc[3 + i] * h;
#pragma unroll
            for (int i = 0; i < 3; ++i) sw[i] = RLR_LIVE(7 + i) ? y[10 + i] + acc[7 + i] * h : y[10 + i];
#if RLR_EXACT
            sm = y[13] + accm * h;
#else
            sm = y[13] + (c_RKC[s] * h) * c.dm;
#endif
#if RLR_EXACT
            if (s <= 5) {
                const double bs = c_RKB[s], es = c_RKE[s];
#pragma unroll
                for (int i = 0; i < 3; ++i) { aB[i] += sv[i] * bs; aE[i] += sv[i] * es; }
            }
#endif
            rhs6r(c, sq, sw, sm, kk);
            if (s <= 5) {
                double2* ko = sKt + s * KROW;
#pragma unroll
                for (int p = 0; p < 5; ++p) ko[p * BS6] = make_double2(kk[2 * p], kk[2 * p + 1]);
            }
        }
        // ---- error norm (rk.py:101-107): components in index order r, v, q, w, m
        double rn[3];
        double ssum = 0.0;
#if RLR_EXACT
#define RLR_ERR_TERM(num, yo, yn) ((num) / (atol + fmax(fabs(yo), fabs(yn)) * rtol))
#else
#define RLR_ERR_TERM(num, yo, yn) ((num) * rcp1_(fma(absmax_(yo, yn), rtol, atol)))
#endif
#if RLR_EXACT
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            rn[i] = y[i] + h * aB[i];
            const double e = RLR_ERR_TERM((aE[i] + sv[i] * E6) * h, y[i], rn[i]);
            ssum += e * e;
        }
#else
        // Fast build: position rows by quadrature of the acceleration rows instead of per-stage accumulators (12 registers less in
        // the stage loop):  sum_s b_s v_s = v + h sum_j beta_j K_j,  beta_j = sum_s b_s a_sj = b_j (1 - c_j)  (= 0 for j = 1, 5), and
        // sum_s e_s v_s = h sum_j eps_j K_j  because the error weights sum to zero (eps_j = sum_s e_s a_sj + e_6 b_j).  Evaluated in
        // the loop below, where the dv rows are loaded anyway.
        constexpr double BT0 = 35.0 / 384, BT2 = 50.0 / 159, BT3 = 25.0 / 192, BT4 = -243.0 / 6784;
        constexpr double EP0 = -611.0 / 230400, EP2 = 514.0 / 83475, EP3 = -391.0 / 38400, EP4 = 4617.0 / 1356800, EP5 = 11.0 / 3360;
        const double hsq = h * h;
#endif
#pragma unroll
        for (int p = 0; p < 5; ++p) {
            const double2 k0 = KS2(0, p), k2 = KS2(2, p), k3 = KS2(3, p), k4 = KS2(4, p), k5 = KS2(5, p);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int cc = 2 * p + hh;
                if (!RLR_LIVE(cc)) continue;          // K rows of omega_x are exactly zero in the fast build
                const double yo = y[3 + cc];
                const double yn = (cc < 3) ? sv[cc] : ((cc < 7) ? sq[cc - 3] : sw[cc - 7]);
                const double c0 = hh ? k0.y : k0.x, c2 = hh ? k2.y : k2.x, c3 = hh ? k3.y : k3.x, c4 = hh ? k4.y : k4.x, c5 = hh ? k5.y : k5.x;
                const double es = ((((c0 * E0 + c2 * E2) + c3 * E3) + c4 * E4) + c5 * E5) + kk[cc] * E6;
                const double e = RLR_ERR_TERM(es * h, yo, yn);
                ssum += e * e;
#if !RLR_EXACT
                if (cc < 3) {          // the position component integrated from this acceleration row
                    const double sb = ((c0 * BT0 + c2 * BT2) + c3 * BT3) + c4 * BT4;
                    const double se = (((c0 * EP0 + c2 * EP2) + c3 * EP3) + c4 * EP4) + c5 * EP5;
                    rn[cc] = y[cc] + (h * y[3 + cc] + hsq * sb);
                    const double er = RLR_ERR_TERM(se * hsq, y[cc], rn[cc]);
                    ssum += er * er;
                }
#endif
            }
        }
#if RLR_EXACT
        {
            const double dm = c.dm;
            const double es = (((((dm * E0 + dm * 0.0) + dm * E2) + dm * E3) + dm * E4) + dm * E5) + dm * E6;
            const double e = RLR_ERR_TERM(es * h, y[13], sm);
            ssum += e * e;
        }
        const double err = RLR_DIV(sqrt_(ssum), sqrt(14.0));
#else
        // the error weights sum to zero: a constant derivative (the mass) contributes |dm| * O(1e-17) -- dropped
        const double err = sqrtc_(ssum) * RLR_INV_RS14;
#endif
#else
        // ---- rk_step (rk.py:14-71), fast build.  d(omega_x)/dt is identically zero (J_y == J_z, see rhs6r), so component 7 of the
        // stage matrix is skipped and omega_x keeps its value; dm is constant, so the mass stage value is m + (c_s h) dm.
        // Stages 1..5 run in a rolled loop (one inlined right-hand side); rows 1 .. s-1 of the stage sum are unrolled behind
        // warp-uniform guards, so every shared-memory / constant-bank address is base + immediate.
#define RLR_LIVE(cc) ((cc) != 7)
#pragma unroll 1
        for (int s = 1; s <= 5; ++s) {
            double acc[10];
            const double a0 = c_RKA[s][0];
#pragma unroll
            for (int p = 0; p < 5; ++p) {
                const double2 k2 = KS2(0, p);
                acc[2 * p] = k2.x * a0;
                if (RLR_LIVE(2 * p + 1)) acc[2 * p + 1] = k2.y * a0;
            }
#pragma unroll
            for (int j = 1; j <= 4; ++j) {
                if (j < s) {
                    const double aj = c_RKA[s][j];
#pragma unroll
                    for (int p = 0; p < 5; ++p) {
                        const double2 k2 = KS2(j, p);
                        acc[2 * p] += k2.x * aj;
                        if (RLR_LIVE(2 * p + 1)) acc[2 * p + 1] += k2.y * aj;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) sv[i] = y[3 + i] + acc[i] * h;
#pragma unroll
            for (int i = 0; i < 4; ++i) sq[i] = y[6 + i] + acc[3 + i] * h;
#pragma unroll
            for (int i = 0; i < 3; ++i) sw[i] = RLR_LIVE(7 + i) ? y[10 + i] + acc[7 + i] * h : y[10 + i];
            sm = y[13] + (c_RKC[s] * h) * c.dm;
            rhs6r(c, sq, sw, sm, kk);
            double2* ko = sKt + s * KROW;
#pragma unroll
            for (int p = 0; p < 5; ++p) ko[p * BS6] = make_double2(kk[2 * p], kk[2 * p + 1]);
        }
        // ---- stage 6 is y_new (row 6 of the tableau is B; b_1 = 0, so row 1 is not read), straight-line and fused with everything
        // else that reads the stage matrix: ONE pass over rows 0, 2, 3, 4, 5 feeds
        //   * the B-weighted sum (y_new),
        //   * the E-weighted sum of the error estimate (rk.py:101-107; K6 = f(y_new) is added from registers after the right-hand side),
        //   * the position rows by quadrature of the acceleration rows instead of per-stage accumulators:
        //       sum_s b_s v_s = v + h sum_j beta_j K_j,  beta_j = sum_s b_s a_sj = b_j (1 - c_j)  (= 0 for j = 1, 5), and
        //       sum_s e_s v_s = h sum_j eps_j K_j  because the error weights sum to zero (eps_j = sum_s e_s a_sj + e_6 b_j).
        // (Before: the error norm re-read five rows per attempt -- 16 % of the kernel's shared-memory wavefronts, and the shared-memory
        // pipe is ~75 % busy while the warps of an SM are in their RK attempts.)
        constexpr double BT0 = 35.0 / 384, BT2 = 50.0 / 159, BT3 = 25.0 / 192, BT4 = -243.0 / 6784;
        constexpr double EP0 = -611.0 / 230400, EP2 = 514.0 / 83475, EP3 = -391.0 / 38400, EP4 = 4617.0 / 1356800, EP5 = 11.0 / 3360;
        double ae[10], sb[3], se[3];
        {
            double acc[10];
#pragma unroll
            for (int p = 0; p < 5; ++p) {
                const double2 k0 = KS2(0, p), k2 = KS2(2, p), k3 = KS2(3, p), k4 = KS2(4, p), k5 = KS2(5, p);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int cc = 2 * p + hh;
                    if (!RLR_LIVE(cc)) continue;
                    const double c0 = hh ? k0.y : k0.x, c2 = hh ? k2.y : k2.x, c3 = hh ? k3.y : k3.x, c4 = hh ? k4.y : k4.x, c5 = hh ? k5.y : k5.x;
                    acc[cc] = (((c0 * B0 + c2 * B2) + c3 * B3) + c4 * B4) + c5 * B5;
                    ae[cc] = (((c0 * E0 + c2 * E2) + c3 * E3) + c4 * E4) + c5 * E5;
                    if (cc < 3) {
                        sb[cc] = ((c0 * BT0 + c2 * BT2) + c3 * BT3) + c4 * BT4;
                        se[cc] = (((c0 * EP0 + c2 * EP2) + c3 * EP3) + c4 * EP4) + c5 * EP5;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) sv[i] = y[3 + i] + acc[i] * h;
#pragma unroll
            for (int i = 0; i < 4; ++i) sq[i] = y[6 + i] + acc[3 + i] * h;
#pragma unroll
            for (int i = 0; i < 3; ++i) sw[i] = RLR_LIVE(7 + i) ? y[10 + i] + acc[7 + i] * h : y[10 + i];
            sm = y[13] + h * c.dm;
        }
        rhs6r(c, sq, sw, sm, kk);
        // ---- error norm (rk.py:101-107): components in index order r, v, q, w (the error weights sum to zero: a constant derivative
        // -- the mass -- contributes |dm| * O(1e-17), dropped)
        double rn[3];
        double ssum = 0.0;
#define RLR_ERR_TERM(num, yo, yn) ((num) * rcp1_(fma(absmax_(yo, yn), rtol, atol)))
        const double hsq = h * h;
#pragma unroll
        for (int cc = 0; cc < 10; ++cc) {
            if (!RLR_LIVE(cc)) continue;          // K rows of omega_x are exactly zero
            const double yo = y[3 + cc];
            const double yn = (cc < 3) ? sv[cc] : ((cc < 7) ? sq[cc - 3] : sw[cc - 7]);
            const double es = ae[cc] + kk[cc] * E6;
            const double e = RLR_ERR_TERM(es * h, yo, yn);
            ssum += e * e;
            if (cc < 3) {          // the position component integrated from this acceleration row
                rn[cc] = y[cc] + (h * y[3 + cc] + hsq * sb[cc]);
                const double er = RLR_ERR_TERM(se[cc] * hsq, y[cc], rn[cc]);
                ssum += er * er;
            }
        }
        const double err = sqrtc_(ssum) * RLR_INV_RS14;
#endif
#undef RLR_ERR_TERM
        if (err < 1) {
#if RLR_EXACT
            double factor = (err == 0) ? 10.0 : py_min(10.0, 0.9 * pow_m02(err));
#else
            double factor = (err == 0) ? 10.0 : py_min(10.0, 0.9 * powc_m02(err));
#endif
            if (rejected) factor = py_min(1.0, factor);
            h_abs *= factor;
            const double g = y[0], g_new = rn[0];
            if (((g <= 0) && (g_new >= 0)) || ((g >= 0) && (g_new <= 0))) {
#if RLR_EXACT
                // ---- terminal ground event: rare (once per landing) -> out-of-line, arguments through a small local buffer
                double ebuf[32];
#pragma unroll
                for (int i = 0; i < 14; ++i) ebuf[i] = y[i];
#pragma unroll
                for (int cc = 0; cc < 10; ++cc) ebuf[14 + cc] = kk[cc];
#pragma unroll
                for (int i = 0; i < 3; ++i) ebuf[24 + i] = sv[i];
                ebuf[27] = c.dm; ebuf[28] = t; ebuf[29] = h; ebuf[30] = t_new;
                RLR_PHASE(2);           // 2: RK attempts (all of them when the step ends in an event)
                event_path6(sKt, ebuf);
#pragma unroll
                for (int i = 0; i < 14; ++i) y[i] = ebuf[i];
#else
                // ---- terminal ground event: rare (once per landing) -> out-of-line; six scalars in, the weighted row sums come back
                // through the lane's own stage-matrix column (see event_path6), the K6 and old-velocity terms are added here
                RLR_PHASE(2);           // 2: RK attempts (all of them when the step ends in an event)
                event_path6(sKt, y[0], y[3], sv[0], h, t, t_new);
                {
                    const double2 r10 = KS2(1, 0), r11 = KS2(1, 1), r12 = KS2(1, 2);
                    const double w6 = r11.y, S = r12.x, xx = r12.y;
                    const double pw[3] = {r10.x, r10.y, r11.x};
#pragma unroll
                    for (int i = 0; i < 3; ++i) y[i] = fma(h, fma(h, pw[i], S * y[3 + i]), y[i]);          // uses the OLD v
#pragma unroll
                    for (int p = 0; p < 5; ++p) {
                        const double2 d2 = KS2(0, p);
                        y[3 + 2 * p] = fma(h, fma(w6, kk[2 * p], d2.x), y[3 + 2 * p]);
                        if (RLR_LIVE(2 * p + 1)) y[4 + 2 * p] = fma(h, fma(w6, kk[2 * p + 1], d2.y), y[4 + 2 * p]);
                    }
                    y[13] = fma(h * xx, c.dm, y[13]);                                                    // sum_j w_j = x: the mass is linear in time
                }
#endif
                status = 1;
                break;
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) { y[i] = rn[i]; y[3 + i] = sv[i]; y[10 + i] = sw[i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i) y[6 + i] = sq[i];
            y[13] = sm;
#pragma unroll
            for (int p = 0; p < 5; ++p) KS2(0, p) = make_double2(kk[2 * p], kk[2 * p + 1]);     // FSAL: f(y_new) becomes K0
            t = t_new;
            if (t - t_bound >= 0) { status = 0; break; }
            rejected = false;
            min_step = min_step_at(t);
            if (h_abs < min_step) h_abs = min_step;
        } else {
#if RLR_EXACT
            h_abs *= py_max(0.2, 0.9 * pow_m02(err));
#else
            h_abs *= py_max(0.2, 0.9 * powc_m02(err));
#endif
            rejected = true;
        }
    }
#undef RLR_LIVE
#undef KS
#undef KS2
    if (status != 1) { RLR_PHASE(2); }
    RLR_PHASE(3);                       // 3: event path (dense output + root) -- ~0 for warps without an event
    return status;
}

// ------------------------------------------------------------------ 3DOF / 1DOF, fast build: specialised solver
// y = [x, z, theta, vx, vz, omega, m].  Only the two acceleration components need a stage matrix (12 doubles, in registers):
//   * d(omega)/dt = -torque / J and dm/dt are constant over the env step (simulator.py:124-126), so omega and m are linear and theta
//     quadratic in time:  theta_s = theta + h (c_s omega + h domega gamma_s),  gamma_s = sum_j a_sj c_j = (0, 0, 9/200, 8/25, 32/81,
//     1/2, 1/2);  their error estimates vanish identically (sum_s e_s = sum_s e_s c_s = 0), as do the higher dense-output
//     coefficients (sum_j P_jc = (1,0,0,0), sum_j c_j P_jc = (0,1/2,0,0));
//   * the right-hand side does not depend on x, z: the position follows by quadrature of the acceleration rows (see rk45_solve6).
// One sincos_fast per right-hand side.  ~2.4 x fewer instructions than the generic solver, which the exact build keeps.
#if !RLR_EXACT
RLR_DEV void rhs3r(const Sys3::Ctx& c, double phi, double vx, double vz, double inv_m, double& dvx, double& dvz) {
    const double v2 = vx * vx + vz * vz;
    const double A = (0.3 * ((0.5 * 1.225) * v2)) * 10.5;
    double sp, cp;
    sincos_fast(phi, &sp, &cp);
    const double sdp = c.sd * cp + c.cd * sp, cdp = c.cd * cp - c.sd * sp;
    const double Acp = A * cp;                 // drag multiplies cos(phi) on BOTH axes -- reference behaviour (simulator.py:122-123)
    dvx = (c.T * cdp - Acp) * inv_m;
    dvz = (c.T * sdp - Acp) * inv_m - 9.81;
}

// Terminal ground event (z = 0) of the 3DOF solve.  buf = [y(7) | kx(7) | kz(7) | dom dm t h t_new]; y(t_e) is written to buf[0..6].
__device__ __noinline__ void event_path3(double* __restrict__ buf) {
    const double dom = buf[21], dm = buf[22], t = buf[23], h = buf[24], t_new = buf[25];
    const double* kx = buf + 7;
    const double* kz = buf + 14;
    double vsx[6], vsz[6];                   // stage velocities 2..5 (K_1 has zero weight in P)
#pragma unroll
    for (int s = 2; s <= 5; ++s) {
        double ax = kx[0] * c_RKA[s][0], az = kz[0] * c_RKA[s][0];
#pragma unroll
        for (int j = 1; j < s; ++j) { ax += kx[j] * c_RKA[s][j]; az += kz[j] * c_RKA[s][j]; }
        vsx[s] = buf[3] + ax * h; vsz[s] = buf[4] + az * h;
    }
    // v_new = v + h sum_j b_j K_j (row 6 of the table)
    double bx = kx[0] * c_RKB[0], bz = kz[0] * c_RKB[0];
#pragma unroll
    for (int j = 2; j <= 5; ++j) { bx += kx[j] * c_RKB[j]; bz += kz[j] * c_RKB[j]; }
    const double vx6 = buf[3] + bx * h, vz6 = buf[4] + bz * h;
#define DENSE_Q(k0, k2, k3, k4, k5, k6, cc) \
    ((((((k0) * c_P[0][cc] + (k2) * c_P[2][cc]) + (k3) * c_P[3][cc]) + (k4) * c_P[4][cc]) + (k5) * c_P[5][cc]) + (k6) * c_P[6][cc])
    const double vz0 = buf[4];
    const double hq0 = vz0, hq1 = DENSE_Q(vz0, vsz[2], vsz[3], vsz[4], vsz[5], vz6, 1), hq2 = DENSE_Q(vz0, vsz[2], vsz[3], vsz[4], vsz[5], vz6, 2),
                 hq3 = DENSE_Q(vz0, vsz[2], vsz[3], vsz[4], vsz[5], vz6, 3);
    double xx;
    {
        const double g0 = buf[1];
        const double g1 = h * (((hq0 + hq1) + hq2) + hq3) + g0;
        double x = RLR_DIV(g0, g0 - g1);
        double gx = 0.0;
#pragma unroll 1
        for (int it = 0; it < 5; ++it) {
            const double s = fma(fma(fma(hq3, x, hq2), x, hq1), x, hq0);
            const double ds = fma(fma(fma(4.0 * hq3, x, 3.0 * hq2), x, 2.0 * hq1), x, hq0);
            gx = fma(h * x, s, g0);
            x = x - RLR_DIV(gx, h * ds);
        }
        const double s = fma(fma(fma(hq3, x, hq2), x, hq1), x, hq0);
        gx = fma(h * x, s, g0);
        const double tol = 64 * 2.220446049250313e-16 * (fabs(g0) + fabs(g1));
        if (x >= 0.0 && x <= 1.0 && fabs(gx) <= tol) xx = x;
        else xx = RLR_DIV(brentq_event(hq0, hq1, hq2, hq3, g0, t, h, t, t_new) - t, h);
    }
    const double p1 = xx, p2 = p1 * xx, p3 = p2 * xx, p4 = p3 * xx;
#define DENSE_Y(yold, k0, k2, k3, k4, k5, k6)                                                                \
    (h * ((((k0) * p1 + DENSE_Q(k0, k2, k3, k4, k5, k6, 1) * p2) + DENSE_Q(k0, k2, k3, k4, k5, k6, 2) * p3) + \
          DENSE_Q(k0, k2, k3, k4, k5, k6, 3) * p4) + (yold))
    const double hx = h * xx;
    buf[0] = DENSE_Y(buf[0], buf[3], vsx[2], vsx[3], vsx[4], vsx[5], vx6);          // uses the OLD v
    buf[1] = DENSE_Y(buf[1], buf[4], vsz[2], vsz[3], vsz[4], vsz[5], vz6);
    buf[2] = buf[2] + hx * (buf[5] + 0.5 * hx * dom);
    buf[3] = DENSE_Y(buf[3], kx[0], kx[2], kx[3], kx[4], kx[5], kx[6]);
    buf[4] = DENSE_Y(buf[4], kz[0], kz[2], kz[3], kz[4], kz[5], kz[6]);
    buf[5] = buf[5] + hx * dom;
    buf[6] = buf[6] + hx * dm;
#undef DENSE_Y
#undef DENSE_Q
}

RLR_DEV int rk45_solve3(const Sys3::Ctx& c, double t0, double t_bound, double (&y)[7], int& natt) {
    constexpr double rtol = 1e-3, atol = 1e-6;
    constexpr double INV_RS7 = 0.37796447300922722721;
    const double dom = (-0.0 - c.torque) * (1.0 / 6.04e6);
    if (y[6] == 0.0) { natt = 1; return -1; }          // zero mass: see rk45_solve6
    double kx[7], kz[7];
    rhs3r(c, y[2], y[3], y[4], rcp_(y[6]), kx[0], kz[0]);
    natt = 0;
    double h_abs;
    {   // ---- select_initial_step (common.py:68-134); f0 = [vx, vz, omega, dvx, dvz, domega, dm]
        const double interval = fabs(t_bound - t0);
        double isc[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) isc[i] = rcp1_(fma(fabs(y[i]), rtol, atol));
        const double f0[7] = {y[3], y[4], y[5], kx[0], kz[0], dom, c.dm};
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int i = 0; i < 7; ++i) { const double u_ = y[i] * isc[i], w_ = f0[i] * isc[i]; a0 += u_ * u_; a1 += w_ * w_; }
        const double d0 = sqrtc_(a0) * INV_RS7, d1 = sqrtc_(a1) * INV_RS7;
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : RLR_DIV(0.01 * d0, d1);
        h0 = py_min(h0, interval);
        const double vx1 = y[3] + h0 * kx[0], vz1 = y[4] + h0 * kz[0], om1 = y[5] + h0 * dom;
        double f1x, f1z;
        rhs3r(c, y[2] + h0 * y[5], vx1, vz1, rcp_(y[6] + h0 * c.dm), f1x, f1z);
        const double df[5] = {vx1 - y[3], vz1 - y[4], om1 - y[5], f1x - kx[0], f1z - kz[0]};
        double a2 = 0.0;
#pragma unroll
        for (int i = 0; i < 5; ++i) { const double u_ = df[i] * isc[i]; a2 += u_ * u_; }
        const double d2 = RLR_DIV(sqrtc_(a2) * INV_RS7, h0);
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = py_max(1e-6, h0 * 1e-3);
        else h1 = powc_p02(0.01 * rcp_(py_max(d1, d2)));
        h_abs = py_min(py_min(100 * h0, h1), interval);
    }
    double t = t0;
    double min_step = min_step_at(t);
    if (h_abs < min_step) h_abs = min_step;
    bool rejected = false;
    int status;
    constexpr double GAM[7] = {0.0, 0.0, 9.0 / 200, 8.0 / 25, 32.0 / 81, 0.5, 0.5};
    constexpr double CS[7] = {0.0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0};
    constexpr double RA[7][6] = {{0, 0, 0, 0, 0, 0}, {A10, 0, 0, 0, 0, 0}, {A20, A21, 0, 0, 0, 0}, {A30, A31, A32, 0, 0, 0},
                                 {A40, A41, A42, A43, 0, 0}, {A50, A51, A52, A53, A54, 0}, {B0, 0, B2, B3, B4, B5}};
#pragma unroll 1
    for (;;) {
        if (!(h_abs >= min_step)) { status = -1; break; }
        double t_new = t + h_abs;
        if (t_new - t_bound > 0) t_new = t_bound;
        const double h = t_new - t;
        h_abs = fabs(h);
        ++natt;
        double svx = y[3], svz = y[4];
        // ---- rk_step (rk.py:14-71), unrolled: compile-time coefficients, stage matrix in registers
#pragma unroll
        for (int s = 1; s <= 6; ++s) {
            double ax = kx[0] * RA[s][0], az = kz[0] * RA[s][0];
#pragma unroll
            for (int j = 1; j < s; ++j) {
                if (RA[s][j] != 0.0) { ax += kx[j] * RA[s][j]; az += kz[j] * RA[s][j]; }
            }
            svx = y[3] + ax * h; svz = y[4] + az * h;
            const double ch = CS[s] * h;
            rhs3r(c, y[2] + h * (CS[s] * y[5] + (h * dom) * GAM[s]), svx, svz, rcp_(y[6] + ch * c.dm), kx[s], kz[s]);
        }
        // ---- error norm (rk.py:101-107): x, z by quadrature of the acceleration rows; theta, omega, m contribute exactly zero
        constexpr double BT0 = 35.0 / 384, BT2 = 50.0 / 159, BT3 = 25.0 / 192, BT4 = -243.0 / 6784;
        constexpr double EP0 = -611.0 / 230400, EP2 = 514.0 / 83475, EP3 = -391.0 / 38400, EP4 = 4617.0 / 1356800, EP5 = 11.0 / 3360;
        const double hsq = h * h;
        const double xn = y[0] + (h * y[3] + hsq * (((kx[0] * BT0 + kx[2] * BT2) + kx[3] * BT3) + kx[4] * BT4));
        const double zn = y[1] + (h * y[4] + hsq * (((kz[0] * BT0 + kz[2] * BT2) + kz[3] * BT3) + kz[4] * BT4));
        const double epx = hsq * ((((kx[0] * EP0 + kx[2] * EP2) + kx[3] * EP3) + kx[4] * EP4) + kx[5] * EP5);
        const double epz = hsq * ((((kz[0] * EP0 + kz[2] * EP2) + kz[3] * EP3) + kz[4] * EP4) + kz[5] * EP5);
        const double evx = h * (((((kx[0] * E0 + kx[2] * E2) + kx[3] * E3) + kx[4] * E4) + kx[5] * E5) + kx[6] * E6);
        const double evz = h * (((((kz[0] * E0 + kz[2] * E2) + kz[3] * E3) + kz[4] * E4) + kz[5] * E5) + kz[6] * E6);
        const double e0 = epx * rcp1_(fma(absmax_(y[0], xn), rtol, atol)), e1 = epz * rcp1_(fma(absmax_(y[1], zn), rtol, atol));
        const double e3 = evx * rcp1_(fma(absmax_(y[3], svx), rtol, atol)), e4 = evz * rcp1_(fma(absmax_(y[4], svz), rtol, atol));
        const double err = sqrtc_((e0 * e0 + e1 * e1) + (e3 * e3 + e4 * e4)) * INV_RS7;
        if (err < 1) {
            double factor = (err == 0) ? 10.0 : py_min(10.0, 0.9 * powc_m02(err));
            if (rejected) factor = py_min(1.0, factor);
            h_abs *= factor;
            const double g = y[1], g_new = zn;
            if (((g <= 0) && (g_new >= 0)) || ((g >= 0) && (g_new <= 0))) {
                double ebuf[26];
#pragma unroll
                for (int i = 0; i < 7; ++i) { ebuf[i] = y[i]; ebuf[7 + i] = kx[i]; ebuf[14 + i] = kz[i]; }
                ebuf[21] = dom; ebuf[22] = c.dm; ebuf[23] = t; ebuf[24] = h; ebuf[25] = t_new;
                event_path3(ebuf);
#pragma unroll
                for (int i = 0; i < 7; ++i) y[i] = ebuf[i];
                status = 1;
                break;
            }
            y[0] = xn; y[1] = zn;
            y[2] = y[2] + h * (y[5] + (0.5 * h) * dom);
            y[3] = svx; y[4] = svz;
            y[5] = y[5] + h * dom;
            y[6] = y[6] + h * c.dm;
            kx[0] = kx[6]; kz[0] = kz[6];          // FSAL
            t = t_new;
            if (t - t_bound >= 0) { status = 0; break; }
            rejected = false;
            min_step = min_step_at(t);
            if (h_abs < min_step) h_abs = min_step;
        } else {
            h_abs *= py_max(0.2, 0.9 * powc_m02(err));
            rejected = true;
        }
    }
    return status;
}
#endif

// ------------------------------------------------------------------ reset RNG (new; Philox4x32-10)
RLR_DEV void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

template <int N>
RLR_DEV void sample_ic(const RlrParams& p, uint64_t seed, int64_t env, int32_t episode, float (&ic)[N]) {
#pragma unroll 1
    for (int b = 0; b * 2 < N; ++b) {      // rolled: the reset path is rare per thread, keep its code small
        uint32_t c[4] = {(uint32_t)b, (uint32_t)episode, (uint32_t)((uint64_t)env & 0xffffffffu), (uint32_t)((uint64_t)env >> 32)};
        philox4x32_10(c, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = b * 2 + h;
            if (j < N) {
                const uint64_t bits = ((uint64_t)c[2 * h] << 32) | (uint64_t)c[2 * h + 1];
                const double uu = (double)(bits >> 11) * 0x1.0p-53;
                const double lo = (double)p.ic_low[j], hi = (double)p.ic_high[j];
                ic[j] = (float)(lo + (hi - lo) * uu);
            }
        }
    }
}

// Warp-cooperative version of sample_ic for the in-kernel auto-reset: only ~1.4 % of the lanes finish an episode in a given step,
// but a lane that does would walk the N/2 Philox blocks one after the other while the other 31 lanes wait (7.4 k cycles per warp
// with a reset, profiles/r01c).  Here lane b computes block b of the resetting lane's initial condition and the 2 x N/2 values
// are handed over with shuffles: one Philox latency per resetting lane.  Same counters, same bits, same arithmetic as sample_ic.
// Must be called by all 32 lanes (warp-uniform); fills `ic` of the lanes with need == true.
template <int N>
RLR_DEV void sample_ic_warp(const RlrParams& p, uint64_t seed, int64_t env, int32_t episode, bool need, float (&ic)[N]) {
    constexpr int NB = (N + 1) / 2;
    const int lane = threadIdx.x & 31;
    unsigned todo = __ballot_sync(0xffffffffu, need);
    while (todo) {                                   // warp-uniform
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const long long env_s = __shfl_sync(0xffffffffu, (long long)env, src);
        const int ep_s = __shfl_sync(0xffffffffu, episode, src);
        float f0 = 0.0f, f1 = 0.0f;
        if (lane < NB) {
            uint32_t c[4] = {(uint32_t)lane, (uint32_t)ep_s, (uint32_t)((uint64_t)env_s & 0xffffffffu), (uint32_t)((uint64_t)env_s >> 32)};
            philox4x32_10(c, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
            const int j0 = 2 * lane, j1 = (2 * lane + 1 < N) ? 2 * lane + 1 : 2 * lane;
            const uint64_t b0 = ((uint64_t)c[0] << 32) | (uint64_t)c[1], b1 = ((uint64_t)c[2] << 32) | (uint64_t)c[3];
            const double u0 = (double)(b0 >> 11) * 0x1.0p-53, u1 = (double)(b1 >> 11) * 0x1.0p-53;
            const double lo0 = (double)p.ic_low[j0], hi0 = (double)p.ic_high[j0], lo1 = (double)p.ic_low[j1], hi1 = (double)p.ic_high[j1];
            f0 = (float)(lo0 + (hi0 - lo0) * u0);
            f1 = (float)(lo1 + (hi1 - lo1) * u1);
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const float v = __shfl_sync(0xffffffffu, (j & 1) ? f1 : f0, j >> 1);
            if (lane == src) ic[j] = v;
        }
    }
}

RLR_DEV void vtarg3(const RlrParams& p, const float (&s)[7], float v0, double (&vt)[4]);
template <int NO>
RLR_DEV void gaudet_obs3(const float (&s)[7], const double (&vt)[4], float (&obs)[NO]);

// reset from an f32 IC: quaternion normalised in f32 (rocket_env.py:673), v0 = f32 norm of IC velocity
template <class S>
RLR_DEV void reset_from_ic(const RlrParams& p, float (&ic)[S::N], double (&y)[S::N], float& v0, float (&obs)[S::N]) {
    if constexpr (S::N == 14) {
        const float nq = f32_norm4(ic[6], ic[7], ic[8], ic[9]);
#pragma unroll
        for (int i = 6; i < 10; ++i) ic[i] = fdiv(ic[i], nq);
        v0 = f32_norm3(ic[3], ic[4], ic[5]);
    } else {
        v0 = f32_norm2(ic[3], ic[4]);
    }
#pragma unroll
    for (int i = 0; i < S::N; ++i) {
        y[i] = (double)ic[i];
#if RLR_EXACT
        obs[i] = (float)((double)ic[i] / p.normalizer[i]);
#else
        obs[i] = (float)((double)ic[i] * p.inv_normalizer[i]);
#endif
    }
    if constexpr (S::N == 7) {
        if (p.obs_mode == RLR_OBS_GAUDET) {      // ObservationWrapper.reset -> observation(): vtarg of the initial condition
            double vt[4];
            vtarg3(p, (const float(&)[7])ic, v0, vt);
            gaudet_obs3((const float(&)[7])ic, vt, obs);
        }
    }
}

// ------------------------------------------------------------------ reward / termination epilogues
struct StepRes {
    double reward;
    double terms[6];
    double vt[4];            // target velocity (3) and t_go of _compute_vtarg
    int viol, landed, status, done;
};

// Rocket6DOF._compute_reward / _compute_vtarg / _check_attitude_limits / _check_landing / bounds
// (rocket_env.py:825-859, 986-1014, 1036-1061); s = f32 copy of the state, u = f32 physical action.
RLR_DEV void reward6(const RlrParams& p, const float (&s)[14], const float* u, float a_thr, float v0, int status, StepRes& o) {
    double rh0, rh1, rh2, vh0, tau;
    if ((double)s[0] > p.waypoint) {
        rh0 = (double)s[0] - p.waypoint; rh1 = (double)s[1]; rh2 = (double)s[2];
        vh0 = (double)s[3] - (-2.0); tau = 20.0;
    } else {
        rh0 = (double)fadd(s[0], 1.0f); rh1 = 0.0; rh2 = 0.0;
        vh0 = (double)s[3] - (-1.0); tau = 100.0;
    }
    const double vh1 = (double)s[4], vh2 = (double)s[5];
    const double nr = sqrt_((rh0 * rh0 + rh1 * rh1) + rh2 * rh2);
    const double nv = sqrt_((vh0 * vh0 + vh1 * vh1) + vh2 * vh2);
    const double t_go = RLR_DIV(nr, nv);
    const double den = py_max(1e-3, nr);
    const double fac = 1 - exp(RLR_DIV(-t_go, tau));
    const double mv0 = (double)(-v0);
#if RLR_EXACT
    const double vt0 = (mv0 * (rh0 / den)) * fac, vt1 = (mv0 * (rh1 / den)) * fac, vt2 = (mv0 * (rh2 / den)) * fac;
#else
    const double iden = rcp_(den);
    const double vt0 = (mv0 * (rh0 * iden)) * fac, vt1 = (mv0 * (rh1 * iden)) * fac, vt2 = (mv0 * (rh2 * iden)) * fac;
#endif
    const double d0 = (double)s[3] - vt0, d1 = (double)s[4] - vt1, d2 = (double)s[5] - vt2;
    o.vt[0] = vt0; o.vt[1] = vt1; o.vt[2] = vt2; o.vt[3] = t_go;
    // Rotation.from_quat(roll(q_f32,-1)).as_euler("zyx")
    const double q0 = s[6], q1 = s[7], q2 = s[8], q3 = s[9];
#if RLR_EXACT
    const double nq = sqrt(((q1 * q1 + q2 * q2) + q3 * q3) + q0 * q0);
    const double x = q1 / nq, yy = q2 / nq, z = q3 / nq, w = q0 / nq;
#else
    const double rnq = rsqrt_(((q1 * q1 + q2 * q2) + q3 * q3) + q0 * q0);
    const double x = q1 * rnq, yy = q2 * rnq, z = q3 * rnq, w = q0 * rnq;
#endif
    const double R00 = x * x - yy * yy - z * z + w * w, R01 = 2 * (x * yy - z * w), R02 = 2 * (x * z + yy * w);
    const double R12 = 2 * (yy * z - x * w), R22 = -x * x - yy * yy + z * z + w * w;
#if RLR_EXACT
    const double ang0 = atan2(-R01, R00);
    const double ang1 = asin(fmin(1.0, fmax(-1.0, R02)));
    const double ang2 = atan2(-R12, R22);
    const int att_viol = (fabs(ang0) > p.att_traj_limit[0]) | (fabs(ang1) > p.att_traj_limit[1]) | (fabs(ang2) > p.att_traj_limit[2]);
    const int att_ok = (fabs(ang0) < p.att_land_limit[0]) | (fabs(ang1) < p.att_land_limit[1]) | (fabs(ang2) < p.att_land_limit[2]);
#else
    // The Euler angles are only compared with limits: |atan2(b, a)| > L  <=>  a < cos(L) * hypot(a, b)  (cos is decreasing on
    // [0, pi]) and |asin(v)| > L  <=>  |v| > sin(L); cos(L) / sin(L) come precomputed from the host (with sentinels for limits
    // outside [0, pi] resp. [0, pi/2]), so no atan2 / asin is evaluated.
    // a < c * hypot(a, b) without the square root:  sign(a) a^2 < sign(c) c^2 (a^2 + b^2)  (x -> sign(x) x^2 is increasing)
    const double h0sq = R00 * R00 + R01 * R01, h2sq = R22 * R22 + R12 * R12, av1 = fabs(R02);
    const double s00 = R00 * fabs(R00), s22 = R22 * fabs(R22);
    const double t0c = p.att_traj_c[0] * fabs(p.att_traj_c[0]), t2c = p.att_traj_c[2] * fabs(p.att_traj_c[2]);
    const double l0c = p.att_land_c[0] * fabs(p.att_land_c[0]), l2c = p.att_land_c[2] * fabs(p.att_land_c[2]);
    const int att_viol = (s00 < t0c * h0sq) | (av1 > p.att_traj_c[1]) | (s22 < t2c * h2sq);
    const int att_ok = (s00 > l0c * h0sq) | (av1 < p.att_land_c[1]) | (s22 > l2c * h2sq);
#endif
    const int om_ok = ((double)fabsf(s[10]) < p.omega_lim[0]) | ((double)fabsf(s[11]) < p.omega_lim[1]) | ((double)fabsf(s[12]) < p.omega_lim[2]);
    const float rn = f32_norm3(s[0], s[1], s[2]), vn = f32_norm3(s[3], s[4], s[5]);
    const int landed = (s[0] <= 1e-3f) && ((double)vn < p.maximum_velocity) && ((double)rn < p.landing_radius) && att_ok && om_ok;
    int inside = 1;
#pragma unroll
    for (int i = 0; i < 3; ++i) inside = inside && (s[i] >= p.pos_low[i]) && (s[i] <= p.pos_high[i]);
    const int viol = !inside;
    o.terms[0] = p.alfa * sqrt_((d0 * d0 + d1 * d1) + d2 * d2);
    o.terms[1] = (double)fmul((float)p.beta, u[2]);
    o.terms[2] = p.eta;
    o.terms[3] = p.gamma * (double)att_viol;
    o.terms[4] = 0.0;
    o.terms[5] = p.kappa * (double)landed;
    double r = (((o.terms[0] + o.terms[1]) + o.terms[2]) + o.terms[3]) + o.terms[5];
    if (viol) r += -50.0;
    if (p.reward_mode == 1) {               // RewardAnnealing (wrappers.py:68-86)
        const float tp = fmul((float)(-p.xi), fadd(a_thr, 1.0f));
        r = (o.terms[3] + o.terms[5]) + (double)tp;
        o.terms[1] = (double)tp;
    }
    o.reward = r; o.viol = viol; o.landed = landed; o.status = status;
    o.done = (status != 0) || viol;
}

// Rocket._compute_vtarg (rocket_env.py:219-247) on the f32 state: vt = [v_targ_x, v_targ_z, 0, t_go]
RLR_DEV void vtarg3(const RlrParams& p, const float (&s)[7], float v0, double (&vt)[4]) {
    double rh0, rh1, vh1, tau;
    if ((double)s[1] > p.waypoint) {
        rh0 = (double)s[0]; rh1 = (double)s[1] - p.waypoint; vh1 = (double)s[4] - (-2.0); tau = 20.0;
    } else {
        rh0 = 0.0; rh1 = (double)s[1]; vh1 = (double)s[4] - (-1.0); tau = 100.0;
    }
    const double vh0 = (double)s[3];
#if RLR_EXACT
    const double nr = sqrt(rh0 * rh0 + rh1 * rh1);
    const double nv = sqrt(vh0 * vh0 + vh1 * vh1);
    const double t_go = nr / nv;
    const double den = py_max(1e-3, nr);
    const double fac = 1 - exp(-t_go / tau);
    const double mv0 = (double)(-v0);
    vt[0] = (mv0 * (rh0 / den)) * fac; vt[1] = (mv0 * (rh1 / den)) * fac;
#else
    const double nr = sqrt_(rh0 * rh0 + rh1 * rh1);
    const double nv = sqrt_(vh0 * vh0 + vh1 * vh1);
    const double t_go = nr * rcp_(nv);
    const double den = py_max(1e-3, nr);
    const double fac = 1 - exp(-t_go * rcp_(tau));
    const double mv0 = (double)(-v0);
    const double iden = rcp_(den);
    vt[0] = (mv0 * (rh0 * iden)) * fac; vt[1] = (mv0 * (rh1 * iden)) * fac;
#endif
    vt[2] = 0.0; vt[3] = t_go;
}

// GaudetStateObs (wrappers.py:51-66), corrected: np.float32([vx - vtarg_x, vz - vtarg_z, t_go, z]) of the physical f32 state
template <int NO>
RLR_DEV void gaudet_obs3(const float (&s)[7], const double (&vt)[4], float (&obs)[NO]) {
    obs[0] = (float)((double)s[3] - vt[0]);
    obs[1] = (float)((double)s[4] - vt[1]);
    obs[2] = (float)vt[3];
    obs[3] = s[1];
}

// Rocket._compute_reward / _compute_vtarg / _check_landing / _check_bounds (rocket_env.py:177-247, 431-476)
RLR_DEV void reward3(const RlrParams& p, const float (&s)[7], const float* u, float a_thr, bool thr_is_row, float v0, int status, StepRes& o) {
    vtarg3(p, s, v0, o.vt);
    const double d0 = (double)s[3] - o.vt[0], d1 = (double)s[4] - o.vt[1];
    const float zeta = fadd(s[2], -1.5707963267948966f);
    const float az = fabsf(zeta);
    const float rn = f32_norm2(s[0], s[1]), vn = f32_norm2(s[3], s[4]);
    const int landed = (s[1] <= 1e-3f) && (vn < 15.0f) && ((double)rn < p.landing_radius) && (az < 0.2f) && (fabsf(s[5]) < 0.2f);
    const int viol = ((double)s[0] <= p.x_left) || ((double)s[0] >= p.x_right) || ((double)s[1] >= p.y_up);
    const float hint_arg = fadd(az, -1.5707963267948966f);
    float hint = hint_arg > 0.0f ? hint_arg : 0.0f;
    if (hint_arg != hint_arg) hint = hint_arg;
#if RLR_EXACT
    o.terms[0] = p.alfa * sqrt(d0 * d0 + d1 * d1);
#else
    o.terms[0] = p.alfa * sqrt_(d0 * d0 + d1 * d1);
#endif
    o.terms[1] = (double)fmul((float)p.beta, u[1]);
    o.terms[2] = p.eta;
    o.terms[3] = p.gamma * (double)(az > 6.283185307179586f);
    o.terms[4] = (double)fmul((float)p.delta, hint);
    o.terms[5] = p.kappa * (double)landed;
    double r = ((((o.terms[0] + o.terms[1]) + o.terms[2]) + o.terms[3]) + o.terms[4]) + o.terms[5];
    if (viol) r += -50.0;
    if (p.reward_mode == 1) {
        if (thr_is_row) {                   // DiscreteActions3DOF(RewardAnnealing(env)): python float * np.int64 -> float64
            const double tp = (-p.xi) * ((double)a_thr + 1.0);
            r = (o.terms[3] + o.terms[5]) + tp;
            o.terms[1] = tp;
        } else {                            // python float + np.float32 -> f32 arithmetic
            const float tp = fmul((float)(-p.xi), fadd(a_thr, 1.0f));
            const float rr = fadd((float)(o.terms[3] + o.terms[5]), tp);
            r = (double)rr;
            o.terms[1] = (double)tp;
        }
    }
    o.reward = r; o.viol = viol; o.landed = landed; o.status = status;
    o.done = (status != 0) || viol;
}

// ------------------------------------------------------------------ one environment step in registers
template <class S>
struct EnvState {
    double y[S::N];
    int32_t k, episode;
    float v0;
    double ep_ret;
};

// action -> physical f32 u (+ the normalised throttle for RewardAnnealing)
// a_thr: normalised throttle RewardAnnealing sees; thr_is_row: it is a DiscreteActions3DOF table row (an int64 in the reference, so
// the annealed penalty is evaluated in float64, wrappers.py:34, :80)
template <class S>
RLR_DEV void map_action(const RlrParams& p, const void* actions, int action_kind, int64_t i, float (&u)[3], float& a_thr, bool& thr_is_row) {
    a_thr = 0.0f;
    thr_is_row = false;
    if constexpr (S::N == 14) {
        const float* a = (const float*)actions + i * 3;
        const float a0 = a[0], a1 = a[1], a2 = a[2];
        if (action_kind == RLR_ACT_NORMALIZED) {
            u[0] = (float)((double)a0 * p.max_gimbal);
            u[1] = (float)((double)a1 * p.max_gimbal);
            u[2] = fmul(fdiv(fadd(a2, 1.0f), 2.0f), p.max_thrust);
            if (p.min_thrust > 0.0f) u[2] = fmaxf(u[2], p.min_thrust);      // (extension: the reference's minThrust TODO, rocket_env.py:980)
            a_thr = a2;
        } else { u[0] = a0; u[1] = a1; u[2] = a2; }
    } else {
        float a0, a1;
        if (action_kind == RLR_ACT_DISCRETE) {
            long long idx = ((const long long*)actions)[i];
            idx = idx < 0 ? 0 : (idx >= p.n_discrete ? p.n_discrete - 1 : idx);
            u[0] = p.discrete_table[2 * idx]; u[1] = p.discrete_table[2 * idx + 1]; u[2] = 0.0f;
            a_thr = p.discrete_thr[idx];
            thr_is_row = true;
            return;
        }
        if (p.dof == 1) { a0 = 0.0f; a1 = ((const float*)actions)[i]; }
        else { const float* a = (const float*)actions + i * 2; a0 = a[0]; a1 = a[1]; }
        if (action_kind == RLR_ACT_NORMALIZED) {
            u[0] = (float)((double)a0 * p.max_gimbal);
            u[1] = fmul(fdiv(fadd(a1, 1.0f), 2.0f), p.max_thrust);
            if (p.min_thrust > 0.0f) u[1] = fmaxf(u[1], p.min_thrust);      // (extension: rocket_env.py:405)
            a_thr = a1;
        } else { u[0] = a0; u[1] = a1; }
        u[2] = 0.0f;
    }
}

struct StepFlags { int done, trunc, viol, landed, status, natt; };

// Integrate one env step, compute reward/obs/flags.  `st` is updated to the post-step state
// (NOT reset).  obs = observation of the post-step state.
template <class S>
RLR_DEV void env_step(const RlrParams& p, EnvState<S>& st, const float (&u)[3], float a_thr, bool thr_is_row, float (&obs)[S::N], StepRes& res,
                      StepFlags& fl, double* sK) {
    typename S::Ctx c;
    S::make_ctx(c, u);
    const double t = div1000_((long long)st.k * (long long)p.dt_ms);            // round(t + dt, 3)  (simulator.py:245)
    int natt;
    int status;
    if constexpr (S::N == 14) {
#if RLR_PARK
        const Park6 park(sK);
        *park.ep_ret = st.ep_ret; *park.k = st.k; *park.episode = st.episode; *park.v0 = st.v0;
#endif
        status = rk45_solve6(c, t, t + p.dt, st.y, natt, reinterpret_cast<double2*>(sK) + threadIdx.x);
#if RLR_PARK
        st.ep_ret = *park.ep_ret; st.k = *park.k; st.episode = *park.episode; st.v0 = *park.v0;
#endif
    }
#if RLR_EXACT
    else status = rk45_solve<S>(c, t, t + p.dt, st.y, natt);
#else
    else {
        if constexpr (S::AERO) {
            if (st.y[6] == 0.0) { natt = 1; status = -1; }        // zero mass: see rk45_solve6
            else status = rk45_solve<S>(c, t, t + p.dt, st.y, natt);
        } else {
            status = rk45_solve3(c, t, t + p.dt, st.y, natt);
        }
    }
#endif
    float s[S::N];
    if constexpr (S::N == 14) {
        const double n2 = ((st.y[6] * st.y[6] + st.y[7] * st.y[7]) + st.y[8] * st.y[8]) + st.y[9] * st.y[9];
#if RLR_EXACT
        const double nq = sqrt(n2);
#pragma unroll
        for (int i = 6; i < 10; ++i) st.y[i] = st.y[i] / nq;        // simulator.py:250
#pragma unroll
        for (int i = 0; i < S::N; ++i) { s[i] = (float)st.y[i]; obs[i] = (float)(st.y[i] / p.normalizer[i]); }   // rocket_env.py:695, :963-964
#else
        const double rnq = rsqrt_(n2);
#pragma unroll
        for (int i = 6; i < 10; ++i) st.y[i] = st.y[i] * rnq;
#pragma unroll
        for (int i = 0; i < S::N; ++i) { s[i] = (float)st.y[i]; obs[i] = (float)(st.y[i] * p.inv_normalizer[i]); }
#endif
        reward6(p, (const float(&)[14])s, u, a_thr, st.v0, status, res);
    } else {
        const double pi2 = 2.0 * 3.141592653589793;
#if RLR_EXACT
        st.y[2] = fmod(fmod(st.y[2], pi2) + pi2, pi2);                // simulator.py:77, :150-163
#else
        {   // the same two exact remainders without libdevice's fmod: r0 = theta - trunc(theta / 2pi) 2pi is one FMA (exact), the
            // second operand fl(r0 + 2pi) lies in (0, 4pi), where the remainder is a conditional subtraction
            const double th = st.y[2];
            double q = trunc(th * (1.0 / pi2));
            double r0 = fma(-q, pi2, th);
            if (r0 <= -pi2 || (th >= 0 && r0 < 0)) r0 += pi2;        // quotient off by one after the rounded multiply
            else if (r0 >= pi2 || (th < 0 && r0 > 0)) r0 -= pi2;
            const double s1 = r0 + pi2;
            double r = (s1 >= pi2) ? s1 - pi2 : s1;
            if (r >= pi2) r -= pi2;                                   // fl(r0 + 2pi) can round up to exactly 4pi
            st.y[2] = r;
        }
#endif
#pragma unroll
        for (int i = 0; i < S::N; ++i) { s[i] = (float)st.y[i]; obs[i] = (float)((double)s[i] / p.normalizer[i]); }   // rocket_env.py:155,:175,:209-210
        reward3(p, (const float(&)[7])s, u, a_thr, thr_is_row, st.v0, status, res);
        if (p.obs_mode == RLR_OBS_GAUDET) gaudet_obs3(s, res.vt, obs);
    }
    st.k += 1;
    st.ep_ret += res.reward;
    fl.trunc = 0;
    fl.done = res.done;
    if (p.max_episode_steps > 0 && st.k >= p.max_episode_steps) { fl.trunc = !res.done; fl.done = 1; }   // gym TimeLimit
    fl.viol = res.viol; fl.landed = res.landed; fl.status = status; fl.natt = natt;
}

// ------------------------------------------------------------------ kernels
// L2 prefetch of everything a step reads for environment i, issued BEFORE griddepcontrol.wait: under programmatic dependent launch the
// blocks of step t+1 become resident while the last warps of step t are still running (the tail of a launch is ~1/3 of its duration) and
// would only sit at the wait; a prefetch returns no data to the thread, so it is safe even when the predecessor is still writing
// these very lines (L2 is the point of coherence), and the loads after the wait then find their lines in L2 instead of paying the
// HBM latency + the 9.4 MB burst of a cold launch (2.4 us of the 15.6 us an average warp needs).
RLR_DEV void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
// One 128-byte line per thread, ONE prefetch instruction: a full 64-thread 6DOF block reads 14 state components x 512 B (56 lines:
// thread t < 56 takes line t & 3 of component t >> 2), 512 B of running returns (threads 56..59) and 768 B of actions (threads 60..63,
// four of its six lines).  A per-environment prefetch needs 21 instructions per thread, which queue ahead of the real loads and cost
// 0.4 us when the state is L2-resident; a generic segment walk costs more in address arithmetic than it saves; a second instruction
// for the k / episode / v0 slices, or the earliest blocks prefetching for the latest ones, gain nothing (all measured).
RLR_DEV void prefetch_block_inputs6(const RlrState& gs, const void* actions, int64_t n, int64_t i0) {
    const int t = threadIdx.x;
    const char* ptr;
    if (t < 56) ptr = (const char*)((const double*)gs.y + (int64_t)(t >> 2) * n + i0) + (t & 3) * 128;
    else if (t < 60) ptr = (const char*)(gs.ep_ret + i0) + (t - 56) * 128;
    else ptr = (const char*)actions + i0 * 12 + (t - 60) * 192;
    prefetch_l2(ptr);
}
template <class S>
RLR_DEV void load_state(const RlrState& gs, int64_t n, int64_t i, EnvState<S>& st) {
    // ld.global.cg: every element is read once (nothing to gain from L1), and a chained launch (rlr_step_chained) may run on an SM
    // whose L1 was last invalidated before the predecessor block stored these very lines -- L2 is the point of coherence
    const double* y = (const double*)gs.y + i;
#pragma unroll
    for (int c = 0; c < S::N; ++c, y += n) st.y[c] = __ldcg(y);
    st.k = __ldcg(gs.k + i); st.episode = __ldcg(gs.episode + i); st.v0 = __ldcg(gs.v0 + i); st.ep_ret = __ldcg(gs.ep_ret + i);
}
template <class S>
RLR_DEV void store_state(const RlrState& gs, int64_t n, int64_t i, const EnvState<S>& st) {
#if RLR_OPAQUE_STORE_INDEX
    // the store addresses are recomputed from an index the compiler cannot identify with the one the loads used: otherwise it keeps the
    // load-time addresses alive across the whole kernel and, at the register cap, spills them (a local-memory reload = an L2 round trip here)
    asm volatile("" : "+l"(i));
#endif
    double* y = (double*)gs.y + i;
#pragma unroll
    for (int c = 0; c < S::N; ++c, y += n) *y = st.y[c];
    gs.k[i] = st.k; gs.episode[i] = st.episode; gs.v0[i] = st.v0; gs.ep_ret[i] = st.ep_ret;
}

template <class S>
RLR_DEV void write_obs(const RlrParams& p, float* dst, int64_t i, const float (&obs)[S::N]) {
    if (p.dof == 1) { dst[i * 2] = obs[1]; dst[i * 2 + 1] = obs[4]; }
    else if (S::N == 7 && p.obs_mode == RLR_OBS_GAUDET) {
#pragma unroll
        for (int c = 0; c < 4; ++c) dst[i * 4 + c] = obs[c];
    } else {
#pragma unroll
        for (int c = 0; c < S::N; ++c) dst[i * S::N + c] = obs[c];
    }
}

// 6DOF observation rows are 56 bytes: per-lane stores touch 32 different sectors per instruction (14 instructions per row).  When the
// whole warp is active the 32 x 14 floats go through the warp's own (now idle) columns of the stage matrix in shared memory and leave
// as 112 coalesced 16-byte stores.  Falls back to the per-lane stores for partial warps or an unaligned destination.  `stage` is
// set by the launcher when the destination is mapped HOST memory (PCIe writes need full lines: 117 vs 257 us per step); for device
// destinations the per-lane stores are ~2 % faster (L2 merges the partial sectors).
RLR_DEV void write_obs6_warp(const RlrParams& p, float* dst, int64_t i, bool active, const float (&obs)[14], double* sK, bool stage) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* tile_dst = dst + (i - lane) * 14;                  // row of the warp's first environment
    const bool staged = stage && __all_sync(0xffffffffu, active) && ((reinterpret_cast<uintptr_t>(tile_dst) & 15) == 0);
    if (!staged) {
        if (active) write_obs<Sys6>(p, dst, i, obs);
        return;
    }
    float* st = reinterpret_cast<float*>(sK) + w * 128;       // rows are 64 threads x 16 B = 256 floats apart; a warp owns 128 of them
    __syncwarp();                                             // every lane is done with its stage-matrix columns
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const int o4 = lane * 14 + 2 * k;
        *reinterpret_cast<float2*>(st + (o4 >> 7) * 256 + (o4 & 127)) = make_float2(obs[2 * k], obs[2 * k + 1]);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int f = k * 32 + lane;
        if (f < 112) reinterpret_cast<float4*>(tile_dst)[f] = *reinterpret_cast<const float4*>(st + k * 256 + 4 * lane);
    }
    __syncwarp();                                             // the columns may be reused (rollout: next step's K0) only after the reads
}

// Step kernel, 6DOF: observation store + auto-reset of the finished environments, written out COOPERATIVELY.  The reset of a finished
// environment (Philox initial condition -> f32 quaternion normalisation -> state, |v0|, observation; rocket_env.py:665-688) used to run
// in the finishing lane alone -- 14 values handed over with shuffles, then ~300 divergent instructions while 31 lanes wait -- and it
// sits at the very end of the critical chain of the last warps of a launch (ground event -> reward -> reset -> stores).  Here lane b < 7 of a
// group of eight owns components 2b, 2b+1 of one finishing lane's initial condition (its Philox block b; four groups = up to four
// resets per pass), the two norms come from seven shuffles, and the seven lanes store the new state and observation of that environment straight to memory; the finishing lane itself stores
// nothing.  Same counters, bits and arithmetic as sample_ic / reset_from_ic (the reset kernel and the oracle).  The fused rollout
// keeps its state in registers across steps and stays with sample_ic_warp + finish_and_reset.
// `stage`: see write_obs6_warp (host destinations: rows leave as coalesced 16-byte stores through the warp's idle stage-matrix columns).
// YT: element type of the state arrays (double; float for the RLR_F32 kernels).  plane(k), k = 0..3: the warp's k-th 128-float staging
// plane in its idle stage-matrix columns (f64 kernels: rows are 64 threads x 16 B = 256 floats apart and a warp owns 128 of them).
template <class YT, class Plane>
RLR_DEV void store_obs_and_reset6_warp(const RlrParams& p, const RlrState& gs, float* dst, int64_t n, int64_t i, bool active, bool finished,
                                         int32_t next_episode, const float (&obs)[14], Plane plane, bool stage, const float* ic_inject,
                                         uint64_t seed, int64_t env) {
    const int lane = threadIdx.x & 31;
    float* tile_dst = dst + (i - lane) * 14;                  // row of the warp's first environment
    const bool staged = stage && __all_sync(0xffffffffu, active) && ((reinterpret_cast<uintptr_t>(tile_dst) & 15) == 0);
    if (staged) {
        __syncwarp();                                         // every lane is done with its stage-matrix columns
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int o4 = lane * 14 + 2 * k;
            *reinterpret_cast<float2*>(plane(o4 >> 7) + (o4 & 127)) = make_float2(obs[2 * k], obs[2 * k + 1]);
        }
        __syncwarp();
    }
    const bool pair_ok = (reinterpret_cast<uintptr_t>(dst) & 7) == 0;
    if (!staged && active && !finished) {
        write_obs<Sys6>(p, dst, i, obs);
    }
    unsigned todo = __ballot_sync(0xffffffffu, finished);
    const int grp = lane >> 3, sub = lane & 7, lead = lane & 24;      // four groups of eight lanes: up to four resets per pass
    while (todo) {                                            // warp-uniform
        int src = -1;
#pragma unroll
        for (int g = 0; g < 4; ++g) {                         // group g takes the g-th finished lane still to do
            const int b = todo ? __ffs(todo) - 1 : -1;
            if (g == grp) src = b;
            todo &= todo - 1;
        }
        const bool have = src >= 0;
        const int srcc = have ? src : 0;
        const long long env_s = __shfl_sync(0xffffffffu, (long long)env, srcc);
        const long long i_s = __shfl_sync(0xffffffffu, (long long)i, srcc);
        const int ep_s = __shfl_sync(0xffffffffu, next_episode, srcc);
        const bool work = have && sub < 7;                    // lane `sub` of the group owns components 2 sub, 2 sub + 1
        float f0 = 0.0f, f1 = 0.0f;
        if (work) {
            if (ic_inject != nullptr) {
                f0 = ic_inject[i_s * 14 + 2 * sub];
                f1 = ic_inject[i_s * 14 + 2 * sub + 1];
            } else {
                uint32_t c[4] = {(uint32_t)sub, (uint32_t)ep_s, (uint32_t)((uint64_t)env_s & 0xffffffffu), (uint32_t)((uint64_t)env_s >> 32)};
                philox4x32_10(c, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
                const uint64_t b0 = ((uint64_t)c[0] << 32) | (uint64_t)c[1], b1 = ((uint64_t)c[2] << 32) | (uint64_t)c[3];
                const double u0 = (double)(b0 >> 11) * 0x1.0p-53, u1 = (double)(b1 >> 11) * 0x1.0p-53;
                const double lo0 = (double)p.ic_low[2 * sub], hi0 = (double)p.ic_high[2 * sub];
                const double lo1 = (double)p.ic_low[2 * sub + 1], hi1 = (double)p.ic_high[2 * sub + 1];
                f0 = (float)(lo0 + (hi0 - lo0) * u0);
                f1 = (float)(lo1 + (hi1 - lo1) * u1);
            }
        }
        // components within the group: lane 1 = (2, 3), lane 2 = (4, 5), lane 3 = (6, 7), lane 4 = (8, 9)
        const float v3 = __shfl_sync(0xffffffffu, f1, lead + 1), v4 = __shfl_sync(0xffffffffu, f0, lead + 2), v5 = __shfl_sync(0xffffffffu, f1, lead + 2);
        const float q6 = __shfl_sync(0xffffffffu, f0, lead + 3), q7 = __shfl_sync(0xffffffffu, f1, lead + 3);
        const float q8 = __shfl_sync(0xffffffffu, f0, lead + 4), q9 = __shfl_sync(0xffffffffu, f1, lead + 4);
        if (work) {
            if (sub == 3 || sub == 4) {                       // quaternion normalised in f32 (rocket_env.py:673)
                const float nq = f32_norm4(q6, q7, q8, q9);
                f0 = fdiv(f0, nq);
                f1 = fdiv(f1, nq);
            }
            YT* y = (YT*)gs.y + i_s;
            y[(int64_t)(2 * sub) * n] = (YT)f0;
            y[(int64_t)(2 * sub + 1) * n] = (YT)f1;
#if RLR_EXACT
            const float o0 = (float)((double)f0 / p.normalizer[2 * sub]), o1 = (float)((double)f1 / p.normalizer[2 * sub + 1]);
#else
            const float o0 = (float)((double)f0 * p.inv_normalizer[2 * sub]), o1 = (float)((double)f1 * p.inv_normalizer[2 * sub + 1]);
#endif
            if (staged) {
                const int o4 = src * 14 + 2 * sub;
                *reinterpret_cast<float2*>(plane(o4 >> 7) + (o4 & 127)) = make_float2(o0, o1);
            } else if (pair_ok) {
                *reinterpret_cast<float2*>(dst + i_s * 14 + 2 * sub) = make_float2(o0, o1);
            } else {
                dst[i_s * 14 + 2 * sub] = o0;
                dst[i_s * 14 + 2 * sub + 1] = o1;
            }
            if (sub == 0) {
                gs.k[i_s] = 0;
                gs.episode[i_s] = ep_s;
                gs.v0[i_s] = f32_norm3(v3, v4, v5);           // |v| of the initial condition in f32 (vtarg, rocket_env.py:989-991)
                gs.ep_ret[i_s] = 0.0;
            }
        }
    }
    if (staged) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int f = k * 32 + lane;
            if (f < 112) reinterpret_cast<float4*>(tile_dst)[f] = *reinterpret_cast<const float4*>(plane(k) + 4 * lane);
        }
        __syncwarp();
    }
}

// warp reduction of the episode statistics (REDUX for the 32-bit counters, shuffles for the 64-bit fixed-point return sum);
// lane 0 issues one atomic per non-zero slot.
RLR_DEV unsigned long long warp_sum(unsigned long long x) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    return x;
}
RLR_DEV void accumulate_stats(unsigned long long* stats, bool active, const StepFlags& fl, int ep_len, double ep_ret) {
    if (stats == nullptr) return;
    const bool fin = active && fl.done;
    // five 6-bit counters (<= 32 lanes each) in one 32-bit word and the attempt count: two REDUX instructions per step
    const unsigned packed = (active ? 1u : 0u) | ((fin ? 1u : 0u) << 6) | (((fin && fl.landed) ? 1u : 0u) << 12) |
                            (((fin && fl.viol) ? 1u : 0u) << 18) | (((fin && fl.trunc) ? 1u : 0u) << 24);
    const unsigned cnt = __reduce_add_sync(0xffffffffu, packed);
    const unsigned n_att = __reduce_add_sync(0xffffffffu, active ? (unsigned)fl.natt : 0u);
    unsigned long long len = 0ull, ret = 0ull;
    if (cnt & (63u << 6)) {       // warp-uniform: some lane finished an episode
        len = (unsigned long long)__reduce_add_sync(0xffffffffu, fin ? (unsigned)ep_len : 0u);
        ret = warp_sum(fin ? (unsigned long long)(long long)llrint(ep_ret * 1048576.0) : 0ull);
    }
    if ((threadIdx.x & 31) == 0) {
        const unsigned long long n_steps = cnt & 63u, n_done = (cnt >> 6) & 63u, n_land = (cnt >> 12) & 63u;
        const unsigned long long n_viol = (cnt >> 18) & 63u, n_trunc = (cnt >> 24) & 63u;
        if (n_steps) atomicAdd(stats + RLR_STAT_STEPS, n_steps);
        if (n_att) atomicAdd(stats + RLR_STAT_ATTEMPTS, (unsigned long long)n_att);
        if (n_done) {
            atomicAdd(stats + RLR_STAT_DONE, n_done);
            if (n_land) atomicAdd(stats + RLR_STAT_LANDED, n_land);
            if (n_viol) atomicAdd(stats + RLR_STAT_VIOL, n_viol);
            if (n_trunc) atomicAdd(stats + RLR_STAT_TRUNC, n_trunc);
            atomicAdd(stats + RLR_STAT_LEN_SUM, len);
            atomicAdd(stats + RLR_STAT_RET_SUM, ret);
        }
    }
}

// one record of RlrStepOut.fin_records per finished episode (rare path: ~1.4 % of the lanes per step)
template <int NS>
RLR_DEV void append_fin_record(const RlrParams& p, const RlrStepOut& out, int64_t n, int64_t i, double ep_ret, int ep_len, const StepFlags& fl,
                                  const float (&obs)[NS]) {
    const int no = (p.dof == 1) ? 2 : ((NS == 7 && p.obs_mode == RLR_OBS_GAUDET) ? 4 : NS);
    const unsigned slot = atomicAdd(out.fin_count, 1u) % (unsigned)n;      // a ring of n records: the counter need not be zeroed per step
    double* r = out.fin_records + (size_t)slot * (RLR_FIN_HEADER + no);
    r[0] = (double)i; r[1] = ep_ret; r[2] = (double)ep_len; r[3] = (double)fl.trunc;
    r[4] = (double)((fl.viol ? 1 : 0) | (fl.landed ? 2 : 0) | (fl.status == 1 ? 4 : 0) | (fl.status < 0 ? 8 : 0));
    if (p.dof == 1) { r[5] = obs[1]; r[6] = obs[4]; }
    else {
#pragma unroll
        for (int c = 0; c < NS; ++c) if (c < no) r[RLR_FIN_HEADER + c] = obs[c];
    }
}

// finished-episode handling shared by step and rollout kernels: reset in registers from an injected or a warp-sampled IC
template <class S>
RLR_DEV void finish_and_reset(const RlrParams& p, EnvState<S>& st, const float* ic_inject, int64_t i, float (&ic)[S::N], float (&obs)[S::N]) {
    st.episode += 1;
    if (ic_inject != nullptr) {
#pragma unroll
        for (int c = 0; c < S::N; ++c) ic[c] = ic_inject[i * S::N + c];
    }
    reset_from_ic<S>(p, ic, st.y, st.v0, obs);
    st.k = 0;
    st.ep_ret = 0.0;
}

// CHAINED: the rlr_step_chained instantiation (per-block ordering between consecutive launches); the plain instantiation carries none
// of that code (its mere presence -- a block barrier in a not-taken branch -- cost the plain step 0.5 us)
template <class S, bool CHAINED = false>
__global__ void RLR_KERNEL_BOUNDS(S) step_kernel(const __grid_constant__ RlrParams p, RlrState gs, const void* actions, int action_kind,
                                                   const float* ic_inject, RlrStepOut out, int64_t n, int64_t i_begin, int64_t i_end,
                                                   int64_t env_offset, uint64_t seed, bool stage_obs, StepChain chain) {
    // environments [i_begin, i_end) of a batch of n (n is the structure-of-arrays stride; rlr_step passes 0, n)
    extern __shared__ double sK[];
    pdl_launch_dependents();
    const int64_t i = i_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < i_end;
    StepFlags fl = {0, 0, 0, 0, 0, 0};
    int ep_len = 0;
    double ep_ret = 0.0;
    EnvState<S> st;
    st.episode = 0;
    float obs[S::N];
#if RLR_PREFETCH
    if constexpr (S::N == 14) {
        const int64_t i0 = i_begin + (int64_t)blockIdx.x * BS6;
        if (i0 + BS6 <= i_end && action_kind != RLR_ACT_DISCRETE) prefetch_block_inputs6(gs, actions, n, i0);
    }
#endif
    if constexpr (CHAINED) chain_wait(chain); else pdl_wait_for_predecessor();
    RLR_PHASE(-1);

    if (active) {
        load_state<S>(gs, n, i, st);
        float u[3], a_thr;
        bool thr_is_row;
        map_action<S>(p, actions, action_kind, i, u, a_thr, thr_is_row);
        StepRes res;
        RLR_PHASE(0);                   // 0: state / action loads
        env_step<S>(p, st, u, a_thr, thr_is_row, obs, res, fl, sK);
        RLR_PHASE(4);                   // 4: normalise, obs, reward
        ep_len = st.k; ep_ret = st.ep_ret;
        if (out.reward) out.reward[i] = (float)res.reward;
        if (out.reward64) out.reward64[i] = res.reward;
        out.done[i] = (uint8_t)fl.done;
        if (out.truncated) out.truncated[i] = (uint8_t)fl.trunc;
        if (out.natt) out.natt[i] = fl.natt;
        if (out.flags) out.flags[i] = (uint8_t)((fl.viol ? 1 : 0) | (fl.landed ? 2 : 0) | (fl.status == 1 ? 4 : 0) | (fl.status < 0 ? 8 : 0));
        if (out.reward_terms) {
#pragma unroll
            for (int c = 0; c < 6; ++c) out.reward_terms[c * n + i] = res.terms[c];
        }
        if (out.y_pre_reset) {
            double* yp = (double*)out.y_pre_reset;
#pragma unroll
            for (int c = 0; c < S::N; ++c) yp[c * n + i] = st.y[c];
        }
        if (out.vtarg) {
#pragma unroll
            for (int c = 0; c < 4; ++c) out.vtarg[c * n + i] = res.vt[c];
        }
        RLR_PHASE(5);                   // 5: per-step output stores
        if (fl.done) {
            if (out.terminal_obs) write_obs<S>(p, out.terminal_obs, i, obs);
            if (out.fin_ret) out.fin_ret[i] = st.ep_ret;
            if (out.fin_len) out.fin_len[i] = st.k;
            if (out.fin_records) append_fin_record<S::N>(p, out, n, i, st.ep_ret, st.k, fl, obs);
        }
    }
    const bool finished = active && fl.done;
    RLR_PHASE(9);                       // 9: terminal-observation / finished-episode stores
    if constexpr (S::N == 14) {
        // ---- 6DOF: the finished lanes' reset (initial condition, new state, new observation) is computed and stored by the whole warp
        if (active && !finished) store_state<S>(gs, n, i, st);
        RLR_PHASE(6);                   // 6: state stores of the continuing environments
        float* const tile = reinterpret_cast<float*>(sK) + (threadIdx.x >> 5) * 128;
        store_obs_and_reset6_warp<double>(p, gs, out.obs, n, i, active, finished, st.episode + 1, (const float(&)[14])obs,
                                          [tile](int k) { return tile + k * 256; }, stage_obs, ic_inject, seed, env_offset + i);
        RLR_PHASE(8);                   // 8: observation stores + cooperative auto-reset
    } else {
        // ---- auto-reset: the finished lanes' initial conditions are sampled by the whole warp (sample_ic_warp)
        float ic[S::N];
        if (ic_inject == nullptr) sample_ic_warp<S::N>(p, seed, env_offset + i, st.episode + 1, finished, ic);
        RLR_PHASE(8);                   // 8: cooperative initial-condition sampling
        if (active) {
            if (finished) finish_and_reset<S>(p, st, ic_inject, i, ic, obs);
            RLR_PHASE(6);               // 6: auto-reset
            write_obs<S>(p, out.obs, i, obs);
            store_state<S>(gs, n, i, st);
        }
    }
    accumulate_stats(out.stats, active, fl, ep_len, ep_ret);
    if constexpr (CHAINED) chain_publish(chain);
    RLR_PHASE(7);                       // 7: obs / state stores, statistics
#if RLR_WARP_TRACE
    {
        const unsigned m_att = __reduce_max_sync(0xffffffffu, (unsigned)fl.natt);
        const unsigned n_ev = __popc(__ballot_sync(0xffffffffu, fl.status == 1)), n_fin = __popc(__ballot_sync(0xffffffffu, fl.done != 0));
        const unsigned long long w = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if ((threadIdx.x & 31) == 0 && w < 8192) {
            g_warp_trace[w * 16 + 0] = smid_();
            g_warp_trace[w * 16 + 2] = globaltimer_ns();
            g_warp_trace[w * 16 + 3] = m_att | (n_ev << 8) | (n_fin << 16);
        }
    }
#endif
}

template <class S>
__global__ void __launch_bounds__(128) reset_kernel(const __grid_constant__ RlrParams p, RlrState gs, const uint8_t* mask, const float* ic_inject,
                                                    float* obs_out, int64_t n, int64_t env_offset, uint64_t seed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mask != nullptr && mask[i] == 0) return;
    EnvState<S> st;
    st.episode = gs.episode[i];
    float ic[S::N], obs[S::N];
    if (ic_inject != nullptr) {
#pragma unroll
        for (int c = 0; c < S::N; ++c) ic[c] = ic_inject[i * S::N + c];
    } else {
        sample_ic<S::N>(p, seed, env_offset + i, st.episode, ic);
    }
    reset_from_ic<S>(p, ic, st.y, st.v0, obs);
    st.k = 0; st.ep_ret = 0.0;
    if (obs_out) write_obs<S>(p, obs_out, i, obs);
    store_state<S>(gs, n, i, st);
}

// T fused steps: state stays in registers between env steps; per-step outputs go to [T][n][..] rollout buffers
template <class S>
__global__ void RLR_KERNEL_BOUNDS(S) rollout_kernel(const __grid_constant__ RlrParams p, RlrState gs, const void* actions, int action_kind, int T,
                                                      float* obs_buf, float* reward_buf, uint8_t* done_buf, uint8_t* trunc_buf, float* last_obs,
                                                      unsigned long long* stats, int64_t n, int64_t env_offset, uint64_t seed) {
    extern __shared__ double sK[];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    EnvState<S> st;
    st.episode = 0;
    if (active) load_state<S>(gs, n, i, st);
    const int na = (p.dof == 1) ? 1 : S::NA;
    const int no = (p.dof == 1) ? 2 : ((S::N == 7 && p.obs_mode == RLR_OBS_GAUDET) ? 4 : S::NOBS);
    float obs[S::N];
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
        StepFlags fl = {0, 0, 0, 0, 0, 0};
        int ep_len = 0;
        double ep_ret = 0.0;
        if (active) {
            const void* a_t = (action_kind == RLR_ACT_DISCRETE) ? (const void*)((const long long*)actions + (int64_t)t * n)
                                                                : (const void*)((const float*)actions + (int64_t)t * n * na);
            float u[3], a_thr;
            bool thr_is_row;
            map_action<S>(p, a_t, action_kind, i, u, a_thr, thr_is_row);
            StepRes res;
            env_step<S>(p, st, u, a_thr, thr_is_row, obs, res, fl, sK);
            ep_len = st.k; ep_ret = st.ep_ret;
            if (reward_buf) reward_buf[(int64_t)t * n + i] = (float)res.reward;
            if (done_buf) done_buf[(int64_t)t * n + i] = (uint8_t)fl.done;
            if (trunc_buf) trunc_buf[(int64_t)t * n + i] = (uint8_t)fl.trunc;
        }
        float ic[S::N];
        const bool finished = active && fl.done;
        sample_ic_warp<S::N>(p, seed, env_offset + i, st.episode + 1, finished, ic);
        if (active && finished) finish_and_reset<S>(p, st, nullptr, i, ic, obs);
        if (obs_buf) {
            if constexpr (S::N == 14) write_obs6_warp(p, obs_buf + (int64_t)t * n * no, i, active, (const float(&)[14])obs, sK, false);
            else if (active) write_obs<S>(p, obs_buf + (int64_t)t * n * no, i, obs);
        }
        accumulate_stats(stats, active, fl, ep_len, ep_ret);
    }
    if (active) {
        if (last_obs && T > 0) write_obs<S>(p, last_obs, i, obs);
        store_state<S>(gs, n, i, st);
    }
}

#if !RLR_EXACT
#include "rocket_f32.cuh"           // RLR_F32 precision of the 6DOF path (fast build only)
#endif

// ------------------------------------------------------------------ host launchers (called from abi.cu)
// Launch geometry.  6DOF: 64-thread blocks, 30 KB of dynamic shared memory each (stage matrix).  The number of co-resident
// blocks per SM comes from the occupancy calculator; when the grid would end in a nearly empty last wave, residency is
// reduced (by padding the dynamic shared-memory request) to the value that maximises  utilisation * sqrt(resident blocks)
// -- e.g. 65 536 envs on 148 SMs: 7 blocks/SM -> 1024 blocks on 1036 slots, one wave.
struct Geometry { int block; size_t smem; int blocks_per_sm; };
struct GeometryEntry { const void* kern; int64_t n; int dev; Geometry g; };
struct DeviceInfo { int sm_count, smem_per_sm; };

// Launch-geometry state is per DEVICE (cudaFuncSetAttribute applies to the current device only) and guarded by a mutex: ctypes
// releases the GIL around every call, so two host threads (or two batches on two GPUs of one process) may be in here at once.
inline std::mutex& geometry_mutex() { static std::mutex m; return m; }

inline DeviceInfo device_info(int dev) {           // call with geometry_mutex held
    static DeviceInfo info[64] = {};
    DeviceInfo& d = info[dev & 63];
    if (d.sm_count == 0) {
        cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&d.smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        if (d.sm_count <= 0) d.sm_count = 148;
        if (d.smem_per_sm <= 0) d.smem_per_sm = 233472;
    }
    return d;
}

template <typename Kern>
Geometry pick_geometry(Kern kern, int64_t n, int block, size_t smem_min) {
    static GeometryEntry cache[16] = {};
    static int cache_next = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    const void* key = (const void*)kern;
    std::lock_guard<std::mutex> lock(geometry_mutex());
    for (int e = 0; e < 16; ++e)
        if (cache[e].kern == key && cache[e].n == n && cache[e].dev == dev) return cache[e].g;
    const DeviceInfo di = device_info(dev);
    const int sm_count = di.sm_count, smem_per_sm = di.smem_per_sm;
    Geometry best = {block, smem_min, 1};
    int nb_max = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb_max, kern, block, smem_min) != cudaSuccess || nb_max <= 0) nb_max = 1;
    best.blocks_per_sm = nb_max;
    if (smem_min > 0) {
        double best_score = -1.0;
        for (int nb = nb_max; nb >= (nb_max + 1) / 2; --nb) {
            const double cap = (double)sm_count * nb * block;
            const double util = (double)n / (ceil((double)n / cap) * cap);
            const double score = util * sqrt((double)nb);
            if (score > best_score * 1.001) { best_score = score; best.blocks_per_sm = nb; }
        }
        if (best.blocks_per_sm < nb_max) {
            // pad the request so that exactly blocks_per_sm blocks fit (1 KB per block is reserved by the driver)
            size_t padded = ((size_t)smem_per_sm / best.blocks_per_sm - 1024) & ~(size_t)1023;
            if (padded > 227 * 1024) padded = 227 * 1024;
            if (padded > smem_min) best.smem = padded;
        }
        // the opt-in limit is a per-device function attribute: raise it to the cap once per (kernel, device) so that every later
        // padded request on this device is legal whatever batch size asked first
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    }
    cudaGetLastError();
    cache[cache_next] = {key, n, dev, best};
    cache_next = (cache_next + 1) % 16;
    return best;
}

inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

// launch with the programmatic-stream-serialization attribute (see pdl_launch_dependents)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// true when `ptr` is (mapped, pinned) host memory.  A small per-thread cache keeps the driver query off the hot path even when a
// caller alternates between device and host destinations (step_tensor / step_host on one batch).
inline bool is_host_pointer(const void* ptr) {
    struct Entry { const void* ptr; bool host; };
    thread_local Entry cache[4] = {};
    thread_local int next = 0;
    for (int e = 0; e < 4; ++e)
        if (cache[e].ptr == ptr && ptr != nullptr) return cache[e].host;
    cudaPointerAttributes attr;
    bool host = false;
    if (cudaPointerGetAttributes(&attr, ptr) == cudaSuccess) host = (attr.type == cudaMemoryTypeHost);
    cudaGetLastError();
    cache[next] = {ptr, host};
    next = (next + 1) & 3;
    return host;
}

int launch_step(const RlrParams* p, RlrState st, const void* actions, int action_kind, const float* ic_inject, RlrStepOut out, int64_t n,
                int64_t begin, int64_t count, int64_t env_offset, uint64_t seed, cudaStream_t stream, uint32_t* block_seq, uint32_t seq_wait,
                uint32_t seq_publish) {
    const int64_t end = begin + count;
    const StepChain chain = {block_seq, block_seq ? seq_wait : 0u, block_seq ? seq_publish : 0u};
    const StepChain nochain = {nullptr, 0u, 0u};
    if (p->dof == 6) {
        if (chain.block_seq != nullptr) {
            const Geometry g = pick_geometry(step_kernel<Sys6, true>, count, BS6, SMEM6_BYTES);
            launch_pdl(step_kernel<Sys6, true>, grid_for(count, g.block), g.block, g.smem, stream, *p, st, actions, action_kind, ic_inject, out, n,
                       begin, end, env_offset, seed, is_host_pointer(out.obs), chain);
        } else {
            const Geometry g = pick_geometry(step_kernel<Sys6>, count, BS6, SMEM6_BYTES);
            launch_pdl(step_kernel<Sys6>, grid_for(count, g.block), g.block, g.smem, stream, *p, st, actions, action_kind, ic_inject, out, n, begin,
                       end, env_offset, seed, is_host_pointer(out.obs), nochain);
        }
    } else if (p->aero_mode != 0) {
        launch_pdl(step_kernel<Sys3Aero>, grid_for(count, 128), 128, 0, stream, *p, st, actions, action_kind, ic_inject, out, n, begin, end,
                   env_offset, seed, false, nochain);
    } else {
        if (count > 148 * 3 * 128)
            launch_pdl(step_kernel<Sys3Dense>, grid_for(count, 128), 128, 0, stream, *p, st, actions, action_kind, ic_inject, out, n, begin, end,
                       env_offset, seed, false, nochain);
        else
            launch_pdl(step_kernel<Sys3>, grid_for(count, 128), 128, 0, stream, *p, st, actions, action_kind, ic_inject, out, n, begin, end,
                       env_offset, seed, false, nochain);
    }
    return (int)cudaGetLastError();
}

int launch_reset(const RlrParams* p, RlrState st, const uint8_t* mask, const float* ic_inject, float* obs, int64_t n, int64_t env_offset,
                 uint64_t seed, cudaStream_t stream) {
    const int block = 128;
    if (p->dof == 6) reset_kernel<Sys6><<<grid_for(n, block), block, 0, stream>>>(*p, st, mask, ic_inject, obs, n, env_offset, seed);
    else reset_kernel<Sys3><<<grid_for(n, block), block, 0, stream>>>(*p, st, mask, ic_inject, obs, n, env_offset, seed);
    return (int)cudaGetLastError();
}

int launch_rollout(const RlrParams* p, RlrState st, const void* actions, int action_kind, int T, float* obs_buf, float* reward_buf,
                   uint8_t* done_buf, uint8_t* trunc_buf, float* last_obs, unsigned long long* stats, int64_t n, int64_t env_offset,
                   uint64_t seed, cudaStream_t stream) {
    if (p->dof == 6) {
        const Geometry g = pick_geometry(rollout_kernel<Sys6>, n, BS6, SMEM6_BYTES);
        rollout_kernel<Sys6><<<grid_for(n, g.block), g.block, g.smem, stream>>>(*p, st, actions, action_kind, T, obs_buf, reward_buf, done_buf,
                                                                              trunc_buf, last_obs, stats, n, env_offset, seed);
    } else if (p->aero_mode != 0) {
        rollout_kernel<Sys3Aero><<<grid_for(n, 128), 128, 0, stream>>>(*p, st, actions, action_kind, T, obs_buf, reward_buf, done_buf, trunc_buf,
                                                                       last_obs, stats, n, env_offset, seed);
    } else {
        if (n > 148 * 3 * 128)
            rollout_kernel<Sys3Dense><<<grid_for(n, 128), 128, 0, stream>>>(*p, st, actions, action_kind, T, obs_buf, reward_buf, done_buf,
                                                                            trunc_buf, last_obs, stats, n, env_offset, seed);
        else
            rollout_kernel<Sys3><<<grid_for(n, 128), 128, 0, stream>>>(*p, st, actions, action_kind, T, obs_buf, reward_buf, done_buf, trunc_buf,
                                                                       last_obs, stats, n, env_offset, seed);
    }
    return (int)cudaGetLastError();
}

#if !RLR_EXACT
// RLR_F32 precision (6DOF only): 128-thread blocks, 27 KB of dynamic shared memory for the binary32 stage matrix
int launch_step_f32(const RlrParams* p, RlrState st, const void* actions, int action_kind, const float* ic_inject, RlrStepOut out, int64_t n,
                    int64_t begin, int64_t count, int64_t env_offset, uint64_t seed, cudaStream_t stream) {
    launch_pdl(f32::step_kernel_f32, grid_for(count, f32::BSF), f32::BSF, f32::SMEMF_BYTES, stream, *p, st, (const float*)actions, action_kind,
               ic_inject, out, n, begin, begin + count, env_offset, seed, is_host_pointer(out.obs));
    return (int)cudaGetLastError();
}
int launch_reset_f32(const RlrParams* p, RlrState st, const uint8_t* mask, const float* ic_inject, float* obs, int64_t n, int64_t env_offset,
                     uint64_t seed, cudaStream_t stream) {
    f32::reset_kernel_f32<<<grid_for(n, 128), 128, 0, stream>>>(*p, st, mask, ic_inject, obs, n, env_offset, seed);
    return (int)cudaGetLastError();
}
int launch_rollout_f32(const RlrParams* p, RlrState st, const void* actions, int action_kind, int T, float* obs_buf, float* reward_buf,
                       uint8_t* done_buf, uint8_t* trunc_buf, float* last_obs, unsigned long long* stats, int64_t n, int64_t env_offset,
                       uint64_t seed, cudaStream_t stream) {
    f32::rollout_kernel_f32<<<grid_for(n, f32::BSF), f32::BSF, f32::SMEMF_BYTES, stream>>>(*p, st, (const float*)actions, action_kind, T, obs_buf,
                                                                                            reward_buf, done_buf, trunc_buf, last_obs, stats, n,
                                                                                            env_offset, seed);
    return (int)cudaGetLastError();
}
#endif

}  // namespace RLR_NS
