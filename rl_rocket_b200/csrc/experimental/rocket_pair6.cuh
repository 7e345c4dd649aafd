// rocket_pair6.cuh -- warp-specialised 6DOF step / rollout kernels of the fast build (included from rocket_kernels.cuh,
// inside namespace RLR_NS, RLR_EXACT == 0 only).
//
// Why.  With one thread per environment the 65 536 environments of BASELINE config 3 are 14 warps per SM: the register file
// (128 registers x 448 threads) and shared memory (the RK stage matrix, 480 B per environment) are both full, and 14 warps
// cannot hide the latency of a 5 400-instruction dependent chain (profiles/r01b: FP64 pipe 35 % busy, issue slots 40 %).
// Here every environment is integrated by TWO threads in two different warps of the same block:
//     role A ("translation"): r, v      -- stage sums of v, thrust rotation R(q) Tb / (m |q|^2), position quadrature, vtarg reward
//     role B ("rotation")   : q, w, m   -- stage sums of q and w, quaternion kinematics, Euler equation, attitude limits
// so the same 443 environments per SM are 28 warps with half-length chains, each warp running one role's straight-line
// code (no divergence between roles, half the instruction footprint per warp).  B never needs anything from A inside an RK
// attempt; A needs the stage quaternion and m |q|^2 from B (5 doubles per stage).  The partial sums of the error norm and of
// select_initial_step are exchanged once per attempt.  The step-size controller is evaluated redundantly by both roles from
// bit-identical inputs with operations that leave the compiler no contraction freedom, so both take the same decisions.
//
// Messages go through two single-slot mailboxes per environment group in shared memory, handed over with named barriers
// (bar.arrive by the producer warp, bar.sync by the consumer warp -- 64 threads per barrier, 4 barriers per pair of warps).
//
// Reference call path reproduced: the same as rocket_kernels.cuh (Rocket6DOF.step rocket_env.py:690-719 ->
// Simulator6DOF.step simulator.py:227-257 -> solve_ivp/RK45).  Algebraic rearrangements of this file (all within the fast
// build's tolerance, tests/goldutil.py): R(q) v as (w^2 - u.u) v + 2 (u.v) u + 2 w (u x v); d(omega_x)/dt == 0 because
// J_y == J_z (simulator.py:216-222), so omega_x is a per-step constant; the mass stage value is m + c_s h dm; the error
// weights sum to zero, so the mass contributes nothing to the error norm.
#pragma once

namespace pair6 {

constexpr int ENVS = 32;                    // environments per block
constexpr int THREADS = 64;                 // 2 warps: one (A, B) pair
// shared-memory layout per block, in doubles (every array is [slot][lane], conflict-free)
constexpr int OFF_KA2 = 0;                  // double2 [6][32]      dv_x, dv_y of stages 0..5 (row 0 unused: K0 stays in registers)
constexpr int OFF_KA1 = 384;                // double  [6][32]      dv_z                              (A private)
constexpr int OFF_KB = 576;                 // double2 [6][3][32]   (dq0,dq1) (dq2,dq3) (dw_y,dw_z)   (B private)
constexpr int OFF_SBQ = 1728;               // double2 [2][32]      mailbox B -> A: stage quaternion / partial sums
constexpr int OFF_SB1 = 1856;               // double  [32]         mailbox B -> A: m |q|^2 / flags
constexpr int OFF_SA = 1888;                // double  [2][32]      mailbox A -> B
constexpr int BLOCK_DOUBLES = 1952;         // 61 doubles per environment
constexpr size_t SMEM_BYTES = BLOCK_DOUBLES * sizeof(double);          // 15 616 B per block -> 14 blocks (448 envs, 28 warps) per SM

__constant__ double c_RKC[7] = {0.0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0};      // nodes c_s of the tableau (rk.py:282)
constexpr double INV_RS14 = 0.2672612419124243846;     // 1 / sqrt(14)
constexpr double RTOL = 1e-3, ATOL = 1e-6;

// Named barriers with IMMEDIATE ids: with ids in registers ptxas reserves all 16 barriers per block, and an SM has 64 -> 4 blocks.
template <int ID> RLR_DEV void bar_sync() { asm volatile("bar.sync %0, 64;" ::"n"(ID) : "memory"); }
template <int ID> RLR_DEV void bar_arrive() { asm volatile("bar.arrive %0, 64;" ::"n"(ID) : "memory"); }

// Mailbox accesses are asm volatile: another warp writes these words, so the compiler must neither hoist nor merge them
// (plain loads through a const / restrict pointer were treated as loop invariant).
RLR_DEV double2 mb_ld2(const double* p) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
RLR_DEV double mb_ld1(const double* p) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
RLR_DEV void mb_st2(double* p, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "d"(x), "d"(y) : "memory");
}
RLR_DEV void mb_st1(double* p, double x) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "d"(x) : "memory");
}

// one direction of the A <-> B exchange: a single-slot mailbox guarded by a full / empty barrier pair
template <int FULL, int EMPTY>
struct Chan {
    // (__syncwarp: the .aligned barrier instructions need the whole warp converged)
    RLR_DEV static void begin_send() { __syncwarp(); bar_sync<EMPTY>(); }        // wait until the previous message was consumed
    RLR_DEV static void end_send() { __syncwarp(); bar_arrive<FULL>(); }
    RLR_DEV static void begin_recv() { __syncwarp(); bar_sync<FULL>(); }
    RLR_DEV static void end_recv() { __syncwarp(); bar_arrive<EMPTY>(); }
};
using ChanBA = Chan<0, 1>;      // B -> A
using ChanAB = Chan<2, 3>;      // A -> B

// ---- scalar step-size logic, evaluated by BOTH roles: only single operations, no a*b+c patterns the compiler could contract
struct Ctl {
    double t, t_bound, h_abs, h, t_new;
    int natt, rejected, status, running;
};

RLR_DEV double sel_h0(double a0, double a1, double interval, double& d1) {
    const double d0 = __dmul_rn(sqrt_(a0), INV_RS14);
    d1 = __dmul_rn(sqrt_(a1), INV_RS14);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : __dmul_rn(__dmul_rn(0.01, d0), rcp_(d1));
    return py_min(h0, interval);
}
RLR_DEV double sel_habs(double a2, double d1, double h0, double interval) {
    const double d2 = __dmul_rn(__dmul_rn(sqrt_(a2), INV_RS14), rcp_(h0));
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = py_max(1e-6, __dmul_rn(h0, 1e-3));
    else h1 = pow_p02(__dmul_rn(0.01, rcp_(py_max(d1, d2))));
    return py_min(py_min(__dmul_rn(100.0, h0), h1), interval);
}
// start of an RK attempt (rk.py:111-127)
constexpr int MAX_ATTEMPTS = 100000;        // a non-finite step size would loop forever (as scipy does); end the episode instead
RLR_DEV void ctl_begin(Ctl& c) {
    if (c.running && (c.h_abs < min_step_at(c.t) || c.natt >= MAX_ATTEMPTS)) { c.status = -1; c.running = 0; }
    double t_new = __dadd_rn(c.t, c.h_abs);
    if (__dsub_rn(t_new, c.t_bound) > 0) t_new = c.t_bound;
    c.t_new = t_new;
    c.h = __dsub_rn(t_new, c.t);
    if (c.running) { c.h_abs = fabs(c.h); ++c.natt; }
}
// end of an attempt (rk.py:143-176 + the event test of ivp.py:678-699).  Returns 0 rejected, 1 accepted, 2 accepted with a ground event
RLR_DEV int ctl_end(Ctl& c, double ssum, int sign_change) {
    if (!c.running) return 0;
    const double err = __dmul_rn(sqrt_(ssum), INV_RS14);
    if (err < 1) {
        double factor = (err == 0) ? 10.0 : py_min(10.0, __dmul_rn(0.9, pow_m02(err)));
        if (c.rejected) factor = py_min(1.0, factor);
        c.h_abs = __dmul_rn(c.h_abs, factor);
        if (sign_change) { c.status = 1; c.running = 0; return 2; }
        c.t = c.t_new;
        if (__dsub_rn(c.t, c.t_bound) >= 0) { c.status = 0; c.running = 0; return 1; }
        c.rejected = 0;
        const double ms = min_step_at(c.t);
        if (c.h_abs < ms) c.h_abs = ms;
        return 1;
    }
    c.h_abs = __dmul_rn(c.h_abs, py_max(0.2, __dmul_rn(0.9, pow_m02(err))));
    c.rejected = 1;
    return 0;
}

RLR_DEV double rcp_scale(double a, double b) {            // 1 / (atol + max(|a|, |b|) rtol)
    const double fa = fabs(a), fb = fabs(b);
    return rcp_(fma((fa > fb) ? fa : fb, RTOL, ATOL));
}

// dense-output helpers (rk.py:554-565): P[.][0] = (1,0,..), P[1][.] = 0
#define P6_Q(k0, k2, k3, k4, k5, k6, cc) \
    ((((((k0) * c_P[0][cc] + (k2) * c_P[2][cc]) + (k3) * c_P[3][cc]) + (k4) * c_P[4][cc]) + (k5) * c_P[5][cc]) + (k6) * c_P[6][cc])
#define P6_Y(yold, k0, k2, k3, k4, k5, k6) \
    (h * ((((k0) * p1 + P6_Q(k0, k2, k3, k4, k5, k6, 1) * p2) + P6_Q(k0, k2, k3, k4, k5, k6, 2) * p3) + P6_Q(k0, k2, k3, k4, k5, k6, 3) * p4) + (yold))

// =============================================================================================== role A: r, v
struct StA {
    double r[3], v[3], ep_ret;
    float v0;
    int k, episode;
};
struct CtxA { double Tb0, Tb1, Tb2; };

// dv/dt = R(q) Tb / (m |q|^2) + g
RLR_DEV void dv_a(const CtxA& c, const double2 qa, const double2 qb, double mn2, double (&d)[3]) {
    const double w = qa.x, x = qa.y, y = qb.x, z = qb.y;
    const double inv = rcp_(mn2);
    const double a = w * w - ((x * x + y * y) + z * z);
    const double b = 2.0 * ((x * c.Tb0 + y * c.Tb1) + z * c.Tb2);
    const double w2 = w + w;
    const double cx = y * c.Tb2 - z * c.Tb1, cy = z * c.Tb0 - x * c.Tb2, cz = x * c.Tb1 - y * c.Tb0;
    const double F0 = (a * c.Tb0 + b * x) + w2 * cx;
    const double F1 = (a * c.Tb1 + b * y) + w2 * cy;
    const double F2 = (a * c.Tb2 + b * z) + w2 * cz;
    d[0] = inv * F0 + (-9.81);
    d[1] = inv * F1;
    d[2] = inv * F2;
}

// Terminal event, A side: root of the height quartic, then r(t_e), v(t_e).  Returns x = (t_e - t) / h.
__device__ __noinline__ double event_a(const double* __restrict__ gp, int lane, double* __restrict__ buf) {
    // buf = [r(3) v(3) | k0(3) | k6(3) | v_new(3) | t h t_new]
    const double t = buf[15], h = buf[16], t_new = buf[17];
    const double2* KA2 = reinterpret_cast<const double2*>(gp + OFF_KA2) + lane;
    const double* KA1 = gp + OFF_KA1 + lane;
    double K[6][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) K[0][i] = buf[6 + i];
#pragma unroll
    for (int j = 1; j <= 5; ++j) { const double2 k2 = KA2[j * 32]; K[j][0] = k2.x; K[j][1] = k2.y; K[j][2] = KA1[j * 32]; }
    double vs[6][3];                         // stage velocities 2..5 (K_1 has zero weight in P)
#pragma unroll
    for (int s = 2; s <= 5; ++s) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double a3 = K[0][i] * c_RKA[s][0];
#pragma unroll
            for (int j = 1; j < s; ++j) a3 += K[j][i] * c_RKA[s][j];
            vs[s][i] = buf[3 + i] + a3 * h;
        }
    }
    const double vx0 = buf[3], vx6 = buf[12];
    const double hq0 = vx0, hq1 = P6_Q(vx0, vs[2][0], vs[3][0], vs[4][0], vs[5][0], vx6, 1),
                 hq2 = P6_Q(vx0, vs[2][0], vs[3][0], vs[4][0], vs[5][0], vx6, 2),
                 hq3 = P6_Q(vx0, vs[2][0], vs[3][0], vs[4][0], vs[5][0], vx6, 3);
    // Newton from the secant point on g(x) = r0 + h x (q0 + x (q1 + x (q2 + x q3))); Brent (scipy's brentq) as the fallback
    double xx;
    {
        const double g0 = buf[0];
        const double g1 = h * (((hq0 + hq1) + hq2) + hq3) + g0;
        double x = RLR_DIV(g0, g0 - g1);
        double gx;
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
#pragma unroll
    for (int i = 0; i < 3; ++i) buf[i] = P6_Y(buf[i], buf[3 + i], vs[2][i], vs[3][i], vs[4][i], vs[5][i], buf[12 + i]);   // uses the OLD v
#pragma unroll
    for (int i = 0; i < 3; ++i) buf[3 + i] = P6_Y(buf[3 + i], K[0][i], K[2][i], K[3][i], K[4][i], K[5][i], buf[9 + i]);
    return xx;
}

// One environment step, A side.  obs[0..5] = observation of the post-step r, v.
RLR_DEV void env_step_a(const RlrParams& p, StA& s, const float (&u)[3], float a_thr, double* gp, int lane,
                        float (&obs)[6], StepRes& res, StepFlags& fl) {
    using rx = ChanBA;
    using tx = ChanAB;
    double2* KA2 = reinterpret_cast<double2*>(gp + OFF_KA2) + lane;
    double* KA1 = gp + OFF_KA1 + lane;
    const double* SBQ0 = gp + OFF_SBQ + 2 * lane;           // mailbox B -> A: two double2 slots and one double
    const double* SBQ1 = gp + OFF_SBQ + 64 + 2 * lane;
    const double* SB1 = gp + OFF_SB1 + lane;
    double* SA0 = gp + OFF_SA + lane;                       // mailbox A -> B: two doubles
    double* SA1 = gp + OFF_SA + 32 + lane;
    CtxA c;
    {
        const float cy = np_cosf(u[0]), sy = np_sinf(u[0]), cz = np_cosf(u[1]), sz = np_sinf(u[1]);
        const double T = (double)u[2];
        c.Tb0 = (double)fmul(cy, cz) * T;        // first column of _rot_mat_thrust_to_body times thrust (simulator.py:316-318)
        c.Tb1 = (double)fmul(sy, cz) * T;
        c.Tb2 = (double)sz * T;
    }
    Ctl ctl;
    ctl.t = (double)((long long)s.k * (long long)p.dt_ms) / 1000.0;      // round(t + dt, 3)  (simulator.py:245)
    ctl.t_bound = ctl.t + p.dt;
    ctl.natt = 0; ctl.rejected = 0; ctl.status = 0; ctl.running = 1;
    double k0[3];                            // K0 = f(y) of the v rows (FSAL), kept in registers
    {
        rx::begin_recv();
        const double2 qa = mb_ld2(SBQ0), qb = mb_ld2(SBQ1);
        const double mn2 = mb_ld1(SB1);
        rx::end_recv();
        dv_a(c, qa, qb, mn2, k0);
    }
    // ---- select_initial_step (common.py:68-134): partial sums over r, v
    {
        const double interval = fabs(ctl.t_bound - ctl.t);
        double isc[6];
#pragma unroll
        for (int i = 0; i < 3; ++i) { isc[i] = rcp_(fma(fabs(s.r[i]), RTOL, ATOL)); isc[3 + i] = rcp_(fma(fabs(s.v[i]), RTOL, ATOL)); }
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) { const double e = s.r[i] * isc[i]; a0 += e * e; }
#pragma unroll
        for (int i = 0; i < 3; ++i) { const double e = s.v[i] * isc[3 + i]; a0 += e * e; }
#pragma unroll
        for (int i = 0; i < 3; ++i) { const double e = s.v[i] * isc[i]; a1 += e * e; }
#pragma unroll
        for (int i = 0; i < 3; ++i) { const double e = k0[i] * isc[3 + i]; a1 += e * e; }
        tx::begin_send(); mb_st1(SA0, a0); mb_st1(SA1, a1); tx::end_send();
        rx::begin_recv();
        const double2 ab = mb_ld2(SBQ0);
        rx::end_recv();
        double d1;
        const double h0 = sel_h0(__dadd_rn(a0, ab.x), __dadd_rn(a1, ab.y), interval, d1);
        double v1[3], f1[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) v1[i] = s.v[i] + h0 * k0[i];
        rx::begin_recv();
        const double2 qa = mb_ld2(SBQ0), qb = mb_ld2(SBQ1);
        const double mn2 = mb_ld1(SB1);
        rx::end_recv();
        dv_a(c, qa, qb, mn2, f1);
        double a2 = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) { const double e = (v1[i] - s.v[i]) * isc[i]; a2 += e * e; }
#pragma unroll
        for (int i = 0; i < 3; ++i) { const double e = (f1[i] - k0[i]) * isc[3 + i]; a2 += e * e; }
        tx::begin_send(); mb_st1(SA0, a2); tx::end_send();
        rx::begin_recv();
        const double a2b = mb_ld2(SBQ0).x;
        rx::end_recv();
        ctl.h_abs = sel_habs(__dadd_rn(a2, a2b), d1, h0, interval);
        const double ms = min_step_at(ctl.t);
        if (ctl.h_abs < ms) ctl.h_abs = ms;
    }
    // ---- RK45 attempts; every lane of the warp walks the same number of attempts as its slowest lane (commits are guarded)
#pragma unroll 1
    for (;;) {
        ctl_begin(ctl);
        if (!__any_sync(0xffffffffu, ctl.running)) break;
        const double h = ctl.h;
        double aB[3], aE[3], eacc[3], sv[3], kv[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { aB[i] = s.v[i] * B0; aE[i] = s.v[i] * E0; eacc[i] = k0[i] * E0; }
        // rk_step (rk.py:14-71), rolled: both roles' stage bodies have to stay resident in the small instruction caches
#pragma unroll 1
        for (int st = 1; st <= 6; ++st) {
            double acc[3];
            {
                const double a = c_RKA[st][0];
#pragma unroll
                for (int i = 0; i < 3; ++i) acc[i] = k0[i] * a;
            }
#pragma unroll 1
            for (int j = 1; j < st; ++j) {
                const double aj = c_RKA[st][j];
                const double2 k2 = KA2[j * 32];
                const double k1 = KA1[j * 32];
                acc[0] += k2.x * aj; acc[1] += k2.y * aj; acc[2] += k1 * aj;
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) sv[i] = s.v[i] + acc[i] * h;
            rx::begin_recv();
            const double2 qa = mb_ld2(SBQ0), qb = mb_ld2(SBQ1);
            const double mn2 = mb_ld1(SB1);
            rx::end_recv();
            if (st <= 5) {
                const double bs = c_RKB[st], es = c_RKE[st];
#pragma unroll
                for (int i = 0; i < 3; ++i) { aB[i] += sv[i] * bs; aE[i] += sv[i] * es; }
            }
            dv_a(c, qa, qb, mn2, kv);
            if (st <= 5) {
                KA2[st * 32] = make_double2(kv[0], kv[1]);
                KA1[st * 32] = kv[2];
                const double es = c_RKE[st];                   // E1 == 0: adds an exact zero
#pragma unroll
                for (int i = 0; i < 3; ++i) eacc[i] += kv[i] * es;
            }
        }
        // ---- error norm, partial sum over r, v (rk.py:101-107)
        double rn[3], ssum = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            rn[i] = s.r[i] + h * aB[i];
            const double e = ((aE[i] + sv[i] * E6) * h) * rcp_scale(s.r[i], rn[i]);
            ssum += e * e;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double e = ((eacc[i] + kv[i] * E6) * h) * rcp_scale(s.v[i], sv[i]);
            ssum += e * e;
        }
        const double g = s.r[0], g_new = rn[0];
        const int sign_change = (((g <= 0) && (g_new >= 0)) || ((g >= 0) && (g_new <= 0))) ? 1 : 0;
        tx::begin_send(); mb_st1(SA0, ssum); mb_st1(SA1, (double)sign_change); tx::end_send();
        rx::begin_recv();
        const double ssum_b = mb_ld2(SBQ0).x;
        rx::end_recv();
        const double t_old = ctl.t;
        const int was_running = ctl.running;
        const int verdict = ctl_end(ctl, __dadd_rn(ssum, ssum_b), sign_change);
        const int ev = (was_running && verdict == 2) ? 1 : 0;
        if (__any_sync(0xffffffffu, ev)) {
            // ---- terminal ground event (rare): A finds the root and sends x = (t_e - t) / h to B
            double xx = 0.0;
            if (ev) {
                double ebuf[18];
#pragma unroll
                for (int i = 0; i < 3; ++i) { ebuf[i] = s.r[i]; ebuf[3 + i] = s.v[i]; ebuf[6 + i] = k0[i]; ebuf[9 + i] = kv[i]; ebuf[12 + i] = sv[i]; }
                ebuf[15] = t_old; ebuf[16] = h; ebuf[17] = ctl.t_new;
                xx = event_a(gp, lane, ebuf);
#pragma unroll
                for (int i = 0; i < 3; ++i) { s.r[i] = ebuf[i]; s.v[i] = ebuf[3 + i]; }
            }
            tx::begin_send(); mb_st1(SA0, xx); tx::end_send();
        }
        if (was_running && verdict == 1) {
#pragma unroll
            for (int i = 0; i < 3; ++i) { s.r[i] = rn[i]; s.v[i] = sv[i]; k0[i] = kv[i]; }     // FSAL: f(y_new) becomes K0
        }
    }
    // ---- reward, termination (rocket_env.py:825-859, 986-1014, 1036-1061)
    float sr[3], sw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        sr[i] = (float)s.r[i]; sw[i] = (float)s.v[i];
        obs[i] = (float)(s.r[i] * p.inv_normalizer[i]);
        obs[3 + i] = (float)(s.v[i] * p.inv_normalizer[3 + i]);
    }
    double rh0, rh1, rh2, vh0, tau;
    if ((double)sr[0] > p.waypoint) {
        rh0 = (double)sr[0] - p.waypoint; rh1 = (double)sr[1]; rh2 = (double)sr[2];
        vh0 = (double)sw[0] - (-2.0); tau = 20.0;
    } else {
        rh0 = (double)fadd(sr[0], 1.0f); rh1 = 0.0; rh2 = 0.0;
        vh0 = (double)sw[0] - (-1.0); tau = 100.0;
    }
    const double vh1 = (double)sw[1], vh2 = (double)sw[2];
    const double nr = sqrt_((rh0 * rh0 + rh1 * rh1) + rh2 * rh2);
    const double nv = sqrt_((vh0 * vh0 + vh1 * vh1) + vh2 * vh2);
    const double t_go = RLR_DIV(nr, nv);
    const double den = py_max(1e-3, nr);
    const double fac = 1 - exp(RLR_DIV(-t_go, tau));
    const double mv0 = (double)(-s.v0);
    const double iden = rcp_(den);
    const double d0 = (double)sw[0] - (mv0 * (rh0 * iden)) * fac;
    const double d1 = (double)sw[1] - (mv0 * (rh1 * iden)) * fac;
    const double d2 = (double)sw[2] - (mv0 * (rh2 * iden)) * fac;
    const float rnorm = f32_norm3(sr[0], sr[1], sr[2]), vnorm = f32_norm3(sw[0], sw[1], sw[2]);
    int inside = 1;
#pragma unroll
    for (int i = 0; i < 3; ++i) inside = inside && (sr[i] >= p.pos_low[i]) && (sr[i] <= p.pos_high[i]);
    const int viol = !inside;
    rx::begin_recv();
    const int bflags = (int)mb_ld1(SB1);          // bit0 attitude-trajectory violation, bit1 landing attitude ok, bit2 omega ok
    rx::end_recv();
    const int att_viol = bflags & 1, att_ok = (bflags >> 1) & 1, om_ok = (bflags >> 2) & 1;
    const int landed = (sr[0] <= 1e-3f) && ((double)vnorm < p.maximum_velocity) && ((double)rnorm < p.landing_radius) && att_ok && om_ok;
    res.terms[0] = p.alfa * sqrt_((d0 * d0 + d1 * d1) + d2 * d2);
    res.terms[1] = (double)fmul((float)p.beta, u[2]);
    res.terms[2] = p.eta;
    res.terms[3] = p.gamma * (double)att_viol;
    res.terms[4] = 0.0;
    res.terms[5] = p.kappa * (double)landed;
    double r = (((res.terms[0] + res.terms[1]) + res.terms[2]) + res.terms[3]) + res.terms[5];
    if (viol) r += -50.0;
    if (p.reward_mode == 1) {               // RewardAnnealing (wrappers.py:68-86)
        const float tp = fmul((float)(-p.xi), fadd(a_thr, 1.0f));
        r = (res.terms[3] + res.terms[5]) + (double)tp;
        res.terms[1] = (double)tp;
    }
    res.reward = r; res.viol = viol; res.landed = landed; res.status = ctl.status;
    res.done = (ctl.status != 0) || viol;
    s.k += 1;
    s.ep_ret += r;
    fl.trunc = 0;
    fl.done = res.done;
    if (p.max_episode_steps > 0 && s.k >= p.max_episode_steps) { fl.trunc = !res.done; fl.done = 1; }      // gym TimeLimit
    fl.viol = viol; fl.landed = landed; fl.status = ctl.status; fl.natt = ctl.natt;
    tx::begin_send(); mb_st1(SA0, (double)fl.done); tx::end_send();
}

// =============================================================================================== role B: q, w, m
struct StB {
    double q[4], w[3], m;
    int k, episode;
};
struct CtxB { double dm, ct1, ct2, hw0; };         // hw0 = omega_x / 2: omega_x is constant over the step (J_y == J_z)

RLR_DEV void dqw_b(const CtxB& c, const double (&q)[4], double wy, double wz, double (&d)[6]) {
    // dq = 0.5 * OMEGA(w) . q with the un-normalised q (simulator.py:287, :362-370)
    // dw = Jinv (tau - w x (J w)) with J = diag(J0, J1, J1): component 0 vanishes, the others are (J0 - J1) wx wz, (J1 - J0) wx wy
    constexpr double J0 = 75350.25, J1 = 6037675.13, J2 = 6037675.13;
    const double hw0 = c.hw0;
    const double hy = 0.5 * wy, hz = 0.5 * wz;
    d[0] = ((-hw0) * q[1] + (-hy) * q[2]) + (-hz) * q[3];
    d[1] = (hw0 * q[0] + hz * q[2]) + (-hy) * q[3];
    d[2] = (hy * q[0] + (-hz) * q[1]) + hw0 * q[3];
    d[3] = (hz * q[0] + hy * q[1]) + (-hw0) * q[2];
    d[4] = c.ct1 - ((2.0 * (J0 - J2) / J1) * hw0) * wz;
    d[5] = c.ct2 - ((2.0 * (J1 - J0) / J2) * hw0) * wy;
}

__device__ __noinline__ void event_b(const double* __restrict__ gp, int lane, double* __restrict__ buf) {
    // buf = [q(4) wy wz m | k6(6) | dm h xx]; overwritten with the state at the event
    const double dm = buf[13], h = buf[14], xx = buf[15];
    const double2* KB = reinterpret_cast<const double2*>(gp + OFF_KB) + lane;
    const double p1 = xx, p2 = p1 * xx, p3 = p2 * xx, p4 = p3 * xx;
#pragma unroll
    for (int pp = 0; pp < 3; ++pp) {
        const double2 k0 = KB[(0 * 3 + pp) * 32], k2 = KB[(2 * 3 + pp) * 32], k3 = KB[(3 * 3 + pp) * 32], k4 = KB[(4 * 3 + pp) * 32],
                      k5 = KB[(5 * 3 + pp) * 32];
        buf[2 * pp] = P6_Y(buf[2 * pp], k0.x, k2.x, k3.x, k4.x, k5.x, buf[7 + 2 * pp]);
        buf[2 * pp + 1] = P6_Y(buf[2 * pp + 1], k0.y, k2.y, k3.y, k4.y, k5.y, buf[8 + 2 * pp]);
    }
    buf[6] = P6_Y(buf[6], dm, dm, dm, dm, dm, dm);
}

RLR_DEV void env_step_b(const RlrParams& p, StB& s, const float (&u)[3], double* gp, int lane, float (&obs)[8], int& done_out,
                        int& natt_out) {
    using rx = ChanAB;
    using tx = ChanBA;
    double2* KB = reinterpret_cast<double2*>(gp + OFF_KB) + lane;
    double* SBQ0 = gp + OFF_SBQ + 2 * lane;
    double* SBQ1 = gp + OFF_SBQ + 64 + 2 * lane;
    double* SB1 = gp + OFF_SB1 + lane;
    const double* SA0 = gp + OFF_SA + lane;
    const double* SA1 = gp + OFF_SA + 32 + lane;
    constexpr double J1 = 6037675.13, J2 = 6037675.13;
    CtxB c;
    {
        const float sy = np_sinf(u[0]), cz = np_cosf(u[1]), sz = np_sinf(u[1]);
        const double T = (double)u[2];
        const double Tb1 = (double)fmul(sy, cz) * T, Tb2 = (double)sz * T;
        c.ct1 = (1.0 / J1) * (15.0 * Tb2);           // r_T_B x T_b with r_T_B = [-15,0,0]  (simulator.py:373-378)
        c.ct2 = (1.0 / J2) * (-15.0 * Tb1);
        c.hw0 = 0.5 * s.w[0];
        c.dm = (double)fdiv(-u[2], 3531.6f);         // f32 division under NEP-50 (simulator.py:291-292)
    }
    Ctl ctl;
    ctl.t = (double)((long long)s.k * (long long)p.dt_ms) / 1000.0;
    ctl.t_bound = ctl.t + p.dt;
    ctl.natt = 0; ctl.rejected = 0; ctl.status = 0; ctl.running = 1;
    double kk[6];
    {
        const double n2 = ((s.q[1] * s.q[1] + s.q[2] * s.q[2]) + s.q[3] * s.q[3]) + s.q[0] * s.q[0];
        tx::begin_send();
        mb_st2(SBQ0, s.q[0], s.q[1]); mb_st2(SBQ1, s.q[2], s.q[3]); mb_st1(SB1, s.m * n2);
        tx::end_send();
        dqw_b(c, s.q, s.w[1], s.w[2], kk);
#pragma unroll
        for (int pp = 0; pp < 3; ++pp) KB[pp * 32] = make_double2(kk[2 * pp], kk[2 * pp + 1]);
    }
    // ---- select_initial_step: partial sums over q, w, m
    {
        const double interval = fabs(ctl.t_bound - ctl.t);
        double isc[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) isc[i] = rcp_(fma(fabs(s.q[i]), RTOL, ATOL));
#pragma unroll
        for (int i = 0; i < 3; ++i) isc[4 + i] = rcp_(fma(fabs(s.w[i]), RTOL, ATOL));
        isc[7] = rcp_(fma(fabs(s.m), RTOL, ATOL));
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const double e = s.q[i] * isc[i]; a0 += e * e; }
#pragma unroll
        for (int i = 0; i < 3; ++i) { const double e = s.w[i] * isc[4 + i]; a0 += e * e; }
        { const double e = s.m * isc[7]; a0 += e * e; }
#pragma unroll
        for (int i = 0; i < 4; ++i) { const double e = kk[i] * isc[i]; a1 += e * e; }
        { const double e = kk[4] * isc[5]; a1 += e * e; }
        { const double e = kk[5] * isc[6]; a1 += e * e; }
        { const double e = c.dm * isc[7]; a1 += e * e; }
        tx::begin_send(); mb_st2(SBQ0, a0, a1); tx::end_send();
        rx::begin_recv();
        const double a0a = mb_ld1(SA0), a1a = mb_ld1(SA1);
        rx::end_recv();
        double d1;
        const double h0 = sel_h0(__dadd_rn(a0a, a0), __dadd_rn(a1a, a1), interval, d1);
        double q1[4], f1[6];
#pragma unroll
        for (int i = 0; i < 4; ++i) q1[i] = s.q[i] + h0 * kk[i];
        const double wy1 = s.w[1] + h0 * kk[4], wz1 = s.w[2] + h0 * kk[5];
        const double m1 = s.m + h0 * c.dm;
        const double n2 = ((q1[1] * q1[1] + q1[2] * q1[2]) + q1[3] * q1[3]) + q1[0] * q1[0];
        tx::begin_send();
        mb_st2(SBQ0, q1[0], q1[1]); mb_st2(SBQ1, q1[2], q1[3]); mb_st1(SB1, m1 * n2);
        tx::end_send();
        dqw_b(c, q1, wy1, wz1, f1);
        double a2 = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const double e = (f1[i] - kk[i]) * isc[i]; a2 += e * e; }
        { const double e = (f1[4] - kk[4]) * isc[5]; a2 += e * e; }
        { const double e = (f1[5] - kk[5]) * isc[6]; a2 += e * e; }
        tx::begin_send(); mb_st2(SBQ0, a2, 0.0); tx::end_send();
        rx::begin_recv();
        const double a2a = mb_ld1(SA0);
        rx::end_recv();
        ctl.h_abs = sel_habs(__dadd_rn(a2a, a2), d1, h0, interval);
        const double ms = min_step_at(ctl.t);
        if (ctl.h_abs < ms) ctl.h_abs = ms;
    }
#pragma unroll 1
    for (;;) {
        ctl_begin(ctl);
        if (!__any_sync(0xffffffffu, ctl.running)) break;
        const double h = ctl.h;
        double sq[4], swy = s.w[1], swz = s.w[2], sm = s.m;
#pragma unroll 1
        for (int st = 1; st <= 6; ++st) {
            double acc[6];
            {
                const double a = c_RKA[st][0];
#pragma unroll
                for (int pp = 0; pp < 3; ++pp) { const double2 k2 = KB[pp * 32]; acc[2 * pp] = k2.x * a; acc[2 * pp + 1] = k2.y * a; }
            }
            const double2* kp = KB + 3 * 32;
#pragma unroll 1
            for (int j = 1; j < st; ++j, kp += 3 * 32) {
                const double aj = c_RKA[st][j];
#pragma unroll
                for (int pp = 0; pp < 3; ++pp) { const double2 k2 = kp[pp * 32]; acc[2 * pp] += k2.x * aj; acc[2 * pp + 1] += k2.y * aj; }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) sq[i] = s.q[i] + acc[i] * h;
            swy = s.w[1] + acc[4] * h; swz = s.w[2] + acc[5] * h;
            sm = s.m + (c_RKC[st] * h) * c.dm;
            const double n2 = ((sq[1] * sq[1] + sq[2] * sq[2]) + sq[3] * sq[3]) + sq[0] * sq[0];
            tx::begin_send();
            mb_st2(SBQ0, sq[0], sq[1]); mb_st2(SBQ1, sq[2], sq[3]); mb_st1(SB1, sm * n2);
            tx::end_send();
            dqw_b(c, sq, swy, swz, kk);
            if (st <= 5) {
                double2* ko = KB + st * 3 * 32;
#pragma unroll
                for (int pp = 0; pp < 3; ++pp) ko[pp * 32] = make_double2(kk[2 * pp], kk[2 * pp + 1]);
            }
        }
        // ---- error norm, partial sum over q, w (omega_x and m contribute exactly zero)
        double ssum = 0.0;
#pragma unroll
        for (int pp = 0; pp < 3; ++pp) {
            const double2 k0 = KB[pp * 32], k2 = KB[(2 * 3 + pp) * 32], k3 = KB[(3 * 3 + pp) * 32], k4 = KB[(4 * 3 + pp) * 32],
                          k5 = KB[(5 * 3 + pp) * 32];
            const double ex = (((((k0.x * E0 + k2.x * E2) + k3.x * E3) + k4.x * E4) + k5.x * E5) + kk[2 * pp] * E6) * h;
            const double ey = (((((k0.y * E0 + k2.y * E2) + k3.y * E3) + k4.y * E4) + k5.y * E5) + kk[2 * pp + 1] * E6) * h;
            const double yo_x = (pp < 2) ? s.q[2 * pp] : s.w[1], yo_y = (pp < 2) ? s.q[2 * pp + 1] : s.w[2];
            const double yn_x = (pp < 2) ? sq[2 * pp] : swy, yn_y = (pp < 2) ? sq[2 * pp + 1] : swz;
            const double e1 = ex * rcp_scale(yo_x, yn_x), e2 = ey * rcp_scale(yo_y, yn_y);
            ssum += e1 * e1;
            ssum += e2 * e2;
        }
        tx::begin_send(); mb_st2(SBQ0, ssum, 0.0); tx::end_send();
        rx::begin_recv();
        const double ssum_a = mb_ld1(SA0);
        const int sign_change = (int)mb_ld1(SA1);
        rx::end_recv();
        const int was_running = ctl.running;
        const int verdict = ctl_end(ctl, __dadd_rn(ssum_a, ssum), sign_change);
        const int ev = (was_running && verdict == 2) ? 1 : 0;
        if (__any_sync(0xffffffffu, ev)) {
            rx::begin_recv();
            const double xx = mb_ld1(SA0);
            rx::end_recv();
            if (ev) {
                double ebuf[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) ebuf[i] = s.q[i];
                ebuf[4] = s.w[1]; ebuf[5] = s.w[2]; ebuf[6] = s.m;
#pragma unroll
                for (int i = 0; i < 6; ++i) ebuf[7 + i] = kk[i];
                ebuf[13] = c.dm; ebuf[14] = h; ebuf[15] = xx;
                event_b(gp, lane, ebuf);
#pragma unroll
                for (int i = 0; i < 4; ++i) s.q[i] = ebuf[i];
                s.w[1] = ebuf[4]; s.w[2] = ebuf[5]; s.m = ebuf[6];
            }
        }
        if (was_running && verdict == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) s.q[i] = sq[i];
            s.w[1] = swy; s.w[2] = swz; s.m = sm;
#pragma unroll
            for (int pp = 0; pp < 3; ++pp) KB[pp * 32] = make_double2(kk[2 * pp], kk[2 * pp + 1]);     // FSAL
        }
    }
    // ---- quaternion normalisation (simulator.py:250), attitude and rate limits, observation of q, w, m
    {
        const double n2 = ((s.q[0] * s.q[0] + s.q[1] * s.q[1]) + s.q[2] * s.q[2]) + s.q[3] * s.q[3];
        const double rnq = rsqrt_(n2);
#pragma unroll
        for (int i = 0; i < 4; ++i) s.q[i] = s.q[i] * rnq;
    }
    float sq32[4], sw32[3];
#pragma unroll
    for (int i = 0; i < 4; ++i) { sq32[i] = (float)s.q[i]; obs[i] = (float)(s.q[i] * p.inv_normalizer[6 + i]); }
#pragma unroll
    for (int i = 0; i < 3; ++i) { sw32[i] = (float)s.w[i]; obs[4 + i] = (float)(s.w[i] * p.inv_normalizer[10 + i]); }
    obs[7] = (float)(s.m * p.inv_normalizer[13]);
    int bflags;
    {
        // Rotation.from_quat(roll(q_f32,-1)).as_euler("zyx") is only compared with limits:  |atan2(b, a)| > L  <=>  a < cos(L) hypot(a, b),
        // |asin(v)| > L  <=>  |v| > sin(L)  (cos / sin of the limits come from the host, with sentinels outside [0, pi] / [0, pi/2])
        const double q0 = sq32[0], q1 = sq32[1], q2 = sq32[2], q3 = sq32[3];
        const double rnq = rsqrt_(((q1 * q1 + q2 * q2) + q3 * q3) + q0 * q0);
        const double x = q1 * rnq, yy = q2 * rnq, z = q3 * rnq, w = q0 * rnq;
        const double R00 = x * x - yy * yy - z * z + w * w, R01 = 2 * (x * yy - z * w), R02 = 2 * (x * z + yy * w);
        const double R12 = 2 * (yy * z - x * w), R22 = -x * x - yy * yy + z * z + w * w;
        const double hy0 = sqrt_(R00 * R00 + R01 * R01), hy2 = sqrt_(R22 * R22 + R12 * R12), av1 = fabs(R02);
        const int att_viol = (R00 < p.att_traj_c[0] * hy0) | (av1 > p.att_traj_c[1]) | (R22 < p.att_traj_c[2] * hy2);
        const int att_ok = (R00 > p.att_land_c[0] * hy0) | (av1 < p.att_land_c[1]) | (R22 > p.att_land_c[2] * hy2);
        const int om_ok = ((double)fabsf(sw32[0]) < p.omega_lim[0]) | ((double)fabsf(sw32[1]) < p.omega_lim[1]) | ((double)fabsf(sw32[2]) < p.omega_lim[2]);
        bflags = att_viol | (att_ok << 1) | (om_ok << 2);
    }
    tx::begin_send(); mb_st1(SB1, (double)bflags); tx::end_send();
    s.k += 1;
    rx::begin_recv();
    done_out = (int)mb_ld1(SA0);
    rx::end_recv();
    natt_out = ctl.natt;
}

// =============================================================================================== reset (Philox blocks split by role)
RLR_DEV void philox_ic(const RlrParams& p, uint64_t seed, int64_t env, int32_t episode, int b, float& f0, float& f1) {
    uint32_t c[4] = {(uint32_t)b, (uint32_t)episode, (uint32_t)((uint64_t)env & 0xffffffffu), (uint32_t)((uint64_t)env >> 32)};
    philox4x32_10(c, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
    const uint64_t b0 = ((uint64_t)c[0] << 32) | (uint64_t)c[1], b1 = ((uint64_t)c[2] << 32) | (uint64_t)c[3];
    const double u0 = (double)(b0 >> 11) * 0x1.0p-53, u1 = (double)(b1 >> 11) * 0x1.0p-53;
    const double lo0 = (double)p.ic_low[2 * b], hi0 = (double)p.ic_high[2 * b], lo1 = (double)p.ic_low[2 * b + 1], hi1 = (double)p.ic_high[2 * b + 1];
    f0 = (float)(lo0 + (hi0 - lo0) * u0);
    f1 = (float)(lo1 + (hi1 - lo1) * u1);
}

RLR_DEV void reset_a(const RlrParams& p, StA& s, const float* ic_inject, int64_t i, int64_t env, uint64_t seed, float (&obs)[6]) {
    s.episode += 1;
    float ic[6];
    if (ic_inject != nullptr) {
#pragma unroll
        for (int c = 0; c < 6; ++c) ic[c] = ic_inject[i * 14 + c];
    } else {
#pragma unroll 1
        for (int b = 0; b < 3; ++b) {
            float f0, f1;
            philox_ic(p, seed, env, s.episode, b, f0, f1);
            if (b == 0) { ic[0] = f0; ic[1] = f1; } else if (b == 1) { ic[2] = f0; ic[3] = f1; } else { ic[4] = f0; ic[5] = f1; }
        }
    }
    s.v0 = f32_norm3(ic[3], ic[4], ic[5]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        s.r[c] = (double)ic[c]; s.v[c] = (double)ic[3 + c];
        obs[c] = (float)((double)ic[c] * p.inv_normalizer[c]);
        obs[3 + c] = (float)((double)ic[3 + c] * p.inv_normalizer[3 + c]);
    }
    s.k = 0;
    s.ep_ret = 0.0;
}

RLR_DEV void reset_b(const RlrParams& p, StB& s, const float* ic_inject, int64_t i, int64_t env, uint64_t seed, float (&obs)[8]) {
    s.episode += 1;
    float ic[8];
    if (ic_inject != nullptr) {
#pragma unroll
        for (int c = 0; c < 8; ++c) ic[c] = ic_inject[i * 14 + 6 + c];
    } else {
#pragma unroll 1
        for (int b = 3; b < 7; ++b) {
            float f0, f1;
            philox_ic(p, seed, env, s.episode, b, f0, f1);
            if (b == 3) { ic[0] = f0; ic[1] = f1; } else if (b == 4) { ic[2] = f0; ic[3] = f1; } else if (b == 5) { ic[4] = f0; ic[5] = f1; }
            else { ic[6] = f0; ic[7] = f1; }
        }
    }
    const float nq = f32_norm4(ic[0], ic[1], ic[2], ic[3]);          // quaternion normalised in f32 (rocket_env.py:673)
#pragma unroll
    for (int c = 0; c < 4; ++c) ic[c] = fdiv(ic[c], nq);
#pragma unroll
    for (int c = 0; c < 4; ++c) s.q[c] = (double)ic[c];
#pragma unroll
    for (int c = 0; c < 3; ++c) s.w[c] = (double)ic[4 + c];
    s.m = (double)ic[7];
#pragma unroll
    for (int c = 0; c < 8; ++c) obs[c] = (float)((double)ic[c] * p.inv_normalizer[6 + c]);
    s.k = 0;
}

// =============================================================================================== kernels
RLR_DEV void load_a(const RlrState& gs, int64_t n, int64_t i, StA& s) {
    const double* y = (const double*)gs.y;
#pragma unroll
    for (int c = 0; c < 3; ++c) { s.r[c] = y[c * n + i]; s.v[c] = y[(3 + c) * n + i]; }
    s.k = gs.k[i]; s.episode = gs.episode[i]; s.v0 = gs.v0[i]; s.ep_ret = gs.ep_ret[i];
}
RLR_DEV void store_a(const RlrState& gs, int64_t n, int64_t i, const StA& s) {
    double* y = (double*)gs.y;
#pragma unroll
    for (int c = 0; c < 3; ++c) { y[c * n + i] = s.r[c]; y[(3 + c) * n + i] = s.v[c]; }
    gs.k[i] = s.k; gs.episode[i] = s.episode; gs.v0[i] = s.v0; gs.ep_ret[i] = s.ep_ret;
}
RLR_DEV void load_b(const RlrState& gs, int64_t n, int64_t i, StB& s) {
    const double* y = (const double*)gs.y;
#pragma unroll
    for (int c = 0; c < 4; ++c) s.q[c] = y[(6 + c) * n + i];
#pragma unroll
    for (int c = 0; c < 3; ++c) s.w[c] = y[(10 + c) * n + i];
    s.m = y[13 * n + i];
    s.k = gs.k[i]; s.episode = gs.episode[i];
}
RLR_DEV void store_b(const RlrState& gs, int64_t n, int64_t i, const StB& s) {
    double* y = (double*)gs.y;
#pragma unroll
    for (int c = 0; c < 4; ++c) y[(6 + c) * n + i] = s.q[c];
#pragma unroll
    for (int c = 0; c < 3; ++c) y[(10 + c) * n + i] = s.w[c];
    y[13 * n + i] = s.m;
}
// obs rows are 56 bytes: A writes floats 0..5, B floats 6..13, as 8-byte stores
RLR_DEV void write_obs_a(float* dst, int64_t i, const float (&o)[6]) {
    float2* d = reinterpret_cast<float2*>(dst + i * 14);
    d[0] = make_float2(o[0], o[1]); d[1] = make_float2(o[2], o[3]); d[2] = make_float2(o[4], o[5]);
}
RLR_DEV void write_obs_b(float* dst, int64_t i, const float (&o)[8]) {
    float2* d = reinterpret_cast<float2*>(dst + i * 14 + 6);
    d[0] = make_float2(o[0], o[1]); d[1] = make_float2(o[2], o[3]); d[2] = make_float2(o[4], o[5]); d[3] = make_float2(o[6], o[7]);
}
RLR_DEV void map_action6(const RlrParams& p, const float* actions, int action_kind, int64_t i, float (&u)[3], float& a_thr) {
    const float* a = actions + i * 3;
    const float a0 = a[0], a1 = a[1], a2 = a[2];
    a_thr = 0.0f;
    if (action_kind == RLR_ACT_NORMALIZED) {
        u[0] = (float)((double)a0 * p.max_gimbal);
        u[1] = (float)((double)a1 * p.max_gimbal);
        u[2] = fmul(fdiv(fadd(a2, 1.0f), 2.0f), p.max_thrust);
        a_thr = a2;
    } else { u[0] = a0; u[1] = a1; u[2] = a2; }
}

// role of a warp: A and B alternate with the block parity so that every scheduler partition gets both roles
RLR_DEV bool warp_is_a() { return (((threadIdx.x >> 5) ^ blockIdx.x) & 1) == 0; }

__global__ void __launch_bounds__(THREADS, 14) pair_step_kernel(const __grid_constant__ RlrParams p, RlrState gs, const float* actions,
                                                                 int action_kind, const float* ic_inject, RlrStepOut out, int64_t n,
                                                                 int64_t env_offset, uint64_t seed) {
    extern __shared__ double smem_pair[];
    const bool is_a = warp_is_a();
    const int lane = threadIdx.x & 31;
    double* gp = smem_pair;
    const int64_t i_raw = (int64_t)blockIdx.x * ENVS + lane;
    const bool active = i_raw < n;
    const int64_t i = active ? i_raw : n - 1;          // tail lanes shadow the last environment (no stores) so that barriers stay uniform
    float u[3], a_thr;
    map_action6(p, actions, action_kind, i, u, a_thr);
    if (is_a) {
        ChanBA::end_recv();                              // prime: the B -> A mailbox is empty
        StA s;
        load_a(gs, n, i, s);
        float obs[6];
        StepRes res;
        StepFlags fl = {0, 0, 0, 0, 0, 0};
        env_step_a(p, s, u, a_thr, gp, lane, obs, res, fl);
        const int ep_len = s.k;
        const double ep_ret = s.ep_ret;
        if (active) {
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
                for (int c = 0; c < 3; ++c) { yp[c * n + i] = s.r[c]; yp[(3 + c) * n + i] = s.v[c]; }
            }
        }
        if (fl.done) {
            if (active) {
                if (out.terminal_obs) write_obs_a(out.terminal_obs, i, obs);
                if (out.fin_ret) out.fin_ret[i] = s.ep_ret;
                if (out.fin_len) out.fin_len[i] = s.k;
            }
            reset_a(p, s, ic_inject, i, env_offset + i, seed, obs);
        }
        if (active) {
            write_obs_a(out.obs, i, obs);
            store_a(gs, n, i, s);
        }
        accumulate_stats(out.stats, active, fl, ep_len, ep_ret);
        ChanAB::begin_send();                            // drain: consume B's last "empty" arrival
    } else {
        ChanAB::end_recv();                              // prime: the A -> B mailbox is empty
        StB s;
        load_b(gs, n, i, s);
        float obs[8];
        int done, natt;
        env_step_b(p, s, u, gp, lane, obs, done, natt);
        if (active && out.y_pre_reset) {
            double* yp = (double*)out.y_pre_reset;
#pragma unroll
            for (int c = 0; c < 4; ++c) yp[(6 + c) * n + i] = s.q[c];
#pragma unroll
            for (int c = 0; c < 3; ++c) yp[(10 + c) * n + i] = s.w[c];
            yp[13 * n + i] = s.m;
        }
        if (done) {
            if (active && out.terminal_obs) write_obs_b(out.terminal_obs, i, obs);
            reset_b(p, s, ic_inject, i, env_offset + i, seed, obs);
        }
        if (active) {
            write_obs_b(out.obs, i, obs);
            store_b(gs, n, i, s);
        }
        ChanBA::begin_send();                            // drain
    }
}

// T fused steps: the state stays in registers between environment steps
__global__ void __launch_bounds__(THREADS, 14) pair_rollout_kernel(const __grid_constant__ RlrParams p, RlrState gs, const float* actions,
                                                                    int action_kind, int T, float* obs_buf, float* reward_buf, uint8_t* done_buf,
                                                                    uint8_t* trunc_buf, float* last_obs, unsigned long long* stats, int64_t n,
                                                                    int64_t env_offset, uint64_t seed) {
    extern __shared__ double smem_pair[];
    const bool is_a = warp_is_a();
    const int lane = threadIdx.x & 31;
    double* gp = smem_pair;
    const int64_t i_raw = (int64_t)blockIdx.x * ENVS + lane;
    const bool active = i_raw < n;
    const int64_t i = active ? i_raw : n - 1;
    if (is_a) {
        ChanBA::end_recv();
        StA s;
        load_a(gs, n, i, s);
        float obs[6];
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
            float u[3], a_thr;
            map_action6(p, actions + (int64_t)t * n * 3, action_kind, i, u, a_thr);
            StepRes res;
            StepFlags fl = {0, 0, 0, 0, 0, 0};
            env_step_a(p, s, u, a_thr, gp, lane, obs, res, fl);
            const int ep_len = s.k;
            const double ep_ret = s.ep_ret;
            if (fl.done) reset_a(p, s, nullptr, i, env_offset + i, seed, obs);
            if (active) {
                if (obs_buf) write_obs_a(obs_buf + (int64_t)t * n * 14, i, obs);
                if (reward_buf) reward_buf[(int64_t)t * n + i] = (float)res.reward;
                if (done_buf) done_buf[(int64_t)t * n + i] = (uint8_t)fl.done;
                if (trunc_buf) trunc_buf[(int64_t)t * n + i] = (uint8_t)fl.trunc;
            }
            accumulate_stats(stats, active, fl, ep_len, ep_ret);
        }
        if (active) {
            if (last_obs && T > 0) write_obs_a(last_obs, i, obs);
            store_a(gs, n, i, s);
        }
        ChanAB::begin_send();
    } else {
        ChanAB::end_recv();
        StB s;
        load_b(gs, n, i, s);
        float obs[8];
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
            float u[3], a_thr;
            map_action6(p, actions + (int64_t)t * n * 3, action_kind, i, u, a_thr);
            int done, natt;
            env_step_b(p, s, u, gp, lane, obs, done, natt);
            if (done) reset_b(p, s, nullptr, i, env_offset + i, seed, obs);
            if (active && obs_buf) write_obs_b(obs_buf + (int64_t)t * n * 14, i, obs);
        }
        if (active) {
            if (last_obs && T > 0) write_obs_b(last_obs, i, obs);
            store_b(gs, n, i, s);
        }
        ChanBA::begin_send();
    }
}

#undef P6_Q
#undef P6_Y

}  // namespace pair6
