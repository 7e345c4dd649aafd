// rocket_f32.cuh -- RLR_F32 precision of the 6DOF environment step (included from rocket_kernels.cuh inside namespace RLR_NS,
// fast build only).
//
// Same algorithm as the f64 path (scipy RK45 semantics: select_initial_step, error-controlled sub-steps with FSAL, terminal ground
// event on the quartic dense output; Rocket6DOF reward / termination / TimeLimit / auto-reset), with the integrated state, the stage
// matrix and the step-size control in binary32.  It is NOT a bit-level restatement of the reference: the contract is the tolerance
// BASELINE.json states for fp32 -- state within 1e-4 (normaliser-scaled) of the fp64 reference after 500 steps, done flags equal
// except where the fp64 margin to a threshold is below that tolerance (tests/test_gpu_f32.py counts and bounds those).
//
// What differs from the f64 kernels besides the scalar type:
//  * the solve runs in LOCAL time tau in [0, dt] (the right-hand side is autonomous), so t + h does not lose the step length to the
//    absolute time's rounding; `k` still counts steps and gives t = k dt_ms / 1000 exactly when it is needed;
//  * reciprocals / rsqrt are the hardware approximations (1-2 ulp in binary32) and x^(-1/5) is exp2(-0.2 log2 x) with one Newton step;
//  * the event root is a bracket-safeguarded Newton iteration (no Brent fallback needed: it bisects when Newton leaves the bracket);
//  * the reward is evaluated by the same `reward6` as the f64 path (binary64 arithmetic on the binary32 state, which is exactly the
//    reference's dtype flow: it casts the state to float32 before computing the reward, rocket_env.py:695).
// Reference call path: Rocket6DOF.step rocket_env.py:690-719 -> Simulator6DOF.step simulator.py:227-257 -> solve_ivp (scipy _ivp/rk.py).
#pragma once

namespace f32 {

__constant__ float cf_A[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {(float)A10, 0, 0, 0, 0, 0},
    {(float)A20, (float)A21, 0, 0, 0, 0},
    {(float)A30, (float)A31, (float)A32, 0, 0, 0},
    {(float)A40, (float)A41, (float)A42, (float)A43, 0, 0},
    {(float)A50, (float)A51, (float)A52, (float)A53, (float)A54, 0},
    {(float)B0, 0, (float)B2, (float)B3, (float)B4, (float)B5}};
__constant__ float cf_B[6] = {(float)B0, 0, (float)B2, (float)B3, (float)B4, (float)B5};
__constant__ float cf_E[7] = {(float)E0, 0, (float)E2, (float)E3, (float)E4, (float)E5, (float)E6};
__constant__ float cf_C[7] = {0, 0.2f, 0.3f, 0.8f, (float)(8.0 / 9), 1, 1};
__constant__ float cf_P[7][4] = {
    {1, (float)(-8048581381.0 / 2820520608), (float)(8663915743.0 / 2820520608), (float)(-12715105075.0 / 11282082432)},
    {0, 0, 0, 0},
    {0, (float)(131558114200.0 / 32700410799), (float)(-68118460800.0 / 10900136933), (float)(87487479700.0 / 32700410799)},
    {0, (float)(-1754552775.0 / 470086768), (float)(14199869525.0 / 1410260304), (float)(-10690763975.0 / 1880347072)},
    {0, (float)(127303824393.0 / 49829197408), (float)(-318862633887.0 / 49829197408), (float)(701980252875.0 / 199316789632)},
    {0, (float)(-282668133.0 / 205662961), (float)(2019193451.0 / 616988883), (float)(-1453857185.0 / 822651844)},
    {0, (float)(40617522.0 / 29380423), (float)(-110615467.0 / 29380423), (float)(69997945.0 / 29380423)}};

constexpr float RTOLF = 1e-3f, ATOLF = 1e-6f;
constexpr float INV_RS14F = 0.26726124f;

RLR_DEV float rcpf_(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
RLR_DEV float rsqrtf_(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * fmaf(-0.5f * x * r, r, 1.5f);               // one Newton step: ~1 ulp
}
RLR_DEV float sqrtf_(float x) { return (x == 0.0f) ? 0.0f : x * rsqrtf_(x); }
RLR_DEV float powf_m02(float x) {                           // x^(-1/5)
    const float xc = fminf(fmaxf(x, 1e-30f), 1e30f);
    float y = __powf(xc, -0.2f);
    const float y2 = y * y;
    return y * fmaf(-0.2f * x, (y2 * y2) * y, 1.2f);
}
RLR_DEV float powf_p02(float x) { const float y = powf_m02(x), y2 = y * y; return x * (y2 * y2); }
RLR_DEV float absmaxf_(float a, float b) { return fmaxf(fabsf(a), fabsf(b)); }
RLR_DEV float min_stepf_at(float t) { return 10.0f * fabsf(__int_as_float(__float_as_int(t) + 1) - t); }      // t >= 0

// Packed binary32 arithmetic (Blackwell FFMA2 / FMUL2: two lanes of a 64-bit register pair per instruction).  The f32 kernels are
// instruction-issue bound (69 % of the issue slots, profiles/r01d_f32_kernels_ncu.txt); the stage sums, stage states and error
// estimates are pairwise-independent, so packing them halves their instruction count.
RLR_DEV float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(reinterpret_cast<unsigned long long&>(d))
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)), "l"(reinterpret_cast<unsigned long long&>(c)));
    return d;
}
RLR_DEV float2 fmul2(float2 a, float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(reinterpret_cast<unsigned long long&>(d))
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return d;
}
RLR_DEV float2 bc2(float x) { return make_float2(x, x); }

struct Ctx6f { float Tb0, Tb1, Tb2, tau1, tau2, dm; };

// d = [dv(3), dq(4), dw_y, dw_z]   (dw_x == 0: J_y == J_z)
RLR_DEV void rhs6f(const Ctx6f& c, const float (&q)[4], float wx, float wy, float wz, float m, float (&d)[9]) {
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    const float uu = (x * x + y * y) + z * z;
    const float inv_m = rcpf_(m * (uu + w * w));
    const float a = w * w - uu;
    const float b = 2.0f * ((x * c.Tb0 + y * c.Tb1) + z * c.Tb2);
    const float w2 = w + w;
    const float cx = y * c.Tb2 - z * c.Tb1, cy = z * c.Tb0 - x * c.Tb2, cz = x * c.Tb1 - y * c.Tb0;
    d[0] = inv_m * ((a * c.Tb0 + b * x) + w2 * cx) + (-9.81f);
    d[1] = inv_m * ((a * c.Tb1 + b * y) + w2 * cy);
    d[2] = inv_m * ((a * c.Tb2 + b * z) + w2 * cz);
    const float hx = 0.5f * wx, hy = 0.5f * wy, hz = 0.5f * wz;
    d[3] = ((-hx) * x + (-hy) * y) + (-hz) * z;
    d[4] = (hx * w + hz * y) + (-hy) * z;
    d[5] = (hy * w + (-hz) * x) + hx * z;
    d[6] = (hz * w + hy * x) + (-hx) * y;
    constexpr float J0 = 75350.25f, J1 = 6037675.13f;
    d[7] = (1.0f / J1) * c.tau1 - (((J0 - J1) / J1) * wx) * wz;
    d[8] = (1.0f / J1) * c.tau2 - (((J1 - J0) / J1) * wx) * wy;
}

constexpr int BSF = 128;                         // threads per block
constexpr int NKF = 9;                           // stage-matrix components per thread and stage
// Stage row of a thread = two float4 (components 0..7) and one float (component 8): rows are [vec][tid] so that consecutive
// threads touch consecutive 16-byte (4-byte) words -- three shared-memory instructions per row instead of nine.
constexpr int ROWF = 9 * BSF;                    // floats per stage (8 * BSF in the float4 planes, BSF in the scalar plane)
constexpr size_t SMEMF_BYTES = 6 * ROWF * sizeof(float);           // 27 KB per block
struct KRow { float4 a, b; float c; };
RLR_DEV KRow kf_load(const float* base, int j) {          // base = block's stage matrix, includes nothing thread specific
    const float* r = base + j * ROWF;
    KRow k;
    k.a = reinterpret_cast<const float4*>(r)[threadIdx.x];
    k.b = reinterpret_cast<const float4*>(r + 4 * BSF)[threadIdx.x];
    k.c = r[8 * BSF + threadIdx.x];
    return k;
}
RLR_DEV void kf_store(float* base, int j, const float (&v)[9]) {
    float* r = base + j * ROWF;
    reinterpret_cast<float4*>(r)[threadIdx.x] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(r + 4 * BSF)[threadIdx.x] = make_float4(v[4], v[5], v[6], v[7]);
    r[8 * BSF + threadIdx.x] = v[8];
}
RLR_DEV void kf_unpack(const KRow& k, float (&v)[9]) {
    v[0] = k.a.x; v[1] = k.a.y; v[2] = k.a.z; v[3] = k.a.w; v[4] = k.b.x; v[5] = k.b.y; v[6] = k.b.z; v[7] = k.b.w; v[8] = k.c;
}

// Terminal ground event: dense output (rk.py:554-565), safeguarded Newton root of the height quartic, state at the event.
// buf = [y(14) | K6(9) | v_new(3) | dm h]; y(t_e) is written to buf[0..13].
__device__ __noinline__ void event_path6f(const float* __restrict__ sKb, float* __restrict__ buf) {
    const float dm = buf[26], h = buf[27];
    float K[6][9];
#pragma unroll
    for (int j = 0; j < 6; ++j) kf_unpack(kf_load(sKb, j), K[j]);
#define KF(j, cc) K[j][cc]
    float vs[6][3];
#pragma unroll
    for (int s = 2; s <= 5; ++s) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float a3 = KF(0, i) * cf_A[s][0];
#pragma unroll
            for (int j = 1; j < s; ++j) a3 += KF(j, i) * cf_A[s][j];
            vs[s][i] = buf[3 + i] + a3 * h;
        }
    }
#define QF(k0, k2, k3, k4, k5, k6, cc) \
    ((((((k0) * cf_P[0][cc] + (k2) * cf_P[2][cc]) + (k3) * cf_P[3][cc]) + (k4) * cf_P[4][cc]) + (k5) * cf_P[5][cc]) + (k6) * cf_P[6][cc])
    const float vx0 = buf[3], vx6 = buf[23];
    const float q0 = vx0, q1 = QF(vx0, vs[2][0], vs[3][0], vs[4][0], vs[5][0], vx6, 1), q2 = QF(vx0, vs[2][0], vs[3][0], vs[4][0], vs[5][0], vx6, 2),
                q3 = QF(vx0, vs[2][0], vs[3][0], vs[4][0], vs[5][0], vx6, 3);
    // g(x) = r0 + h x (q0 + x (q1 + x (q2 + x q3))) on the bracket [0, 1]
    const float g0 = buf[0];
    const float g1 = h * (((q0 + q1) + q2) + q3) + g0;
    float lo = 0.0f, hi = 1.0f, glo = g0;
    float x = (g0 == g1) ? 0.5f : g0 / (g0 - g1);
    x = fminf(fmaxf(x, 0.0f), 1.0f);
#pragma unroll 1
    for (int it = 0; it < 24; ++it) {
        const float s = fmaf(fmaf(fmaf(q3, x, q2), x, q1), x, q0);
        const float ds = fmaf(fmaf(fmaf(4.0f * q3, x, 3.0f * q2), x, 2.0f * q1), x, q0);
        const float gx = fmaf(h * x, s, g0);
        if (gx == 0.0f) break;
        if ((gx < 0.0f) == (glo < 0.0f)) { lo = x; glo = gx; } else hi = x;
        float xn = x - gx / (h * ds);
        if (!(xn > lo && xn < hi)) xn = 0.5f * (lo + hi);          // Newton left the bracket (or NaN): bisect
        if (xn == x) break;
        x = xn;
    }
    const float p1 = x, p2 = p1 * x, p3 = p2 * x, p4 = p3 * x;
#define YF(yold, k0, k2, k3, k4, k5, k6) \
    (h * ((((k0) * p1 + QF(k0, k2, k3, k4, k5, k6, 1) * p2) + QF(k0, k2, k3, k4, k5, k6, 2) * p3) + QF(k0, k2, k3, k4, k5, k6, 3) * p4) + (yold))
#pragma unroll
    for (int i = 0; i < 3; ++i) buf[i] = YF(buf[i], buf[3 + i], vs[2][i], vs[3][i], vs[4][i], vs[5][i], buf[23 + i]);     // uses the OLD v
#pragma unroll
    for (int cc = 0; cc < 7; ++cc) buf[3 + cc] = YF(buf[3 + cc], KF(0, cc), KF(2, cc), KF(3, cc), KF(4, cc), KF(5, cc), buf[14 + cc]);
#pragma unroll
    for (int cc = 7; cc < 9; ++cc) buf[4 + cc] = YF(buf[4 + cc], KF(0, cc), KF(2, cc), KF(3, cc), KF(4, cc), KF(5, cc), buf[14 + cc]);
    buf[13] = YF(buf[13], dm, dm, dm, dm, dm, dm);
#undef YF
#undef QF
#undef KF
}

// solve_ivp over one agent step in local time [0, dt]; returns 0 reached dt, 1 ground event, -1 step size underflow
RLR_DEV int rk45_solve6f(const Ctx6f& c, float dt, float (&y)[14], int& natt, float* __restrict__ sKb) {
    float kk[9];
    {
        const float q[4] = {y[6], y[7], y[8], y[9]};
        rhs6f(c, q, y[10], y[11], y[12], y[13], kk);
        kf_store(sKb, 0, kk);
    }
    natt = 0;
    if (y[13] == 0.0f) { natt = 1; return -1; }        // zero mass: scipy's outcome stated directly (see rk45_solve6)
    // state component of stage-matrix slot cc: 3 + cc for cc < 7 (v, q), 4 + cc for cc = 7, 8 (w_y, w_z)
#define YIDX(cc) ((cc) < 7 ? 3 + (cc) : 4 + (cc))
    float h_abs;
    {   // ---- select_initial_step (common.py:68-134)
        float isc[14];
#pragma unroll
        for (int i = 0; i < 14; ++i) isc[i] = rcpf_(fmaf(fabsf(y[i]), RTOLF, ATOLF));
        float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
        for (int i = 0; i < 14; ++i) { const float e = y[i] * isc[i]; a0 = fmaf(e, e, a0); }
#pragma unroll
        for (int i = 0; i < 3; ++i) { const float e = y[3 + i] * isc[i]; a1 = fmaf(e, e, a1); }
#pragma unroll
        for (int cc = 0; cc < NKF; ++cc) { const float e = kk[cc] * isc[YIDX(cc)]; a1 = fmaf(e, e, a1); }
        { const float e = c.dm * isc[13]; a1 = fmaf(e, e, a1); }
        const float d0 = sqrtf_(a0) * INV_RS14F, d1 = sqrtf_(a1) * INV_RS14F;
        float h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : (0.01f * d0) * rcpf_(d1);
        h0 = fminf(h0, dt);
        float v1[3], q1[4], f1[9];
#pragma unroll
        for (int i = 0; i < 3; ++i) v1[i] = fmaf(h0, kk[i], y[3 + i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) q1[i] = fmaf(h0, kk[3 + i], y[6 + i]);
        rhs6f(c, q1, y[10], fmaf(h0, kk[7], y[11]), fmaf(h0, kk[8], y[12]), fmaf(h0, c.dm, y[13]), f1);
        float a2 = 0.0f;
#pragma unroll
        for (int i = 0; i < 3; ++i) { const float e = (v1[i] - y[3 + i]) * isc[i]; a2 = fmaf(e, e, a2); }
#pragma unroll
        for (int cc = 0; cc < NKF; ++cc) { const float e = (f1[cc] - kk[cc]) * isc[YIDX(cc)]; a2 = fmaf(e, e, a2); }
        const float d2 = (sqrtf_(a2) * INV_RS14F) * rcpf_(h0);
        float h1;
        if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0 * 1e-3f);
        else h1 = powf_p02(0.01f * rcpf_(fmaxf(d1, d2)));
        h_abs = fminf(fminf(100.0f * h0, h1), dt);
    }
    float t = 0.0f;
    float min_step = min_stepf_at(t);
    if (h_abs < min_step) h_abs = min_step;
    bool rejected = false;
    int status;
#pragma unroll 1
    for (;;) {
        if (!(h_abs >= min_step) || natt >= 100000) { status = -1; break; }       // also ends a non-finite step size
        float t_new = t + h_abs;
        if (t_new - dt > 0.0f) t_new = dt;
        const float h = t_new - t;
        h_abs = fabsf(h);
        ++natt;
        // Components with a stage-matrix row, packed in pairs: (v0,v1) (v2,q0) (q1,q2) (q3,wy) and the scalar wz
        const float2 yp0 = make_float2(y[3], y[4]), yp1 = make_float2(y[5], y[6]), yp2 = make_float2(y[7], y[8]), yp3 = make_float2(y[9], y[11]);
        float2 sp0 = yp0, sp1 = yp1, sp2 = yp2, sp3 = yp3;
        float swz = y[12], sm = y[13];
        const float2 h2 = bc2(h);
        // ---- rk_step (rk.py:14-71), rolled; stage 6 is y_new (row 6 of the table = B)
#pragma unroll 1
        for (int s = 1; s <= 6; ++s) {
            float2 a2 = bc2(cf_A[s][0]);
            KRow k = kf_load(sKb, 0);
            float2 acc0 = fmul2(make_float2(k.a.x, k.a.y), a2), acc1 = fmul2(make_float2(k.a.z, k.a.w), a2);
            float2 acc2 = fmul2(make_float2(k.b.x, k.b.y), a2), acc3 = fmul2(make_float2(k.b.z, k.b.w), a2);
            float accs = k.c * a2.x;
#pragma unroll 1
            for (int j = 1; j < s; ++j) {
                a2 = bc2(cf_A[s][j]);
                k = kf_load(sKb, j);
                acc0 = ffma2(make_float2(k.a.x, k.a.y), a2, acc0); acc1 = ffma2(make_float2(k.a.z, k.a.w), a2, acc1);
                acc2 = ffma2(make_float2(k.b.x, k.b.y), a2, acc2); acc3 = ffma2(make_float2(k.b.z, k.b.w), a2, acc3);
                accs = fmaf(k.c, a2.x, accs);
            }
            sp0 = ffma2(acc0, h2, yp0); sp1 = ffma2(acc1, h2, yp1); sp2 = ffma2(acc2, h2, yp2); sp3 = ffma2(acc3, h2, yp3);
            swz = fmaf(accs, h, y[12]);
            sm = fmaf(cf_C[s] * h, c.dm, y[13]);
            const float sq[4] = {sp1.y, sp2.x, sp2.y, sp3.x};
            rhs6f(c, sq, y[10], sp3.y, swz, sm, kk);
            if (s <= 5) kf_store(sKb, s, kk);
        }
        // ---- error norm (rk.py:101-107).  Position rows by quadrature of the acceleration rows (see rk45_solve6):
        //      sum_s b_s v_s = v + h sum_j beta_j K_j,  sum_s e_s v_s = h sum_j eps_j K_j
        float rn[3], ssum = 0.0f;
        {
            constexpr float BT0 = (float)(35.0 / 384), BT2 = (float)(50.0 / 159), BT3 = (float)(25.0 / 192), BT4 = (float)(-243.0 / 6784);
            constexpr float EP0 = (float)(-611.0 / 230400), EP2 = (float)(514.0 / 83475), EP3 = (float)(-391.0 / 38400),
                            EP4 = (float)(4617.0 / 1356800), EP5 = (float)(11.0 / 3360);
            const KRow k0 = kf_load(sKb, 0), k2 = kf_load(sKb, 2), k3 = kf_load(sKb, 3), k4 = kf_load(sKb, 4), k5 = kf_load(sKb, 5);
#define PAIR(k, i) ((i) == 0 ? make_float2(k.a.x, k.a.y) : (i) == 1 ? make_float2(k.a.z, k.a.w) : (i) == 2 ? make_float2(k.b.x, k.b.y) : make_float2(k.b.z, k.b.w))
            float2 es[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 e = fmul2(PAIR(k0, i), bc2((float)E0));
                e = ffma2(PAIR(k2, i), bc2((float)E2), e);
                e = ffma2(PAIR(k3, i), bc2((float)E3), e);
                e = ffma2(PAIR(k4, i), bc2((float)E4), e);
                e = ffma2(PAIR(k5, i), bc2((float)E5), e);
                es[i] = ffma2(make_float2(kk[2 * i], kk[2 * i + 1]), bc2((float)E6), e);
            }
            const float ess = ((((k0.c * (float)E0 + k2.c * (float)E2) + k3.c * (float)E3) + k4.c * (float)E4) + k5.c * (float)E5) + kk[8] * (float)E6;
            // position: pair (r0, r1) from the (dv0, dv1) rows, r2 from dv2
            float2 sb = fmul2(PAIR(k0, 0), bc2(BT0)), se = fmul2(PAIR(k0, 0), bc2(EP0));
            sb = ffma2(PAIR(k2, 0), bc2(BT2), sb); se = ffma2(PAIR(k2, 0), bc2(EP2), se);
            sb = ffma2(PAIR(k3, 0), bc2(BT3), sb); se = ffma2(PAIR(k3, 0), bc2(EP3), se);
            sb = ffma2(PAIR(k4, 0), bc2(BT4), sb); se = ffma2(PAIR(k4, 0), bc2(EP4), se);
            se = ffma2(PAIR(k5, 0), bc2(EP5), se);
            const float sb2 = ((k0.a.z * BT0 + k2.a.z * BT2) + k3.a.z * BT3) + k4.a.z * BT4;
            const float se2 = (((k0.a.z * EP0 + k2.a.z * EP2) + k3.a.z * EP3) + k4.a.z * EP4) + k5.a.z * EP5;
#undef PAIR
            const float hsq = h * h;
            const float sbv[3] = {sb.x, sb.y, sb2}, sev[3] = {se.x, se.y, se2};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                rn[i] = y[i] + fmaf(hsq, sbv[i], h * y[3 + i]);
                const float e = (sev[i] * hsq) * rcpf_(fmaf(absmaxf_(y[i], rn[i]), RTOLF, ATOLF));
                ssum = fmaf(e, e, ssum);
            }
            const float esv[9] = {es[0].x, es[0].y, es[1].x, es[1].y, es[2].x, es[2].y, es[3].x, es[3].y, ess};
            const float ynv[9] = {sp0.x, sp0.y, sp1.x, sp1.y, sp2.x, sp2.y, sp3.x, sp3.y, swz};
#pragma unroll
            for (int cc = 0; cc < NKF; ++cc) {
                const float e = (esv[cc] * h) * rcpf_(fmaf(absmaxf_(y[YIDX(cc)], ynv[cc]), RTOLF, ATOLF));
                ssum = fmaf(e, e, ssum);
            }
        }
        const float sv[3] = {sp0.x, sp0.y, sp1.x};
        const float sq[4] = {sp1.y, sp2.x, sp2.y, sp3.x};
        const float swy = sp3.y;
        const float err = sqrtf_(ssum) * INV_RS14F;
        if (err < 1.0f) {
            float factor = (err == 0.0f) ? 10.0f : fminf(10.0f, 0.9f * powf_m02(err));
            if (rejected) factor = fminf(1.0f, factor);
            h_abs *= factor;
            const float g = y[0], g_new = rn[0];
            if (((g <= 0) && (g_new >= 0)) || ((g >= 0) && (g_new <= 0))) {
                float ebuf[28];
#pragma unroll
                for (int i = 0; i < 14; ++i) ebuf[i] = y[i];
#pragma unroll
                for (int cc = 0; cc < NKF; ++cc) ebuf[14 + cc] = kk[cc];
#pragma unroll
                for (int i = 0; i < 3; ++i) ebuf[23 + i] = sv[i];
                ebuf[26] = c.dm; ebuf[27] = h;
                event_path6f(sKb, ebuf);
#pragma unroll
                for (int i = 0; i < 14; ++i) y[i] = ebuf[i];
                status = 1;
                break;
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) { y[i] = rn[i]; y[3 + i] = sv[i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i) y[6 + i] = sq[i];
            y[11] = swy; y[12] = swz; y[13] = sm;
            kf_store(sKb, 0, kk);                                        // FSAL
            t = t_new;
            if (t - dt >= 0.0f) { status = 0; break; }
            rejected = false;
            min_step = min_stepf_at(t);
            if (h_abs < min_step) h_abs = min_step;
        } else {
            h_abs *= fmaxf(0.2f, 0.9f * powf_m02(err));
            rejected = true;
        }
    }
#undef YIDX
    return status;
}

struct EnvStateF {
    float y[14];
    int32_t k, episode;
    float v0;
    double ep_ret;
};

RLR_DEV void load_state_f(const RlrState& gs, int64_t n, int64_t i, EnvStateF& st) {
    const float* y = (const float*)gs.y + i;
#pragma unroll
    for (int c = 0; c < 14; ++c, y += n) st.y[c] = *y;
    st.k = gs.k[i]; st.episode = gs.episode[i]; st.v0 = gs.v0[i]; st.ep_ret = gs.ep_ret[i];
}
RLR_DEV void store_state_f(const RlrState& gs, int64_t n, int64_t i, const EnvStateF& st) {
    float* y = (float*)gs.y + i;
#pragma unroll
    for (int c = 0; c < 14; ++c, y += n) *y = st.y[c];
    gs.k[i] = st.k; gs.episode[i] = st.episode; gs.v0[i] = st.v0; gs.ep_ret[i] = st.ep_ret;
}

// reset from an f32 IC (rocket_env.py:665-688): the binary32 state holds the IC exactly
RLR_DEV void reset_from_ic_f(const RlrParams& p, float (&ic)[14], EnvStateF& st, float (&obs)[14]) {
    const float nq = f32_norm4(ic[6], ic[7], ic[8], ic[9]);
#pragma unroll
    for (int i = 6; i < 10; ++i) ic[i] = fdiv(ic[i], nq);
    st.v0 = f32_norm3(ic[3], ic[4], ic[5]);
#pragma unroll
    for (int i = 0; i < 14; ++i) { st.y[i] = ic[i]; obs[i] = (float)((double)ic[i] * p.inv_normalizer[i]); }
    st.k = 0;
    st.ep_ret = 0.0;
}

RLR_DEV void env_step_f(const RlrParams& p, EnvStateF& st, const float (&u)[3], float a_thr, float (&obs)[14], StepRes& res, StepFlags& fl,
                        float* sK) {
    Ctx6f c;
    {
        const float cy = np_cosf(u[0]), sy = np_sinf(u[0]), cz = np_cosf(u[1]), sz = np_sinf(u[1]);
        c.Tb0 = fmul(cy, cz) * u[2];
        c.Tb1 = fmul(sy, cz) * u[2];
        c.Tb2 = sz * u[2];
        c.tau1 = 15.0f * c.Tb2;
        c.tau2 = -15.0f * c.Tb1;
        c.dm = fdiv(-u[2], 3531.6f);
    }
    int natt;
    const int status = rk45_solve6f(c, (float)p.dt, st.y, natt, sK);
    {
        const float n2 = ((st.y[6] * st.y[6] + st.y[7] * st.y[7]) + st.y[8] * st.y[8]) + st.y[9] * st.y[9];
        const float rnq = rsqrtf_(n2);
#pragma unroll
        for (int i = 6; i < 10; ++i) st.y[i] *= rnq;               // simulator.py:250
    }
#pragma unroll
    for (int i = 0; i < 14; ++i) obs[i] = st.y[i] * (float)p.inv_normalizer[i];
    reward6(p, st.y, u, a_thr, st.v0, status, res);
    st.k += 1;
    st.ep_ret += res.reward;
    fl.trunc = 0;
    fl.done = res.done;
    if (p.max_episode_steps > 0 && st.k >= p.max_episode_steps) { fl.trunc = !res.done; fl.done = 1; }
    fl.viol = res.viol; fl.landed = res.landed; fl.status = status; fl.natt = natt;
}

// coalesced observation rows through the warp's own (idle) float4 planes of stage rows 0 and 1 -- see write_obs6_warp
RLR_DEV void write_obs6_warp_f(const RlrParams& p, float* dst, int64_t i, bool active, const float (&obs)[14], float* sKb, bool stage) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* tile_dst = dst + (i - lane) * 14;
    const bool staged = stage && __all_sync(0xffffffffu, active) && ((reinterpret_cast<uintptr_t>(tile_dst) & 15) == 0);
    if (!staged) {
        if (active) write_obs<Sys6>(p, dst, i, obs);
        return;
    }
    // plane k (128 floats per warp): row 0 plane a, row 0 plane b, row 1 plane a, row 1 plane b
    auto plane = [&](int k) { return sKb + (k >> 1) * ROWF + (k & 1) * 4 * BSF + w * 128; };
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const int o4 = lane * 14 + 2 * k;
        *reinterpret_cast<float2*>(plane(o4 >> 7) + (o4 & 127)) = make_float2(obs[2 * k], obs[2 * k + 1]);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int f = k * 32 + lane;
        if (f < 112) reinterpret_cast<float4*>(tile_dst)[f] = *reinterpret_cast<const float4*>(plane(k) + 4 * lane);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(BSF) step_kernel_f32(const __grid_constant__ RlrParams p, RlrState gs, const float* actions, int action_kind,
                                                       const float* ic_inject, RlrStepOut out, int64_t n, int64_t i_begin, int64_t i_end,
                                                       int64_t env_offset, uint64_t seed, bool stage_obs) {
    extern __shared__ float sKf[];
    pdl_launch_dependents();
    const int64_t i = i_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < i_end;
    StepFlags fl = {0, 0, 0, 0, 0, 0};
    int ep_len = 0;
    double ep_ret = 0.0;
    EnvStateF st;
    st.episode = 0;
    float obs[14];
#if RLR_PREFETCH
    {   // one line per thread, one instruction, before the wait (see prefetch_block_inputs6): 14 x 512 B of state, 1 KB of returns, 1.5 KB of actions
        const int64_t i0 = i_begin + (int64_t)blockIdx.x * BSF;
        const int t = threadIdx.x;
        if (i0 + BSF <= i_end && t < 76) {
            const char* ptr;
            if (t < 56) ptr = (const char*)((const float*)gs.y + (int64_t)(t >> 2) * n + i0) + (t & 3) * 128;
            else if (t < 64) ptr = (const char*)(gs.ep_ret + i0) + (t - 56) * 128;
            else ptr = (const char*)(actions + i0 * 3) + (t - 64) * 128;
            prefetch_l2(ptr);
        }
    }
#endif
    pdl_wait_for_predecessor();
    if (active) {
        load_state_f(gs, n, i, st);
        float u[3], a_thr;
        bool thr_is_row;
        map_action<Sys6>(p, actions, action_kind, i, u, a_thr, thr_is_row);
        StepRes res;
        env_step_f(p, st, u, a_thr, obs, res, fl, sKf);
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
            float* yp = (float*)out.y_pre_reset;
#pragma unroll
            for (int c = 0; c < 14; ++c) yp[c * n + i] = st.y[c];
        }
        if (out.vtarg) {
#pragma unroll
            for (int c = 0; c < 4; ++c) out.vtarg[c * n + i] = res.vt[c];
        }
        if (fl.done) {
            if (out.terminal_obs) write_obs<Sys6>(p, out.terminal_obs, i, obs);
            if (out.fin_ret) out.fin_ret[i] = st.ep_ret;
            if (out.fin_len) out.fin_len[i] = st.k;
            if (out.fin_records) append_fin_record<14>(p, out, n, i, st.ep_ret, st.k, fl, obs);
        }
    }
    // the finished lanes' reset (initial condition, new state, new observation) is computed and stored by the whole warp
    // (store_obs_and_reset6_warp, rocket_kernels.cuh); staging planes: row 0 plane a, row 0 plane b, row 1 plane a, row 1 plane b
    const bool finished = active && fl.done;
    if (active && !finished) store_state_f(gs, n, i, st);
    float* const sKw = sKf + (threadIdx.x >> 5) * 128;
    store_obs_and_reset6_warp<float>(p, gs, out.obs, n, i, active, finished, st.episode + 1, obs,
                                     [sKw](int k) { return sKw + (k >> 1) * ROWF + (k & 1) * 4 * BSF; }, stage_obs, ic_inject, seed, env_offset + i);
    accumulate_stats(out.stats, active, fl, ep_len, ep_ret);
}

__global__ void __launch_bounds__(128) reset_kernel_f32(const __grid_constant__ RlrParams p, RlrState gs, const uint8_t* mask,
                                                        const float* ic_inject, float* obs_out, int64_t n, int64_t env_offset, uint64_t seed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mask != nullptr && mask[i] == 0) return;
    EnvStateF st;
    st.episode = gs.episode[i];
    float ic[14], obs[14];
    if (ic_inject != nullptr) {
#pragma unroll
        for (int c = 0; c < 14; ++c) ic[c] = ic_inject[i * 14 + c];
    } else {
        sample_ic<14>(p, seed, env_offset + i, st.episode, ic);
    }
    reset_from_ic_f(p, ic, st, obs);
    if (obs_out) write_obs<Sys6>(p, obs_out, i, obs);
    store_state_f(gs, n, i, st);
}

__global__ void __launch_bounds__(BSF) rollout_kernel_f32(const __grid_constant__ RlrParams p, RlrState gs, const float* actions, int action_kind,
                                                          int T, float* obs_buf, float* reward_buf, uint8_t* done_buf, uint8_t* trunc_buf,
                                                          float* last_obs, unsigned long long* stats, int64_t n, int64_t env_offset, uint64_t seed) {
    extern __shared__ float sKf[];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    EnvStateF st;
    st.episode = 0;
    if (active) load_state_f(gs, n, i, st);
    float obs[14];
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
        StepFlags fl = {0, 0, 0, 0, 0, 0};
        int ep_len = 0;
        double ep_ret = 0.0;
        if (active) {
            float u[3], a_thr;
            bool thr_is_row;
            map_action<Sys6>(p, actions + (int64_t)t * n * 3, action_kind, i, u, a_thr, thr_is_row);
            StepRes res;
            env_step_f(p, st, u, a_thr, obs, res, fl, sKf);
            ep_len = st.k; ep_ret = st.ep_ret;
            if (reward_buf) reward_buf[(int64_t)t * n + i] = (float)res.reward;
            if (done_buf) done_buf[(int64_t)t * n + i] = (uint8_t)fl.done;
            if (trunc_buf) trunc_buf[(int64_t)t * n + i] = (uint8_t)fl.trunc;
        }
        float ic[14];
        const bool finished = active && fl.done;
        sample_ic_warp<14>(p, seed, env_offset + i, st.episode + 1, finished, ic);
        if (active) {
            if (finished) { st.episode += 1; reset_from_ic_f(p, ic, st, obs); }
            if (obs_buf) write_obs<Sys6>(p, obs_buf + (int64_t)t * n * 14, i, obs);
        }
        accumulate_stats(stats, active, fl, ep_len, ep_ret);
    }
    if (active) {
        if (last_obs && T > 0) write_obs<Sys6>(p, last_obs, i, obs);
        store_state_f(gs, n, i, st);
    }
}

}  // namespace f32
