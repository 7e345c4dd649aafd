/*
 * rocket_oracle.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C scalar restatement of the RL_rocket environment step as the
 * reference executes it in this image (numpy 2.x "NEP-50" promotion, scipy
 * 1.18 solve_ivp/RK45).  It is the checker for the CUDA path; it is never the
 * product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load it.
 *
 * Parity status: PINNED by execution -- the reference has no asserting tests,
 * so this file is validated against trajectories produced by the reference's
 * own sources run in the dev container (oracle/live_reference.py,
 * oracle/make_golden.py -> tests/golden/), see tests/test_oracle_golden.py.
 *
 * What each block follows (reference = /root/reference, scipy = installed 1.18.1):
 *   rk45_solve            scipy/integrate/_ivp/rk.py:14-71,111-176,541-565,715-737
 *                         scipy/integrate/_ivp/common.py:63-134 (select_initial_step)
 *                         scipy/integrate/_ivp/ivp.py:52-77,134-158,651-700 (events)
 *                         call sites my_environment/utils/simulator.py:64-69,236-241
 *   brentq                scipy/optimize/Zeros/brentq.c (restated)
 *   rhs6 / consts6        my_environment/utils/simulator.py:259-378
 *   rhs3 / consts3        my_environment/utils/simulator.py:88-130
 *   sim6_step / sim3_step my_environment/utils/simulator.py:227-257 / :55-86
 *   env6_step             my_environment/envs/rocket_env.py:690-719,825-859,963-1014,1036-1061
 *   env3_step             my_environment/envs/rocket_env.py:150-247,395-406,431-476
 *   annealing / discrete  my_environment/wrappers/wrappers.py:24-35,68-86
 *   TimeLimit/auto-reset  gym 0.21 TimeLimit, SB3 1.6 DummyVecEnv.step_wait
 *   reset RNG             NEW (Philox4x32-10 keyed by seed/env/episode); the reference's
 *                         gym RandomState stream is not reproduced, parity runs inject ICs.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/build_oracle.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define NMAX 14

/* ------------------------------------------------------------------ params */
typedef struct {
    int    dof;                 /* 6, 3 or 1 */
    int    reward_mode;         /* 0 = shaped (default), 1 = RewardAnnealing */
    int    max_episode_steps;   /* TimeLimit; <=0 = none */
    int    dt_ms;               /* timestep in integer milliseconds */
    double dt;                  /* timestep (python float) */
    double normalizer[NMAX];
    float  ic_low[NMAX], ic_high[NMAX];
    /* 6DOF position bounds (f32 Box) / 3DOF bounds (f64) */
    float  pos_low[3], pos_high[3];
    double x_left, x_right, y_up;
    /* reward */
    double alfa, beta, eta, gamma, delta, kappa, xi;
    double waypoint, landing_radius, maximum_velocity;
    double att_traj_limit[3], att_land_limit[3], omega_lim[3];
    double max_gimbal;          /* deg2rad(20) */
    float  max_thrust;          /* 981e3 */
    int    obs_mode;            /* 0 normalised state; 1 (3DOF) GaudetStateObs, corrected (wrappers.py:51-66): see gaudet_obs3 */
    float  min_thrust;          /* EXTENSION (no reference behaviour: the TODO at rocket_env.py:405, :980): lower bound of the
                                   de-normalised thrust; 0 = off */
    int    aero_mode;           /* EXTENSION, 3DOF / 1DOF: 0 = the reference's behaviour (alfa = 0, simulator.py:103); 1 = the line the
                                   reference keeps commented out (simulator.py:104, `alfa = self._computeAoA(...)`) enabled, with
                                   its two defects fixed: the state is passed (not the height), and the missing `_normalize` is
                                   the wrap to (-pi, pi].  Everything downstream is the reference's own RHS (Cn, N, the N terms of
                                   ax, ay, dom; simulator.py:107-127).  See rhs3. */
} OrParams;

/* ------------------------------------------------------------ small helpers */
static float np_sinf_small(float x) {        /* numpy f32 sin, |x| < pi/4 */
    float x2 = x * x, r;
    r = fmaf(0x1.7d3bbcp-19f, x2, -0x1.a06bbap-13f);
    r = fmaf(r, x2, 0x1.11119ap-07f);
    r = fmaf(r, x2, -0x1.555556p-03f);
    r = fmaf(r, x2, 0.0f);
    r = fmaf(r, x, x);
    return r;
}
static float np_cosf_small(float x) {        /* numpy f32 cos, |x| < pi/4 */
    float x2 = x * x, r;
    r = fmaf(0x1.98e616p-16f, x2, -0x1.6c06dcp-10f);
    r = fmaf(r, x2, 0x1.55553cp-05f);
    r = fmaf(r, x2, -0x1.0p-01f);
    r = fmaf(r, x2, 1.0f);
    return r;
}
/* numpy's float32 sin / cos ufunc loops for any |x| up to the Cody-Waite limit (numpy/_core/src/umath/
 * loops_trigonometric.dispatch: cos 71476.0625, sin 117435.992; call sites simulator.py:351-357, :124):
 *   quadrant  k = rint(x * 2/pi) via the 1.5 * 2^23 magic-number add -- the multiply-add is FUSED
 *   reduction r = fma(k, -pi/2_c, ...) in three f32 terms
 *   both polynomials on r, pick sin or cos by the parity of k (k + 1 for cosine), negate when bit 1 is set.
 * For |x| < pi/4, k = 0 and r = x: the small-argument polynomials above.  Checked in this image against
 * np.sin / np.cos on EVERY float32 with |x| <= 64 (2.2e9 arguments per function, 0 mismatches) and on 8e6
 * random arguments up to +-7e4; tests/test_oracle_golden.py re-checks samples.  Beyond the limit numpy
 * calls libm per element; restated as the correctly rounded double function (not pinned, unreachable
 * for |action| < 2e5). */
static float np_sincosf(float x, int want_cos) {
    const float lim = want_cos ? 71476.0625f : 117435.992f;
    if (!(fabsf(x) <= lim)) return want_cos ? (float)cos((double)x) : (float)sin((double)x);
    const float magic = 0x1.8p23f;
    volatile float t = fmaf(x, 0.636619772367581343075535053490057448f, magic);
    const float k = t - magic;
    float r = fmaf(k, -0x1.921fb0p+00f, x);
    r = fmaf(k, -0x1.5110b4p-22f, r);
    r = fmaf(k, -0x1.846988p-48f, r);
    int ik = (int)k;
    if (want_cos) ik += 1;
    float v = ((ik & 1) == 0) ? np_sinf_small(r) : np_cosf_small(r);
    if ((ik & 2) == 2) v = 0.0f - v;
    return v;
}
static float np_sinf(float x) { return np_sincosf(x, 0); }
static float np_cosf(float x) { return np_sincosf(x, 1); }

/* numpy float32 linalg.norm on a short vector: sdot with a double accumulator of
 * f32-rounded products, result rounded to f32, then sqrtf (SURVEY "hard parts"). */
static float f32_norm(const float* x, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) { float p = x[i] * x[i]; acc += (double)p; }
    return sqrtf((float)acc);
}
static double f64_norm(const double* x, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc += x[i] * x[i];
    return sqrt(acc);
}
static double py_max(double a, double b) { return (b > a) ? b : a; }   /* python max(a, b) */
static double py_min(double a, double b) { return (b < a) ? b : a; }   /* python min(a, b) */

/* ------------------------------------------------------------------- RHS */
typedef struct {
    int    n;           /* 14 or 7 */
    int    event_idx;   /* 0 (6DOF height x) or 1 (3DOF height z) */
    /* 6DOF per-step constants */
    double Tb[3], tau[3], dm;
    /* 3DOF per-step constants */
    float  delta3, T3;
    double torque3;     /* (double)( f32(T*sinf(delta)) * 30f ) */
    int    aero;        /* OrParams.aero_mode */
} RhsCtx;

static const double J6[3]    = {75350.25, 6037675.13, 6037675.13};

static void rhs6(const RhsCtx* c, const double* y, double* d) {
    const double *v = y + 3, *q = y + 6, *w = y + 10;
    double m = y[13];
    /* Rotation.from_quat([q1,q2,q3,q0]) normalises, then as_matrix() */
    /* scipy normalises the (x,y,z,w) array: norm = sqrt(x^2+y^2+z^2+w^2) */
    double nq = sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3] + q[0] * q[0]);
    double x = q[1] / nq, yy = q[2] / nq, z = q[3] / nq, ww = q[0] / nq;
    double x2 = x * x, y2 = yy * yy, z2 = z * z, w2 = ww * ww;
    double xy = x * yy, zw = z * ww, xz = x * z, yw = yy * ww, yz = yy * z, xw = x * ww;
    double R[3][3];
    R[0][0] = x2 - y2 - z2 + w2; R[1][0] = 2 * (xy + zw);     R[2][0] = 2 * (xz - yw);
    R[0][1] = 2 * (xy - zw);     R[1][1] = -x2 + y2 - z2 + w2; R[2][1] = 2 * (yz + xw);
    R[0][2] = 2 * (xz + yw);     R[1][2] = 2 * (yz - xw);     R[2][2] = -x2 - y2 + z2 + w2;
    double F[3];
    for (int i = 0; i < 3; ++i) F[i] = R[i][0] * c->Tb[0] + R[i][1] * c->Tb[1] + R[i][2] * c->Tb[2];
    double inv_m = 1.0 / m;
    d[0] = v[0]; d[1] = v[1]; d[2] = v[2];
    d[3] = inv_m * F[0] + (-9.81);
    d[4] = inv_m * F[1] + 0.0;
    d[5] = inv_m * F[2] + 0.0;
    /* dq = 0.5 * OMEGA(w) . q  (un-normalised q) */
    double wx = w[0], wy = w[1], wz = w[2];
    d[6] = 0.5 * (0.0 * q[0] + (-wx) * q[1] + (-wy) * q[2] + (-wz) * q[3]);
    d[7] = 0.5 * (wx * q[0] + 0.0 * q[1] + wz * q[2] + (-wy) * q[3]);
    d[8] = 0.5 * (wy * q[0] + (-wz) * q[1] + 0.0 * q[2] + wx * q[3]);
    d[9] = 0.5 * (wz * q[0] + wy * q[1] + (-wx) * q[2] + 0.0 * q[3]);
    /* dom = Jinv . (tau - w x (J w)) */
    double Jw[3] = {J6[0] * wx, J6[1] * wy, J6[2] * wz};
    double cr[3] = {wy * Jw[2] - wz * Jw[1], wz * Jw[0] - wx * Jw[2], wx * Jw[1] - wy * Jw[0]};
    for (int i = 0; i < 3; ++i) d[10 + i] = (1.0 / J6[i]) * (c->tau[i] - cr[i]);
    d[13] = c->dm;
}

static void rhs3(const RhsCtx* c, const double* y, double* d) {
    double phi = y[2], vx = y[3], vz = y[4], om = y[5], mass = y[6];
    double T = (double)c->T3;
    double v2 = vx * vx + vz * vz;
    double Q = 0.5 * 1.225 * v2;
    double A = 0.3 * Q * 10.5;
    double N = 0.0;
    if (c->aero) {
        /* _computeAoA (simulator.py:132-148) on the state: alfa = phi - arctan2(vz, vx) unless the velocity is exactly zero, then
         * `_normalize` := _wrapTo2Pi (simulator.py:150-163) folded to (-pi, pi]; Cn = Cnalfa * alfa with Cnalfa = 1, N = Cn * Q * Sref */
        const double pi = 3.141592653589793, pi2 = 2.0 * 3.141592653589793;
        double alfa = 0.0;
        if (!(vx == 0.0 && vz == 0.0)) alfa = phi - atan2(vz, vx);
        alfa = fmod(fmod(alfa, pi2) + pi2, pi2);
        if (alfa > pi) alfa = alfa - pi2;
        double Cn = 1 * alfa;
        N = Cn * Q * 10.5;
    }
    double dphi = (double)c->delta3 + phi;
    double ax = (T * cos(dphi) - N * sin(phi) - A * cos(phi)) / mass;
    double ay = (T * sin(dphi) + N * cos(phi) - A * cos(phi)) / mass - 9.81;
    double dom = (N * (double)(10 - 20) - c->torque3) / 6.04e6;
    d[0] = vx; d[1] = vz; d[2] = om; d[3] = ax; d[4] = ay; d[5] = dom; d[6] = c->dm;
}

static void rhs(const RhsCtx* c, const double* y, double* d) {
    if (c->n == 14) rhs6(c, y, d); else rhs3(c, y, d);
}

/* --------------------------------------------------- Dormand-Prince tableau */
static const double RK_C[6] = {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1};
static const double RK_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double RK_B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double RK_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};
static const double RK_P[7][4] = {
    {1, -8048581381.0 / 2820520608, 8663915743.0 / 2820520608, -12715105075.0 / 11282082432},
    {0, 0, 0, 0},
    {0, 131558114200.0 / 32700410799, -68118460800.0 / 10900136933, 87487479700.0 / 32700410799},
    {0, -1754552775.0 / 470086768, 14199869525.0 / 1410260304, -10690763975.0 / 1880347072},
    {0, 127303824393.0 / 49829197408, -318862633887.0 / 49829197408, 701980252875.0 / 199316789632},
    {0, -282668133.0 / 205662961, 2019193451.0 / 616988883, -1453857185.0 / 822651844},
    {0, 40617522.0 / 29380423, -110615467.0 / 29380423, 69997945.0 / 29380423}};

static double rms(const double* x, int n) { return f64_norm(x, n) / sqrt((double)n); }

/* dense output: y(t) = y_old + h * Q . [x, x^2, x^3, x^4] */
static void dense_eval(int n, const double Q[NMAX][4], const double* y_old, double t_old, double h, double t, double* out) {
    double x = (t - t_old) / h;
    double p[4]; p[0] = x; p[1] = p[0] * x; p[2] = p[1] * x; p[3] = p[2] * x;
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int c = 0; c < 4; ++c) s += Q[i][c] * p[c];
        out[i] = h * s + y_old[i];
    }
}

typedef struct { int n; const double (*Q)[4]; const double* y_old; double t_old, h; int idx; } EvCtx;
static double ev_fun(const EvCtx* e, double t) {
    double tmp[NMAX];
    dense_eval(e->n, e->Q, e->y_old, e->t_old, e->h, t, tmp);
    return tmp[e->idx];
}

static double brentq(const EvCtx* e, double xa, double xb) {
    const double xtol = 4 * 2.220446049250313e-16, rtol = 4 * 2.220446049250313e-16;
    double xpre = xa, xcur = xb, xblk = 0, fblk = 0, spre = 0, scur = 0;
    double fpre = ev_fun(e, xpre), fcur = ev_fun(e, xcur);
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    for (int it = 0; it < 100; ++it) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre; fblk = fpre; spre = scur = xcur - xpre;
        }
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
            double lim = fmin(fabs(spre), 3 * fabs(sbis) - delta);
            if (2 * fabs(stry) < lim) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else { spre = sbis; scur = sbis; }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = ev_fun(e, xcur);
    }
    return xcur;
}

/* solve_ivp(fun, [t0, t_bound], y, events=terminal height event), RK45 defaults.
 * y is replaced by solution.y[:, -1].  Returns status (0 end, 1 event, -1 failed). */
static int rk45_solve(const RhsCtx* c, double t0, double t_bound, double* y, int* nfev_out, int* natt_out) {
    const int n = c->n;
    const double rtol = 1e-3, atol = 1e-6;
    double K[7][NMAX], f[NMAX], scale[NMAX], tmp[NMAX], y1[NMAX], f1[NMAX], y_new[NMAX], f_new[NMAX];
    int nfev = 0, natt = 0;
    double t = t0;
    rhs(c, y, f); nfev++;
    /* --- select_initial_step --- */
    double h_abs;
    {
        double interval = fabs(t_bound - t0);
        if (interval == 0.0) { h_abs = 0.0; }
        else {
            for (int i = 0; i < n; ++i) scale[i] = atol + fabs(y[i]) * rtol;
            for (int i = 0; i < n; ++i) tmp[i] = y[i] / scale[i];
            double d0 = rms(tmp, n);
            for (int i = 0; i < n; ++i) tmp[i] = f[i] / scale[i];
            double d1 = rms(tmp, n);
            double h0;
            if (d0 < 1e-5 || d1 < 1e-5) h0 = 1e-6; else h0 = 0.01 * d0 / d1;
            h0 = py_min(h0, interval);
            for (int i = 0; i < n; ++i) y1[i] = y[i] + h0 * 1.0 * f[i];
            rhs(c, y1, f1); nfev++;
            for (int i = 0; i < n; ++i) tmp[i] = (f1[i] - f[i]) / scale[i];
            double d2 = rms(tmp, n) / h0;
            double h1;
            if (d1 <= 1e-15 && d2 <= 1e-15) h1 = py_max(1e-6, h0 * 1e-3);
            else h1 = pow(0.01 / py_max(d1, d2), 1.0 / 5.0);
            h_abs = py_min(py_min(py_min(100 * h0, h1), interval), INFINITY);
        }
    }
    int status = 2;          /* running */
    double g = y[c->event_idx];
    if (t == t_bound) status = 0;   /* degenerate span: solver finishes without stepping */
    while (status == 2) {
        /* --- _step_impl --- */
        double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
        double ha = h_abs;
        if (ha < min_step) ha = min_step;
        int accepted = 0, rejected = 0, failed = 0;
        double h = 0, t_new = 0;
        while (!accepted) {
            if (ha < min_step) { failed = 1; break; }
            h = ha * 1.0;
            t_new = t + h;
            if (1.0 * (t_new - t_bound) > 0) t_new = t_bound;
            h = t_new - t;
            ha = fabs(h);
            natt++;
            /* rk_step */
            memcpy(K[0], f, sizeof(double) * n);
            for (int s = 1; s < 6; ++s) {
                for (int i = 0; i < n; ++i) {
                    double acc = 0.0;
                    for (int j = 0; j < s; ++j) acc += K[j][i] * RK_A[s][j];
                    tmp[i] = y[i] + acc * h;
                }
                rhs(c, tmp, K[s]); nfev++;
            }
            for (int i = 0; i < n; ++i) {
                double acc = 0.0;
                for (int j = 0; j < 6; ++j) acc += K[j][i] * RK_B[j];
                y_new[i] = y[i] + h * acc;
            }
            rhs(c, y_new, f_new); nfev++;
            memcpy(K[6], f_new, sizeof(double) * n);
            for (int i = 0; i < n; ++i) {
                double sc = atol + fmax(fabs(y[i]), fabs(y_new[i])) * rtol;
                double acc = 0.0;
                for (int j = 0; j < 7; ++j) acc += K[j][i] * RK_E[j];
                tmp[i] = (acc * h) / sc;
            }
            double err = rms(tmp, n);
            if (err < 1) {
                double factor;
                if (err == 0) factor = 10;
                else factor = py_min(10, 0.9 * pow(err, -0.2));
                if (rejected) factor = py_min(1, factor);
                ha *= factor;
                accepted = 1;
            } else {
                ha *= py_max(0.2, 0.9 * pow(err, -0.2));
                rejected = 1;
            }
        }
        if (failed) { status = -1; break; }
        double t_old = t;
        /* accept */
        double y_old[NMAX];
        memcpy(y_old, y, sizeof(double) * n);
        t = t_new;
        memcpy(y, y_new, sizeof(double) * n);
        memcpy(f, f_new, sizeof(double) * n);
        h_abs = ha;
        if (1.0 * (t - t_bound) >= 0) status = 0;
        /* --- events --- */
        double g_new = y[c->event_idx];
        int up = (g <= 0) && (g_new >= 0), down = (g >= 0) && (g_new <= 0);
        if (up || down) {
            double Q[NMAX][4];
            for (int i = 0; i < n; ++i)
                for (int cc = 0; cc < 4; ++cc) {
                    double acc = 0.0;
                    for (int j = 0; j < 7; ++j) acc += K[j][i] * RK_P[j][cc];
                    Q[i][cc] = acc;
                }
            EvCtx e = {n, (const double (*)[4])Q, y_old, t_old, t - t_old, c->event_idx};
            double root = brentq(&e, t_old, t);
            dense_eval(n, (const double (*)[4])Q, y_old, t_old, t - t_old, root, y);
            status = 1;
            t = root;
        }
        g = g_new;
    }
    if (nfev_out) *nfev_out = nfev;
    if (natt_out) *natt_out = natt;
    return status;
}

/* ------------------------------------------------------- action mapping */
/* rocket_env.py:969-981 with an f32 ndarray action (NEP-50): */
static void denorm_action6(const OrParams* p, const float* a, float* u) {
    u[0] = (float)((double)a[0] * p->max_gimbal);
    u[1] = (float)((double)a[1] * p->max_gimbal);
    u[2] = ((a[2] + 1.0f) / 2.0f) * p->max_thrust;
    if (p->min_thrust > 0.0f && !(u[2] >= p->min_thrust)) u[2] = (u[2] != u[2]) ? u[2] : p->min_thrust;   /* fmaxf, NaN passes */
}
static void denorm_action3(const OrParams* p, const float* a, float* u) {
    u[0] = (float)((double)a[0] * p->max_gimbal);
    u[1] = ((a[1] + 1.0f) / 2.0f) * p->max_thrust;
    if (p->min_thrust > 0.0f && !(u[1] >= p->min_thrust)) u[1] = (u[1] != u[1]) ? u[1] : p->min_thrust;
}

/* DiscreteActions3DOF.action (wrappers.py:34): the table row is an int64 ndarray, so Rocket._denormalize_action
 * (rocket_env.py:395-406) runs in float64 and only the final np.float32([...]) rounds.  `row` holds the integers as floats. */
static void denorm_action3_row(const OrParams* p, const float* row, float* u) {
    u[0] = (float)((double)row[0] * p->max_gimbal);
    u[1] = (float)((((double)row[1] + 1.0) / 2.0) * (double)p->max_thrust);
    if (p->min_thrust > 0.0f && u[1] < p->min_thrust) u[1] = p->min_thrust;
}

/* ------------------------------------------------------- simulators */
static double t_of_k(const OrParams* p, int k) { return (double)((int64_t)k * p->dt_ms) / 1000.0; }

static void consts6(RhsCtx* c, const float* u) {
    c->aero = 0;
    float cy = np_cosf(u[0]), sy = np_sinf(u[0]), cz = np_cosf(u[1]), sz = np_sinf(u[1]);
    float cycz = cy * cz, sycz = sy * cz;
    double T = (double)u[2];
    c->n = 14; c->event_idx = 0;
    c->Tb[0] = (double)cycz * T + 0.0 + 0.0;
    c->Tb[1] = (double)sycz * T + 0.0 + 0.0;
    c->Tb[2] = (double)sz * T + 0.0 + 0.0;
    c->tau[0] = 0.0;
    c->tau[1] = 0.0 * c->Tb[0] - (-15.0) * c->Tb[2];
    c->tau[2] = (-15.0) * c->Tb[1] - 0.0 * c->Tb[0];
    c->dm = (double)((-u[2]) / 3531.6f);
}
static void consts3(RhsCtx* c, const float* u) {
    c->n = 7; c->event_idx = 1; c->aero = 0;
    c->delta3 = u[0]; c->T3 = u[1];
    float ts = u[1] * np_sinf(u[0]);
    ts = ts * 30.0f;
    c->torque3 = (double)ts;
    c->dm = (double)((-u[1]) / 3531.6f);
}

/* Simulator6DOF.step: returns status; y in/out; k = step index before the step */
static int sim6_step(const OrParams* p, double* y, int k, const float* u, int* nfev, int* natt) {
    RhsCtx c; consts6(&c, u);
    double t = t_of_k(p, k);
    int st = rk45_solve(&c, t, t + p->dt, y, nfev, natt);
    double nq = sqrt(y[6] * y[6] + y[7] * y[7] + y[8] * y[8] + y[9] * y[9]);
    for (int i = 6; i < 10; ++i) y[i] = y[i] / nq;
    return st;
}
static int sim3_step(const OrParams* p, double* y, int k, const float* u, int* nfev, int* natt) {
    RhsCtx c; consts3(&c, u);
    c.aero = p->aero_mode;
    double t = t_of_k(p, k);
    int st = rk45_solve(&c, t, t + p->dt, y, nfev, natt);
    const double pi2 = 2.0 * 3.141592653589793;
    y[2] = fmod(fmod(y[2], pi2) + pi2, pi2);
    return st;
}

/* ------------------------------------------------------- rewards */
typedef struct {
    double reward;          /* total incl. bounds penalty (or annealed reward) */
    double terms[6];        /* velocity_tracking, thrust_penalty, eta, attitude_constraint, attitude_hint(3DOF)/0, rew_goal */
    int    viol, landed, status, done;
    double vtarg[3], t_go;
} StepOut;

static void euler_zyx(const float* s, double* ang) {
    /* Rotation.from_quat(roll(q_f32, -1)).as_euler("zyx") (extrinsic z, y, x) */
    double q0 = s[6], q1 = s[7], q2 = s[8], q3 = s[9];
    double nq = sqrt(q1 * q1 + q2 * q2 + q3 * q3 + q0 * q0);
    double x = q1 / nq, y = q2 / nq, z = q3 / nq, w = q0 / nq;
    double R00 = x * x - y * y - z * z + w * w, R01 = 2 * (x * y - z * w), R02 = 2 * (x * z + y * w);
    double R12 = 2 * (y * z - x * w), R22 = -x * x - y * y + z * z + w * w;
    ang[0] = atan2(-R01, R00);
    double r02 = R02 > 1 ? 1 : (R02 < -1 ? -1 : R02);
    ang[1] = asin(r02);
    ang[2] = atan2(-R12, R22);
}

static void reward6(const OrParams* p, const float* s, const float* u, const float* a_norm, float v0, int status, StepOut* o) {
    /* --- vtarg (rocket_env.py:986-1014) --- */
    double rh[3], vh[3], tau;
    if ((double)s[0] > p->waypoint) {
        rh[0] = (double)s[0] - p->waypoint; rh[1] = (double)s[1] - 0; rh[2] = (double)s[2] - 0;
        vh[0] = (double)s[3] - (-2); vh[1] = (double)s[4] - 0; vh[2] = (double)s[5] - 0; tau = 20;
    } else {
        rh[0] = (double)(s[0] + 1.0f); rh[1] = 0; rh[2] = 0;
        vh[0] = (double)s[3] - (-1); vh[1] = (double)s[4] - 0; vh[2] = (double)s[5] - 0; tau = 100;
    }
    double nr = f64_norm(rh, 3), nv = f64_norm(vh, 3);
    double t_go = nr / nv;
    double den = py_max(1e-3, nr);
    double fac = 1 - exp(-t_go / tau);
    double mv0 = (double)(-v0);
    double vt[3], dv[3];
    for (int i = 0; i < 3; ++i) { vt[i] = (mv0 * (rh[i] / den)) * fac; dv[i] = (double)s[3 + i] - vt[i]; o->vtarg[i] = vt[i]; }
    o->t_go = t_go;
    /* --- attitude / landing --- */
    double ang[3]; euler_zyx(s, ang);
    int att_viol = 0, att_ok = 0, om_ok = 0;
    for (int i = 0; i < 3; ++i) {
        if (fabs(ang[i]) > p->att_traj_limit[i]) att_viol = 1;
        if (fabs(ang[i]) < p->att_land_limit[i]) att_ok = 1;
        if ((double)fabsf(s[10 + i]) < p->omega_lim[i]) om_ok = 1;
    }
    float rn = f32_norm(s, 3), vn = f32_norm(s + 3, 3);
    int landed = (s[0] <= 1e-3f) && ((double)vn < p->maximum_velocity) && ((double)rn < p->landing_radius) && att_ok && om_ok;
    /* --- bounds (rocket_env.py:1036-1038) --- */
    int inside = 1;
    for (int i = 0; i < 3; ++i) inside = inside && (s[i] >= p->pos_low[i]) && (s[i] <= p->pos_high[i]);
    int viol = !inside;
    o->terms[0] = p->alfa * f64_norm(dv, 3);
    o->terms[1] = (double)((float)p->beta * u[2]);
    o->terms[2] = p->eta;
    o->terms[3] = p->gamma * (double)att_viol;
    o->terms[4] = 0.0;
    o->terms[5] = p->kappa * (double)landed;
    double r = 0.0;
    r = r + o->terms[0]; r = r + o->terms[1]; r = r + o->terms[2]; r = r + o->terms[3]; r = r + o->terms[5];
    if (viol) r += -50;
    if (p->reward_mode == 1) {
        /* wrappers.py:68-86: attitude_constraint + rew_goal + (-xi*(a2+1)) [f32 product], int + f32 -> f64 */
        float tp = (float)(-p->xi) * (a_norm[2] + 1.0f);
        r = (o->terms[3] + o->terms[5]) + (double)tp;
        o->terms[1] = (double)tp;
    }
    o->reward = r; o->viol = viol; o->landed = landed; o->status = status;
    o->done = (status != 0) || viol;
}

/* Rocket._compute_vtarg (rocket_env.py:219-247) on the f32 state */
static void vtarg3(const OrParams* p, const float* s, float v0, double* vt, double* t_go_out) {
    double rh[2], vh[2], tau;
    if ((double)s[1] > p->waypoint) {
        rh[0] = (double)s[0] - 0; rh[1] = (double)s[1] - p->waypoint;
        vh[0] = (double)s[3] - 0; vh[1] = (double)s[4] - (-2); tau = 20;
    } else {
        rh[0] = 0; rh[1] = (double)s[1];
        vh[0] = (double)s[3] - 0; vh[1] = (double)s[4] - (-1); tau = 100;
    }
    double nr = f64_norm(rh, 2), nv = f64_norm(vh, 2);
    double t_go = nr / nv;
    double den = py_max(1e-3, nr);
    double fac = 1 - exp(-t_go / tau);
    double mv0 = (double)(-v0);
    for (int i = 0; i < 2; ++i) vt[i] = (mv0 * (rh[i] / den)) * fac;
    *t_go_out = t_go;
}

/* GaudetStateObs.observation (wrappers.py:55-66) with its two defects fixed -- it calls `compute_vtarg`, which does not exist
 * (the method is `_compute_vtarg`, rocket_env.py:219), and passes it the NORMALISED r, v although the method is written for
 * the physical state (waypoint in metres, v_0 in m/s).  Fixed semantics: np.float32([vx - vt_x, vz - vt_z, t_go, z]) of the
 * physical float32 state the environment itself hands to _compute_vtarg (rocket_env.py:155, :180-184).  Pinned by
 * tests/golden/traj_3dof_gaudet.npz, generated with that corrected wrapper over the live reference env. */
static void gaudet_obs3(const OrParams* p, const float* s, float v0, float* obs) {
    double vt[2], t_go;
    vtarg3(p, s, v0, vt, &t_go);
    obs[0] = (float)((double)s[3] - vt[0]);
    obs[1] = (float)((double)s[4] - vt[1]);
    obs[2] = (float)t_go;
    obs[3] = s[1];
}

static void reward3(const OrParams* p, const float* s, const float* u, const float* a_norm, int a_is_row, float v0, int status, StepOut* o) {
    double vt[2], dv[2], t_go;
    vtarg3(p, s, v0, vt, &t_go);
    for (int i = 0; i < 2; ++i) { dv[i] = (double)s[3 + i] - vt[i]; o->vtarg[i] = vt[i]; }
    o->vtarg[2] = 0; o->t_go = t_go;
    float zeta = s[2] - 1.5707963267948966f;        /* f32 scalar - python float -> f32 */
    float az = fabsf(zeta);
    float rn = f32_norm(s, 2), vn = f32_norm(s + 3, 2);
    int landed = (s[1] <= 1e-3f) && (vn < 15.0f) && ((double)rn < p->landing_radius) && (az < 0.2f) && (fabsf(s[5]) < 0.2f);
    int viol = ((double)s[0] <= p->x_left) || ((double)s[0] >= p->x_right) || ((double)s[1] >= p->y_up);
    float hint_arg = az - 1.5707963267948966f;
    float hint = hint_arg > 0.0f ? hint_arg : 0.0f;   /* np.maximum(0, f32) */
    if (hint_arg != hint_arg) hint = hint_arg;
    o->terms[0] = p->alfa * f64_norm(dv, 2);
    o->terms[1] = (double)((float)p->beta * u[1]);
    o->terms[2] = p->eta;
    o->terms[3] = p->gamma * (double)(az > 6.283185307179586f);
    o->terms[4] = (double)((float)p->delta * hint);
    o->terms[5] = p->kappa * (double)landed;
    double r = 0.0;
    for (int i = 0; i < 6; ++i) r = r + o->terms[i];
    if (viol) r += -50;
    if (p->reward_mode == 1) {
        /* 3DOF: python float (ac + rg) + np.float32 -> f32 arithmetic */
        if (a_is_row) {
            /* DiscreteActions3DOF(RewardAnnealing(env)): the wrapper sees the int64 table row -> -xi * (row[1] + 1) in float64 */
            double tp = (-p->xi) * ((double)a_norm[1] + 1.0);
            r = (o->terms[3] + o->terms[5]) + tp;
            o->terms[1] = tp;
        } else {
            float tp = (float)(-p->xi) * (a_norm[1] + 1.0f);
            float rr = (float)(o->terms[3] + o->terms[5]) + tp;
            r = (double)rr;
            o->terms[1] = (double)tp;
        }
    }
    o->reward = r; o->viol = viol; o->landed = landed; o->status = status;
    o->done = (status != 0) || viol;
}

/* --------------------------------------------------------- exported: RNG */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
/* IC for (seed, global env index, episode index): uniform in [low, high] per component,
 * computed in f64 as low + (high-low)*u with u = 53 random bits * 2^-53, then cast to f32. */
void or_sample_ic(const OrParams* p, uint64_t seed, int64_t env, int32_t episode, float* ic) {
    int n = (p->dof == 6) ? 14 : 7;
    for (int b = 0; b * 2 < n; ++b) {
        uint32_t c[4] = {(uint32_t)b, (uint32_t)episode, (uint32_t)((uint64_t)env & 0xffffffffu), (uint32_t)((uint64_t)env >> 32)};
        philox4x32_10(c, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
        for (int h = 0; h < 2 && b * 2 + h < n; ++h) {
            int j = b * 2 + h;
            uint64_t bits = ((uint64_t)c[2 * h] << 32) | (uint64_t)c[2 * h + 1];
            double uu = (double)(bits >> 11) * 0x1.0p-53;
            double lo = (double)p->ic_low[j], hi = (double)p->ic_high[j];
            ic[j] = (float)(lo + (hi - lo) * uu);
        }
    }
}

/* reset: IC (f32) -> state (f64 holding f32 values), quaternion normalised in f32 (6DOF),
 * v0 (f32 norm of the IC velocity), obs. */
void or_reset_from_ic(const OrParams* p, const float* ic_in, double* y, float* v0, float* obs) {
    int n = (p->dof == 6) ? 14 : 7;
    float ic[NMAX];
    memcpy(ic, ic_in, sizeof(float) * n);
    if (p->dof == 6) {
        float nq = f32_norm(ic + 6, 4);
        for (int i = 6; i < 10; ++i) ic[i] = ic[i] / nq;
        *v0 = f32_norm(ic + 3, 3);
    } else {
        *v0 = f32_norm(ic + 3, 2);
    }
    for (int i = 0; i < n; ++i) { y[i] = (double)ic[i]; obs[i] = (float)((double)ic[i] / p->normalizer[i]); }
    if (p->dof == 3 && p->obs_mode == 1) gaudet_obs3(p, ic, *v0, obs);      /* ObservationWrapper.reset -> observation() */
}

/* One environment step WITHOUT TimeLimit / auto-reset.
 *  action_kind 0: normalised f32 action (3 / 2 / 1 components); 1: physical f32 u (3 / 2);
 *              2: (3DOF) a DiscreteActions3DOF table row [gimbal, throttle] given as floats holding the table's integers.
 *  y in/out (f64), k = episode step index before this step.
 *  outputs: obs f32[n] (6DOF: (float)(y/norm); 3DOF: (float)((double)s_f32/norm)); out->* */
void or_env_step(const OrParams* p, double* y, int k, float v0, const float* action, int action_kind,
                 float* obs, double* reward, double* terms, int* flags /* status, viol, landed, done, nfev, natt */,
                 double* vtarg_out) {
    StepOut o; memset(&o, 0, sizeof(o));
    int nfev = 0, natt = 0;
    float a_norm[3] = {0, 0, 0};
    if (p->dof == 6) {
        float u[3];
        if (action_kind == 0) { denorm_action6(p, action, u); memcpy(a_norm, action, sizeof(float) * 3); }
        else memcpy(u, action, sizeof(float) * 3);
        int st = sim6_step(p, y, k, u, &nfev, &natt);
        float s[14];
        for (int i = 0; i < 14; ++i) s[i] = (float)y[i];
        reward6(p, s, u, a_norm, v0, st, &o);
        for (int i = 0; i < 14; ++i) obs[i] = (float)(y[i] / p->normalizer[i]);
    } else {
        float u[2], a2[2];
        if (p->dof == 1) { a2[0] = 0.0f; a2[1] = action[0]; } else { a2[0] = action[0]; a2[1] = action[1]; }
        if (action_kind == 0) { denorm_action3(p, a2, u); a_norm[0] = a2[0]; a_norm[1] = a2[1]; }
        else if (action_kind == 2) { denorm_action3_row(p, a2, u); a_norm[0] = a2[0]; a_norm[1] = a2[1]; }
        else { u[0] = a2[0]; u[1] = a2[1]; }
        int st = sim3_step(p, y, k, u, &nfev, &natt);
        float s[7];
        for (int i = 0; i < 7; ++i) s[i] = (float)y[i];
        reward3(p, s, u, a_norm, action_kind == 2, v0, st, &o);
        for (int i = 0; i < 7; ++i) obs[i] = (float)((double)s[i] / p->normalizer[i]);
        if (p->dof == 3 && p->obs_mode == 1) gaudet_obs3(p, s, v0, obs);
    }
    *reward = o.reward;
    if (terms) memcpy(terms, o.terms, sizeof(o.terms));
    flags[0] = o.status; flags[1] = o.viol; flags[2] = o.landed; flags[3] = o.done; flags[4] = nfev; flags[5] = natt;
    if (vtarg_out) { vtarg_out[0] = o.vtarg[0]; vtarg_out[1] = o.vtarg[1]; vtarg_out[2] = o.vtarg[2]; vtarg_out[3] = o.t_go; }
}

/* Vectorised step with TimeLimit + auto-reset (DummyVecEnv.step_wait semantics).
 * State arrays are SoA [n][N]; obs is AoS [N][n_obs] (n_obs = n, or 2 for 1DOF).
 * ic_inject: optional [N][n] f32 ICs used for envs that reset in this call (else Philox). */
void or_vec_step(const OrParams* p, int64_t N, int64_t env_offset, uint64_t seed,
                 double* y, int32_t* k, int32_t* episode, float* v0, double* ep_ret,
                 const float* actions, int action_kind, const float* ic_inject,
                 float* obs, double* reward, uint8_t* done, uint8_t* truncated, float* terminal_obs,
                 double* fin_ret, int32_t* fin_len, uint8_t* fin_landed, int32_t* natt_out, double* y_pre_reset) {
    int n = (p->dof == 6) ? 14 : 7;
    int na = (p->dof == 6) ? 3 : (p->dof == 3 ? 2 : 1);
    int no = (p->dof == 1) ? 2 : ((p->dof == 3 && p->obs_mode == 1) ? 4 : n);
    for (int64_t i = 0; i < N; ++i) {
        double yy[NMAX]; float ob[NMAX]; double r; int fl[6];
        for (int c = 0; c < n; ++c) yy[c] = y[c * N + i];
        or_env_step(p, yy, k[i], v0[i], actions + i * na, action_kind, ob, &r, 0, fl, 0);
        int d = fl[3], tr = 0;
        int kk = k[i] + 1;
        if (p->max_episode_steps > 0 && kk >= p->max_episode_steps) { tr = !d; d = 1; }
        double er = ep_ret[i] + r;
        reward[i] = r; done[i] = (uint8_t)d; truncated[i] = (uint8_t)tr;
        if (natt_out) natt_out[i] = fl[5];
        if (y_pre_reset) for (int c = 0; c < n; ++c) y_pre_reset[c * N + i] = yy[c];
        float* oo = obs + i * no;
        float* to = terminal_obs ? terminal_obs + i * no : 0;
        if (p->dof == 1) { ob[0] = ob[1]; ob[1] = ob[4]; }
        if (d) {
            if (to) memcpy(to, ob, sizeof(float) * no);
            if (fin_ret) fin_ret[i] = er;
            if (fin_len) fin_len[i] = kk;
            if (fin_landed) fin_landed[i] = (uint8_t)fl[2];
            float ic[NMAX];
            episode[i] += 1;
            if (ic_inject) memcpy(ic, ic_inject + i * n, sizeof(float) * n);
            else or_sample_ic(p, seed, env_offset + i, episode[i], ic);
            or_reset_from_ic(p, ic, yy, &v0[i], ob);
            if (p->dof == 1) { ob[0] = ob[1]; ob[1] = ob[4]; }
            kk = 0; er = 0.0;
        }
        memcpy(oo, ob, sizeof(float) * no);
        for (int c = 0; c < n; ++c) y[c * N + i] = yy[c];
        k[i] = kk; ep_ret[i] = er;
    }
}

void or_vec_reset(const OrParams* p, int64_t N, int64_t env_offset, uint64_t seed,
                  double* y, int32_t* k, int32_t* episode, float* v0, double* ep_ret,
                  const float* ic_inject, float* obs) {
    int n = (p->dof == 6) ? 14 : 7;
    int no = (p->dof == 1) ? 2 : ((p->dof == 3 && p->obs_mode == 1) ? 4 : n);
    for (int64_t i = 0; i < N; ++i) {
        float ic[NMAX], ob[NMAX]; double yy[NMAX];
        if (ic_inject) memcpy(ic, ic_inject + i * n, sizeof(float) * n);
        else or_sample_ic(p, seed, env_offset + i, episode[i], ic);
        or_reset_from_ic(p, ic, yy, &v0[i], ob);
        if (p->dof == 1) { ob[0] = ob[1]; ob[1] = ob[4]; }
        memcpy(obs + i * no, ob, sizeof(float) * no);
        for (int c = 0; c < n; ++c) y[c * N + i] = yy[c];
        k[i] = 0; ep_ret[i] = 0.0;
    }
}

/* Bare simulator step (KAT-1: Simulator6DOF / Simulator3DOF with an arbitrary f64 state and f32 u). */
int or_sim_step(const OrParams* p, double* y, int k, const float* u, int* nfev, int* natt) {
    return (p->dof == 6) ? sim6_step(p, y, k, u, nfev, natt) : sim3_step(p, y, k, u, nfev, natt);
}

/* test_vtarg.py:47-70 (simplified target, f64 throughout) */
void or_vtarg_simple(const double* r, const double* v, double v0, double* vt, double* tgo) {
    double nr = f64_norm(r, 2), nv = f64_norm(v, 2);
    double t_go = nr / nv;
    for (int i = 0; i < 2; ++i) vt[i] = ((-v0) * (r[i] / nr)) * (1 - exp(-t_go / 20));
    *tgo = t_go;
}

float or_f32_norm(const float* x, int n) { return f32_norm(x, n); }
float or_np_sinf(float x) { return np_sinf(x); }
float or_np_cosf(float x) { return np_cosf(x); }
int   or_sizeof_params(void) { return (int)sizeof(OrParams); }
