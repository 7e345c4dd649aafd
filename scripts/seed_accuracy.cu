// Development probe: relative error of the hardware reciprocal / rsqrt seeds (rcp.approx.ftz.f64, rsqrt.approx.ftz.f64) and of the
// refinement schemes built on them.  nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/_bin/seed_accuracy scripts/seed_accuracy.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(int n, double* out) {
    double m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned long long s = 0x9E3779B97F4A7C15ull * (i + 1);
        s ^= s >> 31; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 29;
        const double u = (double)(s >> 11) * 0x1.0p-53;
        const double x = exp2(40.0 * u - 20.0) * (1.0 + u);
        double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        const double ex = 1.0 / x;
        m[0] = fmax(m[0], fabs(r - ex) / ex);
        double e = fma(-x, r, 1.0), r1 = fma(r, e, r);
        m[1] = fmax(m[1], fabs(r1 - ex) / ex);
        double e2 = fma(-x, r1, 1.0), r2 = fma(r1, e2, r1);
        m[2] = fmax(m[2], fabs(r2 - ex) / ex);
        double p = fma(e, e, e), rc = fma(r, p, r);
        m[3] = fmax(m[3], fabs(rc - ex) / ex);
        double q; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(x));
        const double eq = 1.0 / sqrt(x);
        m[4] = fmax(m[4], fabs(q - eq) / eq);
        double f = fma(-x * q, q, 1.0), q1 = fma(0.5 * q, f, q);
        m[5] = fmax(m[5], fabs(q1 - eq) / eq);
        double f2 = fma(-x * q1, q1, 1.0), q2 = fma(0.5 * q1, f2, q1);
        m[6] = fmax(m[6], fabs(q2 - eq) / eq);
        // x^(-1/5): f32 seed + Newton
        const float xf = fminf(fmaxf((float)x, 1e-30f), 1e30f);
        double y = (double)__powf(xf, -0.2f);
        const double ep = pow(x, -0.2);
        double y2 = y * y, y5 = (y2 * y2) * y; y = y * fma(-0.2 * x, y5, 1.2);
        m[7] = fmax(m[7], fabs(y - ep) / ep);
    }
    for (int j = 0; j < 8; ++j) atomicMax((unsigned long long*)&out[j], (unsigned long long)__double_as_longlong(m[j]));
}
int main() {
    double* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
    k<<<592, 256>>>(1 << 26, d);
    double h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    const char* nm[8] = {"rcp seed", "rcp + 1 Newton", "rcp + 2 Newton", "rcp + 1 cubic", "rsqrt seed", "rsqrt + 1 Newton", "rsqrt + 2 Newton", "pow(-0.2): powf seed + 1 Newton"};
    for (int j = 0; j < 8; ++j) printf("%-34s max rel err %.3e  (2^%.1f)\n", nm[j], h[j], log2(h[j]));
    return 0;
}
