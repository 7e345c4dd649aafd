// Development probe: DFMA / LDS / MUFU.RCP64H dependent-issue latency and DFMA issue interval of ONE warp on B200.
// nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/_bin/dfma_latency scripts/dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double* out, long long* cyc, double a, double b) {
    double x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lds_chain(double* out, long long* cyc) {
    __shared__ int idx[256];
    idx[threadIdx.x] = (threadIdx.x + 32) & 255; idx[threadIdx.x + 32] = (threadIdx.x + 64) & 255;
    for (int i = threadIdx.x; i < 256; i += 32) idx[i] = (i + 32) & 255;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 1024; ++it) p = idx[p];
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void rcp_chain(double* out, long long* cyc) {
    double x = 1.5 + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 1024; ++it) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 4096); cudaMalloc(&c, 8);
    long long h;
#define RUN(K, N, label) K<<<1, 32>>>(d, c, 1.0000001, 1e-9); K<<<1, 32>>>(d, c, 1.0000001, 1e-9); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
    printf("%-28s %.2f cycles per DFMA issue (one warp)\n", label, (double)h / (1024.0 * N));
    RUN(chain<1>, 1, "DFMA dependent chain"); RUN(chain<2>, 2, "DFMA ILP 2"); RUN(chain<4>, 4, "DFMA ILP 4"); RUN(chain<8>, 8, "DFMA ILP 8");
    lds_chain<<<1, 32>>>(d, c); lds_chain<<<1, 32>>>(d, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %.2f cycles\n", "LDS dependent (pointer chase)", (double)h / 1024.0);
    rcp_chain<<<1, 32>>>(d, c); rcp_chain<<<1, 32>>>(d, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %.2f cycles\n", "MUFU.RCP64H dependent", (double)h / 1024.0);
    return 0;
}
