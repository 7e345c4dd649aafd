// fma_peak.cu -- FP64 / FP32 FMA roofline calibration (MEASURED_PEAKS.json carries no ALU figure).
// 8 independent FMA chains per thread so that the pipe, not the dependency latency, is the limit.
#include <cuda_runtime.h>

#include "rl_rocket_b200.h"

namespace rlr_peak {

template <typename T>
__global__ void fma_chain_kernel(int iters, T seed, double* sink) {
    T a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const T m = (T)0.999999, c = (T)1e-7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
            a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
        }
    }
    T s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == (T)-12345.678) sink[0] = (double)s;   // never true; keeps the chains alive
}

int launch_fma_peak(int precision, int blocks, int threads, int iters, double* sink, double* flops_out, cudaStream_t stream) {
    if (precision == RLR_F64) fma_chain_kernel<double><<<blocks, threads, 0, stream>>>(iters, 1.0, sink);
    else fma_chain_kernel<float><<<blocks, threads, 0, stream>>>(iters, 1.0f, sink);
    if (flops_out) *flops_out = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;   // 64 FMAs per iteration per thread
    return (int)cudaGetLastError();
}

}  // namespace rlr_peak
