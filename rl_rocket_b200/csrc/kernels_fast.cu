// Fast-math instantiation: FMA contraction on (nvcc default), reciprocal-sqrt quaternion normalisation.
#define RLR_NS rlr_fast
#define RLR_EXACT 0
#include "rocket_kernels.cuh"

#if RLR_PHASE_TIMING
// tuning builds only: read (and clear) the per-phase cycle counters; out = [sum(16) | max(16) | count(16)]
extern "C" int rlr_debug_phases(unsigned long long* out48) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out48, rlr_fast::g_phase_sum, 16 * 8);
    cudaMemcpyFromSymbol(out48 + 16, rlr_fast::g_phase_max, 16 * 8);
    cudaMemcpyFromSymbol(out48 + 32, rlr_fast::g_phase_cnt, 16 * 8);
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(rlr_fast::g_phase_sum, z, 16 * 8);
    cudaMemcpyToSymbol(rlr_fast::g_phase_max, z, 16 * 8);
    cudaMemcpyToSymbol(rlr_fast::g_phase_cnt, z, 16 * 8);
    return (int)cudaGetLastError();
}
#endif

#if RLR_WARP_TRACE
// tuning builds only: the per-warp records of the last step launch; out = [8192][16] u64
extern "C" int rlr_debug_warp_trace(unsigned long long* out) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, rlr_fast::g_warp_trace, 8192 * 16 * 8);
}
#endif
