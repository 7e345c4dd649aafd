// Development aid: what an event pair around ONE launch measures besides the kernel -- empty kernels with small / 2 KB parameter
// blocks, with and without the programmatic-stream-serialization attribute, after a memory-bound "flush" kernel (as bench.py does).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/launch_overhead scripts/launch_overhead.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <algorithm>
#include <vector>
struct Big { double v[240]; };
struct Small { double v[4]; };
__global__ void k_small(Small p, double* out) { if (out && p.v[0] == 123.0) out[0] = p.v[1]; }
__global__ void k_big(const __grid_constant__ Big p, double* out) { if (out && p.v[0] == 123.0) out[0] = p.v[1]; }
__global__ void k_big_pdl(const __grid_constant__ Big p, double* out) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (out && p.v[0] == 123.0) out[0] = p.v[1];
}
__global__ void flush(const float4* __restrict__ a, size_t n, float* sink) {
    float s = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { float4 v = a[i]; s += v.x + v.y + v.z + v.w; }
    if (s == 12345.f) *sink = s;
}
template <class F> static void run(const char* name, F launch, float4* buf, size_t n, float* sink, bool do_flush) {
    const int R = 200;
    std::vector<cudaEvent_t> e0(R), e1(R);
    for (int i = 0; i < R; ++i) { cudaEventCreate(&e0[i]); cudaEventCreate(&e1[i]); }
    for (int rep = 0; rep < 2; ++rep) {
        for (int i = 0; i < R; ++i) {
            if (do_flush) flush<<<148 * 8, 256>>>(buf, n, sink);
            cudaEventRecord(e0[i]); launch(); cudaEventRecord(e1[i]);
        }
        cudaDeviceSynchronize();
    }
    std::vector<float> t(R);
    for (int i = 0; i < R; ++i) cudaEventElapsedTime(&t[i], e0[i], e1[i]);
    std::sort(t.begin(), t.end());
    double m = 0; for (float x : t) m += x;
    printf("%-44s flush=%d  mean %.2f us  median %.2f us  min %.2f us\n", name, (int)do_flush, m / R * 1e3, t[R / 2] * 1e3, t[0] * 1e3);
}
int main() {
    size_t n = (256u << 20) / 16; float4* buf; float* sink; double* out;
    cudaMalloc(&buf, n * 16); cudaMemset(buf, 0, n * 16); cudaMalloc(&sink, 4); cudaMalloc(&out, 8);
    Big b = {}; Small s = {};
    for (int fl = 0; fl < 2; ++fl) {
        run("empty kernel, 32 B params, 1024 x 64", [&] { k_small<<<1024, 64>>>(s, out); }, buf, n, sink, fl);
        run("empty kernel, 1920 B params, 1024 x 64", [&] { k_big<<<1024, 64>>>(b, out); }, buf, n, sink, fl);
        run("empty kernel, 1920 B params, 31 KB smem", [&] { k_big<<<1024, 64, 31 * 1024>>>(b, out); }, buf, n, sink, fl);
        run("empty, 1920 B params, 31 KB smem, PDL attr", [&] {
            cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(1024); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = 31 * 1024; cfg.stream = 0;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1; cudaLaunchKernelEx(&cfg, k_big_pdl, b, out); }, buf, n, sink, fl);
    }
    cudaFuncSetAttribute(k_big, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    for (int fl = 0; fl < 2; ++fl) {
        run("empty, 1920 B params, 512 x 128, 62 KB smem", [&] { k_big<<<512, 128, 62 * 1024>>>(b, out); }, buf, n, sink, fl);
        run("empty, 1920 B params, 256 x 256, 110 KB smem", [&] { k_big<<<256, 256, 110 * 1024>>>(b, out); }, buf, n, sink, fl);
        run("empty, 1920 B params, 148 x 448, 219 KB smem", [&] { k_big<<<148, 448, 219 * 1024>>>(b, out); }, buf, n, sink, fl);
        run("empty, 1920 B params, 148 x 448, 0 smem", [&] { k_big<<<148, 448>>>(b, out); }, buf, n, sink, fl);
        run("empty, 32 B params, 1 x 32", [&] { k_small<<<1, 32>>>(s, out); }, buf, n, sink, fl);
        run("no kernel (two event records)", [&] {}, buf, n, sink, fl);
    }
    return 0;
}
