// host_pipe.cu -- the host-buffer pipeline of the C ABI (rlr_pipe_*): one call per environment step that
//   copies the step's actions from pinned host memory to the device       (copy stream "h2d"),
//   enqueues the step kernel on the caller's stream                        (rlr_step),
//   copies the packed obs | reward | done result back to pinned host memory (copy stream "d2h"),
// with `depth` slots of caller-owned buffers, so that the PCIe transfers of step t overlap the integration of step t + 1
// (the device-to-host copy of 61 bytes per 6DOF environment is ~2.5x longer than the kernel at 65 536 environments).
// Events order the three streams; nothing here synchronises the host except rlr_pipe_wait.
#include <cuda_runtime.h>
#include <stdio.h>

#include <vector>

#include "rl_rocket_b200.h"

struct RlrHostPipe {
    int device = 0, depth = 0;
    cudaStream_t h2d = nullptr, d2h = nullptr;
    std::vector<cudaEvent_t> actions_ready, kernel_done, result_ready;
    std::vector<char> used;
};

extern "C" int rlr_set_error_(int code, const char* msg);      // abi.cu

static int cu(cudaError_t e) {
    if (e == cudaSuccess) return RLR_OK;
    char buf[200];
    snprintf(buf, sizeof(buf), "CUDA error %d: %s", (int)e, cudaGetErrorString(e));
    return rlr_set_error_(RLR_E_CUDA, buf);
}
#define CU(x) do { int rc_ = cu(x); if (rc_) return rc_; } while (0)

extern "C" {

int rlr_pipe_destroy(RlrHostPipe* hp);

static int pipe_init(RlrHostPipe* hp, int depth) {
    hp->depth = depth;
    CU(cudaGetDevice(&hp->device));
    CU(cudaStreamCreateWithFlags(&hp->h2d, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&hp->d2h, cudaStreamNonBlocking));
    hp->actions_ready.assign(depth, nullptr); hp->kernel_done.assign(depth, nullptr); hp->result_ready.assign(depth, nullptr);
    hp->used.assign(depth, 0);
    for (int s = 0; s < depth; ++s) {
        CU(cudaEventCreateWithFlags(&hp->actions_ready[s], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&hp->kernel_done[s], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&hp->result_ready[s], cudaEventDisableTiming));
    }
    return RLR_OK;
}

int rlr_pipe_create(RlrHostPipe** pipe, int depth) {
    if (!pipe || depth < 1 || depth > 16) return rlr_set_error_(RLR_E_INVALID, "rlr_pipe_create: need a pipe pointer and 1 <= depth <= 16");
    RlrHostPipe* hp = new RlrHostPipe();
    const int rc = pipe_init(hp, depth);
    if (rc) { rlr_pipe_destroy(hp); return rc; }      // (the error text of the failing call stays in rlr_last_error)
    *pipe = hp;
    return RLR_OK;
}

int rlr_pipe_destroy(RlrHostPipe* hp) {
    if (!hp) return RLR_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(hp->device);
    for (size_t s = 0; s < hp->actions_ready.size(); ++s) {
        if (hp->actions_ready[s]) cudaEventDestroy(hp->actions_ready[s]);
        if (hp->kernel_done[s]) cudaEventDestroy(hp->kernel_done[s]);
        if (hp->result_ready[s]) cudaEventDestroy(hp->result_ready[s]);
    }
    if (hp->h2d) cudaStreamDestroy(hp->h2d);
    if (hp->d2h) cudaStreamDestroy(hp->d2h);
    cudaSetDevice(prev);
    delete hp;
    return RLR_OK;
}

int rlr_pipe_submit(RlrHostPipe* hp, int slot, const RlrParams* p, int precision, int math_mode, RlrState st, const void* h_actions,
                    void* d_actions, size_t action_bytes, int action_kind, RlrStepOut out, const void* d_pack, void* h_pack,
                    size_t pack_bytes, int64_t n, int64_t env_offset, uint64_t seed, void* compute_stream) {
    if (!hp || slot < 0 || slot >= hp->depth) return rlr_set_error_(RLR_E_INVALID, "rlr_pipe_submit: bad pipe or slot");
    if (!h_actions || !d_actions || !d_pack || !h_pack) return rlr_set_error_(RLR_E_INVALID, "rlr_pipe_submit: NULL buffer");
    cudaStream_t cs = (cudaStream_t)compute_stream;
    // the slot's device action buffer is free once the kernel that read it last has finished
    if (hp->used[slot]) CU(cudaStreamWaitEvent(hp->h2d, hp->kernel_done[slot], 0));
    CU(cudaMemcpyAsync(d_actions, h_actions, action_bytes, cudaMemcpyHostToDevice, hp->h2d));
    CU(cudaEventRecord(hp->actions_ready[slot], hp->h2d));
    CU(cudaStreamWaitEvent(cs, hp->actions_ready[slot], 0));
    // ... and its device result buffer once the previous device-to-host copy out of it has finished
    if (hp->used[slot]) CU(cudaStreamWaitEvent(cs, hp->result_ready[slot], 0));
    int rc = rlr_step(p, precision, math_mode, st, d_actions, action_kind, nullptr, out, n, env_offset, seed, compute_stream);
    if (rc) return rc;
    CU(cudaEventRecord(hp->kernel_done[slot], cs));
    CU(cudaStreamWaitEvent(hp->d2h, hp->kernel_done[slot], 0));
    CU(cudaMemcpyAsync(h_pack, d_pack, pack_bytes, cudaMemcpyDeviceToHost, hp->d2h));
    CU(cudaEventRecord(hp->result_ready[slot], hp->d2h));
    hp->used[slot] = 1;
    return RLR_OK;
}

int rlr_pipe_wait(RlrHostPipe* hp, int slot) {
    if (!hp || slot < 0 || slot >= hp->depth) return rlr_set_error_(RLR_E_INVALID, "rlr_pipe_wait: bad pipe or slot");
    if (!hp->used[slot]) return RLR_OK;
    CU(cudaEventSynchronize(hp->result_ready[slot]));
    return RLR_OK;
}

}  // extern "C"
