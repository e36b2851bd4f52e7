// peer.cu -- frame buffers in NVLink peer memory: the multi-GPU exchange of the render path without a
// collective kernel in the data path.
//
// Rays are independent, so N GPUs render disjoint row blocks of a frame with a replicated read-only model and every rank ends
// up with the whole composited frame (the role of the reference's dead eval-time `dist.all_gather(preds)`,
// nerf/trainer.py:1582-1601).  One process per GPU; every rank owns a full-frame buffer allocated here (cudaMalloc, exported
// with cudaIpcGetMemHandle, opened by the other ranks with lazy peer access over NVLink / NVSwitch).  A rank's pixels reach the
// peers in one of two ways, both asynchronous and stream ordered:
//   * narrow per-ray outputs (image, depth, weights_sum: 20 B per ray): the fused render kernel stores them straight into every
//     peer's buffer next to its own (sanerf_render_args_t::peer_*), i.e. the kernel's final stores ARE the all-gather;
//   * wide outputs (the 256-d SAM feature, 1 KB per ray): copy-engine pushes (sanerf_peer_push = one cudaMemcpyAsync per peer on
//     side streams) of a row group while the SMs already render the next group -- no SM is spent on communication.
// A frame ends with sanerf_peer_barrier: every rank signals "my pushes and stores are complete" into each peer's flag array and
// waits for the peers' signals (release / acquire at system scope).  torch.distributed is used only to exchange the IPC handles.
#include <cuda_runtime.h>

#include <cstring>

#include "common.cuh"

namespace sanerf {

struct PeerFlags {
    uint32_t* flags[SANERF_MAX_PEERS];   // flags[r] = rank r's flag array (SANERF_MAX_PEERS words), mapped into this process
};

// thread p: signal peer p (slot [rank] of ITS array), then wait for peer p's signal (slot [p] of MY array).
// Everything this rank enqueued before the barrier on the same stream has completed (kernel boundary); the system-scope fence +
// release store order those writes before the flag for the peer that acquires it.
__global__ void peer_barrier_kernel(PeerFlags f, uint32_t rank, uint32_t world, uint32_t epoch, long long timeout_cycles, uint32_t* status) {
    const uint32_t p = threadIdx.x;
    if (p >= world || p == rank) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.flags[p] + rank), "r"(epoch) : "memory");
    const uint32_t* mine = f.flags[rank] + p;
    const long long t0 = clock64();
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        if (clock64() - t0 > timeout_cycles) {   // a peer died: report instead of hanging the GPU
            if (status) atomicExch(status, 1u);
            break;
        }
        __nanosleep(200);
    }
}

}  // namespace sanerf

using namespace sanerf;

extern "C" {

int sanerf_peer_alloc(size_t bytes, void** ptr) {
    if (!ptr || bytes == 0) return SANERF_E_NULL;
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*ptr, 0, bytes);
    return (int)e;
}

int sanerf_peer_free(void* ptr) { return ptr ? (int)cudaFree(ptr) : 0; }

int sanerf_peer_export(const void* ptr, uint8_t handle[SANERF_PEER_HANDLE_BYTES]) {
    if (!ptr || !handle) return SANERF_E_NULL;
    static_assert(sizeof(cudaIpcMemHandle_t) == SANERF_PEER_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(ptr));
    if (e != cudaSuccess) return (int)e;
    memcpy(handle, &h, sizeof(h));
    return 0;
}

int sanerf_peer_open(const uint8_t handle[SANERF_PEER_HANDLE_BYTES], void** ptr) {
    if (!ptr || !handle) return SANERF_E_NULL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

int sanerf_peer_close(void* ptr) { return ptr ? (int)cudaIpcCloseMemHandle(ptr) : 0; }

int sanerf_peer_push(void* const* dst, const void* src, size_t bytes, uint32_t n_dst, const sanerf_stream_t* streams) {
    if (bytes == 0 || n_dst == 0) return 0;
    if (!dst || !src || !streams) return SANERF_E_NULL;
    for (uint32_t i = 0; i < n_dst; i++) {
        if (!dst[i]) return SANERF_E_NULL;
        cudaError_t e = cudaMemcpyAsync(dst[i], src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)streams[i]);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

int sanerf_peer_barrier(uint32_t* const* flags, uint32_t rank, uint32_t world, uint32_t epoch, float timeout_s, uint32_t* status,
                        sanerf_stream_t stream) {
    if (world <= 1) return 0;
    if (!flags || world > SANERF_MAX_PEERS || rank >= world) return SANERF_E_CONFIG;
    PeerFlags f;
    for (uint32_t r = 0; r < SANERF_MAX_PEERS; r++) f.flags[r] = r < world ? flags[r] : nullptr;
    for (uint32_t r = 0; r < world; r++)
        if (!f.flags[r]) return SANERF_E_NULL;
    int dev = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const long long cycles = (long long)((timeout_s > 0 ? timeout_s : 10.0f) * 1e3 * (double)(khz > 0 ? khz : 1500000));
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, rank, world, epoch, cycles, status);
    return check_launch();
}

}  // extern "C"
