// pnm_host.cu -- entry points for callers whose particles live in HOST memory (what a pynmap user
// has after snapshot.read, pynmap/pynmap.py:182-232).
//
//   pnm_accumulate_host  streams host arrays through a small device staging ring in chunks; the
//                        H2D copy of chunk c+1 (copy engine) overlaps the pnm_project_bin kernel of
//                        chunk c (SMs).  All chunks add into the caller's device accumulators.
//   pnm_project_host     grid + accumulate + finalise (+ optional smoothing) + D2H of the results:
//                        snapshot.rotate -> velocity_moments -> create_losvds -> convolve_spatial in
//                        one call, host buffers in and out.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>

#include <cuda_runtime.h>

#include "../../include/pnm_b200.h"

namespace {

constexpr int kSlots = 3;
constexpr int64_t kChunk = 1 << 22;   // particles per chunk: 7 x 16 MiB (float32) per slot

struct HostRing {
    int device = -1;
    size_t slot_bytes = 0;            // bytes per array per slot
    void* buf[kSlots][7] = {};
    cudaStream_t stream[kSlots] = {};
    void* arena = nullptr;            // pnm_project_host: accumulators + outputs + smoothing scratch
    size_t arena_bytes = 0;
};

HostRing g_ring;
std::mutex g_mutex;
thread_local char g_host_err[512] = "";

// publishes g_host_err through pnm_last_error() (defined in pnm_api.cu)
extern "C" void pnm_set_last_error_(const char* msg);

int host_fail(int code) { pnm_set_last_error_(g_host_err); return code; }

void release_locked() {
    if (g_ring.device < 0) return;
    cudaSetDevice(g_ring.device);
    for (int s = 0; s < kSlots; ++s) {
        for (int a = 0; a < 7; ++a) { cudaFree(g_ring.buf[s][a]); g_ring.buf[s][a] = nullptr; }
        if (g_ring.stream[s]) { cudaStreamDestroy(g_ring.stream[s]); g_ring.stream[s] = nullptr; }
    }
    cudaFree(g_ring.arena);
    g_ring = HostRing();
}

#define HOST_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            snprintf(g_host_err, sizeof(g_host_err), "%s failed: %s (%s:%d)", #call,      \
                     cudaGetErrorString(e_), __FILE__, __LINE__);                         \
            return host_fail(PNM_ECUDA);                                                  \
        }                                                                                 \
    } while (0)

int ensure_ring(int device, size_t slot_bytes, size_t arena_bytes) {
    if (g_ring.device != device || g_ring.slot_bytes < slot_bytes) {
        release_locked();
        HOST_CUDA(cudaSetDevice(device));
        g_ring.device = device;
        for (int s = 0; s < kSlots; ++s) {
            HOST_CUDA(cudaStreamCreateWithFlags(&g_ring.stream[s], cudaStreamNonBlocking));
            for (int a = 0; a < 7; ++a) HOST_CUDA(cudaMalloc(&g_ring.buf[s][a], slot_bytes));
        }
        g_ring.slot_bytes = slot_bytes;
    }
    if (g_ring.arena_bytes < arena_bytes) {
        cudaFree(g_ring.arena);
        g_ring.arena = nullptr; g_ring.arena_bytes = 0;
        HOST_CUDA(cudaMalloc(&g_ring.arena, arena_bytes));
        g_ring.arena_bytes = arena_bytes;
    }
    return PNM_OK;
}

// Enqueue copy + bin of every chunk on the ring's streams; the caller's stream `after` (may be
// null = legacy stream) is made to wait for all of them.  `before`: work the chunks must wait for.
int accumulate_locked(const pnm_grid* grid, int device, int64_t n, int dtype, const void* const src[7],
                      const float* mats_host, int nchain, int sums, int acc_mode, const double* scales3_host,
                      void* acc_maps, void* acc_cube, cudaStream_t user) {
    const size_t esz = dtype == PNM_F32 ? 4 : 8;
    int rc = ensure_ring(device, (size_t)kChunk * 8, 0);
    if (rc != PNM_OK) return rc;
    cudaEvent_t ready = nullptr, done = nullptr;
    HOST_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    HOST_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ready, user);            // accumulators zeroed / earlier work on `user`
    for (int s = 0; s < kSlots && e == cudaSuccess; ++s) e = cudaStreamWaitEvent(g_ring.stream[s], ready, 0);
    const int narr = src[6] ? 7 : 6;
    int64_t off = 0;
    for (int c = 0; off < n && e == cudaSuccess && rc == PNM_OK; ++c, off += kChunk) {
        const int slot = c % kSlots;
        const int64_t m = std::min<int64_t>(kChunk, n - off);
        cudaStream_t st = g_ring.stream[slot];
        for (int a = 0; a < narr && e == cudaSuccess; ++a)
            e = cudaMemcpyAsync(g_ring.buf[slot][a], static_cast<const char*>(src[a]) + (size_t)off * esz,
                                (size_t)m * esz, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) break;
        void** b = g_ring.buf[slot];
        rc = pnm_project_bin(grid, m, dtype, b[0], b[1], b[2], b[3], b[4], b[5], src[6] ? b[6] : nullptr, nullptr,
                             PNM_MASK_NONE, mats_host, nchain, 1, sums, acc_mode, scales3_host, acc_maps, acc_cube, st);
    }
    for (int s = 0; s < kSlots && e == cudaSuccess; ++s) {    // join the ring into the user's stream
        e = cudaEventRecord(done, g_ring.stream[s]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(user, done, 0);
    }
    // the host arrays may be reused by the caller as soon as we return: wait for the copies
    for (int s = 0; s < kSlots; ++s) { cudaError_t t = cudaStreamSynchronize(g_ring.stream[s]); if (e == cudaSuccess) e = t; }
    cudaEventDestroy(ready);
    cudaEventDestroy(done);
    if (rc != PNM_OK) return rc;                              // message set by pnm_project_bin
    if (e != cudaSuccess) {
        snprintf(g_host_err, sizeof(g_host_err), "pnm_accumulate_host: %s", cudaGetErrorString(e));
        return host_fail(PNM_ECUDA);
    }
    return PNM_OK;
}

bool bad_arrays(int64_t n, int dtype, const void* const src[7], const float* mats) {
    if (n < 0 || (dtype != PNM_F32 && dtype != PNM_F64) || !mats) return true;
    for (int a = 0; a < 6; ++a) if (n > 0 && !src[a]) return true;
    return false;
}

}  // namespace

extern "C" {

int pnm_host_release(void) {
    std::lock_guard<std::mutex> lock(g_mutex);
    release_locked();
    return PNM_OK;
}

int pnm_accumulate_host(const pnm_grid* g, int device, int64_t n, int dtype,
                        const void* x, const void* y, const void* z,
                        const void* vx, const void* vy, const void* vz, const void* w,
                        const float* mats_host, int nchain, int sums, int acc_mode, const double* scales3_host,
                        void* acc_maps, void* acc_cube, void* stream) {
    std::lock_guard<std::mutex> lock(g_mutex);
    const void* src[7] = {x, y, z, vx, vy, vz, w};
    if (!g || bad_arrays(n, dtype, src, mats_host)) {
        snprintf(g_host_err, sizeof(g_host_err), "pnm_accumulate_host: bad argument");
        return host_fail(PNM_EINVAL);
    }
    HOST_CUDA(cudaSetDevice(device));
    return accumulate_locked(g, device, n, dtype, src, mats_host, nchain, sums, acc_mode, scales3_host, acc_maps,
                             acc_cube, reinterpret_cast<cudaStream_t>(stream));
}

int pnm_project_host(int device, int64_t n, int dtype,
                     const void* x, const void* y, const void* z,
                     const void* vx, const void* vy, const void* vz, const void* w,
                     const float* mats_host, int nchain,
                     int32_t nx, const double* ex, int32_t ny, const double* ey, int32_t nv, const double* ev,
                     int acc_mode, const double* scales3_host,
                     const double* taps0_host, int32_t ntaps0, const double* taps1_host, int32_t ntaps1,
                     double* M_host, double* V_host, double* S_host, double* cube_host) {
    std::lock_guard<std::mutex> lock(g_mutex);
    g_host_err[0] = 0;
    const void* src[7] = {x, y, z, vx, vy, vz, w};
    if (bad_arrays(n, dtype, src, mats_host)) {
        snprintf(g_host_err, sizeof(g_host_err), "pnm_project_host: bad argument");
        return host_fail(PNM_EINVAL);
    }
    const bool want_maps = M_host || V_host || S_host;
    const bool want_cube = cube_host != nullptr;
    const bool smooth = want_cube && ntaps0 > 0 && ntaps1 > 0 && taps0_host && taps1_host;
    if (!want_maps && !want_cube) {
        snprintf(g_host_err, sizeof(g_host_err), "pnm_project_host: no output requested");
        return host_fail(PNM_EINVAL);
    }
    if (want_cube && (nv < 1 || !ev)) {
        snprintf(g_host_err, sizeof(g_host_err), "pnm_project_host: cube requested without V axis");
        return host_fail(PNM_EINVAL);
    }
    HOST_CUDA(cudaSetDevice(device));

    pnm_grid* grid = nullptr;
    int rc = pnm_grid_create(nx, ex, ny, ey, want_cube ? nv : 0, want_cube ? ev : nullptr, &grid);
    if (rc != PNM_OK) return rc;   // message already set by pnm_grid_create

    const int64_t npix = (int64_t)nx * ny;
    const int64_t cube_elems = pnm_cube_acc_elems(grid);
    const int64_t maps_elems = (want_maps && want_cube) ? (int64_t)npix * PNM_MAP_SLOTS : pnm_maps_acc_elems(grid);   // one replica with the interleaved cube
    // maps + cube together: interleaved cube with the moments of the in-range particles (one L2 request each)
    const int64_t cacc_elems = (want_maps && want_cube) ? pnm_cube_moments_acc_elems(grid) : cube_elems;
    // arena: [maps acc | cube acc | M V S | cube out | smoothing out | smoothing scratch]
    const size_t arena_bytes = (size_t)(maps_elems + cacc_elems + 3 * npix + cube_elems + (smooth ? 2 * cube_elems : 0)) * 8;
    rc = ensure_ring(device, (size_t)kChunk * 8, arena_bytes);
    if (rc != PNM_OK) { pnm_grid_destroy(grid); return rc; }
    char* arena = static_cast<char*>(g_ring.arena);
    void* acc_maps = arena;
    void* acc_cube = arena + (size_t)maps_elems * 8;
    double* out_maps = reinterpret_cast<double*>(arena + (size_t)(maps_elems + cacc_elems) * 8);
    double* cube_out = out_maps + 3 * npix;
    double* smooth_out = cube_out + cube_elems;
    double* smooth_work = smooth_out + cube_elems;

    int sums = 0;
    if (want_maps) sums |= PNM_SUM_W | PNM_SUM_WD | PNM_SUM_WD2;
    if (want_cube) sums |= PNM_SUM_CUBE;
    if (want_maps && want_cube) sums |= PNM_W_FROM_CUBE | PNM_CUBE_MOMENTS;

    cudaStream_t s0 = g_ring.stream[0];
    cudaError_t e = cudaMemsetAsync(arena, 0, (size_t)(maps_elems + cacc_elems) * 8, s0);
    if (e == cudaSuccess)
        rc = accumulate_locked(grid, device, n, dtype, src, mats_host, nchain, sums, acc_mode, scales3_host,
                               want_maps ? acc_maps : nullptr, want_cube ? acc_cube : nullptr, s0);
    if (e == cudaSuccess && rc == PNM_OK && want_maps) {
        rc = pnm_finalize_vmaps(grid, acc_maps, want_cube ? acc_cube : nullptr, sums, acc_mode, scales3_host,
                                out_maps, out_maps + npix, out_maps + 2 * npix, nullptr, nullptr, s0);
        double* hosts[3] = {M_host, V_host, S_host};
        for (int k = 0; k < 3 && e == cudaSuccess && rc == PNM_OK; ++k)
            if (hosts[k]) e = cudaMemcpyAsync(hosts[k], out_maps + k * npix, (size_t)npix * 8, cudaMemcpyDeviceToHost, s0);
    }
    if (e == cudaSuccess && rc == PNM_OK && want_cube) {
        rc = pnm_finalize_cube(grid, acc_cube, sums, acc_mode, acc_mode == PNM_ACC_FIXED ? scales3_host[0] : 1.0, cube_out, s0);
        const double* result = cube_out;
        if (rc == PNM_OK && smooth) {
            rc = pnm_smooth_cube(result, smooth_out, smooth_work, nx, ny, nv, taps0_host, ntaps0, taps1_host, ntaps1, s0);
            result = smooth_out;
        }
        if (rc == PNM_OK) e = cudaMemcpyAsync(cube_host, result, (size_t)cube_elems * 8, cudaMemcpyDeviceToHost, s0);
    }
    cudaError_t e2 = cudaStreamSynchronize(s0);
    pnm_grid_destroy(grid);
    if (rc != PNM_OK) return rc;
    if (e == cudaSuccess) e = e2;
    if (e != cudaSuccess) {
        snprintf(g_host_err, sizeof(g_host_err), "pnm_project_host: %s", cudaGetErrorString(e));
        return host_fail(PNM_ECUDA);
    }
    return PNM_OK;
}

}  // extern "C"
