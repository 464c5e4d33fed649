// pnm_host.cu -- pnm_project_host: the end-to-end entry point for callers whose particles live in
// HOST memory (what a pynmap user has after snapshot.read, pynmap/pynmap.py:182-232).
//
// The arrays are cut into chunks; each chunk is copied H2D into one slot of a small staging ring
// and binned by pnm_project_bin on the same stream, so that the PCIe copy of chunk c+1 overlaps
// the kernel of chunk c (copy engine vs SMs).  All chunks add into one set of accumulators (the
// scatter is atomic), then the maps / cube are finalised and copied back.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>

#include <cuda_runtime.h>

#include "../../include/pnm_b200.h"

namespace {

constexpr int kSlots = 3;
constexpr int64_t kChunk = 1 << 22;   // particles per chunk: 7 x 16 MiB (f32) per slot

struct HostRing {
    int device = -1;
    size_t slot_bytes = 0;            // bytes per array per slot
    void* buf[kSlots][7] = {};
    cudaStream_t stream[kSlots] = {};
    void* acc = nullptr;              // accumulators + outputs arena
    size_t acc_bytes = 0;
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
    cudaFree(g_ring.acc);
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

int ensure_ring(int device, size_t slot_bytes, size_t acc_bytes) {
    if (g_ring.device != device || g_ring.slot_bytes < slot_bytes || g_ring.acc_bytes < acc_bytes) {
        release_locked();
        HOST_CUDA(cudaSetDevice(device));
        g_ring.device = device;
        for (int s = 0; s < kSlots; ++s) {
            HOST_CUDA(cudaStreamCreateWithFlags(&g_ring.stream[s], cudaStreamNonBlocking));
            for (int a = 0; a < 7; ++a) HOST_CUDA(cudaMalloc(&g_ring.buf[s][a], slot_bytes));
        }
        g_ring.slot_bytes = slot_bytes;
        HOST_CUDA(cudaMalloc(&g_ring.acc, acc_bytes));
        g_ring.acc_bytes = acc_bytes;
    }
    return PNM_OK;
}

}  // namespace

extern "C" {

int pnm_host_release(void) {
    std::lock_guard<std::mutex> lock(g_mutex);
    release_locked();
    return PNM_OK;
}

int pnm_project_host(int device, int64_t n, int dtype,
                     const void* x, const void* y, const void* z,
                     const void* vx, const void* vy, const void* vz, const void* w,
                     const float* mats_host, int nchain,
                     int32_t nx, const double* ex, int32_t ny, const double* ey, int32_t nv, const double* ev,
                     int acc_mode, const double* scales3_host,
                     double* M_host, double* V_host, double* S_host, double* cube_host) {
    std::lock_guard<std::mutex> lock(g_mutex);
    g_host_err[0] = 0;
    if (n < 0 || (dtype != PNM_F32 && dtype != PNM_F64) || !x || !y || !z || !vx || !vy || !vz || !mats_host) {
        snprintf(g_host_err, sizeof(g_host_err), "pnm_project_host: bad argument");
        return host_fail(PNM_EINVAL);
    }
    const bool want_maps = M_host || V_host || S_host;
    const bool want_cube = cube_host != nullptr;
    if (!want_maps && !want_cube) {
        snprintf(g_host_err, sizeof(g_host_err), "pnm_project_host: no output requested");
        return host_fail(PNM_EINVAL);
    }
    if (want_cube && (nv < 1 || !ev)) {
        snprintf(g_host_err, sizeof(g_host_err), "pnm_project_host: cube requested without V axis");
        return host_fail(PNM_EINVAL);
    }
    HOST_CUDA(cudaSetDevice(device));
    const size_t esz = dtype == PNM_F32 ? 4 : 8;
    const int64_t chunk = std::min<int64_t>(kChunk, std::max<int64_t>(n, 1));

    pnm_grid* grid = nullptr;
    int rc = pnm_grid_create(nx, ex, ny, ey, want_cube ? nv : 0, want_cube ? ev : nullptr, &grid);
    if (rc != PNM_OK) return rc;   // message already set by pnm_grid_create

    const int64_t maps_elems = pnm_maps_acc_elems(grid), cube_elems = pnm_cube_acc_elems(grid);
    const int64_t npix = (int64_t)nx * ny;
    // arena: [maps acc | cube acc | M V S]
    const size_t acc_bytes = (size_t)(maps_elems + cube_elems + 3 * npix) * 8;
    rc = ensure_ring(device, (size_t)kChunk * 8, acc_bytes);
    if (rc != PNM_OK) { pnm_grid_destroy(grid); return rc; }
    char* arena = static_cast<char*>(g_ring.acc);
    void* acc_maps = arena;
    void* acc_cube = arena + (size_t)maps_elems * 8;
    double* out_maps = reinterpret_cast<double*>(arena + (size_t)(maps_elems + cube_elems) * 8);

    int sums = 0;
    if (want_maps) sums |= PNM_SUM_W | PNM_SUM_WD | PNM_SUM_WD2;
    if (want_cube) sums |= PNM_SUM_CUBE;
    if (want_maps && want_cube) sums |= PNM_W_FROM_CUBE;

    cudaStream_t s0 = g_ring.stream[0];
    cudaError_t e = cudaMemsetAsync(arena, 0, (size_t)(maps_elems + cube_elems) * 8, s0);
    cudaEvent_t zeroed = nullptr, done[kSlots] = {};
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&zeroed, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(zeroed, s0);
    for (int s = 1; s < kSlots && e == cudaSuccess; ++s) e = cudaStreamWaitEvent(g_ring.stream[s], zeroed, 0);

    const void* src[7] = {x, y, z, vx, vy, vz, w};
    const int narr = w ? 7 : 6;
    int64_t off = 0;
    for (int c = 0; off < n && e == cudaSuccess && rc == PNM_OK; ++c, off += chunk) {
        const int slot = c % kSlots;
        const int64_t m = std::min(chunk, n - off);
        cudaStream_t st = g_ring.stream[slot];
        for (int a = 0; a < narr && e == cudaSuccess; ++a)
            e = cudaMemcpyAsync(g_ring.buf[slot][a], static_cast<const char*>(src[a]) + (size_t)off * esz,
                                (size_t)m * esz, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) break;
        void** b = g_ring.buf[slot];
        rc = pnm_project_bin(grid, m, dtype, b[0], b[1], b[2], b[3], b[4], b[5], w ? b[6] : nullptr, nullptr,
                             PNM_MASK_NONE, mats_host, nchain, 1, sums, acc_mode, scales3_host,
                             want_maps ? acc_maps : nullptr, want_cube ? acc_cube : nullptr, st);
        if (rc != PNM_OK) snprintf(g_host_err, sizeof(g_host_err), "%s", pnm_last_error());
    }
    // join the slot streams into s0
    for (int s = 1; s < kSlots && e == cudaSuccess && rc == PNM_OK; ++s) {
        e = cudaEventCreateWithFlags(&done[s], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(done[s], g_ring.stream[s]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s0, done[s], 0);
    }
    if (e == cudaSuccess && rc == PNM_OK && want_maps) {
        rc = pnm_finalize_vmaps(grid, acc_maps, want_cube ? acc_cube : nullptr, sums, acc_mode, scales3_host,
                                out_maps, out_maps + npix, out_maps + 2 * npix, nullptr, nullptr, s0);
        if (rc != PNM_OK) snprintf(g_host_err, sizeof(g_host_err), "%s", pnm_last_error());
        double* hosts[3] = {M_host, V_host, S_host};
        for (int k = 0; k < 3 && e == cudaSuccess && rc == PNM_OK; ++k)
            if (hosts[k]) e = cudaMemcpyAsync(hosts[k], out_maps + k * npix, (size_t)npix * 8, cudaMemcpyDeviceToHost, s0);
    }
    if (e == cudaSuccess && rc == PNM_OK && want_cube) {
        rc = pnm_finalize_cube(grid, acc_cube, acc_mode, acc_mode == PNM_ACC_FIXED ? scales3_host[0] : 1.0,
                               static_cast<double*>(acc_cube), s0);
        if (rc != PNM_OK) snprintf(g_host_err, sizeof(g_host_err), "%s", pnm_last_error());
        if (rc == PNM_OK) e = cudaMemcpyAsync(cube_host, acc_cube, (size_t)cube_elems * 8, cudaMemcpyDeviceToHost, s0);
    }
    cudaError_t e2 = cudaSuccess;
    for (int s = 0; s < kSlots; ++s) { cudaError_t t = cudaStreamSynchronize(g_ring.stream[s]); if (e2 == cudaSuccess) e2 = t; }
    if (zeroed) cudaEventDestroy(zeroed);
    for (int s = 1; s < kSlots; ++s) if (done[s]) cudaEventDestroy(done[s]);
    pnm_grid_destroy(grid);
    if (e == cudaSuccess) e = e2;
    if (e != cudaSuccess) {
        snprintf(g_host_err, sizeof(g_host_err), "pnm_project_host: %s", cudaGetErrorString(e));
        return host_fail(PNM_ECUDA);
    }
    if (rc != PNM_OK) return host_fail(rc);
    return PNM_OK;
}

}  // extern "C"
