// pnm_api.cu -- extern "C" surface of libpnm_b200.so (see include/pnm_b200.h for the contract).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pnm_bin.cuh"
#include "pnm_post.cuh"
#include "pnm_fit.cuh"
#include "pnm_profile.cuh"
#include "pnm_cube.cuh"
#include "pnm_plan.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define PNM_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail(PNM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                        __FILE__, __LINE__);                                                 \
    } while (0)

// development knobs (microbenchmarks only; defaults are the tuned values)
inline int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline int grid_for(int64_t work_items, int threads, int ctas_per_sm) {
    int64_t blocks = (work_items + threads - 1) / threads;
    const int64_t cap = (int64_t)pnm::kNumSMs * ctas_per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// smallest T >= e (lower edges) / the exclusive upper bound equivalent to "<= e" (last edge)
float thr_lower_f32(double e) {
    float f = (float)e;
    if ((double)f < e) f = nextafterf(f, INFINITY);
    return f;
}
// Guard band of the uniform bin guess g = (x - e0) * n / (en - e0), in units of bins: if the
// fractional part of the COMPUTED g is further than this from 0 and 1, floor(g) is the true bin.
// Two contributions: (1) how far the given edges deviate from the affine grid (np.linspace edges
// are fl(fl(i*step)+start), a few ulps off), (2) the rounding of x0, inv, the subtraction and the
// product in the particle dtype (unit roundoff u): |dg| <= 4u * (|x - x0| + |x0|) * inv.
double guard_band(const double* e, int n, double u) {
    if (n < 1 || !e) return 1.0;
    const double inv = n / (e[n] - e[0]);
    double dev = 0.0, amax = 0.0;
    for (int i = 0; i <= n; ++i) {
        dev = std::max(dev, std::fabs((e[i] - e[0]) * inv - i));
        amax = std::max(amax, std::fabs(e[i]));
    }
    const double arith = 4.0 * u * ((e[n] - e[0]) + 2.0 * amax) * inv;
    const double eps = 2.0 * (dev + arith) + 1e-9;       // x2 safety
    return (eps < 0.25 && std::isfinite(eps)) ? eps : 1.0;   // 1.0: always decide on the table
}

template <typename T> T thr_last(double e);
template <> float thr_last<float>(double e) {
    float f = (float)e;
    if ((double)f > e) f = nextafterf(f, -INFINITY);   // largest float <= e
    return nextafterf(f, INFINITY);
}
template <> double thr_last<double>(double e) { return nextafter(e, INFINITY); }

}  // namespace

struct pnm_grid {
    int device;
    int nx, ny, nv;
    int replicas;       // copies of the maps accumulator (power of two), see pnm_bin.cuh
    // per axis: thresholds in both particle dtypes, laid out [x | y | v]
    float* thr32;
    double* thr64;
    double x0, inv_dx, y0, inv_dy, v0, inv_dv;
    double eps32[3], eps64[3];   // guard band (in bins) of the uniform guess for float32 / float64 data
};

extern "C" {

// internal: lets pnm_host.cu publish its message through pnm_last_error()
void pnm_set_last_error_(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg ? msg : ""); }

int pnm_version(void) { return PNM_VERSION; }
const char* pnm_last_error(void) { return g_err; }

int pnm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

// ------------------------------------------------------------------------------------ rotate
int pnm_rotate(const void* x, const void* y, const void* z, void* ox, void* oy, void* oz,
               int64_t n, int dtype, const float* mat9_host, void* stream) {
    if (n < 0 || !mat9_host) return fail(PNM_EINVAL, "pnm_rotate: bad n or matrix");
    if (n == 0) return PNM_OK;
    if (!x || !y || !z || !ox || !oy || !oz) return fail(PNM_EINVAL, "pnm_rotate: null array");
    pnm::Mat9 m;
    memcpy(m.m, mat9_host, sizeof(m.m));
    const int grid = grid_for(n, 256, 16);
    if (dtype == PNM_F32)
        pnm::k_rotate<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)x, (const float*)y, (const float*)z,
                                                                  (float*)ox, (float*)oy, (float*)oz, n, m);
    else if (dtype == PNM_F64)
        pnm::k_rotate<double><<<grid, 256, 0, as_stream(stream)>>>((const double*)x, (const double*)y, (const double*)z,
                                                                   (double*)ox, (double*)oy, (double*)oz, n, m);
    else
        return fail(PNM_EDTYPE, "pnm_rotate: dtype %d", dtype);
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

// ------------------------------------------------------------------------------------ grid
int pnm_grid_create(int32_t nx, const double* ex, int32_t ny, const double* ey,
                    int32_t nv, const double* ev, pnm_grid** out) {
    if (!out) return fail(PNM_EINVAL, "pnm_grid_create: out is null");
    *out = nullptr;
    if (nx < 1 || ny < 1 || !ex || !ey || nv < 0 || (nv > 0 && !ev))
        return fail(PNM_EINVAL, "pnm_grid_create: bad sizes nx=%d ny=%d nv=%d", nx, ny, nv);
    if (nx + 1 > PNM_MAX_EDGES || ny + 1 > PNM_MAX_EDGES || nv + 1 > PNM_MAX_EDGES)
        return fail(PNM_ELIMIT, "pnm_grid_create: more than %d edges on an axis", PNM_MAX_EDGES);
    const int32_t ns[3] = {nx, ny, nv};
    const double* es[3] = {ex, ey, ev};
    for (int a = 0; a < 3; ++a)
        for (int i = 0; i < ns[a]; ++i)
            if (!(es[a][i] < es[a][i + 1]))   // numpy raises on non-monotonic bins as well
                return fail(PNM_EINVAL, "pnm_grid_create: axis %d edges must increase strictly", a);

    const int total = (nx + 1) + (ny + 1) + (nv + 1);
    std::vector<float> t32(total);
    std::vector<double> t64(total);
    int o = 0;
    for (int a = 0; a < 3; ++a) {
        if (ns[a] == 0) { t32[o] = 0.f; t64[o] = 0.0; o += 1; continue; }
        for (int i = 0; i < ns[a]; ++i) { t32[o + i] = thr_lower_f32(es[a][i]); t64[o + i] = es[a][i]; }
        t32[o + ns[a]] = thr_last<float>(es[a][ns[a]]);
        t64[o + ns[a]] = thr_last<double>(es[a][ns[a]]);
        o += ns[a] + 1;
    }
    pnm_grid* g = new pnm_grid();
    g->nx = nx; g->ny = ny; g->nv = nv;
    g->x0 = ex[0]; g->inv_dx = nx / (ex[nx] - ex[0]);
    g->y0 = ey[0]; g->inv_dy = ny / (ey[ny] - ey[0]);
    g->v0 = nv ? ev[0] : 0.0; g->inv_dv = nv ? nv / (ev[nv] - ev[0]) : 0.0;
    for (int a = 0; a < 3; ++a) {
        g->eps32[a] = guard_band(es[a], ns[a], 0x1p-24);
        g->eps64[a] = guard_band(es[a], ns[a], 0x1p-53);
    }
    if ((int64_t)nx * ny * std::max(nv, 1) >= (1ll << 31))
        { delete g; return fail(PNM_ELIMIT, "pnm_grid_create: more than 2^31 voxels"); }
    // map replicas: as many as fit in 32 MiB of L2 (32 B per pixel), at most 64 (measured: 128^2 maps
    // 9.2 ms with 1 replica, 2.3 ms with 32, 2.0 ms with 64 for 10^8 disk+bulge particles)
    g->replicas = 1;
    const int max_rep = env_int("PNM_MAX_REPLICAS", 64);
    while (g->replicas < max_rep && (int64_t)nx * ny * PNM_MAP_SLOTS * 8 * (g->replicas * 2) <= ((int64_t)env_int("PNM_REPLICA_MIB", 32) << 20)) g->replicas *= 2;
    g->thr32 = nullptr; g->thr64 = nullptr;
    cudaError_t e = cudaGetDevice(&g->device);
    if (e == cudaSuccess) e = cudaMalloc(&g->thr32, total * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&g->thr64, total * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(g->thr32, t32.data(), total * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->thr64, t64.data(), total * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(g->thr32); cudaFree(g->thr64); delete g;
        return fail(PNM_ECUDA, "pnm_grid_create: %s", cudaGetErrorString(e));
    }
    *out = g;
    return PNM_OK;
}

int pnm_grid_destroy(pnm_grid* g) {
    if (!g) return PNM_OK;
    cudaFree(g->thr32);
    cudaFree(g->thr64);
    delete g;
    return PNM_OK;
}

int64_t pnm_maps_acc_elems(const pnm_grid* g) { return g ? (int64_t)g->nx * g->ny * PNM_MAP_SLOTS * g->replicas : 0; }
int pnm_grid_map_replicas(const pnm_grid* g) { return g ? g->replicas : 0; }
int64_t pnm_cube_acc_elems(const pnm_grid* g) { return g ? (int64_t)g->nx * g->ny * g->nv : 0; }
int64_t pnm_cube_moments_acc_elems(const pnm_grid* g) { return g ? (int64_t)g->nx * g->ny * ((g->nv + 1) / 2 + 1) * 4 : 0; }

}  // extern "C"

// ------------------------------------------------------------------------------------ binning
namespace {

template <typename T, bool FUSED>
int launch_bin(const pnm::BinParams& p, int acc_mode, cudaStream_t st) {
    bool aligned = true;
    const void* ptrs[7] = {p.x, p.y, p.z, p.vx, p.vy, p.vz, p.w};
    for (const void* q : ptrs) aligned = aligned && ((reinterpret_cast<uintptr_t>(q) & 31) == 0);   // 256-bit loads
    const int64_t items = aligned ? (p.n + pnm::Pack<T>::N - 1) / pnm::Pack<T>::N : p.n;
    const int grid = grid_for(items, pnm::kBinThreads, env_int("PNM_CTAS_PER_SM", pnm::kCtasPerSM));
    // SIMPLE: one view, one matrix, no mask -> matrix in registers, cross-particle run merging, hybrid scatter
    const bool simple = p.nviews == 1 && p.nchain == 1 && p.mask_mode == PNM_MASK_NONE;
    // shared memory: edge tables (+ the TMA staging area when the hybrid scatter can run)
    size_t smem = pnm::bin_table_bytes<T>(p.nx, p.ny, p.nv);
    if (aligned && simple && p.tma_slots > 0) smem += pnm::bin_stage_bytes(pnm::kBinThreads, p.tma_slots);
    if (smem > 227 * 1024) return fail(PNM_ELIMIT, "k_project_bin: %zu bytes of shared memory", smem);
    // hot configurations get the set of sums as a compile-time constant (fused path only)
    constexpr int kFull = PNM_SUM_W | PNM_SUM_WD | PNM_SUM_WD2 | PNM_SUM_CUBE | PNM_W_FROM_CUBE;
    constexpr int kMaps = PNM_SUM_W | PNM_SUM_WD | PNM_SUM_WD2;
    constexpr int kFullM = kFull | PNM_CUBE_MOMENTS;
#define PNM_LAUNCH(ACC, VL, SIMPLE, SUMS)                                                            \
    do {                                                                                             \
        auto kern = pnm::k_project_bin<T, FUSED, ACC, VL, SIMPLE, SUMS>;                             \
        if (smem > 48 * 1024)                                                                        \
            PNM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, 256, smem, st>>>(p);                                                            \
    } while (0)
#define PNM_LAUNCH2(ACC)                                                                             \
    do {                                                                                             \
        if (aligned && simple && FUSED && p.sums == kFullM) PNM_LAUNCH(ACC, true, true, kFullM);     \
        else if (aligned && simple && FUSED && p.sums == kFull) PNM_LAUNCH(ACC, true, true, kFull);  \
        else if (aligned && simple && FUSED && p.sums == kMaps) PNM_LAUNCH(ACC, true, true, kMaps);  \
        else if (aligned && simple && FUSED && p.sums == PNM_SUM_CUBE) PNM_LAUNCH(ACC, true, true, PNM_SUM_CUBE); \
        else if (aligned && simple) PNM_LAUNCH(ACC, true, true, -1);                                 \
        else if (aligned) PNM_LAUNCH(ACC, true, false, -1);                                          \
        else PNM_LAUNCH(ACC, false, false, -1);                                                      \
    } while (0)
    if (acc_mode == PNM_ACC_F64) PNM_LAUNCH2(PNM_ACC_F64); else PNM_LAUNCH2(PNM_ACC_FIXED);
#undef PNM_LAUNCH2
#undef PNM_LAUNCH
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int bin_common(const char* who, bool fused, const pnm_grid* g, int64_t n, int dtype,
               const void* x, const void* y, const void* z, const void* vx, const void* vy, const void* vz,
               const void* w, const uint8_t* mask, int mask_mode, const float* mats_host, int nchain, int nviews,
               int sums, int acc_mode, const double* scales3_host, void* acc_maps, void* acc_cube, void* stream) {
    if (!g) return fail(PNM_EINVAL, "%s: grid is null", who);
    if (n < 0) return fail(PNM_EINVAL, "%s: n < 0", who);
    if (dtype != PNM_F32 && dtype != PNM_F64) return fail(PNM_EDTYPE, "%s: dtype %d", who, dtype);
    if (acc_mode != PNM_ACC_F64 && acc_mode != PNM_ACC_FIXED) return fail(PNM_EINVAL, "%s: acc_mode %d", who, acc_mode);
    if (acc_mode == PNM_ACC_FIXED && !scales3_host) return fail(PNM_EINVAL, "%s: FIXED mode needs scales", who);
    const int map_sums = sums & (PNM_SUM_W | PNM_SUM_WD | PNM_SUM_WD2);
    if (!(sums & (PNM_SUM_W | PNM_SUM_WD | PNM_SUM_WD2 | PNM_SUM_CUBE))) return fail(PNM_EINVAL, "%s: no sums requested", who);
    if (map_sums && !acc_maps && !(sums & PNM_CUBE_MOMENTS)) return fail(PNM_EINVAL, "%s: acc_maps is null", who);
    if ((sums & PNM_SUM_CUBE) && (!acc_cube || g->nv < 1)) return fail(PNM_EINVAL, "%s: cube requested without cube accumulator / V axis", who);
    if ((sums & PNM_W_FROM_CUBE) && !((sums & PNM_SUM_CUBE) && (sums & PNM_SUM_W)))
        return fail(PNM_EINVAL, "%s: PNM_W_FROM_CUBE needs PNM_SUM_W|PNM_SUM_CUBE", who);
    if ((sums & PNM_CUBE_MOMENTS) &&
        (sums & (PNM_W_FROM_CUBE | PNM_SUM_WD | PNM_SUM_WD2)) != (PNM_W_FROM_CUBE | PNM_SUM_WD | PNM_SUM_WD2))
        return fail(PNM_EINVAL, "%s: PNM_CUBE_MOMENTS needs PNM_SUM_W|PNM_SUM_WD|PNM_SUM_WD2|PNM_SUM_CUBE|PNM_W_FROM_CUBE", who);
    if ((sums & PNM_CUBE_MOMENTS) && pnm_cube_moments_acc_elems(g) >= (1LL << 31))
        return fail(PNM_ELIMIT, "%s: interleaved cube of 2^31 elements or more", who);
    if ((sums & (PNM_VAL_D32 | PNM_VAL_W32)) && (fused || dtype != PNM_F64))
        return fail(PNM_EINVAL, "%s: PNM_VAL_D32 / PNM_VAL_W32 describe float32 values carried as float64 (pnm_bin_points, PNM_F64)", who);
    if (mask_mode != PNM_MASK_NONE && !mask) return fail(PNM_EINVAL, "%s: mask_mode without mask", who);
    if (mask_mode < 0 || mask_mode > 2) return fail(PNM_EINVAL, "%s: mask_mode %d", who, mask_mode);
    if (fused) {
        if (nchain < 1 || nchain > PNM_MAX_CHAIN) return fail(PNM_ELIMIT, "%s: nchain %d not in 1..%d", who, nchain, PNM_MAX_CHAIN);
        if (nviews < 1 || nviews > PNM_MAX_VIEWS) return fail(PNM_ELIMIT, "%s: nviews %d not in 1..%d", who, nviews, PNM_MAX_VIEWS);
        if (nchain * nviews > pnm::kMaxMats) return fail(PNM_ELIMIT, "%s: nchain*nviews > %d", who, pnm::kMaxMats);
        if (!mats_host) return fail(PNM_EINVAL, "%s: mats is null", who);
    }
    if (n == 0) return PNM_OK;
    if (!x || !y || !vx || (fused && (!z || !vy || !vz))) return fail(PNM_EINVAL, "%s: null particle array", who);

    pnm::BinParams p;
    memset(&p, 0, sizeof(p));
    p.n = n; p.x = x; p.y = y; p.z = z; p.vx = vx; p.vy = vy; p.vz = vz; p.w = w;
    p.mask = mask; p.mask_mode = mask_mode;
    p.nx = g->nx; p.ny = g->ny; p.nv = g->nv;
    const int ox = 0, oy = g->nx + 1, ov = oy + g->ny + 1;
    if (dtype == PNM_F32) { p.thr_x = g->thr32 + ox; p.thr_y = g->thr32 + oy; p.thr_v = g->thr32 + ov; }
    else { p.thr_x = g->thr64 + ox; p.thr_y = g->thr64 + oy; p.thr_v = g->thr64 + ov; }
    p.x0 = g->x0; p.inv_dx = g->inv_dx; p.y0 = g->y0; p.inv_dy = g->inv_dy; p.v0 = g->v0; p.inv_dv = g->inv_dv;
    const double* eps = dtype == PNM_F32 ? g->eps32 : g->eps64;
    p.eps_x = eps[0]; p.eps_y = eps[1]; p.eps_v = eps[2];
    p.sums = sums; p.nchain = fused ? nchain : 1; p.nviews = fused ? nviews : 1;
    p.replicas = (sums & PNM_CUBE_MOMENTS) ? 1 : g->replicas;   // (the interleaved cube does not use the maps accumulator)
    p.rep_shift = env_int("PNM_REP_SHIFT", 5);
    // hybrid scatter: 16-byte TMA bulk reductions need a 16-byte aligned maps accumulator
    p.tma_slots = (acc_maps && !(sums & PNM_CUBE_MOMENTS) && (reinterpret_cast<uintptr_t>(acc_maps) & 15) == 0 && (sums & PNM_SUM_WD) && (sums & PNM_SUM_WD2))
                      ? std::max(0, std::min(env_int("PNM_TMA_SLOTS", 0), pnm::kTmaSlots)) : 0;
    // (measured with lane-paired / lane-quad reductions, 1e8 particles: maps + cube 1.28 / 1.30 / 1.38 / 1.53 ms for
    //  0 / 2 / 3 / 4 of a tile's 8 updates through the TMA unit; maps only 0.86 / 0.88 ms for 0 / 2 -- a paired or quad
    //  reduction is one L2 request per pixel, like a bulk reduction, so the TMA unit has nothing left to save and the
    //  path is off unless PNM_TMA_SLOTS asks for it)
    for (int k = 0; k < 3; ++k) p.scale[k] = (acc_mode == PNM_ACC_FIXED) ? scales3_host[k] : 1.0;
    p.acc_maps = acc_maps; p.acc_cube = acc_cube;
    // per-view strides; with the interleaved cube the maps are a single replica and the cube holds 4-slot records
    p.maps_stride = (sums & PNM_CUBE_MOMENTS) ? (int64_t)g->nx * g->ny * PNM_MAP_SLOTS : pnm_maps_acc_elems(g);
    p.cube_stride = (sums & PNM_CUBE_MOMENTS) ? pnm_cube_moments_acc_elems(g) : pnm_cube_acc_elems(g);
    if (fused) memcpy(p.mats, mats_host, sizeof(float) * 9 * nchain * nviews);

    cudaStream_t st = as_stream(stream);
    if (dtype == PNM_F32) return fused ? launch_bin<float, true>(p, acc_mode, st) : launch_bin<float, false>(p, acc_mode, st);
    return fused ? launch_bin<double, true>(p, acc_mode, st) : launch_bin<double, false>(p, acc_mode, st);
}

}  // namespace

extern "C" {

int pnm_project_bin(const pnm_grid* g, int64_t n, int dtype, const void* x, const void* y, const void* z,
                    const void* vx, const void* vy, const void* vz, const void* w, const uint8_t* mask,
                    int mask_mode, const float* mats_host, int nchain, int nviews, int sums, int acc_mode,
                    const double* scales3_host, void* acc_maps, void* acc_cube, void* stream) {
    return bin_common("pnm_project_bin", true, g, n, dtype, x, y, z, vx, vy, vz, w, mask, mask_mode, mats_host,
                      nchain, nviews, sums, acc_mode, scales3_host, acc_maps, acc_cube, stream);
}

int pnm_bin_points(const pnm_grid* g, int64_t n, int dtype, const void* x, const void* y, const void* d,
                   const void* w, const uint8_t* mask, int mask_mode, int sums, int acc_mode,
                   const double* scales3_host, void* acc_maps, void* acc_cube, void* stream) {
    return bin_common("pnm_bin_points", false, g, n, dtype, x, y, nullptr, d, nullptr, nullptr, w, mask, mask_mode,
                      nullptr, 1, 1, sums, acc_mode, scales3_host, acc_maps, acc_cube, stream);
}

// ------------------------------------------------------------------------------------ finalise
int pnm_finalize_vmaps(const pnm_grid* g, const void* acc_maps, const void* acc_cube, int sums, int acc_mode,
                       const double* scales3_host, double* M, double* V, double* S, double* S1, double* S2,
                       void* stream) {
    if (!g || (!acc_maps && !(sums & PNM_CUBE_MOMENTS))) return fail(PNM_EINVAL, "pnm_finalize_vmaps: null grid / accumulator");
    const int wfc = (sums & PNM_CUBE_MOMENTS) ? 2 : (sums & PNM_W_FROM_CUBE) ? 1 : 0;
    if (wfc && !acc_cube) return fail(PNM_EINVAL, "pnm_finalize_vmaps: PNM_W_FROM_CUBE without cube");
    const int npix = g->nx * g->ny;
    const int grid = (npix + 31) / 32;
    const dim3 block(32, 8);
    const int replicas = (sums & (PNM_MAPS_FOLDED | PNM_CUBE_MOMENTS)) ? 1 : g->replicas;
    if (acc_mode == PNM_ACC_F64) {
        pnm::k_finalize_vmaps<PNM_ACC_F64><<<grid, block, 0, as_stream(stream)>>>(acc_maps, acc_cube, g->nx, g->ny, g->nv,
                                                                              replicas, wfc, 1.0, 1.0, 1.0, M, V, S, S1, S2);
    } else if (acc_mode == PNM_ACC_FIXED) {
        if (!scales3_host) return fail(PNM_EINVAL, "pnm_finalize_vmaps: FIXED mode needs scales");
        pnm::k_finalize_vmaps<PNM_ACC_FIXED><<<grid, block, 0, as_stream(stream)>>>(
            acc_maps, acc_cube, g->nx, g->ny, g->nv, replicas, wfc, 1.0 / scales3_host[0], 1.0 / scales3_host[1],
            1.0 / scales3_host[2], M, V, S, S1, S2);
    } else {
        return fail(PNM_EINVAL, "pnm_finalize_vmaps: acc_mode %d", acc_mode);
    }
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int pnm_finalize_map(const pnm_grid* g, const void* acc_maps, int acc_mode, const double* scales3_host,
                     double* W, double* D, double* DW, void* stream) {
    if (!g || !acc_maps) return fail(PNM_EINVAL, "pnm_finalize_map: null grid / accumulator");
    const int npix = g->nx * g->ny;
    const int grid = (npix + 255) / 256;
    if (acc_mode == PNM_ACC_F64) {
        pnm::k_finalize_map<PNM_ACC_F64><<<grid, 256, 0, as_stream(stream)>>>(acc_maps, npix, g->replicas, 1.0, 1.0, W, D, DW);
    } else if (acc_mode == PNM_ACC_FIXED) {
        if (!scales3_host) return fail(PNM_EINVAL, "pnm_finalize_map: FIXED mode needs scales");
        pnm::k_finalize_map<PNM_ACC_FIXED><<<grid, 256, 0, as_stream(stream)>>>(acc_maps, npix, g->replicas, 1.0 / scales3_host[0],
                                                                              1.0 / scales3_host[1], W, D, DW);
    } else {
        return fail(PNM_EINVAL, "pnm_finalize_map: acc_mode %d", acc_mode);
    }
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int pnm_finalize_cube(const pnm_grid* g, const void* acc_cube, int sums, int acc_mode, double scale_w, double* cube_out,
                      void* stream) {
    const int moments = (sums & PNM_CUBE_MOMENTS) ? 1 : 0;
    if (!g || !acc_cube || !cube_out) return fail(PNM_EINVAL, "pnm_finalize_cube: null argument");
    if (static_cast<const void*>(cube_out) == acc_cube) return fail(PNM_EINVAL, "pnm_finalize_cube: cannot run in place");
    if (g->nv < 1) return fail(PNM_EINVAL, "pnm_finalize_cube: grid has no V axis");
    const int npix = g->nx * g->ny;
    const dim3 grid((npix + 31) / 32, (g->nv + 31) / 32), block(32, 8);
    if (acc_mode == PNM_ACC_F64)
        pnm::k_cube_out<PNM_ACC_F64><<<grid, block, 0, as_stream(stream)>>>(acc_cube, cube_out, npix, g->nv, 1.0, moments);
    else if (acc_mode == PNM_ACC_FIXED)
        pnm::k_cube_out<PNM_ACC_FIXED><<<grid, block, 0, as_stream(stream)>>>(acc_cube, cube_out, npix, g->nv, 1.0 / scale_w, moments);
    else
        return fail(PNM_EINVAL, "pnm_finalize_cube: acc_mode %d", acc_mode);
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

// ------------------------------------------------------------------------------------ slab-sharded finalisation
namespace {
int slab_args(const char* who, const pnm_grid* g, const void* slabs, int g0, int g1, int acc_mode) {
    if (!g || !slabs) return fail(PNM_EINVAL, "%s: null argument", who);
    if (g->nv < 1) return fail(PNM_EINVAL, "%s: the grid has no V axis", who);
    if (g0 < 0 || g1 < g0) return fail(PNM_EINVAL, "%s: slab range [%d, %d)", who, g0, g1);
    if (acc_mode != PNM_ACC_F64 && acc_mode != PNM_ACC_FIXED) return fail(PNM_EINVAL, "%s: acc_mode %d", who, acc_mode);
    return PNM_OK;
}
}  // namespace

int pnm_copy_async(void* dst, const void* src, int64_t nbytes, void* stream) {
    if (nbytes < 0 || (nbytes > 0 && (!dst || !src))) return fail(PNM_EINVAL, "pnm_copy_async: bad argument");
    if (nbytes == 0) return PNM_OK;
    PNM_CUDA(cudaMemcpyAsync(dst, src, (size_t)nbytes, cudaMemcpyDefault, as_stream(stream)));
    return PNM_OK;
}

int pnm_pull_chunks(void* stage, const void* const* bases_host, int32_t nranks, int32_t rank, int64_t chunk_bytes, void* stream) {
    if (!stage || !bases_host || nranks < 1 || nranks > PNM_MAX_PEERS || rank < 0 || rank >= nranks || chunk_bytes < 0)
        return fail(PNM_EINVAL, "pnm_pull_chunks: bad argument (at most %d ranks)", PNM_MAX_PEERS);
    if (chunk_bytes == 0) return PNM_OK;
    // chunk `rank` of every rank's buffer -> stage[p]; own buffer first, then the peers in ring order (every link busy)
    for (int k = 0; k < nranks; ++k) {
        const int p = (rank + k) % nranks;
        if (!bases_host[p]) return fail(PNM_EINVAL, "pnm_pull_chunks: null base %d", p);
        PNM_CUDA(cudaMemcpyAsync(static_cast<char*>(stage) + (size_t)p * chunk_bytes,
                                 static_cast<const char*>(bases_host[p]) + (size_t)rank * chunk_bytes, (size_t)chunk_bytes,
                                 cudaMemcpyDefault, as_stream(stream)));
    }
    return PNM_OK;
}

int pnm_zero_async(void* dst, int64_t nbytes, void* stream) {
    if (nbytes < 0 || (nbytes > 0 && !dst)) return fail(PNM_EINVAL, "pnm_zero_async: bad argument");
    if (nbytes == 0) return PNM_OK;
    PNM_CUDA(cudaMemsetAsync(dst, 0, (size_t)nbytes, as_stream(stream)));
    return PNM_OK;
}

int pnm_sum_peers(const void* const* ptrs_host, int32_t nptrs, int64_t n, int acc_mode, void* out, void* stream) {
    if (!ptrs_host || !out || nptrs < 1 || nptrs > PNM_MAX_PEERS || n < 0) return fail(PNM_EINVAL, "pnm_sum_peers: bad argument (at most %d arrays)", PNM_MAX_PEERS);
    if (acc_mode != PNM_ACC_F64 && acc_mode != PNM_ACC_FIXED) return fail(PNM_EINVAL, "pnm_sum_peers: acc_mode %d", acc_mode);
    pnm::PeerPtrs src;
    memset(&src, 0, sizeof(src));
    src.n = nptrs;
    for (int r = 0; r < nptrs; ++r) {
        if (!ptrs_host[r]) return fail(PNM_EINVAL, "pnm_sum_peers: null array %d", r);
        src.p[r] = ptrs_host[r];
    }
    if (n == 0) return PNM_OK;
    const int grid = (int)((n + 255) / 256);
    if (acc_mode == PNM_ACC_F64) pnm::k_sum_peers<PNM_ACC_F64><<<grid, 256, 0, as_stream(stream)>>>(src, n, out);
    else pnm::k_sum_peers<PNM_ACC_FIXED><<<grid, 256, 0, as_stream(stream)>>>(src, n, out);
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int pnm_sum_chunks(const void* chunks, int32_t nchunks, int64_t per, int acc_mode, void* out, void* stream) {
    if (!chunks || !out || nchunks < 1 || per < 0) return fail(PNM_EINVAL, "pnm_sum_chunks: bad argument");
    if (acc_mode != PNM_ACC_F64 && acc_mode != PNM_ACC_FIXED) return fail(PNM_EINVAL, "pnm_sum_chunks: acc_mode %d", acc_mode);
    if ((reinterpret_cast<uintptr_t>(chunks) | reinterpret_cast<uintptr_t>(out)) & 15) return fail(PNM_EINVAL, "pnm_sum_chunks: arrays must be 16-byte aligned");
    if (per == 0) return PNM_OK;
    const int grid = grid_for((per + 1) / 2, 256, 8);
    if (acc_mode == PNM_ACC_F64) pnm::k_sum_chunks<PNM_ACC_F64><<<grid, 256, 0, as_stream(stream)>>>(chunks, nchunks, per, out);
    else pnm::k_sum_chunks<PNM_ACC_FIXED><<<grid, 256, 0, as_stream(stream)>>>(chunks, nchunks, per, out);
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int pnm_slab_sums(const pnm_grid* g, const void* slabs, int32_t g0, int32_t g1, int acc_mode, void* sums3, void* stream) {
    const int rc = slab_args("pnm_slab_sums", g, slabs, g0, g1, acc_mode);
    if (rc != PNM_OK) return rc;
    if (!sums3) return fail(PNM_EINVAL, "pnm_slab_sums: sums3 is null");
    const int npix = g->nx * g->ny, G = (g->nv + 1) / 2;
    const int grid = (npix + 255) / 256;
    if (acc_mode == PNM_ACC_F64) pnm::k_slab_sums<PNM_ACC_F64><<<grid, 256, 0, as_stream(stream)>>>(slabs, npix, g0, g1, G, sums3);
    else pnm::k_slab_sums<PNM_ACC_FIXED><<<grid, 256, 0, as_stream(stream)>>>(slabs, npix, g0, g1, G, sums3);
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int pnm_finalize_sums(const pnm_grid* g, const void* sums3, int acc_mode, const double* scales3_host, double* M, double* V,
                      double* S, void* stream) {
    if (!g || !sums3) return fail(PNM_EINVAL, "pnm_finalize_sums: null argument");
    const int npix = g->nx * g->ny;
    const int grid = (npix + 255) / 256;
    if (acc_mode == PNM_ACC_F64) {
        pnm::k_finalize_sums<PNM_ACC_F64><<<grid, 256, 0, as_stream(stream)>>>(sums3, g->nx, g->ny, 1.0, 1.0, 1.0, M, V, S);
    } else if (acc_mode == PNM_ACC_FIXED) {
        if (!scales3_host) return fail(PNM_EINVAL, "pnm_finalize_sums: FIXED mode needs scales");
        pnm::k_finalize_sums<PNM_ACC_FIXED><<<grid, 256, 0, as_stream(stream)>>>(sums3, g->nx, g->ny, 1.0 / scales3_host[0],
                                                                              1.0 / scales3_host[1], 1.0 / scales3_host[2], M, V, S);
    } else {
        return fail(PNM_EINVAL, "pnm_finalize_sums: acc_mode %d", acc_mode);
    }
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int pnm_finalize_cube_slabs(const pnm_grid* g, const void* slabs, int32_t g0, int32_t g1, int acc_mode, double scale_w,
                            double* cube_out, void* stream) {
    const int rc = slab_args("pnm_finalize_cube_slabs", g, slabs, g0, g1, acc_mode);
    if (rc != PNM_OK) return rc;
    const int iv0 = std::min(2 * g0, g->nv), iv1 = std::min(2 * g1, g->nv);
    const int nvl = iv1 - iv0;
    if (nvl <= 0) return PNM_OK;                     // this rank holds no velocity bins (only padding / the rest slab)
    if (!cube_out) return fail(PNM_EINVAL, "pnm_finalize_cube_slabs: cube_out is null");
    const int npix = g->nx * g->ny;
    const dim3 grid((npix + 31) / 32, (nvl + 31) / 32), block(32, 8);
    if (acc_mode == PNM_ACC_F64)
        pnm::k_cube_out_slabs<PNM_ACC_F64><<<grid, block, 0, as_stream(stream)>>>(slabs, cube_out, npix, g0, iv0, nvl, 1.0);
    else
        pnm::k_cube_out_slabs<PNM_ACC_FIXED><<<grid, block, 0, as_stream(stream)>>>(slabs, cube_out, npix, g0, iv0, nvl, 1.0 / scale_w);
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int pnm_fold_maps(const pnm_grid* g, void* acc_maps, int acc_mode, int clear_others, void* stream) {
    if (!g || !acc_maps) return fail(PNM_EINVAL, "pnm_fold_maps: null argument");
    if (acc_mode != PNM_ACC_F64 && acc_mode != PNM_ACC_FIXED) return fail(PNM_EINVAL, "pnm_fold_maps: acc_mode %d", acc_mode);
    const int64_t per = (int64_t)g->nx * g->ny * PNM_MAP_SLOTS;
    const int grid = (int)((per + 255) / 256);
    cudaStream_t s = as_stream(stream);
    if (acc_mode == PNM_ACC_F64) {
        if (clear_others) pnm::k_fold_maps<PNM_ACC_F64, true><<<grid, 256, 0, s>>>(acc_maps, per, g->replicas);
        else pnm::k_fold_maps<PNM_ACC_F64, false><<<grid, 256, 0, s>>>(acc_maps, per, g->replicas);
    } else {
        if (clear_others) pnm::k_fold_maps<PNM_ACC_FIXED, true><<<grid, 256, 0, s>>>(acc_maps, per, g->replicas);
        else pnm::k_fold_maps<PNM_ACC_FIXED, false><<<grid, 256, 0, s>>>(acc_maps, per, g->replicas);
    }
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

// ------------------------------------------------------------------------------------ smoothing
int pnm_smooth_cube(const double* cube_in, double* cube_out, double* work, int32_t nx, int32_t ny, int32_t nv,
                    const double* taps0_host, int32_t ntaps0, const double* taps1_host, int32_t ntaps1,
                    void* stream) {
    if (!cube_in || !cube_out || !work || !taps0_host || !taps1_host) return fail(PNM_EINVAL, "pnm_smooth_cube: null argument");
    if (cube_in == cube_out || work == cube_in || work == cube_out) return fail(PNM_EINVAL, "pnm_smooth_cube: buffers must be distinct");
    if (nx < 1 || ny < 1 || nv < 1) return fail(PNM_EINVAL, "pnm_smooth_cube: bad shape");
    if (ntaps0 < 1 || ntaps1 < 1 || !(ntaps0 & 1) || !(ntaps1 & 1)) return fail(PNM_EINVAL, "pnm_smooth_cube: tap counts must be odd");
    if (ntaps0 > PNM_MAX_TAPS || ntaps1 > PNM_MAX_TAPS) return fail(PNM_ELIMIT, "pnm_smooth_cube: more than %d taps", PNM_MAX_TAPS);
    pnm::Taps t0, t1;
    memset(&t0, 0, sizeof(t0)); memset(&t1, 0, sizeof(t1));
    t0.n = ntaps0; memcpy(t0.t, taps0_host, sizeof(double) * ntaps0);
    t1.n = ntaps1; memcpy(t1.t, taps1_host, sizeof(double) * ntaps1);
    const int64_t total = (int64_t)nx * ny * nv;
    const int grid = grid_for(total, 256, 8);
    // one thread per kSmoothOut outputs along the convolved axis
    // (second pass: outputs with a NaN in their footprint are redone in 2-D with astropy's nan_treatment='interpolate')
    pnm::k_smooth_axis<0><<<grid, 256, 0, as_stream(stream)>>>(cube_in, work, nx, ny, nv, t0, nullptr, t0);
    pnm::k_smooth_axis<1><<<grid, 256, 0, as_stream(stream)>>>(work, cube_out, nx, ny, nv, t1, cube_in, t0);
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int pnm_cube_moments(const double* cube, int32_t nx, int32_t ny, int32_t nv, const double* binv_dev,
                     double empty_m2, double* mom0, double* mom1, double* mom2, void* stream) {
    if (!cube || !binv_dev || !mom0 || !mom1 || !mom2) return fail(PNM_EINVAL, "pnm_cube_moments: null argument");
    const int nspax = nx * ny;
    const int grid = (nspax * 32 + 255) / 256;
    pnm::k_cube_moments<<<grid, 256, 0, as_stream(stream)>>>(cube, nspax, nv, binv_dev, empty_m2, mom0, mom1, mom2);
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int pnm_fit_gh(const double* cube, const double* err_dev, int64_t nspax, int32_t nv, const double* binv_dev,
               int32_t deg_gh, const double* start, const double* lo, const double* hi, const int32_t* fixed,
               double guess_h, double empty_m2, double ftol, double xtol, double gtol, int32_t maxiter, double* gh,
               double* bestfit, int32_t* status, int32_t* niter, double* chi2, void* stream) {
    if (nspax < 0 || nv < 1) return fail(PNM_EINVAL, "pnm_fit_gh: bad sizes");
    if (deg_gh < 2 || deg_gh > pnm::kFitMaxPar)
        return fail(PNM_ELIMIT, "pnm_fit_gh: deg_gh must be in 2..%d (got %d)", pnm::kFitMaxPar, deg_gh);
    if (!lo || !hi) return fail(PNM_EINVAL, "pnm_fit_gh: null limits");
    if (!(ftol > 0.0) || !(xtol > 0.0) || !(gtol > 0.0) || maxiter < 0)
        return fail(PNM_EINVAL, "pnm_fit_gh: input keywords are inconsistent");            // mpfit.py:985-988
    int nfree = 0;
    for (int k = 0; k < deg_gh; ++k) {
        const bool fx = fixed && fixed[k];
        nfree += fx ? 0 : 1;
        if (lo[k] != lo[k] || hi[k] != hi[k] || (!fx && std::isfinite(lo[k]) && std::isfinite(hi[k]) && lo[k] >= hi[k]))
            return fail(PNM_EINVAL, "pnm_fit_gh: parameter limits are not consistent");      // mpfit.py:959-963
    }
    if (nfree == 0) return fail(PNM_EINVAL, "pnm_fit_gh: no free parameters");               // mpfit.py:942-944
    if (nv < nfree) return fail(PNM_EINVAL, "pnm_fit_gh: number of parameters must not exceed data");   // mpfit.py:1016-1018
    if (nspax == 0) return PNM_OK;
    if (!cube || !binv_dev || !gh || !status || !niter || !chi2) return fail(PNM_EINVAL, "pnm_fit_gh: null argument");
    pnm::FitParams P;
    P.cube = cube; P.err = err_dev; P.binv = binv_dev; P.nv = nv; P.nspax = nspax;
    P.fixed_mask = 0;
    for (int k = 0; k < pnm::kFitMaxPar; ++k) {
        const bool used = k < deg_gh;
        P.start[k] = (used && start) ? start[k] : std::nan("");
        P.lo[k] = used ? lo[k] : 0.0;
        P.hi[k] = used ? hi[k] : 0.0;
        if (used && fixed && fixed[k]) P.fixed_mask |= 1u << k;
    }
    P.guess_h = guess_h; P.empty_m2 = empty_m2; P.ftol = ftol; P.xtol = xtol; P.gtol = gtol; P.maxiter = maxiter;
    P.gh = gh; P.bestfit = bestfit; P.status = status; P.niter = niter; P.chi2 = chi2;
    const unsigned grid = (unsigned)((nspax + pnm::kFitWarps - 1) / pnm::kFitWarps);
    const int threads = pnm::kFitWarps * 32;
    cudaStream_t s = as_stream(stream);
    switch (deg_gh) {
        case 2: pnm::k_fit_gh<2><<<grid, threads, 0, s>>>(P); break;
        case 3: pnm::k_fit_gh<3><<<grid, threads, 0, s>>>(P); break;
        case 4: pnm::k_fit_gh<4><<<grid, threads, 0, s>>>(P); break;
        case 5: pnm::k_fit_gh<5><<<grid, threads, 0, s>>>(P); break;
        default: pnm::k_fit_gh<6><<<grid, threads, 0, s>>>(P); break;
    }
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int pnm_amr_jitter(const void* coord, int dtype, const double* scale_dev, const double* uniform_dev, int64_t n,
                   int32_t nsample, double stretch, double* out, void* stream) {
    if (n < 0 || nsample < 1) return fail(PNM_EINVAL, "pnm_amr_jitter: bad sizes");
    if (n == 0) return PNM_OK;
    if (!coord || !scale_dev || !uniform_dev || !out) return fail(PNM_EINVAL, "pnm_amr_jitter: null argument");
    const int64_t total = n * nsample;
    const int grid = grid_for(total, 256, 16);
    if (dtype == PNM_F32) pnm::k_amr_jitter<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(coord), scale_dev, uniform_dev, total, nsample, stretch, out);
    else if (dtype == PNM_F64) pnm::k_amr_jitter<double><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const double*>(coord), scale_dev, uniform_dev, total, nsample, stretch, out);
    else return fail(PNM_EINVAL, "pnm_amr_jitter: dtype %d", dtype);
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int pnm_amr_repeat(const void* values, int dtype, int64_t n, int32_t nsample, int divide, void* out, void* stream) {
    if (n < 0 || nsample < 1) return fail(PNM_EINVAL, "pnm_amr_repeat: bad sizes");
    if (n == 0) return PNM_OK;
    if (!values || !out || values == out) return fail(PNM_EINVAL, "pnm_amr_repeat: bad buffers");
    const int64_t total = n * nsample;
    const int grid = grid_for(total, 256, 16);
    if (dtype == PNM_F32) pnm::k_amr_repeat<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(values), total, nsample, divide ? 1 : 0, static_cast<float*>(out));
    else if (dtype == PNM_F64) pnm::k_amr_repeat<double><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const double*>(values), total, nsample, divide ? 1 : 0, static_cast<double*>(out));
    else return fail(PNM_EINVAL, "pnm_amr_repeat: dtype %d", dtype);
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int64_t pnm_profile_acc_elems(int32_t nbins) {
    return nbins < 1 ? 0 : (int64_t)pnm::kProfileReplicas * nbins * pnm::kProfileSlots;
}

int pnm_profile_rmax(const void* x, const void* y, const void* z, const uint8_t* mask, int64_t n, int dtype,
                     double z_lo, double z_hi, double* out_dev, void* stream) {
    if (!out_dev || n < 0 || (n > 0 && (!x || !y || !z))) return fail(PNM_EINVAL, "pnm_profile_rmax: bad argument");
    PNM_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double), as_stream(stream)));
    if (n == 0) return PNM_OK;
    const int grid = grid_for(n, 256, 8);
    auto* out = reinterpret_cast<unsigned long long*>(out_dev);
    if (dtype == PNM_F32) pnm::k_profile_rmax<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)x, (const float*)y, (const float*)z, mask, n, z_lo, z_hi, out);
    else if (dtype == PNM_F64) pnm::k_profile_rmax<double><<<grid, 256, 0, as_stream(stream)>>>((const double*)x, (const double*)y, (const double*)z, mask, n, z_lo, z_hi, out);
    else return fail(PNM_EDTYPE, "pnm_profile_rmax: dtype %d", dtype);
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int pnm_radial_profiles(const void* x, const void* y, const void* z, const void* vx, const void* vy, const void* vz,
                        const uint8_t* mask, int64_t n, int dtype, double z_lo, double z_hi, const double* rbin_dev,
                        int32_t nbins, double* acc, float* v, float* s, int64_t* counts, void* stream) {
    if (n < 0 || nbins < 1) return fail(PNM_EINVAL, "pnm_radial_profiles: bad sizes");
    if (nbins > pnm::kProfileMaxBins) return fail(PNM_ELIMIT, "pnm_radial_profiles: more than %d bins", pnm::kProfileMaxBins);
    if (!rbin_dev || !acc || !v || !s) return fail(PNM_EINVAL, "pnm_radial_profiles: null argument");
    if (n > 0 && (!x || !y || !z || !vx || !vy || !vz)) return fail(PNM_EINVAL, "pnm_radial_profiles: null particle array");
    cudaStream_t st = as_stream(stream);
    PNM_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * pnm_profile_acc_elems(nbins), st));
    if (n > 0) {
        const int grid = grid_for(n, 256, 8);
        const size_t smem = sizeof(double) * nbins;
        if (dtype == PNM_F32)
            pnm::k_profile_acc<float><<<grid, 256, smem, st>>>((const float*)x, (const float*)y, (const float*)z, (const float*)vx,
                                                              (const float*)vy, (const float*)vz, mask, n, z_lo, z_hi, rbin_dev, nbins, acc);
        else if (dtype == PNM_F64)
            pnm::k_profile_acc<double><<<grid, 256, smem, st>>>((const double*)x, (const double*)y, (const double*)z, (const double*)vx,
                                                               (const double*)vy, (const double*)vz, mask, n, z_lo, z_hi, rbin_dev, nbins, acc);
        else return fail(PNM_EDTYPE, "pnm_radial_profiles: dtype %d", dtype);
        PNM_CUDA(cudaGetLastError());
    }
    pnm::k_profile_final<<<(nbins + 127) / 128, 128, 0, st>>>(acc, nbins, v, s, reinterpret_cast<long long*>(counts));
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int pnm_rebin_cube(const double* cube_in, double* cube_out, int32_t nx, int32_t ny, int32_t nv,
                   int32_t factor_x, int32_t factor_y, int mean, void* stream) {
    if (!cube_in || !cube_out || cube_in == cube_out) return fail(PNM_EINVAL, "pnm_rebin_cube: bad buffers");
    if (nx < 1 || ny < 1 || nv < 1 || factor_x < 1 || factor_y < 1 || factor_x > nx || factor_y > ny)
        return fail(PNM_EINVAL, "pnm_rebin_cube: bad shape / factors");
    const int n0 = nx / factor_x, n1 = ny / factor_y;        // trailing rows / columns are dropped by the caller's reshape rule
    const int64_t total = (int64_t)n0 * n1 * nv;
    pnm::k_rebin_cube<<<grid_for(total, 256, 8), 256, 0, as_stream(stream)>>>(cube_in, cube_out, ny, nv, n0, n1,
                                                                             factor_x, factor_y, mean ? 1 : 0);
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int pnm_ct_interp(const void* qx, const void* qy, int64_t n, int dtype, const double* tri_xform, const double* tri_coef,
                  int32_t ntri, const int32_t* cell_start, const int32_t* cell_tris, int32_t ncell,
                  double lo0, double lo1, double inv0, double inv1, const double* nodes, int32_t nnode,
                  int fill_nearest, double* out, void* stream) {
    if (n < 0 || ntri < 1 || ncell < 1 || nnode < 1) return fail(PNM_EINVAL, "pnm_ct_interp: bad sizes");
    if (nnode > pnm::kCtMaxNodes) return fail(PNM_ELIMIT, "pnm_ct_interp: more than %d table nodes", pnm::kCtMaxNodes);
    if (n == 0) return PNM_OK;
    if (!qx || !qy || !tri_xform || !tri_coef || !cell_start || !cell_tris || !nodes || !out)
        return fail(PNM_EINVAL, "pnm_ct_interp: null argument");
    const int grid = grid_for(n, 256, 8);
    const size_t smem = (size_t)nnode * 3 * sizeof(double);
    if (dtype == PNM_F32) {
        if (smem > 48 * 1024) PNM_CUDA(cudaFuncSetAttribute(pnm::k_ct_interp<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pnm::k_ct_interp<float><<<grid, 256, smem, as_stream(stream)>>>((const float*)qx, (const float*)qy, n, tri_xform, tri_coef,
            cell_start, cell_tris, ncell, lo0, lo1, inv0, inv1, nodes, nnode, fill_nearest, out);
    } else if (dtype == PNM_F64) {
        if (smem > 48 * 1024) PNM_CUDA(cudaFuncSetAttribute(pnm::k_ct_interp<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pnm::k_ct_interp<double><<<grid, 256, smem, as_stream(stream)>>>((const double*)qx, (const double*)qy, n, tri_xform, tri_coef,
            cell_start, cell_tris, ncell, lo0, lo1, inv0, inv1, nodes, nnode, fill_nearest, out);
    } else {
        return fail(PNM_EDTYPE, "pnm_ct_interp: dtype %d", dtype);
    }
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int pnm_order_probe(const void* x, const void* y, const void* z, int64_t n, int dtype, int64_t stride, int64_t nsamples,
                    double limit, uint64_t* out2_dev, void* stream) {
    if (!out2_dev || n < 0 || stride < 1 || nsamples < 0 || (n > 0 && (!x || !y || !z))) return fail(PNM_EINVAL, "pnm_order_probe: bad argument");
    PNM_CUDA(cudaMemsetAsync(out2_dev, 0, 2 * sizeof(uint64_t), as_stream(stream)));
    if (n < 2 || nsamples == 0) return PNM_OK;
    const int grid = (int)((nsamples + 255) / 256);
    unsigned long long* out = reinterpret_cast<unsigned long long*>(out2_dev);
    if (dtype == PNM_F32)
        pnm::k_order_probe<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)x, (const float*)y, (const float*)z, n, stride, nsamples, limit * limit, out);
    else if (dtype == PNM_F64)
        pnm::k_order_probe<double><<<grid, 256, 0, as_stream(stream)>>>((const double*)x, (const double*)y, (const double*)z, n, stride, nsamples, limit * limit, out);
    else return fail(PNM_EDTYPE, "pnm_order_probe: dtype %d", dtype);
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

int pnm_absmax(const void* a, int64_t n, int dtype, double* out_dev, void* stream) {
    if (!out_dev || n < 0 || (n > 0 && !a)) return fail(PNM_EINVAL, "pnm_absmax: bad argument");
    PNM_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double), as_stream(stream)));
    if (n == 0) return PNM_OK;
    const int grid = grid_for(n, 256, 8);
    if (dtype == PNM_F32) pnm::k_absmax<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)a, n, (unsigned long long*)out_dev);
    else if (dtype == PNM_F64) pnm::k_absmax<double><<<grid, 256, 0, as_stream(stream)>>>((const double*)a, n, (unsigned long long*)out_dev);
    else return fail(PNM_EDTYPE, "pnm_absmax: dtype %d", dtype);
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

// ------------------------------------------------------------------------------------ privatised scatter (pnm_cube.cuh)
}  // extern "C"

struct pnm_cube_plan {
    int device;
    int nx, ny, nv;
    void* buf;                      // one allocation: [zeroed per build: colsum | rowsum | tot_len | tot_cap | voxhist] | rest
    size_t zero_bytes;
    pnm::CubePlan* plan;
    pnm::PlanScratch s;
    unsigned long long* work;       // chunk counter of k_project_cube
    int reserved_sms;               // SMs k_project_cube leaves to other streams
    bool built;
};

namespace {

float round_toward_zero_f32(double v) {
    float f = (float)v;
    if (std::fabs((double)f) > std::fabs(v)) f = nextafterf(f, 0.0f);
    return f;
}

// the part of CubeParams that depends on the grid, the particles and the view
int cube_params(const char* who, const pnm_grid* g, int64_t n, const float* x, const float* y, const float* z,
                const float* vx, const float* vy, const float* vz, const float* w, const float* mat9_host,
                pnm::CubeParams& p) {
    if (!g) return fail(PNM_EINVAL, "%s: grid is null", who);
    if (g->nv < 1) return fail(PNM_EINVAL, "%s: the grid has no V axis", who);
    if (n < 0) return fail(PNM_EINVAL, "%s: n < 0", who);
    if (!mat9_host) return fail(PNM_EINVAL, "%s: matrix is null", who);
    if (n > 0 && (!x || !y || !z || !vx || !vy || !vz)) return fail(PNM_EINVAL, "%s: null particle array", who);
    const void* ptrs[7] = {x, y, z, vx, vy, vz, w};
    for (const void* q : ptrs)
        if (reinterpret_cast<uintptr_t>(q) & 15) return fail(PNM_EINVAL, "%s: particle arrays must be 16-byte aligned", who);
    if (pnm_cube_moments_acc_elems(g) >= (1LL << 31)) return fail(PNM_ELIMIT, "%s: interleaved cube of 2^31 elements or more", who);
    memset(&p, 0, sizeof(p));
    p.n = n; p.x = x; p.y = y; p.z = z; p.vx = vx; p.vy = vy; p.vz = vz; p.w = w;
    p.nx = g->nx; p.ny = g->ny; p.nv = g->nv; p.npix = g->nx * g->ny; p.G = (g->nv + 1) / 2;
    p.ax = (float)g->inv_dx; p.cx = (float)(-g->x0 * g->inv_dx - 0.5); p.hx = round_toward_zero_f32(0.5 - g->eps32[0]);
    p.ay = (float)g->inv_dy; p.cy = (float)(-g->y0 * g->inv_dy - 0.5); p.hy = round_toward_zero_f32(0.5 - g->eps32[1]);
    p.av = (float)g->inv_dv; p.cv = (float)(-g->v0 * g->inv_dv - 0.5); p.hv = round_toward_zero_f32(0.5 - g->eps32[2]);
    p.thr_x = g->thr32; p.thr_y = g->thr32 + g->nx + 1; p.thr_v = g->thr32 + g->nx + 1 + g->ny + 1;
    memcpy(p.m, mat9_host, sizeof(p.m));
    return PNM_OK;
}

}  // namespace

extern "C" {

int32_t pnm_cube_slots_max(void) { return pnm::kCubeSlotsMax; }

int pnm_cube_plan_create(const pnm_grid* g, pnm_cube_plan** out) {
    if (!out) return fail(PNM_EINVAL, "pnm_cube_plan_create: out is null");
    *out = nullptr;
    if (!g || g->nv < 1) return fail(PNM_EINVAL, "pnm_cube_plan_create: needs a grid with a V axis");
    pnm_cube_plan* pl = new pnm_cube_plan();
    pl->nx = g->nx; pl->ny = g->ny; pl->nv = g->nv; pl->built = false; pl->reserved_sms = 0;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_col = up(sizeof(uint32_t) * g->nx), b_row = up(sizeof(uint32_t) * g->ny), b_tot = up(sizeof(uint32_t) * pnm::kPlanLevels);
    const size_t b_sum = up(sizeof(double) * 4);
    const size_t b_vox = up(sizeof(uint32_t) * (size_t)pnm::kBoxPix * g->nv);
    const size_t b_keys = up(sizeof(uint2) * pnm::kPlanSample), b_lo = up(sizeof(uint16_t) * pnm::kBoxPix * pnm::kPlanLevels);
    const size_t b_len = up((size_t)pnm::kBoxPix * pnm::kPlanLevels), b_cap = up(sizeof(uint32_t) * pnm::kBoxPix * pnm::kPlanLevels);
    const size_t b_plan = up(sizeof(pnm::CubePlan)), b_work = 256;
    pl->zero_bytes = b_col + b_row + 2 * b_tot + b_sum + b_vox;
    const size_t total = pl->zero_bytes + b_keys + b_lo + b_len + b_cap + b_plan + b_work;
    cudaError_t e = cudaGetDevice(&pl->device);
    if (e == cudaSuccess) e = cudaMalloc(&pl->buf, total);
    if (e == cudaSuccess) e = cudaMemset(pl->buf, 0, total);
    if (e != cudaSuccess) { delete pl; return fail(PNM_ECUDA, "pnm_cube_plan_create: %s", cudaGetErrorString(e)); }
    unsigned char* b = static_cast<unsigned char*>(pl->buf);
    pl->s.colsum = reinterpret_cast<uint32_t*>(b); b += b_col;
    pl->s.rowsum = reinterpret_cast<uint32_t*>(b); b += b_row;
    pl->s.tot_len = reinterpret_cast<uint32_t*>(b); b += b_tot;
    pl->s.tot_cap = reinterpret_cast<uint32_t*>(b); b += b_tot;
    pl->s.sum_abs = reinterpret_cast<double*>(b); b += b_sum;
    pl->s.voxhist = reinterpret_cast<uint32_t*>(b); b += b_vox;
    pl->s.keys = reinterpret_cast<uint2*>(b); b += b_keys;
    pl->s.lo = reinterpret_cast<uint16_t*>(b); b += b_lo;
    pl->s.len = reinterpret_cast<uint8_t*>(b); b += b_len;
    pl->s.cap = reinterpret_cast<uint32_t*>(b); b += b_cap;
    pl->plan = reinterpret_cast<pnm::CubePlan*>(b); b += b_plan;
    pl->work = reinterpret_cast<unsigned long long*>(b);
    *out = pl;
    return PNM_OK;
}

int pnm_cube_plan_set_reserved_sms(pnm_cube_plan* pl, int32_t nsm) {
    if (!pl || nsm < 0 || nsm > 64) return fail(PNM_EINVAL, "pnm_cube_plan_set_reserved_sms: plan is null or nsm outside 0..64");
    pl->reserved_sms = nsm;
    return PNM_OK;
}

int pnm_cube_plan_destroy(pnm_cube_plan* pl) {
    if (!pl) return PNM_OK;
    cudaFree(pl->buf);
    delete pl;
    return PNM_OK;
}

int pnm_cube_plan_build(const pnm_grid* g, pnm_cube_plan* pl, int64_t n, const float* x, const float* y, const float* z,
                        const float* vx, const float* vy, const float* vz, const float* w, const float* mat9_host, void* stream) {
    if (!pl) return fail(PNM_EINVAL, "pnm_cube_plan_build: plan is null");
    if (!g || g->nx != pl->nx || g->ny != pl->ny || g->nv != pl->nv) return fail(PNM_EINVAL, "pnm_cube_plan_build: the plan belongs to another grid");
    pnm::CubeParams p;
    const int rc = cube_params("pnm_cube_plan_build", g, n, x, y, z, vx, vy, vz, w, mat9_host, p);
    if (rc != PNM_OK) return rc;
    cudaStream_t st = as_stream(stream);
    PNM_CUDA(cudaMemsetAsync(pl->buf, 0, pl->zero_bytes, st));
    // groups of 8 consecutive particles, spread evenly over the snapshot
    const int64_t groups = std::max<int64_t>(1, std::min<int64_t>(pnm::kPlanSample / 8, (n + 7) / 8));
    const int64_t stride8 = std::max<int64_t>(8, ((n / groups) / 8) * 8);
    const int64_t nsample = groups * 8;
    pnm::k_plan_sample<<<grid_for(nsample, 256, 8), 256, 0, st>>>(p, pl->s, nsample, stride8);
    pnm::k_plan_box<<<1, 1024, 0, st>>>(p, pl->s, pl->plan);
    pnm::k_plan_vox<<<grid_for(nsample, 256, 8), 256, 0, st>>>(p, pl->s, pl->plan, nsample);
    pnm::k_plan_levels<<<(pnm::kBoxPix * 32 + 255) / 256, 256, 0, st>>>(p, pl->s, pl->plan);
    pnm::k_plan_finish<<<1, 1024, 0, st>>>(pl->s, pl->plan, pnm::kCubeSlotsMax, (int)nsample);
    PNM_CUDA(cudaGetLastError());
    pl->built = true;
    return PNM_OK;
}

/* what the plan found (synchronises the stream): out8 = { bx0, by0, bw, bh, nslots, level, captured, sampled },
 * stats6 = sample means of |w|, |w*d|, |w*d*d|, then the 2^k the float64 mode scales them with in shared memory */
int pnm_cube_plan_info(const pnm_cube_plan* pl, int32_t* out8_host, double* stats6_host, void* stream) {
    if (!pl || !out8_host || !stats6_host) return fail(PNM_EINVAL, "pnm_cube_plan_info: null argument");
    if (!pl->built) return fail(PNM_EINVAL, "pnm_cube_plan_info: the plan has not been built");
    float sc[3];
    PNM_CUDA(cudaMemcpyAsync(out8_host, pl->plan, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, as_stream(stream)));
    PNM_CUDA(cudaMemcpyAsync(stats6_host, pl->plan->mean_abs, 3 * sizeof(double), cudaMemcpyDeviceToHost, as_stream(stream)));
    PNM_CUDA(cudaMemcpyAsync(sc, pl->plan->sc_smem, 3 * sizeof(float), cudaMemcpyDeviceToHost, as_stream(stream)));
    PNM_CUDA(cudaStreamSynchronize(as_stream(stream)));
    for (int k = 0; k < 3; ++k) stats6_host[3 + k] = (double)sc[k];
    return PNM_OK;
}

int pnm_project_cube(const pnm_grid* g, const pnm_cube_plan* pl, int64_t n, const float* x, const float* y, const float* z,
                     const float* vx, const float* vy, const float* vz, const float* w, const float* mat9_host,
                     int acc_mode, const double* scales3_host, void* acc_cube, void* stream) {
    pnm::CubeParams p;
    const int rc = cube_params("pnm_project_cube", g, n, x, y, z, vx, vy, vz, w, mat9_host, p);
    if (rc != PNM_OK) return rc;
    if (!pl || !pl->built) return fail(PNM_EINVAL, "pnm_project_cube: needs a built plan (pnm_cube_plan_build)");
    if (g->nx != pl->nx || g->ny != pl->ny || g->nv != pl->nv) return fail(PNM_EINVAL, "pnm_project_cube: the plan belongs to another grid");
    if (!acc_cube) return fail(PNM_EINVAL, "pnm_project_cube: acc_cube is null");
    if (acc_mode != PNM_ACC_F64 && acc_mode != PNM_ACC_FIXED) return fail(PNM_EINVAL, "pnm_project_cube: acc_mode %d", acc_mode);
    if (acc_mode == PNM_ACC_FIXED && !scales3_host) return fail(PNM_EINVAL, "pnm_project_cube: FIXED mode needs scales");
    if (n >= (1LL << 32)) return fail(PNM_ELIMIT, "pnm_project_cube: 2^32 particles or more in one call");   // carries of one CTA fit the signed high limb
    if (n == 0) return PNM_OK;
    for (int k = 0; k < 3; ++k) {
        const double sg = acc_mode == PNM_ACC_FIXED ? scales3_host[k] : 1.0;
        int e = 0;
        // powers of two inside the float32 exponent range (the kernel scales in float32, see to_acc_f)
        if (!(sg > 0.0) || std::frexp(sg, &e) != 0.5 || e < -119 || e > 121)
            return fail(PNM_EINVAL, "pnm_project_cube: scales must be powers of two between 2^-120 and 2^120");
        p.sc_glob[k] = (float)sg;
    }
    p.cube = acc_cube; p.plan = pl->plan; p.work = pl->work;
    const int64_t nchunks = (n + pnm::kCubeChunk - 1) / pnm::kCubeChunk;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(pnm::kNumSMs - pl->reserved_sms, (nchunks + pnm::kCubeWarps - 1) / pnm::kCubeWarps));
#ifdef PNM_CUBE_KNOBS
    p.dbg = env_int("PNM_CUBE_DBG", 0);
#endif
    cudaStream_t st = as_stream(stream);     // (pl->work is zero: pnm_cube_plan_create, then the kernel's last CTA)
#define PNM_LAUNCH_CUBE(ACC, HW)                                                                                    \
    do {                                                                                                            \
        auto kern = pnm::k_project_cube<ACC, HW>;                                                                   \
        PNM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pnm::kCubeSmemBytes));     \
        kern<<<grid, pnm::kCubeThreads, pnm::kCubeSmemBytes, st>>>(p);                                              \
    } while (0)
    if (acc_mode == PNM_ACC_F64) { if (w) PNM_LAUNCH_CUBE(PNM_ACC_F64, true); else PNM_LAUNCH_CUBE(PNM_ACC_F64, false); }
    else { if (w) PNM_LAUNCH_CUBE(PNM_ACC_FIXED, true); else PNM_LAUNCH_CUBE(PNM_ACC_FIXED, false); }
#undef PNM_LAUNCH_CUBE
    PNM_CUDA(cudaGetLastError());
    return PNM_OK;
}

}  // extern "C"
