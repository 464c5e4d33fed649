// pnm_cube.cuh -- k_project_cube: the fused rotate + project + bin kernel for the hot configuration (float32
// particles, one view, no mask, maps + LOSVD cube), with the hot part of the accumulators PRIVATISED IN SHARED MEMORY.
// Reference semantics: rotation.py:115-117, grid.py:126-131 (moment histograms), grid.py:252-259 (3-D histogram).
//
// Why (round-2 measurements, scripts/red_limits_bench.cu -> profiles/r02_micro_red_limits_bench.log):
//   * scattered f64 reductions at L2 cost ~5.9 ps per 32-byte sector request chip-wide (7.1 ps with three lanes in
//     the sector) and this time ADDS to the 0.45 ms the 2.8 GB particle stream needs, even when independent warps
//     issue the two: one request per particle (round 1's interleaved cube) cannot go below ~0.45 + 0.55 ms;
//   * shared-memory 32-bit integer atomics (ATOMS.ADD) are native and fast (0.09 clk per lane operation per SM);
//     64-bit and floating-point shared atomics are CAS loops (0.6 clk, and a same-address chain of ~70 clk per
//     update on a hot pixel);
//   * a disk + bulge snapshot is concentrated: 16 K of the 4.2 M voxels of the 128 x 128 x 256 cube take half of the
//     particles, the central 64 x 64 pixels 89 %.
// So every CTA (one per SM, persistent) keeps, in shared memory, as 64-bit fixed-point numbers made of two 32-bit
// limbs updated with native ATOMS.ADD:
//   * sum w*d and sum w*d*d of the pixels of a box of at most 64 x 64 pixels (the hot region of the map), and
//   * sum w of the hottest voxels: per box pixel one contiguous window of velocity bins (CubePlan::tab), chosen by
//     the plan kernels (pnm_plan.cuh) from a sample of the particles,
// and flushes them to the global accumulator once, at its end.  Only what the plan does not cover goes to L2 per
// particle: w of a box pixel outside its window (one single-lane RED), and all three sums of particles outside the
// box (one three-lane RED into the particle's record of the interleaved cube, gathered through a per-warp queue in
// shared memory so that the three lanes sit in one instruction).  The plan only decides WHERE a sum is accumulated
// first, never what is summed: with int64 accumulators the result is bit-identical for any plan; with float64
// accumulators the privatised terms are rounded to a fixed-point grid of 2^-62 x (terms per CTA) x max|term|.
//
// Global accumulator (PNM_CUBE_MOMENTS layout): records [G + 1][ix][iy][4], G = ceil(nv / 2);
//   record (g < G, pixel) = { w of velocity bin 2g, w of bin 2g+1, sum w*d, sum w*d*d } (the moments are spread
//   over the records of a pixel; pnm_finalize_vmaps sums them over g), record (G, pixel) = the same for particles
//   outside the V range (slot 0) and the flushed shared-memory moments (slots 2, 3).
#pragma once
#include "pnm_bin.cuh"

namespace pnm {

constexpr int kBoxMax = 64;                       // privatised box: at most kBoxMax x kBoxMax pixels
constexpr int kBoxPix = kBoxMax * kBoxMax;
#ifndef PNM_CUBE_THREADS
#define PNM_CUBE_THREADS 1024
#endif
#ifndef PNM_CUBE_PREFETCH
#define PNM_CUBE_PREFETCH 0
#endif
constexpr int kCubeThreads = PNM_CUBE_THREADS;
constexpr int kCubeWarps = kCubeThreads / 32;
constexpr int kQueueCap = 40;                     // records per warp queue: < 8 pending + at most 32 pushed at once
constexpr int kCubeSub = 8;                       // 4-particle vectors per lane and chunk: a chunk = 1024 particles
constexpr int kCubeSmemBytes = 227 * 1024;
// shared memory: table | pixel moments (2 x 8 B per box pixel) | warp queues | voxel slots (8 B each)
constexpr int kCubeSmemFixed = kBoxPix * 4 + kBoxPix * 16 + kCubeWarps * kQueueCap * 16 + 64 + 32 * 8;   // + counter, dummy slots
constexpr int kCubeSlotsMax = (kCubeSmemBytes - kCubeSmemFixed) / 8;

// table entry: base slot (15 bits) | window length (7 bits) << 15 | first velocity bin (10 bits) << 22
constexpr int kTabBaseBits = 15, kTabLenBits = 7, kTabLenMax = (1 << kTabLenBits) - 1, kTabVloMax = 1023;
static_assert(kCubeSlotsMax < (1 << kTabBaseBits), "slot index must fit the table entry");

struct CubePlan {                                 // device resident; written by the plan kernels
    int bx0, by0, bw, bh;                         // box origin / size in pixels (bw * bh == 0: nothing privatised)
    int nslots;                                   // voxel slots in use (<= kCubeSlotsMax)
    int level, captured, sampled;                 // diagnostics: threshold level, sample particles inside the windows, sample size
    uint32_t tab[kBoxPix];                        // per box pixel bxi * bh + byi
    double mean_abs[3];                           // sample means of |w|, |w*d|, |w*d*d| (typical term sizes)
};

struct CubeParams {
    int64_t n;
    const float *x, *y, *z, *vx, *vy, *vz, *w;    // w may be null (unit weights)
    int nx, ny, nv, npix, G;
    float ax, cx, hx, ay, cy, hy, av, cv, hv;     // bin guess: fma(x, a, c) = (x - x0) * inv - 0.5; sure iff |frac - 0.5| < h
    const float *thr_x, *thr_y, *thr_v;           // exact threshold tables (global memory; rare path)
    float m[9];
    float sc_glob[3];                             // int64 accumulators: 2^k of (w, w*d, w*d*d), |k| <= 120; float64: 1
    float sc_smem[3];                             // fixed-point scales of the shared-memory sums (== sc_glob for int64)
    double inv_smem[3];                           // float64 accumulators: 1 / sc_smem (flush)
    void* cube;
    const CubePlan* plan;
    unsigned long long* work;                     // chunk counter, zeroed by the launcher
    int cta_cap;                                  // chunks one CTA may take (bounds the terms per shared-memory sum)
    int dbg;                                      // development builds (-DPNM_CUBE_KNOBS): 1 drop the shared-memory adds, 2 the in-box REDs, 4 the queue
};

// 128-bit streaming load: no L1 allocation, L2 evict-first through a cache policy (the .L2::evict_first qualifier itself
// only exists for 256-bit loads), so that the particle stream does not push the accumulators out of L2
__device__ __forceinline__ uint64_t evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void ldg4(const float* p, float* o, uint64_t pol) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "l"(p), "l"(pol));
}

// bin guess without conversion instructions: gp = (x - x0) * inv - 0.5 (one FMA); t = gp + 1.5 * 2^23 rounds gp to the
// nearest integer nf = t - 1.5 * 2^23 (exact for |gp| < 2^22), whose value also sits in the low mantissa bits of t.
// r = gp - nf in [-0.5, 0.5] is frac((x - x0) * inv) - 0.5, so |r| < 0.5 - eps means the fraction is further than eps
// from 0 and 1 and nf is PROVABLY the numpy bin (eps: pnm_api.cu guard_band, which covers this formula's rounding:
// u * (2 max|edge| * inv + n + 1) <= the 4u * (n + 2 max|edge| * inv) budgeted there).  NaN fails the test; huge |gp|
// gives an index outside [0, n) (the caller's unsigned range check), as the true bin is.
__device__ __forceinline__ int fast_bin(float x, float a, float c, float h, bool& sure) {
    const float kMagic = 12582912.0f;             // 1.5 * 2^23
    const float gp = __fmaf_rn(x, a, c);
    const float t = __fadd_rn(gp, kMagic);
    const float nf = __fsub_rn(t, kMagic);
    const float r = __fsub_rn(gp, nf);
    sure = sure && (fabsf(r) < h);
    return __float_as_int(t) - 0x4B400000;
}

// Exact bin of a value the guess was not sure about: the guess is at most one bin off, so bins c-1, c, c+1 are tested
// against their thresholds (four independent loads) before the whole table is searched.  Every answer is verified by
// thr[b] <= x < thr[b+1], i.e. it is exact whatever the guess was.
template <typename T>
__device__ __forceinline__ int near_bin(T x, const T* __restrict__ thr, int n, int c) {
    if (c >= 1 && c + 2 <= n) {
        const T t0 = __ldg(thr + c - 1), t1 = __ldg(thr + c), t2 = __ldg(thr + c + 1), t3 = __ldg(thr + c + 2);
        if (x >= t1 && x < t2) return c;
        if (x >= t2 && x < t3) return c + 1;
        if (x >= t0 && x < t1) return c - 1;
    }
    return exact_bin<T>(x, thr, n);
}

// 64-bit fixed-point sums in shared memory as two 32-bit limbs, native ATOMS.ADD only.  The particle body is straight
// line: a warp's lanes disagree about nearly every condition, so instead of branching a lane that has nothing to add
// sends its (meaningless) term to a private dummy slot.  The low limb's old value tells this thread whether ITS addition wrapped (add.cc / addc); the
// carries of all additions then land in the high limb in any order.
__device__ __forceinline__ void smem_add64(uint32_t saddr, long long a) {
    const uint32_t alo = (uint32_t)a, ahi = (uint32_t)((unsigned long long)a >> 32);
    asm volatile(
        "{\n\t.reg .pred c;\n\t.reg .u32 old, t, h;\n\t"
        "atom.shared.add.u32 old, [%0], %1;\n\t"
        "add.cc.u32 t, old, %1;\n\t"
        "addc.u32 h, %2, 0;\n\t"
        "setp.ne.u32 c, h, 0;\n\t"
        "@c red.shared.add.u32 [%0+4], h;\n\t}"
        :: "r"(saddr), "r"(alo), "r"(ahi) : "memory");
}
__device__ __forceinline__ void red_f64_if(bool pred, double* addr, double v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q red.global.add.f64 [%1], %2;\n\t}"
                 :: "r"((uint32_t)pred), "l"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void red_u64_if(bool pred, long long* addr, long long v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q red.global.add.u64 [%1], %2;\n\t}"
                 :: "r"((uint32_t)pred), "l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void red_add_if(bool pred, double* addr, double v) { red_f64_if(pred, addr, v); }
__device__ __forceinline__ void red_add_if(bool pred, long long* addr, long long v) { red_u64_if(pred, addr, v); }
__device__ __forceinline__ long long smem_get64(const uint2* slot) {
    const uint2 v = *slot;
    return (long long)(((unsigned long long)v.y << 32) | v.x);
}
// rint(v * 2^k) of a finite float.  The product with a power of two is exact in float32 as long as it neither overflows
// (the host bounds |k| by 120 and sizes 2^k so that n * max|term| * 2^k <= 2^62) nor falls below 2^-126 (where it
// rounds to 0 like the exact product does), so this equals the oracle's rint((double)v * 2^k).
__device__ __forceinline__ long long to_fix(float v, float scale) { return __float2ll_rn(__fmul_rn(v, scale)); }

template <int ACC>
__device__ __forceinline__ typename AccT<ACC>::type to_acc_f(float v, float scale) {
    if (ACC == PNM_ACC_F64) return (typename AccT<ACC>::type)v;
    const float s = __fmul_rn(v, scale);
    return (typename AccT<ACC>::type)(is_finite(s) ? __float2ll_rn(s) : 0ll);
}

#ifdef PNM_CUBE_KNOBS
#define PNM_CUBE_DBG(bit) (p.dbg & (bit))
#else
#define PNM_CUBE_DBG(bit) false
#endif

template <int ACC, bool HAS_W>
__global__ void __launch_bounds__(kCubeThreads, 1) k_project_cube(const __grid_constant__ CubeParams p) {
    using A = typename AccT<ACC>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* const s_tab = reinterpret_cast<uint32_t*>(smem_raw);
    uint2* const s_pm = reinterpret_cast<uint2*>(smem_raw + kBoxPix * 4);
    uint4* const s_queue = reinterpret_cast<uint4*>(smem_raw + kBoxPix * 4 + kBoxPix * 16);
    int* const s_taken = reinterpret_cast<int*>(smem_raw + kBoxPix * 4 + kBoxPix * 16 + kCubeWarps * kQueueCap * 16);
    uint2* const s_ws = reinterpret_cast<uint2*>(smem_raw + kCubeSmemFixed);

    const CubePlan* plan = p.plan;
    const int bx0 = plan ? plan->bx0 : 0, by0 = plan ? plan->by0 : 0;
    const int bw = plan ? plan->bw : 0, bh = plan ? plan->bh : 0;
    const int nslots = plan ? plan->nslots : 0;
    const int nbox = bw * bh;
    for (int i = threadIdx.x; i < nbox; i += kCubeThreads) s_tab[i] = plan->tab[i];
    for (int i = threadIdx.x; i < 2 * nbox; i += kCubeThreads) s_pm[i] = make_uint2(0u, 0u);
    for (int i = threadIdx.x; i < nslots; i += kCubeThreads) s_ws[i] = make_uint2(0u, 0u);
    if (threadIdx.x == 0) *s_taken = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint4* const queue = s_queue + warp * kQueueCap;
    int q_head = 0, q_count = 0;                       // warp-uniform
    A* const cube = reinterpret_cast<A*>(p.cube);
    const int nx = p.nx, ny = p.ny, nv = p.nv, npix = p.npix, G2 = 2 * p.G;
    const float m0 = p.m[0], m1 = p.m[1], m2 = p.m[2], m3 = p.m[3], m4 = p.m[4], m5 = p.m[5], m6 = p.m[6], m7 = p.m[7], m8 = p.m[8];
    const int ql = lane & 3, qj = lane >> 2;
    const float q_scale = p.sc_glob[ql == 0 ? 0 : ql == 1 ? 0 : ql - 1];      // lane 0: w, lane 2: w*d, lane 3: w*d*d

    // eight queued records -> one RED instruction: lane (qj, ql) sends value ql of record qj; the three lanes of a
    // record hit one 32-byte sector, which L2 takes as ONE request
    auto drain8 = [&](int avail) {
        int pos = q_head + qj;
        if (pos >= kQueueCap) pos -= kQueueCap;
        const uint4 r = queue[pos];
        const int off = (int)r.x;
        const float f = __uint_as_float(ql == 0 ? r.y : ql == 2 ? r.z : r.w);
        red_add_if(qj < avail && ql != 1 && !PNM_CUBE_DBG(4), cube + (ql == 0 ? off : (off & ~3) + ql), to_acc_f<ACC>(f, q_scale));
        q_head += 8;
        if (q_head >= kQueueCap) q_head -= kQueueCap;
        q_count -= 8;
    };
    const uint32_t a_pm = (uint32_t)__cvta_generic_to_shared(s_pm), a_ws = (uint32_t)__cvta_generic_to_shared(s_ws);
    const uint32_t a_tab = (uint32_t)__cvta_generic_to_shared(s_tab);
    const uint32_t a_dummy = (uint32_t)__cvta_generic_to_shared(s_taken) + 64u + 8u * (unsigned)lane;   // this lane's dummy slot
    const float sm0 = p.sc_smem[0], sm1 = p.sc_smem[1], sm2 = p.sc_smem[2], sg0 = p.sc_glob[0];

    // one particle per lane; warp-synchronous (every lane calls it; a lane without a particle passes x = NaN, which no
    // bin takes).  Straight-line: every accumulation is predicated.
    auto particle = [&](float x, float y, float z, float vx, float vy, float vz, float w) {
        const float px = rot_row(m0, m1, m2, x, y, z);
        const float py = rot_row(m3, m4, m5, x, y, z);
        const float d = rot_row(m6, m7, m8, vx, vy, vz);
        bool sure = true;
        int ix = fast_bin(px, p.ax, p.cx, p.hx, sure);
        int iy = fast_bin(py, p.ay, p.cy, p.hy, sure);
        int iv = fast_bin(d, p.av, p.cv, p.hv, sure);
        if (!sure) {                                   // within eps of an edge, NaN: the threshold tables decide
            ix = near_bin<float>(px, p.thr_x, nx, ix);
            iy = near_bin<float>(py, p.thr_y, ny, iy);
            iv = near_bin<float>(d, p.thr_v, nv, iv);
        }
        const bool valid = (unsigned)ix < (unsigned)nx && (unsigned)iy < (unsigned)ny;
        const int ivp = (unsigned)iv < (unsigned)nv ? iv : G2;        // outside the V range: slot 0 of record G
        const float wd = __fmul_rn(w, d), wdd = __fmul_rn(w, __fmul_rn(d, d));
        const int rec = ((ivp >> 1) * npix + ix * ny + iy) * 4 + (ivp & 1);
        const unsigned bxi = (unsigned)(ix - bx0), byi = (unsigned)(iy - by0);
        const bool fin = is_finite(wdd);              // implies finite w and w*d (w*d*d is not smaller than both)
        const bool inbox = valid && fin && bxi < (unsigned)bw && byi < (unsigned)bh;
        const int bp = inbox ? (int)(bxi * (unsigned)bh + byi) : 0;
        uint32_t e;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(a_tab + 4u * (unsigned)bp));
        const bool in_s = inbox && !PNM_CUBE_DBG(1);
        smem_add64(in_s ? a_pm + 16u * (unsigned)bp : a_dummy, to_fix(wd, sm1));
        smem_add64(in_s ? a_pm + 16u * (unsigned)bp + 8u : a_dummy, to_fix(wdd, sm2));
        const unsigned rel = (unsigned)(ivp - (int)(e >> 22));
        const bool cap = inbox && rel < ((e >> kTabBaseBits) & (unsigned)kTabLenMax);
        smem_add64(cap && !PNM_CUBE_DBG(1) ? a_ws + 8u * ((e & ((1u << kTabBaseBits) - 1u)) + rel) : a_dummy, to_fix(w, sm0));
        red_add_if(inbox && !cap && !PNM_CUBE_DBG(2), cube + rec, to_acc_f<ACC>(w, sg0));
        const bool queued = valid && !inbox;
        const unsigned mask = __ballot_sync(0xffffffffu, queued);
        if (mask) {
            if (queued) {
                int pos = q_head + q_count + __popc(mask & lt_mask);
                if (pos >= kQueueCap) pos -= kQueueCap;
                queue[pos] = make_uint4((unsigned)rec, __float_as_uint(w), __float_as_uint(wd), __float_as_uint(wdd));
            }
            q_count += __popc(mask);
            if (q_count >= 8) {
                __syncwarp();
#pragma unroll 1
                while (q_count >= 8) drain8(8);
                __syncwarp();
            }
        }
    };

    // ---- persistent loop: warps take 1024-particle chunks from a global counter (SMs differ in their distance to the
    // hot L2 slices; a static partition leaves the slow ones behind), the next chunk is requested one chunk ahead
    const int64_t nvec = p.n / 4;
    const int64_t nchunks = (nvec + 32 * kCubeSub - 1) / (32 * kCubeSub);
    auto grab = [&]() -> int64_t {
        long long c = 0x7fffffffffffffffll;
        if (lane == 0 && atomicAdd(s_taken, 1) < p.cta_cap) c = (long long)atomicAdd(p.work, 1ull);
        return __shfl_sync(0xffffffffu, c, 0);
    };
    const uint64_t pol = evict_first_policy();
    struct Tile { float x[4], y[4], z[4], vx[4], vy[4], vz[4], w[4]; };
    auto load_tile = [&](Tile& t, int64_t chunk, int sub) {
        const int64_t v = (chunk * kCubeSub + sub) * 32 + lane;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                                   // past the end: x = NaN, which no bin takes
            t.x[k] = __int_as_float(0x7fc00000); t.y[k] = t.z[k] = t.vx[k] = t.vy[k] = t.vz[k] = 0.f; t.w[k] = 1.f;
        }
        if (v < nvec) {
            ldg4(p.x + v * 4, t.x, pol); ldg4(p.y + v * 4, t.y, pol); ldg4(p.z + v * 4, t.z, pol);
            ldg4(p.vx + v * 4, t.vx, pol); ldg4(p.vy + v * 4, t.vy, pol); ldg4(p.vz + v * 4, t.vz, pol);
            if (HAS_W) ldg4(p.w + v * 4, t.w, pol);
        }
    };
    auto process_tile = [&](const Tile& t) {
#pragma unroll
        for (int k = 0; k < 4; ++k) particle(t.x[k], t.y[k], t.z[k], t.vx[k], t.vy[k], t.vz[k], HAS_W ? t.w[k] : 1.0f);
    };
    int64_t chunk = grab();
#if PNM_CUBE_PREFETCH
    // software pipeline: the loads of the next tile are in flight while the current one is processed
    int64_t next = chunk < nchunks ? grab() : chunk;
    int sub = 0;
    auto advance = [&]() {
        if (++sub == kCubeSub) { sub = 0; chunk = next; if (chunk < nchunks) next = grab(); }
    };
    Tile ta, tb;
    if (chunk < nchunks) load_tile(ta, chunk, 0);
    while (chunk < nchunks) {
        advance();
        if (chunk < nchunks) load_tile(tb, chunk, sub);
        process_tile(ta);
        if (!(chunk < nchunks)) break;
        advance();
        if (chunk < nchunks) load_tile(ta, chunk, sub);
        process_tile(tb);
    }
#else
    while (chunk < nchunks) {
        const int64_t next = grab();
#pragma unroll 1
        for (int sub = 0; sub < kCubeSub; ++sub) {
            Tile t;
            load_tile(t, chunk, sub);
            process_tile(t);
        }
        chunk = next;
    }
#endif
    // tail: the last n % 4 particles, by the first warp of the first CTA
    if (blockIdx.x == 0 && warp == 0) {
        const int64_t i = nvec * 4 + lane;
        const bool live = lane < 4 && i < p.n;
        float t[7] = {__int_as_float(0x7fc00000), 0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
        if (live) {
            t[0] = p.x[i]; t[1] = p.y[i]; t[2] = p.z[i]; t[3] = p.vx[i]; t[4] = p.vy[i]; t[5] = p.vz[i];
            if (HAS_W) t[6] = p.w[i];
        }
        particle(t[0], t[1], t[2], t[3], t[4], t[5], t[6]);
    }
    __syncwarp();
    if (q_count > 0) { const int left = q_count; drain8(left); }

    // ---- flush the shared-memory sums: one RED per non-zero sum
    __syncthreads();
    const double i0 = p.inv_smem[0], i1 = p.inv_smem[1], i2 = p.inv_smem[2];
    for (int bp = threadIdx.x; bp < nbox; bp += kCubeThreads) {
        const int bxi = bp / bh, byi = bp - bxi * bh;
        const int pix = (bx0 + bxi) * ny + (by0 + byi);
        A* const rest = cube + ((int64_t)p.G * npix + pix) * 4;
        const long long s1 = smem_get64(s_pm + 2 * bp), s2 = smem_get64(s_pm + 2 * bp + 1);
        if (s1) red_add(rest + 2, ACC == PNM_ACC_F64 ? (A)((double)s1 * i1) : (A)s1);
        if (s2) red_add(rest + 3, ACC == PNM_ACC_F64 ? (A)((double)s2 * i2) : (A)s2);
        const uint32_t e = s_tab[bp];
        const int base = (int)(e & ((1u << kTabBaseBits) - 1u)), len = (int)((e >> kTabBaseBits) & (unsigned)kTabLenMax), vlo = (int)(e >> 22);
        for (int r = 0; r < len; ++r) {
            const long long sw = smem_get64(s_ws + base + r);
            const int iv = vlo + r;
            if (sw) red_add(cube + ((int64_t)(iv >> 1) * npix + pix) * 4 + (iv & 1), ACC == PNM_ACC_F64 ? (A)((double)sw * i0) : (A)sw);
        }
    }
}

}  // namespace pnm
