// pnm_cube.cuh -- k_project_cube: the fused rotate + project + bin kernel for the hot configuration (float32
// particles, one view, no mask, maps + LOSVD cube), with the hot part of the accumulators PRIVATISED IN SHARED MEMORY.
// Reference semantics: rotation.py:115-117, grid.py:126-131 (moment histograms), grid.py:252-259 (3-D histogram).
//
// Why (round-2 measurements, scripts/red_limits_bench.cu -> profiles/r02_micro_red_limits_bench.log):
//   * scattered f64 reductions at L2 cost ~5.9 ps per 32-byte sector request chip-wide (7.1 ps with three lanes in
//     the sector) on top of the 0.45 ms the 2.8 GB particle stream needs: one request per particle (round 1's
//     interleaved cube) cannot go below ~0.45 + 0.55 ms;
//   * shared-memory 32-bit integer atomics (ATOMS.ADD) are native and fast (0.09 clk per lane operation per SM);
//     64-bit and floating-point shared atomics are CAS loops (0.6 clk, and a same-address chain of ~70 clk per
//     update on a hot pixel);
//   * a disk + bulge snapshot is concentrated: 16 K of the 4.2 M voxels of the 128 x 128 x 256 cube take half of the
//     particles, the central 64 x 64 pixels 89 %.
// So every CTA (one per SM, persistent) keeps, in shared memory, as fixed-point numbers whose 32-bit low limb is
// updated with native ATOMS.ADD (the carries out of it go to the global accumulator directly):
//   * sum w*d and sum w*d*d of the pixels of a box of at most 64 x 64 pixels (the hot region of the map), and
//   * sum w of the hottest voxels: per box pixel one contiguous window of velocity bins (CubePlan::tab), chosen by
//     the plan kernels (pnm_plan.cuh) from a sample of the particles,
// and flushes them to the global accumulator once, at its end.  Only what the plan does not cover goes to L2 per
// particle: w of a box pixel outside its window (one single-lane RED), and all three sums of particles outside the
// box (one three-lane RED into the particle's record of the interleaved cube, gathered through a per-warp queue in
// shared memory so that the three lanes sit in one instruction).
//
// Arithmetic of the privatised sums.  A term t (w, w*d or w*d*d, each rounded in float32 like the reference's
// products) enters shared memory as the 32-bit integer a = rint(t * 2^k); a particle whose three scaled terms do
// not all fit 32 bits (or are not finite) takes the global path with all three.  The low limb takes a with one
// ATOMS.ADD whose return value tells this thread whether ITS addition carried; the (rare, ~1 in 256) carry or borrow
// h = +-1 is sent to the global accumulator at once as h * 2^32 (one RED), so shared memory holds 4 bytes per sum and
// the flush adds the unsigned low limbs.
//   * int64 accumulators (PNM_ACC_FIXED): 2^k are the caller's scales, the global path adds the very same integers
//     -> the result is bit-identical for any plan, any CTA count, any order;
//   * float64 accumulators: 2^k is chosen by the plan so that the sample mean of |t| lands in [2^24, 2^25): the
//     privatised terms are rounded to 2^-25 of the typical term (the reference's float32 products already carry that
//     much rounding each), terms up to 64 x the mean stay in shared memory, and particles with a small term (below
//     1/16 of the typical w or w*d, see kCubeWMin) take the exact global path, so that every term is rounded by less
//     than 4.8e-7 of itself and every pixel / voxel sum stays within 4.8e-7 * sum |term| of the float64 sum.
// The plan only decides WHERE a sum is accumulated first, never what is summed.
//
// Global accumulator (PNM_CUBE_MOMENTS layout): records [G + 1][ix][iy][4], G = ceil(nv / 2);
//   record (g < G, pixel) = { w of velocity bin 2g, w of bin 2g+1, sum w*d, sum w*d*d } (the moments are spread
//   over the records of a pixel; pnm_finalize_vmaps sums them over g), record (G, pixel) = the same for particles
//   outside the V range (slot 0) and the flushed shared-memory moments (slots 2, 3).
//
// What bounds it (ncu, round 2; the final kernel is summarised in profiles/r02_ncu_full_summary.txt: 162 warp
// instructions per 32 particles, 67 % issue utilisation, 0.45 L2 reduction requests per particle, DRAM traffic 1.008 x
// the algorithmic bytes).  The first version executed 180 warp instructions per
// particle at 72 % issue utilisation and kept the LSU data pipe 84 % busy (l1tex__data_pipe_lsu_wavefronts): per warp
// and particle 20 wavefronts of shared atomics (the { low, high } pairs put every low limb into 8 or 16 of the 32
// banks), 14 of particle loads (a no-allocate LDG.128 delivers 64 bytes per wavefront), 7 of REDs, 5 of table
// look-ups.  This version: limb arrays with a 4-byte stride (12 wavefronts of atomics), high limbs in global memory
// (4 bytes per sum -> 2.5 x the voxel slots -> 62 % of the particles' w captured instead of 40 %), loop-invariant
// parameters pinned in registers through shared memory, 32-bit conversions, full tiles only in the vector loop, and
// the particles of a tile processed in groups of PNM_CUBE_GROUP with the phases interleaved.  Thread-level
// parallelism beats everything else here: 1024 threads with 64 registers and no software prefetch (0.66 ms per
// 1e8 particles) against 512 threads, 128 registers and a double-buffered tile (0.79 ms).
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
#ifndef PNM_CUBE_PIN
#define PNM_CUBE_PIN 1
#endif
constexpr int kCubeThreads = PNM_CUBE_THREADS;
constexpr int kCubeWarps = kCubeThreads / 32;
#ifndef PNM_CUBE_GROUP
#define PNM_CUBE_GROUP 2
#endif
constexpr int kG = PNM_CUBE_GROUP;                // particles per lane processed together (1, 2 or 4)
constexpr int kQueueCap = kG >= 4 ? 256 : kG == 2 ? 128 : 64;   // records per warp queue (ring, power of two): < 8 pending + at most 32 kG pushed at once
constexpr int kCubeSub = 8;                       // 4-particle vectors per lane and chunk: a chunk = 1024 particles
constexpr int kCubeChunk = 128 * kCubeSub;        // particles per chunk
constexpr int kCubeSmemBytes = 227 * 1024;
// shared memory: table | pixel moments | warp queues | voxel slots -- 32-bit LOW limbs only; the carries go straight
// to the global accumulator (see the kernel header).  One 4-byte array per sum (sum w*d, sum w*d*d, voxel slots):
// consecutive entries fall into consecutive banks, so the 32 scattered ATOMS of a warp spread over all 32 banks (with
// { low, high } pairs every low limb sat in one of 8 or 16 banks: ncu counted 6.7 wavefronts per ATOMS, now ~4).
// Every array ends with 32 dummy entries, one per lane, for the lanes that have nothing to add.
constexpr int kPmStride = (kBoxPix + 32) * 4;     // bytes from the sum w*d array to the sum w*d*d array
constexpr int kCubeSmemFixed = kBoxPix * 4 + 2 * kPmStride + kCubeWarps * kQueueCap * 16;
constexpr int kCubeSlotsMax = (kCubeSmemBytes - kCubeSmemFixed) / 4 - 32 - 4;   // 32 dummy entries + 4 words of carry constants behind the slots
#ifndef PNM_CUBE_TYP_BITS
#define PNM_CUBE_TYP_BITS 24
#endif
constexpr int kCubeTypBits = PNM_CUBE_TYP_BITS;   // float64 mode: the sample mean of |term| * 2^k lies in [2^kCubeTypBits, 2 * 2^kCubeTypBits)
// float64 mode: a particle whose scaled |w| or |w*d| is below 2^20 (1/16 .. 1/32 of the mean) or whose scaled w*d*d is
// below 2^16 takes the exact global path with all three terms.  Every privatised term of sum w and sum w*d is
// therefore rounded by at most 2^-21 of ITSELF (0.5 / 2^20), whatever its size: each pixel / voxel sum is within
// 4.8e-7 * sum |term| of the unrounded float64 sum -- the 1e-6 contract, also for single-particle pixels.
constexpr float kCubeWMin = 1048576.0f, kCubeS1Min = 1048576.0f, kCubeS2Min = 65536.0f;

// table entry: base slot (16 bits) | window length (6 bits) << 16 | first velocity bin (10 bits) << 22
constexpr int kTabBaseBits = 16, kTabLenBits = 6, kTabLenMax = (1 << kTabLenBits) - 1, kTabVloMax = 1023;
static_assert(kCubeSlotsMax < (1 << kTabBaseBits), "slot index must fit the table entry");
static_assert((kQueueCap & (kQueueCap - 1)) == 0 && kQueueCap >= 8 + 32 * kG, "queue: power of two, room for 7 + 32 kG records");
static_assert(kG == 1 || kG == 2 || kG == 4, "a tile holds four particles per lane");

struct CubePlan {                                 // device resident; written by the plan kernels
    int bx0, by0, bw, bh;                         // box origin / size in pixels (bw * bh == 0: nothing privatised)
    int nslots;                                   // voxel slots in use (<= kCubeSlotsMax)
    int level, captured, sampled;                 // diagnostics: threshold level, sample particles inside the windows, sample size
    uint32_t tab[kBoxPix];                        // per box pixel bxi * bh + byi
    double mean_abs[3];                           // sample means of |w|, |w*d|, |w*d*d| (typical term sizes)
    double inv_smem[3];                           // float64 accumulators: 2^-k of the shared-memory sums (flush)
    float sc_smem[3];                             // float64 accumulators: 2^k (see kCubeTypBits)
};

struct CubeParams {
    int64_t n;
    const float *x, *y, *z, *vx, *vy, *vz, *w;    // w may be null (unit weights)
    int nx, ny, nv, npix, G;
    float ax, cx, hx, ay, cy, hy, av, cv, hv;     // bin guess: fma(x, a, c) = (x - x0) * inv - 0.5; sure iff |frac - 0.5| < h
    const float *thr_x, *thr_y, *thr_v;           // exact threshold tables (global memory; rare path)
    float m[9];
    float sc_glob[3];                             // int64 accumulators: 2^k of (w, w*d, w*d*d), |k| <= 120 (also the shared-memory scales)
    void* cube;
    const CubePlan* plan;
    unsigned long long* work;                     // work[0]: chunk counter, work[1]: finished CTAs; both zero between launches (the last CTA resets them)
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

// Exact bin of a value the guess was not sure about: binary search in the threshold table (thr[i] <= x < thr[i+1];
// below thr[0], at or above thr[n] and NaN give -1).  Inline on purpose: a call in the particle loop makes the compiler
// re-load every kernel parameter after it.  ~2 % of the warps take it for some lane.
__device__ __forceinline__ int search_bin(float x, const float* __restrict__ thr, int n) {
    int lo = 0, hi = n;                                  // invariant: thr[lo] <= x < thr[hi]
    const bool inside = x >= __ldg(thr) && x < __ldg(thr + n);
#pragma unroll 1
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= __ldg(thr + mid)) lo = mid; else hi = mid;
    }
    return inside ? lo : -1;
}

// 64-bit fixed-point sums in shared memory as two 32-bit limbs { low (unsigned), high (signed) }, native ATOMS.ADD only.
// The particle body is straight line: a warp's lanes disagree about nearly every condition, so instead of branching a
// lane that has nothing to add sends its (meaningless) term to a private dummy slot.
__device__ __forceinline__ uint32_t smem_add_lo(uint32_t saddr, int a) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(saddr), "r"(a) : "memory");
    return old;
}
// what this thread's addition of the signed term `a` owes the high limb: sign extension + carry out of the low limb
__device__ __forceinline__ int smem_carry(int a, uint32_t old) {
    int h;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.s32 %0, %3, 0;\n\t}" : "=r"(h) : "r"(old), "r"(a), "r"(a >> 31));
    return h;
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
// shared-memory accesses through 32-bit window addresses computed once (a generic pointer makes the compiler rebuild
// the window base from %cluster_ctaid for every access)
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));      // the table is read-only after the prologue: may be moved freely
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
    return v;
}
// Loop-invariant kernel parameters are staged through shared memory and read back with volatile loads: a value that
// comes straight from the parameter bank is re-loaded by ptxas for every particle (a dozen uniform-datapath
// instructions per particle); a value loaded from shared memory has to stay in its register.
__device__ __forceinline__ float lds_f32_pinned(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}

// global-path value of a float32 term: the float itself (float64 accumulators) or rint(v * 2^k) (int64; non-finite -> 0).
// The product with a power of two is exact in float32 as long as it neither overflows (then |v * 2^k| >= 2^127 cannot be
// a legal term: the host sizes 2^k so that n * max|term| * 2^k <= 2^62) nor falls below 2^-126 (where it rounds to 0
// like the exact product does), so this equals rint((double)v * 2^k), the definition of the fixed-point term.
template <int ACC>
__device__ __forceinline__ typename AccT<ACC>::type to_acc_f(float v, float scale) {
    if (ACC == PNM_ACC_F64) return (typename AccT<ACC>::type)v;
    const float s = __fmul_rn(v, scale);
    return (typename AccT<ACC>::type)(is_finite(s) ? __float2ll_rn(s) : 0ll);
}

// h = +-1 carried out of a low limb, as a value of the global accumulator: h * 2^32 (int64) or h * 2^32 * 2^-k (float64).
// 2^32 * 2^-k is a power of two: `hi` is the high word of that double (low word 0, staged in shared memory by the
// prologue), the sign of h goes into its sign bit -- no conversion, no multiplication, no load from the plan.
template <int ACC>
__device__ __forceinline__ typename AccT<ACC>::type carry_value(int h, uint32_t hi) {
    if (ACC == PNM_ACC_F64) return (typename AccT<ACC>::type)__hiloint2double((int)(hi ^ ((uint32_t)h & 0x80000000u)), 0);
    return (typename AccT<ACC>::type)((long long)h << 32);
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
    uint32_t* const s_pm = reinterpret_cast<uint32_t*>(smem_raw + kBoxPix * 4);              // [2][kBoxPix + 32]
    uint4* const s_queue = reinterpret_cast<uint4*>(smem_raw + kBoxPix * 4 + 2 * kPmStride);
    uint32_t* const s_ws = reinterpret_cast<uint32_t*>(smem_raw + kCubeSmemFixed);           // [kCubeSlotsMax + 32]

    const CubePlan* plan = p.plan;
    const int bx0 = plan->bx0, by0 = plan->by0;
    const unsigned bw = (unsigned)plan->bw, bh = (unsigned)plan->bh;
    const int nslots = plan->nslots;
    const int nbox = (int)(bw * bh);
    for (int i = threadIdx.x; i < nbox; i += kCubeThreads) s_tab[i] = plan->tab[i];
    for (int i = threadIdx.x; i < 2 * (kBoxPix + 32); i += kCubeThreads) s_pm[i] = 0u;
    for (int i = threadIdx.x; i < nslots; i += kCubeThreads) s_ws[i] = 0u;
    if (threadIdx.x < 32) s_ws[kCubeSlotsMax + threadIdx.x] = 0u;
    if (threadIdx.x < 3)                                         // carry constants: high word of the double 2^32 * 2^-k
        s_ws[kCubeSlotsMax + 32 + threadIdx.x] = (uint32_t)__double2hiint(4294967296.0 * (ACC == PNM_ACC_F64 ? plan->inv_smem[threadIdx.x] : 1.0));
    if (nbox == 0 && threadIdx.x == 0) s_tab[0] = 0u;           // entry 0 is read by lanes outside the box
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned q_head = 0, q_tail = 0;                   // warp-uniform, free running, in BYTES (16 per record)
    A* const cube = reinterpret_cast<A*>(p.cube);
#if PNM_CUBE_PIN
    // (the queue area is free until the first particle: 32 floats of it stage the parameters)
    const uint32_t a_par = (uint32_t)__cvta_generic_to_shared(s_queue);
    __syncthreads();
    if (threadIdx.x < 9) reinterpret_cast<float*>(s_queue)[threadIdx.x] = p.m[threadIdx.x];
    if (threadIdx.x == 9) {
        float* f = reinterpret_cast<float*>(s_queue) + 9;
        f[0] = p.ax; f[1] = p.cx; f[2] = p.hx; f[3] = p.ay; f[4] = p.cy; f[5] = p.hy; f[6] = p.av; f[7] = p.cv; f[8] = p.hv;
        f[9] = __int_as_float(p.nx); f[10] = __int_as_float(p.ny); f[11] = __int_as_float(p.nv); f[12] = __int_as_float(p.npix);
        f[13] = __int_as_float(2 * p.G);
        f[14] = __int_as_float((int)__cvta_generic_to_shared(smem_raw));
    }
    __syncthreads();
#define PNM_PAR(i, expr) lds_f32_pinned(a_par + 4u * (i))
#else
#define PNM_PAR(i, expr) (expr)
#endif
    const float m0 = PNM_PAR(0, p.m[0]), m1 = PNM_PAR(1, p.m[1]), m2 = PNM_PAR(2, p.m[2]), m3 = PNM_PAR(3, p.m[3]), m4 = PNM_PAR(4, p.m[4]);
    const float m5 = PNM_PAR(5, p.m[5]), m6 = PNM_PAR(6, p.m[6]), m7 = PNM_PAR(7, p.m[7]), m8 = PNM_PAR(8, p.m[8]);
    const float ax = PNM_PAR(9, p.ax), cx = PNM_PAR(10, p.cx), hx = PNM_PAR(11, p.hx), ay = PNM_PAR(12, p.ay), cy = PNM_PAR(13, p.cy);
    const float hy = PNM_PAR(14, p.hy), av = PNM_PAR(15, p.av), cv = PNM_PAR(16, p.cv), hv = PNM_PAR(17, p.hv);
    const unsigned nx = (unsigned)__float_as_int(PNM_PAR(18, __int_as_float(p.nx))), ny = (unsigned)__float_as_int(PNM_PAR(19, __int_as_float(p.ny)));
    const unsigned nv = (unsigned)__float_as_int(PNM_PAR(20, __int_as_float(p.nv)));
    const int npix = __float_as_int(PNM_PAR(21, __int_as_float(p.npix))), G2 = __float_as_int(PNM_PAR(22, __int_as_float(2 * p.G)));
    // shared-memory window addresses (opaque for the same reason: otherwise the window base is rebuilt per access)
    const uint32_t a_smem = (uint32_t)__float_as_int(PNM_PAR(23, __int_as_float((int)__cvta_generic_to_shared(smem_raw))));
#undef PNM_PAR
    const uint32_t a_tab = a_smem, a_pm = a_smem + kBoxPix * 4, a_ws = a_smem + kCubeSmemFixed;
    const uint32_t a_queue = a_smem + kBoxPix * 4 + 2 * kPmStride + (unsigned)warp * (kQueueCap * 16);
    const unsigned pm_dummy = (unsigned)(kBoxPix + lane), ws_dummy = (unsigned)(kCubeSlotsMax + lane);   // this lane's dummy entries
    const uint32_t a_carry = a_ws + 4u * (unsigned)(kCubeSlotsMax + 32);
    unsigned lt_mask;
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
#if PNM_CUBE_PIN
    __syncthreads();                                   // the staging area becomes the queues again
#endif
    const int ql = lane & 3, qj = lane >> 2;
    const float q_scale = p.sc_glob[ql == 0 ? 0 : ql == 1 ? 0 : ql - 1];      // lane 0: w, lane 2: w*d, lane 3: w*d*d
    const float sg0 = p.sc_glob[0];
    // shared-memory scales: the caller's (int64) or the plan's (float64)
    const float sm0 = ACC == PNM_ACC_FIXED ? p.sc_glob[0] : plan->sc_smem[0];
    const float sm1 = ACC == PNM_ACC_FIXED ? p.sc_glob[1] : plan->sc_smem[1];
    const float sm2 = ACC == PNM_ACC_FIXED ? p.sc_glob[2] : plan->sc_smem[2];
    const float kFit = 2147483648.0f;                  // |t * 2^k| < 2^31: rint fits 32 bits

    // eight queued records -> one RED instruction: lane (qj, ql) sends value ql of record qj; the three lanes of a
    // record hit one 32-byte sector, which L2 takes as ONE request
    auto drain8 = [&](unsigned avail) {                // avail: bytes queued, at most 8 records are taken
        const uint4 r = lds_v4(a_queue + ((q_head + 16u * (unsigned)qj) & (kQueueCap * 16 - 1)));
        const int off = (int)r.x;
        const float f = __uint_as_float(ql == 0 ? r.y : ql == 2 ? r.z : r.w);
        red_add_if(16u * (unsigned)qj < avail && ql != 1 && !PNM_CUBE_DBG(4), cube + (ql == 0 ? off : (off & ~3) + ql), to_acc_f<ACC>(f, q_scale));
        q_head += avail < 128u ? avail : 128u;
    };

    // kG particles per lane at a time; warp-synchronous (every lane calls it; a lane without a particle passes
    // x = NaN, which no bin takes).  Straight-line and written in PHASES over the kG particles (all bins, then all
    // table look-ups, then all shared-memory additions, ...): the particles are independent, so every phase is a
    // block of kG independent dependency chains -- with four warps per scheduler the instruction-level parallelism
    // has to come from inside the warp.  Every accumulation is predicated.
    auto group = [&](const float* X, const float* Y, const float* Z, const float* VX, const float* VY, const float* VZ, const float* W) {
        float px[kG], py[kG], d[kG];
        int ix[kG], iy[kG], iv[kG];
        bool sure[kG], all_sure = true;
#pragma unroll
        for (int k = 0; k < kG; ++k) {
            px[k] = rot_row(m0, m1, m2, X[k], Y[k], Z[k]);
            py[k] = rot_row(m3, m4, m5, X[k], Y[k], Z[k]);
            d[k] = rot_row(m6, m7, m8, VX[k], VY[k], VZ[k]);
            sure[k] = true;
            ix[k] = fast_bin(px[k], ax, cx, hx, sure[k]);
            iy[k] = fast_bin(py[k], ay, cy, hy, sure[k]);
            iv[k] = fast_bin(d[k], av, cv, hv, sure[k]);
            all_sure = all_sure && sure[k];
        }
        if (!all_sure) {                                   // within eps of an edge, NaN: the threshold tables decide
#pragma unroll
            for (int k = 0; k < kG; ++k)
                if (!sure[k]) {
                    ix[k] = search_bin(px[k], p.thr_x, (int)nx); iy[k] = search_bin(py[k], p.thr_y, (int)ny);
                    iv[k] = search_bin(d[k], p.thr_v, (int)nv);
                }
        }
        float wd[kG], wdd[kG];
        int rec[kG], pix[kG], a0[kG], a1[kG], a2[kG];
        uint32_t addr0[kG], addr1[kG];
        bool in_s[kG], cap[kG], queued[kG];
#pragma unroll
        for (int k = 0; k < kG; ++k) {
            const float w = HAS_W ? W[k] : 1.0f;
            const bool valid = (unsigned)ix[k] < nx && (unsigned)iy[k] < ny;
            const int ivp = (unsigned)iv[k] < nv ? iv[k] : G2;            // outside the V range: slot 0 of record G
            wd[k] = HAS_W ? __fmul_rn(w, d[k]) : d[k];
            wdd[k] = HAS_W ? __fmul_rn(w, __fmul_rn(d[k], d[k])) : __fmul_rn(d[k], d[k]);
            pix[k] = ix[k] * (int)ny + iy[k];
            rec[k] = ((ivp >> 1) * npix + pix[k]) * 4 + (ivp & 1);
            const unsigned bxi = (unsigned)(ix[k] - bx0), byi = (unsigned)(iy[k] - by0);
            // scaled terms; all three must fit 32 bits (NaN / inf fail), and in float64 mode w must not be tiny
            const float s0 = __fmul_rn(w, sm0), s1 = __fmul_rn(wd[k], sm1), s2 = __fmul_rn(wdd[k], sm2);
            const bool big = ACC == PNM_ACC_FIXED || (fabsf(s0) >= kCubeWMin && fabsf(s1) >= kCubeS1Min && fabsf(s2) >= kCubeS2Min);
            const bool fits = fabsf(s0) < kFit && fabsf(s1) < kFit && fabsf(s2) < kFit && big;
            in_s[k] = valid && fits && bxi < bw && byi < bh;
            queued[k] = valid && !in_s[k];
            const bool st = in_s[k] && !PNM_CUBE_DBG(1);
            const unsigned bp = in_s[k] ? bxi * bh + byi : 0u;
            const uint32_t e = lds_u32(a_tab + 4u * bp);
            a0[k] = __float2int_rn(s0); a1[k] = __float2int_rn(s1); a2[k] = __float2int_rn(s2);
            const unsigned rel = (unsigned)ivp - (e >> 22);
            cap[k] = in_s[k] && rel < ((e >> kTabBaseBits) & (unsigned)kTabLenMax);
            addr1[k] = a_pm + 4u * (st ? bp : pm_dummy);   // (a lane that adds to a dummy entry ignores the carry)
            addr0[k] = a_ws + 4u * (cap[k] && st ? (e & ((1u << kTabBaseBits) - 1u)) + rel : ws_dummy);
        }
        uint32_t old0[kG], old1[kG], old2[kG];
#pragma unroll
        for (int k = 0; k < kG; ++k) {
            old1[k] = smem_add_lo(addr1[k], a1[k]);
            old2[k] = smem_add_lo(addr1[k] + kPmStride, a2[k]);
            old0[k] = smem_add_lo(addr0[k], a0[k]);
        }
        // w of a box pixel outside its window: one single-lane RED
#pragma unroll
        for (int k = 0; k < kG; ++k)
            red_add_if(in_s[k] && !cap[k] && !PNM_CUBE_DBG(2), cube + rec[k], ACC == PNM_ACC_F64 ? (A)(HAS_W ? W[k] : 1.0f) : (A)a0[k]);
        // everything else of this grid: the warp queue
#pragma unroll
        for (int k = 0; k < kG; ++k) {
            const unsigned mask = __ballot_sync(0xffffffffu, queued[k]);
            if (queued[k])
                sts_v4(a_queue + ((q_tail + 16u * (unsigned)__popc(mask & lt_mask)) & (kQueueCap * 16 - 1)),
                       (unsigned)rec[k], __float_as_uint(HAS_W ? W[k] : 1.0f), __float_as_uint(wd[k]), __float_as_uint(wdd[k]));
            q_tail += 16u * (unsigned)__popc(mask);
        }
        // carries of the additions (rare: the scaled terms are ~2^-8 of the limb) -> h * 2^32 to the global accumulator
        int h0[kG], h1[kG], h2[kG], h_any = 0;
#pragma unroll
        for (int k = 0; k < kG; ++k) {
            h1[k] = in_s[k] ? smem_carry(a1[k], old1[k]) : 0; h2[k] = in_s[k] ? smem_carry(a2[k], old2[k]) : 0;
            h0[k] = cap[k] ? smem_carry(a0[k], old0[k]) : 0;
            h_any |= h0[k] | h1[k] | h2[k];
        }
        if (h_any != 0 && !PNM_CUBE_DBG(1)) {
#pragma unroll
            for (int k = 0; k < kG; ++k) {
                A* const rest = cube + ((int64_t)(G2 >> 1) * npix + pix[k]) * 4;
                if (h1[k]) red_add(rest + 2, carry_value<ACC>(h1[k], lds_u32(a_carry + 4u)));
                if (h2[k]) red_add(rest + 3, carry_value<ACC>(h2[k], lds_u32(a_carry + 8u)));
                if (h0[k]) red_add(cube + rec[k], carry_value<ACC>(h0[k], lds_u32(a_carry)));
            }
        }
        if (q_tail - q_head >= 128u) {
            __syncwarp();
#pragma unroll 1
            while (q_tail - q_head >= 128u) drain8(128u);
            __syncwarp();
        }
    };

    // ---- persistent loop: warps take 1024-particle chunks from a global counter (SMs differ in their distance to the
    // hot L2 slices; a static partition leaves the slow ones behind), the next chunk is requested one chunk ahead.
    // Only FULL chunks run through the vector loop; the ragged end (< 1024 + 3 particles) is one scalar chunk.
    const int64_t nfull = p.n / kCubeChunk;
    const bool has_rest = p.n > nfull * kCubeChunk;
    auto grab = [&]() -> int64_t {
        long long c = 0;
        if (lane == 0) c = (long long)atomicAdd(p.work, 1ull);
        return __shfl_sync(0xffffffffu, c, 0);
    };
    const uint64_t pol = evict_first_policy();
    struct Tile { float x[4], y[4], z[4], vx[4], vy[4], vz[4], w[4]; };
    auto load_tile = [&](Tile& t, int64_t chunk, int sub) {
        const int64_t o = ((chunk * kCubeSub + sub) * 32 + lane) * 4;
        ldg4(p.x + o, t.x, pol); ldg4(p.y + o, t.y, pol); ldg4(p.z + o, t.z, pol);
        ldg4(p.vx + o, t.vx, pol); ldg4(p.vy + o, t.vy, pol); ldg4(p.vz + o, t.vz, pol);
        if (HAS_W) ldg4(p.w + o, t.w, pol);
    };
    auto process_tile = [&](const Tile& t) {
#pragma unroll
        for (int k = 0; k < 4; k += kG) group(t.x + k, t.y + k, t.z + k, t.vx + k, t.vy + k, t.vz + k, t.w + k);
    };
    int64_t chunk = grab();
#if PNM_CUBE_PREFETCH
    // software pipeline: the loads of the next tile are in flight while the current one is processed
    int64_t next = chunk < nfull ? grab() : chunk;
    int sub = 0;
    auto advance = [&]() {
        if (++sub == kCubeSub) { sub = 0; chunk = next; if (chunk < nfull) next = grab(); }
    };
    Tile ta, tb;
    if (chunk < nfull) load_tile(ta, chunk, 0);
    while (chunk < nfull) {
        advance();
        if (chunk < nfull) load_tile(tb, chunk, sub);
        process_tile(ta);
        if (!(chunk < nfull)) break;
        advance();
        if (chunk < nfull) load_tile(ta, chunk, sub);
        process_tile(tb);
    }
#else
    while (chunk < nfull) {
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
    // the ragged end: exactly one warp drew the index nfull
    if (has_rest && chunk == nfull) {
#pragma unroll 1
        for (int64_t i0 = nfull * kCubeChunk; i0 < p.n; i0 += 32) {
            const int64_t i = i0 + lane;
            float t[7][kG];
#pragma unroll
            for (int k = 0; k < kG; ++k) {
                t[0][k] = __int_as_float(0x7fc00000); t[1][k] = t[2][k] = t[3][k] = t[4][k] = t[5][k] = 0.f; t[6][k] = 1.f;
            }
            if (i < p.n) {
                t[0][0] = p.x[i]; t[1][0] = p.y[i]; t[2][0] = p.z[i]; t[3][0] = p.vx[i]; t[4][0] = p.vy[i]; t[5][0] = p.vz[i];
                if (HAS_W) t[6][0] = p.w[i];
            }
            group(t[0], t[1], t[2], t[3], t[4], t[5], t[6]);
        }
    }
    __syncwarp();
    if (q_tail != q_head) drain8(q_tail - q_head);

    // ---- flush the shared-memory sums (unsigned low limbs): one RED per non-zero sum
    __syncthreads();
    // every warp of this CTA has drawn its last chunk index: the last CTA to get here leaves the counters at zero
    // for the next launch (no memset between two scatter kernels of a stream)
    if (threadIdx.x == 0 && atomicInc(reinterpret_cast<unsigned*>(p.work + 1), gridDim.x - 1) == gridDim.x - 1) *p.work = 0ull;
    const double i0 = ACC == PNM_ACC_F64 ? plan->inv_smem[0] : 1.0, i1 = ACC == PNM_ACC_F64 ? plan->inv_smem[1] : 1.0;
    const double i2 = ACC == PNM_ACC_F64 ? plan->inv_smem[2] : 1.0;
    for (int bp = threadIdx.x; bp < nbox; bp += kCubeThreads) {
        const int bxi = bp / (int)bh, byi = bp - bxi * (int)bh;
        const int pixel = (bx0 + bxi) * (int)ny + (by0 + byi);
        A* const rest = cube + ((int64_t)p.G * npix + pixel) * 4;
        const uint32_t s1 = s_pm[bp], s2 = s_pm[bp + kBoxPix + 32];
        if (s1) red_add(rest + 2, ACC == PNM_ACC_F64 ? (A)((double)s1 * i1) : (A)s1);
        if (s2) red_add(rest + 3, ACC == PNM_ACC_F64 ? (A)((double)s2 * i2) : (A)s2);
        const uint32_t e = s_tab[bp];
        const int base = (int)(e & ((1u << kTabBaseBits) - 1u)), len = (int)((e >> kTabBaseBits) & (unsigned)kTabLenMax), vlo = (int)(e >> 22);
        for (int r = 0; r < len; ++r) {
            const uint32_t sw = s_ws[base + r];
            const int v = vlo + r;
            if (sw) red_add(cube + ((int64_t)(v >> 1) * npix + pixel) * 4 + (v & 1), ACC == PNM_ACC_F64 ? (A)((double)sw * i0) : (A)sw);
        }
    }
}

}  // namespace pnm
