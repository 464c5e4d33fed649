// pnm_bin.cuh -- fused rotate + project + bin kernel (K1+K2+K3 of SURVEY.md 2.2).
//
// One pass over the SoA particle arrays produces, per viewing angle, the (opaque) accumulators
//   maps  [view][replica][iy][ix][4] = { sum w*d, sum w*d*d, sum w (or its out-of-V remainder), - }
//   cube  [view][iv][ix][iy]         =   sum w
// which pnm_finalize_* turn into the reference's layouts.  Reference semantics:
// rotation.py:115-117 (rotation), grid.py:121-131 / :200-201 (2-D moment histograms),
// grid.py:252-259 (3-D histogram).
//
// Roofline: HBM (28 B per float32 particle).  What actually bounds the kernel is the rate at which L2
// accepts reduction REQUESTS: one per distinct 32-byte sector of a RED instruction (~1.9e11/s chip-wide,
// scripts/red_pair_bench.cu), shared with the particle stream.  The design therefore minimises requests
// per particle, keeps the instruction stream short and spreads same-address traffic:
//   * accumulators (<= 66 MB for 128x128x256) stay resident in the 126 MB L2; the particle stream
//     is loaded with 256-bit LDG, no L1 allocation and an L2 evict-first policy;
//   * maps are replicated R times (one replica per warp id) so that the galaxy-centre pixels, which
//     receive ~2 % of all particles, do not serialise on one L2 sector;
//   * the cube is stored velocity-major, which spreads the hot central spaxel over ~nv L2 lines
//     instead of one 2 KB column;
//   * consecutive particles of a thread that fall in the same pixel are merged in registers before
//     the reduction (free for random order, ~1 per run for space-filling-curve ordered snapshots);
//   * with PNM_W_FROM_CUBE the map's sum w costs no RED at all for particles inside the V range;
//   * LANE PAIRS: (sum w*d, sum w*d*d) of a pixel sit in one sector, so lanes 2j / 2j+1 swap one value
//     (3 SHFL) and every RED instruction carries BOTH moments of one pixel in adjacent lanes: two
//     instructions as before, half the L2 requests (MapRun::flush_paired).  Per particle: one request
//     for the cube voxel, one for the pixel;
//   * TMA BULK REDUCTIONS for maps without a cube: a 32-byte record (sum w*d, sum w*d*d, sum w, -) staged
//     in shared memory goes out as ONE cp.reduce.async.bulk.global.shared::cta ... add.f64/.u64 (SASS
//     UBLKRED) and also carries sum w; 2 of a tile's 8 updates take this path (the TMA unit sustains
//     ~7.5e10 such operations/s).  With the cube the paired REDs leave it nothing to save;
//   * bin lookup: the uniform guess floor((x-x0)*inv) is PROVABLY the numpy bin whenever its
//     fractional part is further than `eps` from an integer (eps bounds the float rounding of the
//     guess plus the deviation of np.linspace edges from an affine grid, computed at grid
//     creation); only the ~2*eps fraction of near-edge particles consults the edge table.
#pragma once
#ifndef PNM_PAIRED
#define PNM_PAIRED 1     // lane pairs share the (S1, S2) reductions of a pixel (MapRun::flush_paired); 0: one lane, one pixel
#endif
#include "pnm_common.cuh"

namespace pnm {

struct BinParams {
    int64_t n;
    const void* x; const void* y; const void* z;     // FUSED: positions; PLAIN: x, y, (z unused)
    const void* vx; const void* vy; const void* vz;  // FUSED: velocities; PLAIN: vx = data d
    const void* w;                                    // weights or nullptr
    const uint8_t* mask;
    int mask_mode;
    int nx, ny, nv;
    const void* thr_x; const void* thr_y; const void* thr_v;  // device tables, dtype of particles
    double x0, inv_dx, y0, inv_dy, v0, inv_dv;        // uniform-guess hints
    double eps_x, eps_y, eps_v;                       // guard band of the guess, in bins (>= 0.5: table only)
    int sums;
    int nchain, nviews;
    int replicas;                                     // power of two
    int rep_shift;                                    // replica = (thread id >> rep_shift) & (replicas-1)
    int tma_slots;                                    // map updates per tile sent through the TMA unit (0..kTmaSlots)
    double scale[3];
    void* acc_maps; void* acc_cube;
    int64_t maps_stride, cube_stride;                 // 8-byte elements per view (maps: all replicas)
    float mats[kMaxMats * 9];                         // [view][chain][9]
};

// ---- 256-bit streaming loads (sm_100: LDG.E.256): read once -> no L1 allocation, L2 evict-first,
// so the 2.8 GB stream does not push the L2-resident accumulators out.
template <typename T> struct Pack;
template <> struct Pack<float> {
    static constexpr int N = 8;
    __device__ __forceinline__ static void load(const float* p, float* o) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]), "=f"(o[4]), "=f"(o[5]), "=f"(o[6]), "=f"(o[7])
                     : "l"(p));
    }
};
template <> struct Pack<double> {
    static constexpr int N = 4;
    __device__ __forceinline__ static void load(const double* p, double* o) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(o[0]), "=d"(o[1]), "=d"(o[2]), "=d"(o[3]) : "l"(p));
    }
};

// ---- bin lookup --------------------------------------------------------------------------------
template <typename T> struct Axis { T x0, inv, eps, one_m_eps; int n; const T* thr; };

// exact path: binary search in the table of thresholds (thr[i] <= x < thr[i+1], see pnm_grid_create),
// independent of the guess, so it is exact for any monotone edges.  Out of line: it is rare.
template <typename T>
__device__ __noinline__ int exact_bin(T x, const T* thr, int n) {
    if (!(x >= thr[0]) || !(x < thr[n])) return -1;     // below, above (closed last edge folded into thr[n]), NaN
    int lo = 0, hi = n;                                  // invariant: thr[lo] <= x < thr[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= thr[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ float floor_t(float v) { return floorf(v); }
__device__ __forceinline__ double floor_t(double v) { return floor(v); }

// uniform guess: floor((x - x0) * inv).  `sure` is cleared when the fractional part is within eps
// of an integer (x within eps bins of an edge, first and last included), for NaN (every comparison
// fails) and for |g| >= 2^23 (2^52), where the fraction is 0; the exact table search then decides.
// Out-of-range guesses (negative, >= n; the conversion saturates) are dropped by the caller's
// unsigned range check.  7 instructions per axis.
template <typename T>
__device__ __forceinline__ int guess_bin(T x, const Axis<T>& a, bool& sure) {
    const T g = (x - a.x0) * a.inv;
    const T f = floor_t(g);
    const T r = g - f;
    sure = sure && (r > a.eps) && (r < a.one_m_eps);
    return (int)f;
}

// ---- accumulator arithmetic --------------------------------------------------------------------
template <int ACC> struct AccT { using type = double; };
template <> struct AccT<PNM_ACC_FIXED> { using type = long long; };

// FIXED: rint(value * 2^k), the product formed in float64 (exact: a power of two).  Non-finite values quantise to 0 (documented deviation: the
// float64 path propagates NaN like the reference does).
template <int ACC>
__device__ __forceinline__ typename AccT<ACC>::type to_acc(float v, double scale) {
    if (ACC == PNM_ACC_F64) return (typename AccT<ACC>::type)v;
    const double s = (double)v * scale;            // (float)scale would overflow for 2^k with k > 127
    return (typename AccT<ACC>::type)(is_finite(s) ? __double2ll_rn(s) : 0ll);
}
template <int ACC>
__device__ __forceinline__ typename AccT<ACC>::type to_acc(double v, double scale) {
    if (ACC == PNM_ACC_F64) return (typename AccT<ACC>::type)v;
    const double s = v * scale;
    return (typename AccT<ACC>::type)(is_finite(s) ? __double2ll_rn(s) : 0ll);
}

// result unused -> REDG (RED.E.ADD.F64.RN / RED.E.ADD.64), no return trip
__device__ __forceinline__ void red_add(double* p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(long long* p, long long v) {
    atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}

// ---- TMA staging of (S1, S2) pairs ---------------------------------------------------------------
// Shared-memory layout per CTA: kTmaStages x tma_slots x blockDim.x records of 32 bytes
// (S1, S2, S0, pad; consecutive threads, consecutive records) + the matching destination offsets.  A thread
// only ever touches its own records, so ordering needs no CTA barrier: generic-proxy writes ->
// fence.proxy.async -> bulk reductions -> commit_group; before a stage is rewritten two tiles
// later, wait_group.read 1 guarantees the TMA unit has read it.  BinParams::tma_slots of the (up to) 8
// map updates of a tile go this way; the rest take the (lane-paired) LSU path.
constexpr int kTmaSlots = 6;      // upper bound; the launcher picks how many are used (BinParams::tma_slots)
constexpr int kTmaStages = 2;

__device__ __forceinline__ void tma_fence_writes() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// 16 bytes = (S1, S2); 32 bytes = (S1, S2, S0, pad) when the record also carries sum w
__device__ __forceinline__ void tma_reduce(double* dst, const void* src_smem, int bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                 :: "l"(dst), "r"((uint32_t)__cvta_generic_to_shared(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_reduce(long long* dst, const void* src_smem, int bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;"
                 :: "l"(dst), "r"((uint32_t)__cvta_generic_to_shared(src_smem)), "r"(bytes) : "memory");
}

template <int ACC>
struct TmaStage {
    using A = typename AccT<ACC>::type;
    A* vals;        // this thread's record 0 of the current stage (record k at vals + 4*k*stride)
    int* offs;      // destination offsets (negated - 1 when the record carries S0), offs + k*stride
    int stride;     // blockDim.x
    int used, cap;  // records filled in this tile, capacity (0 disables the TMA path)

    __device__ __forceinline__ bool push(int off, A s1, A s2, A s0, bool with0) {
        if (used >= cap) return false;
        A* rec = vals + 4 * used * stride;
        rec[0] = s1; rec[1] = s2;                       // STS.128
        if (with0) { rec[2] = s0; rec[3] = 0; }
        offs[used * stride] = with0 ? -off - 1 : off;
        ++used;
        return true;
    }
    // end of tile: hand the records to the TMA unit
    __device__ __forceinline__ void drain(A* maps) {
        if (cap == 0) return;
        if (used > 0) {
            tma_fence_writes();
            for (int k = 0; k < used; ++k) {
                const int o = offs[k * stride];
                tma_reduce(maps + (o < 0 ? -o - 1 : o), vals + 4 * k * stride, o < 0 ? 32 : 16);
            }
        }
        tma_commit();                                   // (possibly empty) group: keeps the group count in step
        used = 0;
    }
};

// ---- lane quads ------------------------------------------------------------------------------------
// Four lanes hold four records (off, a0..a3) whose values belong to ONE 32-byte sector each.  Two butterfly stages
// transpose the 4 x 4 values, so that lane q of the quad holds value #q of all four records, and four REDs follow,
// each carrying one record in the quad's four lanes: L2 sees one request per record (red_pair_bench: one request
// per distinct sector of an instruction).  off < 0: no record.  `lanes`: bit q set = value #q is wanted.
// SLOT0_IN_OFF: value #0 goes to slot (off & 3) instead of slot 0 (the record itself starts at off & ~3).
template <typename A, bool SLOT0_IN_OFF = false>
__device__ __forceinline__ void quad_reduce(A* base, int off, A a0, A a1, A a2, A a3, int lane, int lanes) {
    const int q = lane & 3;
    {   // stage 1: partner lane ^ 1 -- even lanes keep (0, 2) and receive (1, 3), odd lanes the other way round
        const bool odd = q & 1;
        A x = odd ? a0 : a1, y = odd ? a2 : a3;
        x = __shfl_xor_sync(0xffffffffu, x, 1); y = __shfl_xor_sync(0xffffffffu, y, 1);
        if (odd) { a0 = x; a2 = y; } else { a1 = x; a3 = y; }
    }
    {   // stage 2: partner lane ^ 2
        const bool hi = q & 2;
        A x = hi ? a0 : a2, y = hi ? a1 : a3;
        x = __shfl_xor_sync(0xffffffffu, x, 2); y = __shfl_xor_sync(0xffffffffu, y, 2);
        if (hi) { a0 = x; a1 = y; } else { a2 = x; a3 = y; }
    }
    const int qb = lane & ~3;
    const int o0 = __shfl_sync(0xffffffffu, off, qb), o1 = __shfl_sync(0xffffffffu, off, qb + 1);
    const int o2 = __shfl_sync(0xffffffffu, off, qb + 2), o3 = __shfl_sync(0xffffffffu, off, qb + 3);
    if ((lanes >> q) & 1) {
        if (SLOT0_IN_OFF) {
            const int keep = q == 0 ? ~0 : ~3;          // lane 0: the slot in the low bits; lanes 2, 3: slots 2, 3
            if (o0 >= 0) red_add(base + (o0 & keep) + q, a0);
            if (o1 >= 0) red_add(base + (o1 & keep) + q, a1);
            if (o2 >= 0) red_add(base + (o2 & keep) + q, a2);
            if (o3 >= 0) red_add(base + (o3 & keep) + q, a3);
        } else {
            if (o0 >= 0) red_add(base + o0 + q, a0);
            if (o1 >= 0) red_add(base + o1 + q, a1);
            if (o2 >= 0) red_add(base + o2 + q, a2);
            if (o3 >= 0) red_add(base + o3 + q, a3);
        }
    }
}

// pending run of map contributions that share one pixel (merged in registers)
template <int ACC>
struct MapRun {
    using A = typename AccT<ACC>::type;
    int off = -1;
    A s0 = 0, s1 = 0, s2 = 0;
    bool has0 = false;

    __device__ __forceinline__ void flush(A* maps, int sums, TmaStage<ACC>& st) {
        if (off < 0) return;
        const bool with0 = (sums & PNM_SUM_W) && has0;
        const bool both = (sums & PNM_SUM_WD) && (sums & PNM_SUM_WD2);
        if (!(both && st.push(off, s1, s2, s0, with0))) {
            if (with0) red_add(maps + off + kSlotS0, s0);
            if (sums & PNM_SUM_WD) red_add(maps + off + kSlotS1, s1);
            if (sums & PNM_SUM_WD2) red_add(maps + off + kSlotS2, s2);
        }
        off = -1;
    }
    // Warp-synchronous flush (all 32 lanes call it together, a lane without a pending run passes off < 0).
    // L2 takes one request per distinct 32-byte sector of a RED instruction, not per lane
    // (scripts/red_pair_bench.cu: 1.8e11 lane-reductions/s with one lane per sector, 3.6e11 with two), and the
    // (S1, S2) slots of a pixel share a sector.  Lanes 2j and 2j+1 therefore swap one value: the first RED of
    // the pair carries S1 and S2 of the even lane's pixel, the second those of the odd lane's pixel -- two
    // instructions as before, half the requests.
    __device__ __forceinline__ void flush_paired(A* maps, int sums, TmaStage<ACC>& st, int lane) {
        const bool with0 = (sums & PNM_SUM_W) && has0 && off >= 0;
        bool valid = off >= 0;
        const bool both = (sums & PNM_SUM_WD) && (sums & PNM_SUM_WD2);  // else the unused slot just receives zeros / unread sums
        if (valid && both && st.push(off, s1, s2, s0, with0)) valid = false;
        else if (with0) red_add(maps + off + kSlotS0, s0);
        const bool odd = lane & 1;
        const int my_off = valid ? off : -1;
        const int x_off = __shfl_xor_sync(0xffffffffu, my_off, 1);
        const A x_val = __shfl_xor_sync(0xffffffffu, odd ? s1 : s2, 1);     // the partner issues this lane's other half
        {   // the even lane's pixel
            const int o = odd ? x_off : my_off;
            if (o >= 0) red_add(maps + o + (odd ? kSlotS2 : kSlotS1), odd ? x_val : s1);
        }
        {   // the odd lane's pixel
            const int o = odd ? my_off : x_off;
            if (o >= 0) red_add(maps + o + (odd ? kSlotS2 : kSlotS1), odd ? s2 : x_val);
        }
        off = -1;
    }
    // Same idea for runs that carry sum w with every pixel (no cube to recover it from): the three sums of a
    // pixel share a sector, so a quad of lanes transposes its 4 x (S1, S2, S0, -) values (two butterfly stages)
    // and issues four REDs, each carrying all three sums of ONE pixel in lanes q = 0, 1, 2 of the quad (lane 3
    // idles): one L2 request per pixel instead of two or three.  Slot index == lane-in-quad by construction.
    __device__ __forceinline__ void flush_quad(A* maps, int sums, int lane) {
        static_assert(kSlotS1 == 0 && kSlotS2 == 1 && kSlotS0 == 2, "slot order is the lane order of the quad");
        const int lanes = ((sums & PNM_SUM_WD) ? 1 : 0) | ((sums & PNM_SUM_WD2) ? 2 : 0) | ((sums & PNM_SUM_W) ? 4 : 0);
        quad_reduce<A>(maps, off, s1, s2, has0 ? s0 : A(0), A(0), lane, lanes);
        off = -1;
    }
    // add one contribution; returns true when the pixel changed, i.e. the previous run (moved to `prev`) is due
    __device__ __forceinline__ bool add_deferred(MapRun& prev, int o, bool use0, A a0, A a1, A a2) {
        bool due = false;
        if (o != off) {
            prev = *this; due = prev.off >= 0;
            off = o; s0 = 0; s1 = a1; s2 = a2; has0 = false;
        } else {
            s1 += a1; s2 += a2;
        }
        if (use0) { s0 += a0; has0 = true; }
        return due;
    }
    __device__ __forceinline__ void add(A* maps, int sums, TmaStage<ACC>& st, int o, bool use0, A a0, A a1, A a2) {
        if (o != off) {
            flush(maps, sums, st);
            off = o; s0 = 0; s1 = a1; s2 = a2; has0 = false;
        } else {
            s1 += a1; s2 += a2;
        }
        if (use0) { s0 += a0; has0 = true; }
    }
};

// one particle's record in the interleaved cube (PNM_CUBE_MOMENTS)
template <int ACC>
struct CubeRec {
    using A = typename AccT<ACC>::type;
    int off = -1;          // record offset | (iv & 1): which of the record's two velocity bins takes w
    A w = 0, s1 = 0, s2 = 0;
};

// ---- per-thread binning context: everything loop invariant sits in registers -------------------
// SUMS >= 0: the set of sums is a compile-time constant (hot configurations); -1: read at run time.
template <typename T, int ACC, int SUMS>
struct Binner {
    using A = typename AccT<ACC>::type;
    Axis<T> ax, ay, av;
    int rsums, nx, ny, npix, mask_mode;
    double sc0, sc1, sc2;

    __device__ __forceinline__ Binner(const BinParams& p, const T* sx, const T* sy, const T* sv) {
        ax = {(T)p.x0, (T)p.inv_dx, (T)p.eps_x, (T)1 - (T)p.eps_x, p.nx, sx};
        ay = {(T)p.y0, (T)p.inv_dy, (T)p.eps_y, (T)1 - (T)p.eps_y, p.ny, sy};
        av = {(T)p.v0, (T)p.inv_dv, (T)p.eps_v, (T)1 - (T)p.eps_v, p.nv, sv};
        rsums = p.sums; nx = p.nx; ny = p.ny; npix = p.nx * p.ny; mask_mode = p.mask_mode;
        sc0 = p.scale[0]; sc1 = p.scale[1]; sc2 = p.scale[2];
    }
    __device__ __forceinline__ int sums() const { return SUMS >= 0 ? SUMS : rsums; }

    // w*d and w*(d*d), each product rounded in the particle dtype (grid.py:128, :131) -- or in float32 when the
    // values are float32 carried as float64 (PNM_VAL_D32 / PNM_VAL_W32: numpy's result types)
    __device__ __forceinline__ static void weighted_terms(int sm, T w, T d, T& wd, T& wdd) {
        T dd = mul_rn(d, d);
        if (sizeof(T) == 8 && (sm & PNM_VAL_D32)) {
            dd = (T)__fmul_rn((float)d, (float)d);
            if (sm & PNM_VAL_W32) {
                wd = (T)__fmul_rn((float)w, (float)d);
                wdd = (T)__fmul_rn((float)w, (float)dd);
                return;
            }
        }
        wd = mul_rn(w, d); wdd = mul_rn(w, dd);
    }

    // (px, py, d) already projected: locate the bins, then reduce
    __device__ __forceinline__ void scatter(A* maps, A* cube, MapRun<ACC>& run, TmaStage<ACC>& st,
                                            T px, T py, T d, T w, bool masked_in,
                                            MapRun<ACC>* prev = nullptr, bool* due = nullptr, CubeRec<ACC>* mom = nullptr) {
        if (masked_in) {
            if (mask_mode == PNM_MASK_PLAIN) return;
            if (mask_mode == PNM_MASK_AND_NONFINITE && !is_finite(d)) return;
        }
        const int sm = sums();
        const bool want_v = (sm & PNM_SUM_CUBE) != 0;
        bool sure = true;
        int ix = guess_bin<T>(px, ax, sure);
        int iy = guess_bin<T>(py, ay, sure);
        int iv = want_v ? guess_bin<T>(d, av, sure) : -1;
        if (!sure) {                                     // ~1e-3 of the particles
            ix = exact_bin<T>(px, ax.thr, ax.n);
            iy = exact_bin<T>(py, ay.thr, ay.n);
            if (want_v) iv = exact_bin<T>(d, av.thr, av.n);
        }
        if ((unsigned)ix >= (unsigned)nx || (unsigned)iy >= (unsigned)ny) return;
        const bool in_v = want_v && (unsigned)iv < (unsigned)av.n;
        if (sm & PNM_CUBE_MOMENTS) {
            // interleaved cube: record (iv / 2, pixel) = { w of the even bin, w of the odd bin, sum w*d, sum w*d*d };
            // a particle outside the V range goes to slot 0 of the extra record slab G = ceil(nv / 2).  Every particle
            // touches ONE sector and the maps accumulator is not used at all.
            T wd, wdd;
            weighted_terms(sm, w, d, wd, wdd);
            const A aw = to_acc<ACC>(w, sc0), a1 = to_acc<ACC>(wd, sc1), a2 = to_acc<ACC>(wdd, sc2);
            const int ivp = in_v ? iv : 2 * ((av.n + 1) >> 1);
            const int coff = ((ivp >> 1) * npix + ix * ny + iy) * 4;
            if (PNM_PAIRED && mom) {
                mom->off = coff | (ivp & 1); mom->w = aw; mom->s1 = a1; mom->s2 = a2;
            } else {
                red_add(cube + coff + (ivp & 1), aw); red_add(cube + coff + 2, a1); red_add(cube + coff + 3, a2);
            }
            return;
        }
        if (sm & (PNM_SUM_W | PNM_SUM_WD | PNM_SUM_WD2)) {
            const bool use0 = (sm & PNM_SUM_W) && !((sm & PNM_W_FROM_CUBE) && in_v);
            T wd, wdd;
            weighted_terms(sm, w, d, wd, wdd);
            const int off = (iy * nx + ix) * PNM_MAP_SLOTS;
            if (PNM_PAIRED && due)                              // paired mode: a finished run goes back to the caller
                *due = run.add_deferred(*prev, off, use0, to_acc<ACC>(w, sc0), to_acc<ACC>(wd, sc1), to_acc<ACC>(wdd, sc2));
            else
                run.add(maps, sm, st, off, use0, to_acc<ACC>(w, sc0), to_acc<ACC>(wd, sc1), to_acc<ACC>(wdd, sc2));
        }
        if (in_v) red_add(cube + (iv * npix + ix * ny + iy), to_acc<ACC>(w, sc0));
    }
};

// rotation chain of one view (general case: matrices read from the parameter bank)
template <typename T>
__device__ __forceinline__ void project_chain(const float* m, int nchain, T x, T y, T z, T vx, T vy, T vz,
                                              T& px, T& py, T& d) {
    T pz = z, qx = vx, qy = vy, qz = vz;
    px = x; py = y;
    for (int c = 0; c < nchain - 1; ++c, m += 9) {   // compounded rotations keep all rows
        const T ax = rot_row(m[0], m[1], m[2], px, py, pz), ay = rot_row(m[3], m[4], m[5], px, py, pz),
                az = rot_row(m[6], m[7], m[8], px, py, pz);
        const T bx = rot_row(m[0], m[1], m[2], qx, qy, qz), by = rot_row(m[3], m[4], m[5], qx, qy, qz),
                bz = rot_row(m[6], m[7], m[8], qx, qy, qz);
        px = ax; py = ay; pz = az; qx = bx; qy = by; qz = bz;
    }
    const T ax = rot_row(m[0], m[1], m[2], px, py, pz);   // last step: only sky x, sky y and v_los are needed
    const T ay = rot_row(m[3], m[4], m[5], px, py, pz);
    d = rot_row(m[6], m[7], m[8], qx, qy, qz);
    px = ax; py = ay;
}

// Grid: far more CTAs than resident slots (kCtasPerSM per SM, 2 resident).  Measured on 10^8
// particles: 1.98 ms with a grid of exactly the resident CTAs, 1.69 ms with 32 per SM -- SMs differ
// in their distance to the L2 slices that hold the hot pixels, and the block scheduler balances
// what a static grid-stride partition cannot.
#ifndef PNM_MINBLOCKS
#define PNM_MINBLOCKS 2
#endif
constexpr int kCtasPerSM = 32;
constexpr int kBinThreads = 256;

// dynamic shared memory of k_project_bin: edge tables, then (SIMPLE only) the TMA staging area
template <typename T>
__host__ __device__ inline size_t bin_table_bytes(int nx, int ny, int nv) {
    return (((size_t)(nx + 1) + (ny + 1) + (nv + 1)) * sizeof(T) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t bin_stage_bytes(int threads, int slots) {
    return (size_t)kTmaStages * slots * threads * (32 + 4);
}

// ---- kernel: persistent grid-stride loop; each thread takes Pack<T>::N consecutive particles per
// array with one 256-bit load (VECLOAD), scalar path for the tail and for unaligned pointers.
// SIMPLE = one view, one matrix, no mask: the matrix lives in registers, consecutive particles of a
// thread merge their map contributions, lane pairs share the map reductions and the TMA path is available. --
template <typename T, bool FUSED, int ACC, bool VECLOAD, bool SIMPLE, int SUMS>
__global__ void __launch_bounds__(kBinThreads, PNM_MINBLOCKS) k_project_bin(const __grid_constant__ BinParams p) {
    using A = typename AccT<ACC>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sx = reinterpret_cast<T*>(smem_raw);
    T* sy = sx + (p.nx + 1);
    T* sv = sy + (p.ny + 1);
    for (int i = threadIdx.x; i <= p.nx; i += blockDim.x) sx[i] = static_cast<const T*>(p.thr_x)[i];
    for (int i = threadIdx.x; i <= p.ny; i += blockDim.x) sy[i] = static_cast<const T*>(p.thr_y)[i];
    if (p.sums & PNM_SUM_CUBE)
        for (int i = threadIdx.x; i <= p.nv; i += blockDim.x) sv[i] = static_cast<const T*>(p.thr_v)[i];
    __syncthreads();

    constexpr int VN = Pack<T>::N;
    const T* X = static_cast<const T*>(p.x);  const T* Y = static_cast<const T*>(p.y);
    const T* Z = static_cast<const T*>(p.z);  const T* VX = static_cast<const T*>(p.vx);
    const T* VY = static_cast<const T*>(p.vy); const T* VZ = static_cast<const T*>(p.vz);
    const T* W = static_cast<const T*>(p.w);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    // replica by warp id: lanes of a warp stay together, neighbouring warps spread out
    const int rep = (int)((tid >> p.rep_shift) & (p.replicas - 1));
    A* const maps0 = reinterpret_cast<A*>(p.acc_maps) + (int64_t)rep * p.nx * p.ny * PNM_MAP_SLOTS;
    A* const cube0 = reinterpret_cast<A*>(p.acc_cube);
    Binner<T, ACC, SUMS> B(p, sx, sy, sv);
    float m[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = p.mats[k];

    // TMA staging (SIMPLE + vector path only): [stage][slot][thread] records, then the offsets
    unsigned char* stage_raw = smem_raw + bin_table_bytes<T>(p.nx, p.ny, p.nv);
    A* const stage_vals = reinterpret_cast<A*>(stage_raw);
    int* const stage_offs = reinterpret_cast<int*>(stage_raw + (size_t)kTmaStages * p.tma_slots * blockDim.x * 32);
    TmaStage<ACC> st;
    st.stride = blockDim.x; st.used = 0;
    st.cap = (SIMPLE && VECLOAD) ? p.tma_slots : 0;
    st.vals = stage_vals + 4 * threadIdx.x;
    st.offs = stage_offs + threadIdx.x;

    auto one = [&](int64_t idx, T x, T y, T z, T vx, T vy, T vz, T w, MapRun<ACC>& run) {
        if (SIMPLE) {
            T px = x, py = y, d = vx;
            if (FUSED) {
                px = rot_row(m[0], m[1], m[2], x, y, z);
                py = rot_row(m[3], m[4], m[5], x, y, z);
                d = rot_row(m[6], m[7], m[8], vx, vy, vz);
            }
            B.scatter(maps0, cube0, run, st, px, py, d, w, false);
        } else {
            const bool masked = p.mask_mode != PNM_MASK_NONE && p.mask[idx] != 0;
            for (int view = 0; view < p.nviews; ++view) {
                T px = x, py = y, d = vx;
                if (FUSED) project_chain<T>(p.mats + (size_t)view * p.nchain * 9, p.nchain, x, y, z, vx, vy, vz, px, py, d);
                MapRun<ACC> vr;       // no cross-particle merging in the general path
                A* maps = maps0 + view * p.maps_stride;
                B.scatter(maps, cube0 + view * p.cube_stride, vr, st, px, py, d, w, masked);
                vr.flush(maps, B.sums(), st);
            }
        }
    };

    int64_t done = 0;
    if (VECLOAD) {
        const int64_t nvec = p.n / VN;
        int parity = 0;
        // Rotated loop: the loads of the NEXT tile are issued after the current tile's particles have been
        // consumed (their registers are free again) and BEFORE the TMA hand-off, so that the load latency runs
        // under the bulk-reduction burst and the group wait instead of after them.
        T x[VN], y[VN], z[VN], vx[VN], vy[VN], vz[VN], w[VN];
        auto load_tile = [&](int64_t i) {
            Pack<T>::load(X + i * VN, x);
            Pack<T>::load(Y + i * VN, y);
            Pack<T>::load(VX + i * VN, vx);
            if (FUSED) {
                Pack<T>::load(Z + i * VN, z);
                Pack<T>::load(VY + i * VN, vy);
                Pack<T>::load(VZ + i * VN, vz);
            }
            if (W) Pack<T>::load(W + i * VN, w);
        };
        if (PNM_PAIRED && SIMPLE) {
            // warp-uniform trip count: the paired flush is warp-synchronous; lanes past the end pass empty runs
            const int lane = threadIdx.x & 31;
            // sum w travels with every pixel (no cube to recover it from) -> quads, else pairs
            const bool quad = (B.sums() & PNM_SUM_W) && !(B.sums() & PNM_W_FROM_CUBE);
            int64_t ib = tid - lane;
            bool live = ib + lane < nvec;
            if (live) load_tile(ib + lane);
            while (ib < nvec) {
                if (st.cap) {
                    tma_wait_read_1();
                    st.vals = stage_vals + ((size_t)parity * p.tma_slots * blockDim.x + threadIdx.x) * 4;
                    st.offs = stage_offs + (size_t)parity * p.tma_slots * blockDim.x + threadIdx.x;
                    parity ^= 1;
                }
                MapRun<ACC> run;
                if (B.sums() & PNM_CUBE_MOMENTS) {
#pragma unroll
                    for (int k = 0; k < VN; ++k) {
                        CubeRec<ACC> rec;
                        if (live) {
                            T px = x[k], py = y[k], d = vx[k];
                            if (FUSED) {
                                px = rot_row(m[0], m[1], m[2], x[k], y[k], z[k]);
                                py = rot_row(m[3], m[4], m[5], x[k], y[k], z[k]);
                                d = rot_row(m[6], m[7], m[8], vx[k], vy[k], vz[k]);
                            }
                            B.scatter(maps0, cube0, run, st, px, py, d, W ? w[k] : (T)1, false, nullptr, nullptr, &rec);
                        }
                        // three lanes per particle: lane 0 sends w to its velocity bin, lanes 2 and 3 the two moments
                        quad_reduce<A, true>(cube0, rec.off, rec.w, A(0), rec.s1, rec.s2, lane, 13);
                    }
                } else
#pragma unroll
                for (int k = 0; k < VN; ++k) {
                    MapRun<ACC> prev;
                    bool due = false;
                    if (live) {
                        T px = x[k], py = y[k], d = vx[k];
                        if (FUSED) {
                            px = rot_row(m[0], m[1], m[2], x[k], y[k], z[k]);
                            py = rot_row(m[3], m[4], m[5], x[k], y[k], z[k]);
                            d = rot_row(m[6], m[7], m[8], vx[k], vy[k], vz[k]);
                        }
                        B.scatter(maps0, cube0, run, st, px, py, d, W ? w[k] : (T)1, false, &prev, &due);
                    }
                    if (!due) prev.off = -1;
                    if (quad) prev.flush_quad(maps0, B.sums(), lane); else prev.flush_paired(maps0, B.sums(), st, lane);
                }
                if (quad) run.flush_quad(maps0, B.sums(), lane); else run.flush_paired(maps0, B.sums(), st, lane);
                ib += nthreads;
                live = ib + lane < nvec;
                if (live) load_tile(ib + lane);
                st.drain(maps0);
            }
        }
        int64_t i = (PNM_PAIRED && SIMPLE) ? nvec : tid;
        if (i < nvec) load_tile(i);
        while (i < nvec) {
            if (st.cap) {                                   // the stage used two tiles ago must have been read
                tma_wait_read_1();
                st.vals = stage_vals + ((size_t)parity * p.tma_slots * blockDim.x + threadIdx.x) * 4;
                st.offs = stage_offs + (size_t)parity * p.tma_slots * blockDim.x + threadIdx.x;
                parity ^= 1;
            }
            MapRun<ACC> run;
#pragma unroll
            for (int k = 0; k < VN; ++k)
                one(i * VN + k, x[k], y[k], FUSED ? z[k] : (T)0, vx[k], FUSED ? vy[k] : (T)0, FUSED ? vz[k] : (T)0,
                    W ? w[k] : (T)1, run);
            run.flush(maps0, B.sums(), st);
            i += nthreads;
            if (i < nvec) load_tile(i);
            st.drain(maps0);
        }
        done = nvec * VN;
        if (st.cap) { tma_wait_all(); st.cap = 0; }         // every bulk reduction has landed before the kernel ends
    }
    for (int64_t i = done + tid; i < p.n; i += nthreads) {
        MapRun<ACC> run;
        one(i, X[i], Y[i], FUSED ? Z[i] : (T)0, VX[i], FUSED ? VY[i] : (T)0, FUSED ? VZ[i] : (T)0, W ? W[i] : (T)1, run);
        run.flush(maps0, B.sums(), st);
    }
}

}  // namespace pnm
