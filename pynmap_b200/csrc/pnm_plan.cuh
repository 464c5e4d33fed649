// pnm_plan.cuh -- builds the CubePlan of k_project_cube (pnm_cube.cuh) on the device from a strided sample of the
// particles: WHICH pixels' moments and WHICH voxels' weights are accumulated in shared memory first.  The plan affects
// speed only -- every particle is summed exactly once whatever the plan says -- so the sample uses the plain bin guess
// (no exact edge handling) and unit weights.
//   1 k_plan_sample   sample particle -> (pixel, velocity bin) key; column / row counts of the map
//   2 k_plan_box      box of at most 64 x 64 pixels with the largest column x row counts (sliding windows)
//   3 k_plan_vox      keys inside the box -> per-voxel counts of the box
//   4 k_plan_levels   per box pixel and threshold level: the hull of the velocity bins with count >= threshold
//   5 k_plan_finish   smallest threshold whose windows fit the voxel slots; exclusive scan -> table
#pragma once
#include "pnm_cube.cuh"

namespace pnm {

constexpr int kPlanSample = 1 << 20;              // sample particles (groups of 8 consecutive ones: one 32-byte sector per array)
constexpr int kPlanLevels = 16;
__constant__ int c_plan_tau[kPlanLevels] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32, 48, 64, 96};

struct PlanScratch {                              // device pointers into one allocation owned by pnm_cube_plan
    uint2* keys;                                  // [kPlanSample] (pixel = ix * ny + iy, iv) or (0xffffffff, -)
    uint32_t* colsum;                             // [nx] sample particles per map column
    uint32_t* rowsum;                             // [ny]
    uint32_t* voxhist;                            // [kBoxPix][nv]
    uint16_t* lo;                                 // [kBoxPix][kPlanLevels]
    uint8_t* len;                                 // [kBoxPix][kPlanLevels]
    uint32_t* tot_len;                            // [kPlanLevels]
    uint32_t* tot_cap;                            // [kPlanLevels]
    uint32_t* cap;                                // [kBoxPix][kPlanLevels] sample particles inside the window
    double* sum_abs;                              // [4]: sum |w|, sum |w*d|, sum |w*d*d| over the binned sample particles, their count
};

__global__ void __launch_bounds__(256) k_plan_sample(const CubeParams p, PlanScratch s, int64_t nsample, int64_t stride8) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, cnt = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nsample; j += nthreads) {
        const int64_t i = (j >> 3) * stride8 + (j & 7);
        uint2 key = make_uint2(0xffffffffu, 0u);
        if (i < p.n) {
            const float x = p.x[i], y = p.y[i], z = p.z[i], vx = p.vx[i], vy = p.vy[i], vz = p.vz[i];
            const float px = rot_row(p.m[0], p.m[1], p.m[2], x, y, z);
            const float py = rot_row(p.m[3], p.m[4], p.m[5], x, y, z);
            const float d = rot_row(p.m[6], p.m[7], p.m[8], vx, vy, vz);
            bool sure = true;
            const int ix = fast_bin(px, p.ax, p.cx, p.hx, sure);
            const int iy = fast_bin(py, p.ay, p.cy, p.hy, sure);
            const int iv = fast_bin(d, p.av, p.cv, p.hv, sure);
            if ((unsigned)ix < (unsigned)p.nx && (unsigned)iy < (unsigned)p.ny) {
                atomicAdd(s.colsum + ix, 1u);
                atomicAdd(s.rowsum + iy, 1u);
                if ((unsigned)iv < (unsigned)p.nv) key = make_uint2((unsigned)(ix * p.ny + iy), (unsigned)iv);
                // typical term sizes (they set the fixed-point scale of the shared-memory sums under float64 accumulation)
                const float w = p.w ? p.w[i] : 1.0f;
                const float wd = w * d, wdd = wd * d;
                if (is_finite(wdd) && is_finite(w)) { a0 += fabsf(w); a1 += fabsf(wd); a2 += fabsf(wdd); cnt += 1.0; }
            }
        }
        s.keys[j] = key;
    }
    for (int o = 16; o; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0.0) {
        atomicAdd(s.sum_abs, a0); atomicAdd(s.sum_abs + 1, a1); atomicAdd(s.sum_abs + 2, a2); atomicAdd(s.sum_abs + 3, cnt);
    }
}

// best window of `w` consecutive entries of a[0..n): returns its start (block-wide; n <= PNM_MAX_EDGES)
__device__ int best_window(const uint32_t* a, int n, int w, unsigned long long* s_best) {
    if (threadIdx.x == 0) *s_best = 0ull;
    __syncthreads();
    for (int start = threadIdx.x; start + w <= n; start += blockDim.x) {
        unsigned long long sum = 0;
        for (int k = 0; k < w; ++k) sum += a[start + k];
        // largest sum wins, ties go to the smallest start
        atomicMax(s_best, (sum << 16) | (unsigned long long)(0xffff - start));
    }
    __syncthreads();
    const int start = 0xffff - (int)(*s_best & 0xffffull);
    __syncthreads();
    return start;
}

__global__ void __launch_bounds__(1024) k_plan_box(const CubeParams p, PlanScratch s, CubePlan* plan) {
    __shared__ unsigned long long s_best;
    const int bw = min(p.nx, kBoxMax), bh = min(p.ny, kBoxMax);
    const int bx0 = best_window(s.colsum, p.nx, bw, &s_best);
    const int by0 = best_window(s.rowsum, p.ny, bh, &s_best);
    if (threadIdx.x == 0) { plan->bx0 = bx0; plan->by0 = by0; plan->bw = bw; plan->bh = bh; }
}

__global__ void __launch_bounds__(256) k_plan_vox(const CubeParams p, PlanScratch s, const CubePlan* plan, int64_t nsample) {
    const int bx0 = plan->bx0, by0 = plan->by0, bw = plan->bw, bh = plan->bh;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nsample; j += nthreads) {
        const uint2 key = s.keys[j];
        if (key.x == 0xffffffffu) continue;
        const int ix = (int)(key.x / (unsigned)p.ny), iy = (int)(key.x - (unsigned)ix * (unsigned)p.ny);
        const unsigned bxi = (unsigned)(ix - bx0), byi = (unsigned)(iy - by0);
        if (bxi < (unsigned)bw && byi < (unsigned)bh) atomicAdd(s.voxhist + (size_t)(bxi * bh + byi) * p.nv + key.y, 1u);
    }
}

// one warp per box pixel
__global__ void __launch_bounds__(256) k_plan_levels(const CubeParams p, PlanScratch s, const CubePlan* plan) {
    const int bp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (bp >= plan->bw * plan->bh) return;
    const uint32_t* h = s.voxhist + (size_t)bp * p.nv;
    // peak bin (clipping centre for hulls longer than a window may be)
    unsigned long long best = 0;
    for (int iv = lane; iv < p.nv; iv += 32) {
        const unsigned long long c = ((unsigned long long)h[iv] << 16) | (unsigned)(0xffff - iv);
        best = c > best ? c : best;
    }
    for (int o = 16; o; o >>= 1) { const unsigned long long b = __shfl_xor_sync(0xffffffffu, best, o); best = b > best ? b : best; }
    const int peak = 0xffff - (int)(best & 0xffffull);
    for (int j = 0; j < kPlanLevels; ++j) {
        const unsigned tau = (unsigned)c_plan_tau[j];
        int lo = 1 << 30, hi = -1;
        for (int iv = lane; iv < p.nv; iv += 32)
            if (h[iv] >= tau) { lo = min(lo, iv); hi = max(hi, iv); }
        for (int o = 16; o; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
        int len = hi >= lo ? hi - lo + 1 : 0;
        if (len > kTabLenMax) {                       // keep the window around the peak
            lo = max(lo, min(peak - kTabLenMax / 2, hi - kTabLenMax + 1));
            len = kTabLenMax;
        }
        if (len > 0 && lo > kTabVloMax) len = 0;      // the table entry has 10 bits for the first bin
        unsigned cap = 0;
        for (int iv = lo + lane; iv < lo + len; iv += 32) cap += h[iv];
        for (int o = 16; o; o >>= 1) cap += __shfl_xor_sync(0xffffffffu, cap, o);
        if (lane == 0) {
            s.lo[bp * kPlanLevels + j] = (uint16_t)(len ? lo : 0);
            s.len[bp * kPlanLevels + j] = (uint8_t)len;
            s.cap[bp * kPlanLevels + j] = cap;
            if (len) { atomicAdd(s.tot_len + j, (unsigned)len); atomicAdd(s.tot_cap + j, cap); }
        }
    }
}

__global__ void __launch_bounds__(1024) k_plan_finish(PlanScratch s, CubePlan* plan, int slots_max, int nsample) {
    __shared__ int s_scan[1024];
    __shared__ int s_carry;
    const int nbox = plan->bw * plan->bh;
    int level = kPlanLevels - 1;
    for (int j = kPlanLevels - 1; j >= 0; --j)
        if ((int)s.tot_len[j] <= slots_max) level = j;          // smallest threshold that fits
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    // exclusive scan of the window lengths, 1024 box pixels per round; windows past the capacity are dropped
    for (int b0 = 0; b0 < nbox; b0 += 1024) {
        const int bp = b0 + threadIdx.x;
        const int len = bp < nbox ? (int)s.len[bp * kPlanLevels + level] : 0;
        s_scan[threadIdx.x] = len;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int v = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += v;
            __syncthreads();
        }
        const int base = s_carry + s_scan[threadIdx.x] - len;
        if (bp < nbox) {
            const bool fits = base + len <= slots_max;
            const uint32_t lo = s.lo[bp * kPlanLevels + level];
            plan->tab[bp] = fits ? ((uint32_t)base | ((uint32_t)len << kTabBaseBits) | (lo << 22)) : 0u;
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry += s_scan[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        plan->nslots = min(s_carry, slots_max);
        plan->level = level;
        plan->captured = (int)s.tot_cap[level];
        plan->sampled = nsample;
        const double cnt = s.sum_abs[3] > 0.0 ? s.sum_abs[3] : 1.0;
        for (int k = 0; k < 3; ++k) {
            const double m = s.sum_abs[k] / cnt;
            plan->mean_abs[k] = m;
            // float64 accumulators: 2^k that puts the mean |term| into [2^kCubeTypBits, 2^(kCubeTypBits+1)); applied in float32
            int e = 0;
            if (m > 0.0 && m <= 1.7976931348623157e+308) e = ilogb(m);
            const int kk = max(-120, min(120, kCubeTypBits - e));
            plan->sc_smem[k] = (float)ldexp(1.0, kk);
            plan->inv_smem[k] = ldexp(1.0, -kk);
        }
    }
}

}  // namespace pnm
