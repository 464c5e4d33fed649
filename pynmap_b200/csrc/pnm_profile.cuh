// pnm_profile.cuh -- radial profiles of the velocity ellipsoid in an equatorial slab (sm_100a only).
//
// Reference: point_2vprofiles (pynmap/grid.py:16-68; caller snapshot.get_tensor_profiles,
// pynmap/pynmap.py:457-474): select dz[0] < z < dz[1], R = sqrt(x^2 + y^2), theta = arctan2(y, x)
// (rotation.py:121-144), VR = Vx cos + Vy sin, VT = -Vx sin + Vy cos, np.digitize(R, Rbin) on
// log-spaced radii, then mean() and std() of VR, VT, Vz per bin in a Python loop.
//
// Here: one pass over the particles.  R is rounded like numpy's (multiply, add, sqrt in the particle
// dtype), the bin is an exact search over the float64 radii, so the per-bin membership (counts) is
// bit-exact; cos / sin are x / R and y / R instead of libm's cos(arctan2()), and the moments are
// one-pass float64 sums (the reference: float32 pairwise two-pass), i.e. the profiles agree to
// float32 rounding, not bitwise.
#pragma once
#include "pnm_common.cuh"

namespace pnm {

constexpr int kProfileSlots = 8;         // n, sum VR, sum VR^2, sum VT, sum VT^2, sum Vz, sum Vz^2, pad
constexpr int kProfileReplicas = 64;     // selected by warp: keeps the float64 reductions off one address
constexpr int kProfileMaxBins = 4096;

__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float quot_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double quot_rn(double a, double b) { return __ddiv_rn(a, b); }

template <typename T>
__device__ __forceinline__ bool profile_select(const T* z, const uint8_t* mask, int64_t i, double z_lo, double z_hi) {
    if (mask && !mask[i]) return false;
    const double zz = (double)z[i];
    return z_lo < zz && zz < z_hi;                       // grid.py:41 (NaN fails both)
}

template <typename T>
__device__ __forceinline__ T profile_radius(T x, T y) { return sqrt_rn(add_rn(mul_rn(x, x), mul_rn(y, y))); }

// R.max() over the selected particles (grid.py:51); NaN radii are skipped.
template <typename T>
__global__ void __launch_bounds__(256) k_profile_rmax(const T* __restrict__ x, const T* __restrict__ y,
                                                      const T* __restrict__ z, const uint8_t* __restrict__ mask,
                                                      int64_t n, double z_lo, double z_hi, unsigned long long* out) {
    double m = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (!profile_select(z, mask, i, z_lo, z_hi)) continue;
        const double r = (double)profile_radius(x[i], y[i]);
        if (r > m) m = r;
    }
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

template <typename T>
__global__ void __launch_bounds__(256) k_profile_acc(const T* __restrict__ x, const T* __restrict__ y,
                                                     const T* __restrict__ z, const T* __restrict__ vx,
                                                     const T* __restrict__ vy, const T* __restrict__ vz,
                                                     const uint8_t* __restrict__ mask, int64_t n, double z_lo, double z_hi,
                                                     const double* __restrict__ rbin, int nbins, double* __restrict__ acc) {
    extern __shared__ double s_rbin[];
    for (int k = threadIdx.x; k < nbins; k += blockDim.x) s_rbin[k] = rbin[k];
    __syncthreads();
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double* mine = acc + (int64_t)((gtid >> 5) & (kProfileReplicas - 1)) * nbins * kProfileSlots;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = gtid; i < n; i += stride) {
        if (!profile_select(z, mask, i, z_lo, z_hi)) continue;
        const T xi = x[i], yi = y[i];
        const T r = profile_radius(xi, yi);
        const double rd = (double)r;
        if (!(rd == rd)) continue;                       // NaN radius: np.digitize would put it past the end
        int lo = 0, hi = nbins;                          // np.digitize(R, Rbin) = #{k : Rbin[k] <= R}  (grid.py:54)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_rbin[mid] <= rd) lo = mid + 1; else hi = mid;
        }
        if (lo >= nbins) continue;                       // the loop of grid.py:55 never visits digit == nbins
        T c = (T)1, s = (T)0;                            // arctan2(0, 0) = 0
        if (r > (T)0) { c = quot_rn(xi, r); s = quot_rn(yi, r); }
        const T vr = add_rn(mul_rn(vx[i], c), mul_rn(vy[i], s));          // grid.py:47
        const T vt = add_rn(mul_rn(-vx[i], s), mul_rn(vy[i], c));         // grid.py:48
        const double a = (double)vr, b = (double)vt, w = (double)vz[i];
        double* slot = mine + (int64_t)lo * kProfileSlots;
        atomicAdd(slot + 0, 1.0);
        atomicAdd(slot + 1, a); atomicAdd(slot + 2, a * a);
        atomicAdd(slot + 3, b); atomicAdd(slot + 4, b * b);
        atomicAdd(slot + 5, w); atomicAdd(slot + 6, w * w);
    }
}

// mean() and std() (ddof = 0) per bin as float32, like the reference's s / v arrays (grid.py:37-38, :57-62).
__global__ void __launch_bounds__(128) k_profile_final(const double* __restrict__ acc, int nbins,
                                                       float* __restrict__ v, float* __restrict__ s,
                                                       long long* __restrict__ counts) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbins) return;
    double t[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < kProfileReplicas; ++r)
#pragma unroll
        for (int k = 0; k < 7; ++k) t[k] += acc[((int64_t)r * nbins + b) * kProfileSlots + k];
    if (counts) counts[b] = (long long)t[0];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double mean = t[1 + 2 * c] / t[0];         // 0 / 0 = NaN for an empty bin, like numpy
        const double var = t[2 + 2 * c] / t[0] - mean * mean;
        v[c * nbins + b] = (float)mean;
        s[c * nbins + b] = t[0] > 0.0 ? (float)sqrt(fmax(var, 0.0)) : __int_as_float(0x7fc00000);
    }
}

}  // namespace pnm
