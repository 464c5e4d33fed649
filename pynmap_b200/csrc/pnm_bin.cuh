// pnm_bin.cuh -- fused rotate + project + bin kernel (K1+K2+K3 of SURVEY.md 2.2).
//
// One pass over the SoA particle arrays produces, per viewing angle,
//   maps accumulator  [iy][ix][4] = { sum w (or the out-of-V-range remainder), sum w*d, sum w*d*d, - }
//   cube accumulator  [ix][iy][iv] =   sum w
// Reference semantics: rotation.py:115-117 (rotation), grid.py:121-131 / :200-201 (2-D moment
// histograms, result transposed to [iy][ix]), grid.py:252-259 (3-D histogram, [ix][iy][iv]).
//
// Roofline: HBM.  Algorithmic bytes = 7 arrays x sizeof(T) per particle (28 B for float32);
// the accumulators (<= 34 MB for the 128x128x256 cube) stay resident in the 126 MB L2, so the
// scatter is served by L2 reduction units (REDG) and never touches HBM until eviction.
#pragma once
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
    int sums;
    int nchain, nviews;
    double scale[3];
    void* acc_maps; void* acc_cube;
    int64_t maps_stride, cube_stride;                 // 8-byte elements per view
    float mats[kMaxMats * 9];                         // [view][chain][9]
};

template <typename T> struct Vec;
template <> struct Vec<float> { using type = float4; static constexpr int N = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };

template <typename T>
__device__ __forceinline__ void unpack(const float4& v, T* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
template <typename T>
__device__ __forceinline__ void unpack(const double2& v, T* o) { o[0] = v.x; o[1] = v.y; }

// streaming loads: read once -> no L1 allocation, L2 evict-first policy, so the stream does not
// push the L2-resident accumulators out (a bare .L2::evict_first needs 256-bit loads; the
// createpolicy / cache_hint form works at any width)
__device__ __forceinline__ uint64_t make_stream_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ld_stream(const float4* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double2 ld_stream(const double2* p, uint64_t pol) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
                 : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol));
    return r;
}

// --------------------------------------------------------------------------------------------
// One particle, all views.
// --------------------------------------------------------------------------------------------
template <typename T, bool FUSED, int ACC>
__device__ __forceinline__ void bin_particle(const BinParams& p, const T* sx, const T* sy, const T* sv,
                                             T x, T y, T z, T vx, T vy, T vz, T w, bool masked_in) {
    const T x0 = (T)p.x0, idx = (T)p.inv_dx, y0 = (T)p.y0, idy = (T)p.inv_dy, v0 = (T)p.v0, idv = (T)p.inv_dv;
    for (int view = 0; view < p.nviews; ++view) {
        T px = x, py = y, d = vx;
        if (FUSED) {
            const float* m = p.mats + (size_t)view * p.nchain * 9;
            T pz = z, qx = vx, qy = vy, qz = vz;
            for (int c = 0; c < p.nchain - 1; ++c, m += 9) {   // compounded rotations keep all rows
                T ax = rot_row(m[0], m[1], m[2], px, py, pz);
                T ay = rot_row(m[3], m[4], m[5], px, py, pz);
                T az = rot_row(m[6], m[7], m[8], px, py, pz);
                T bx = rot_row(m[0], m[1], m[2], qx, qy, qz);
                T by = rot_row(m[3], m[4], m[5], qx, qy, qz);
                T bz = rot_row(m[6], m[7], m[8], qx, qy, qz);
                px = ax; py = ay; pz = az; qx = bx; qy = by; qz = bz;
            }
            T ax = rot_row(m[0], m[1], m[2], px, py, pz);       // last step: sky x, sky y, v_los
            T ay = rot_row(m[3], m[4], m[5], px, py, pz);
            d = rot_row(m[6], m[7], m[8], qx, qy, qz);
            px = ax; py = ay;
        }
        bool masked = masked_in;
        if (p.mask_mode == PNM_MASK_AND_NONFINITE) masked = masked_in && !is_finite(d);
        if (masked) continue;
        const int ix = find_bin<T>(px, sx, p.nx, x0, idx);
        const int iy = find_bin<T>(py, sy, p.ny, y0, idy);
        if ((ix | iy) < 0) continue;
        int iv = -1;
        if (p.sums & PNM_SUM_CUBE) iv = find_bin<T>(d, sv, p.nv, v0, idv);

        const int64_t pix = ((int64_t)iy * p.nx + ix) * PNM_MAP_SLOTS + view * p.maps_stride;
        if (p.sums & PNM_SUM_W) {
            if (!((p.sums & PNM_W_FROM_CUBE) && iv >= 0)) acc_add<ACC>(p.acc_maps, pix + 0, (double)w, p.scale[0]);
        }
        if (p.sums & PNM_SUM_WD) acc_add<ACC>(p.acc_maps, pix + 1, (double)mul_rn(w, d), p.scale[1]);
        if (p.sums & PNM_SUM_WD2) acc_add<ACC>(p.acc_maps, pix + 2, (double)mul_rn(w, mul_rn(d, d)), p.scale[2]);
        if (iv >= 0) {
            const int64_t vox = ((int64_t)ix * p.ny + iy) * p.nv + iv + view * p.cube_stride;
            acc_add<ACC>(p.acc_cube, vox, (double)w, p.scale[0]);
        }
    }
}

// --------------------------------------------------------------------------------------------
// Kernel: persistent grid-stride loop, 16-byte vector loads (4 float32 / 2 float64 particles per
// thread per array), scalar tail.  VECLOAD=false is the path for unaligned array pointers.
// --------------------------------------------------------------------------------------------
template <typename T, bool FUSED, int ACC, bool VECLOAD>
__global__ void __launch_bounds__(256) k_project_bin(const __grid_constant__ BinParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sx = reinterpret_cast<T*>(smem_raw);
    T* sy = sx + (p.nx + 1);
    T* sv = sy + (p.ny + 1);
    for (int i = threadIdx.x; i <= p.nx; i += blockDim.x) sx[i] = static_cast<const T*>(p.thr_x)[i];
    for (int i = threadIdx.x; i <= p.ny; i += blockDim.x) sy[i] = static_cast<const T*>(p.thr_y)[i];
    if (p.sums & PNM_SUM_CUBE)
        for (int i = threadIdx.x; i <= p.nv; i += blockDim.x) sv[i] = static_cast<const T*>(p.thr_v)[i];
    __syncthreads();

    using V = typename Vec<T>::type;
    constexpr int VN = Vec<T>::N;
    const T* X = static_cast<const T*>(p.x);  const T* Y = static_cast<const T*>(p.y);
    const T* Z = static_cast<const T*>(p.z);  const T* VX = static_cast<const T*>(p.vx);
    const T* VY = static_cast<const T*>(p.vy); const T* VZ = static_cast<const T*>(p.vz);
    const T* W = static_cast<const T*>(p.w);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;

    int64_t done = 0;
    if (VECLOAD) {
        const uint64_t pol = make_stream_policy();
        const int64_t nvec = p.n / VN;
        for (int64_t i = tid; i < nvec; i += nthreads) {
            T x[VN], y[VN], z[VN], vx[VN], vy[VN], vz[VN], w[VN];
            unpack<T>(ld_stream(reinterpret_cast<const V*>(X) + i, pol), x);
            unpack<T>(ld_stream(reinterpret_cast<const V*>(Y) + i, pol), y);
            unpack<T>(ld_stream(reinterpret_cast<const V*>(VX) + i, pol), vx);
            if (FUSED) {
                unpack<T>(ld_stream(reinterpret_cast<const V*>(Z) + i, pol), z);
                unpack<T>(ld_stream(reinterpret_cast<const V*>(VY) + i, pol), vy);
                unpack<T>(ld_stream(reinterpret_cast<const V*>(VZ) + i, pol), vz);
            }
            if (W) unpack<T>(ld_stream(reinterpret_cast<const V*>(W) + i, pol), w);
#pragma unroll
            for (int k = 0; k < VN; ++k) {
                const bool m = p.mask_mode != PNM_MASK_NONE && p.mask[i * VN + k] != 0;
                bin_particle<T, FUSED, ACC>(p, sx, sy, sv, x[k], y[k], FUSED ? z[k] : (T)0, vx[k],
                                            FUSED ? vy[k] : (T)0, FUSED ? vz[k] : (T)0, W ? w[k] : (T)1, m);
            }
        }
        done = nvec * VN;
    }
    for (int64_t i = done + tid; i < p.n; i += nthreads) {
        const bool m = p.mask_mode != PNM_MASK_NONE && p.mask[i] != 0;
        bin_particle<T, FUSED, ACC>(p, sx, sy, sv, X[i], Y[i], FUSED ? Z[i] : (T)0, VX[i],
                                    FUSED ? VY[i] : (T)0, FUSED ? VZ[i] : (T)0, W ? W[i] : (T)1, m);
    }
}

}  // namespace pnm
