// pnm_post.cuh -- everything after the scatter: finalisation (K5), smoothing (K4), cube moments (f1),
// plus the materialising rotation (K1) and the abs-max reduction used to size fixed-point scales.
#pragma once
#ifndef PNM_MAX_PEERS
#define PNM_MAX_PEERS 16
#endif
#include "pnm_common.cuh"
#include "pnm_bin.cuh"

namespace pnm {

// ---------------------------------------------------------------- K1: materialising rotation
struct Mat9 { float m[9]; };

template <typename T>
__global__ void __launch_bounds__(256) k_rotate(const T* __restrict__ x, const T* __restrict__ y,
                                                const T* __restrict__ z, T* ox, T* oy, T* oz,
                                                int64_t n, const __grid_constant__ Mat9 mat) {
    const float* m = mat.m;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T a = x[i], b = y[i], c = z[i];      // read all three before writing: in-place is allowed
        const T r0 = rot_row(m[0], m[1], m[2], a, b, c);
        const T r1 = rot_row(m[3], m[4], m[5], a, b, c);
        const T r2 = rot_row(m[6], m[7], m[8], a, b, c);
        ox[i] = r0; oy[i] = r1; oz[i] = r2;
    }
}

// ---------------------------------------------------------------- accumulator -> float64
// maps accumulator: [replica][iy][ix][4]; cube accumulator: [iv][ix][iy] (velocity-major).
template <int ACC>
__device__ __forceinline__ double maps_slot(const void* acc_maps, int64_t npix, int replicas, int pix, int slot,
                                            double inv_scale) {
    const int64_t stride = npix * PNM_MAP_SLOTS, at = (int64_t)pix * PNM_MAP_SLOTS + slot;
    if (ACC == PNM_ACC_F64) {
        double t = 0.0;
        for (int r = 0; r < replicas; ++r) t = __dadd_rn(t, reinterpret_cast<const double*>(acc_maps)[r * stride + at]);
        return t;
    }
    long long t = 0;
    for (int r = 0; r < replicas; ++r) t += reinterpret_cast<const long long*>(acc_maps)[r * stride + at];
    return __dmul_rn((double)t, inv_scale);
}

// K5 (grid.py:133-138).  Block = 32 spaxels (q = ix*ny + iy) x 8 velocity slabs: with
// PNM_W_FROM_CUBE the V-sum of the cube is a coalesced column sum split over the 8 slabs and
// combined in a fixed order through shared memory.  Separate roundings (no FMA contraction) so
// that equal sums give bit-equal V and S: numpy evaluates s2/s0, V*V and the subtraction as three
// ufunc calls.
template <int ACC>
__global__ void __launch_bounds__(256) k_finalize_vmaps(const void* acc_maps, const void* acc_cube,
                                                        int nx, int ny, int nv, int replicas, int w_from_cube,
                                                        double is0, double is1, double is2,
                                                        double* M, double* V, double* S, double* S1o, double* S2o) {
    using A = typename AccT<ACC>::type;
    __shared__ A part[8][32];
    __shared__ A part1[8][32];
    __shared__ A part2[8][32];
    const int q = blockIdx.x * 32 + threadIdx.x;             // = ix*ny + iy
    const int npix = nx * ny;
    A t = 0, t1 = 0, t2 = 0;
    if (w_from_cube == 2 && q < npix) {
        // interleaved cube (PNM_CUBE_MOMENTS): records [g][q] = { w even, w odd, sum w*d, sum w*d*d }, g <= G;
        // record slab G = ceil(nv / 2) holds the particles outside the V range
        const A* c = reinterpret_cast<const A*>(acc_cube);
        for (int g = threadIdx.y; g <= (nv + 1) / 2; g += 8) {
            const A* rec = c + ((int64_t)g * npix + q) * 4;
            t += rec[0] + rec[1]; t1 += rec[2]; t2 += rec[3];
        }
    } else if (w_from_cube && q < npix) {
        const A* c = reinterpret_cast<const A*>(acc_cube);
        for (int k = threadIdx.y; k < nv; k += 8) t += c[(int64_t)k * npix + q];
    }
    part[threadIdx.y][threadIdx.x] = t;
    if (w_from_cube == 2) { part1[threadIdx.y][threadIdx.x] = t1; part2[threadIdx.y][threadIdx.x] = t2; }
    __syncthreads();
    if (threadIdx.y != 0 || q >= npix) return;
    const int ix = q / ny, iy = q - ix * ny;
    const int pix = iy * nx + ix;                            // map layout [iy][ix] (the reference's .T)
    double s0;
    if (w_from_cube == 2) {
        // all three sums = V-sums of the cube records (the maps accumulator is not used with this layout)
        A tot0 = 0, tot1 = 0, tot2 = 0;
        for (int k = 0; k < 8; ++k) { tot0 += part[k][threadIdx.x]; tot1 += part1[k][threadIdx.x]; tot2 += part2[k][threadIdx.x]; }
        const double s0m = (ACC == PNM_ACC_F64) ? (double)tot0 : __dmul_rn((double)tot0, is0);
        const double s1m = (ACC == PNM_ACC_F64) ? (double)tot1 : __dmul_rn((double)tot1, is1);
        const double s2m = (ACC == PNM_ACC_F64) ? (double)tot2 : __dmul_rn((double)tot2, is2);
        double vm = 0.0, sm = 0.0;
        if (s0m != 0.0) {
            vm = __ddiv_rn(s1m, s0m);
            sm = __dsqrt_rn(__dsub_rn(__ddiv_rn(s2m, s0m), __dmul_rn(vm, vm)));
        }
        if (M) M[pix] = s0m;
        if (V) V[pix] = vm;
        if (S) S[pix] = sm;
        if (S1o) S1o[pix] = s1m;
        if (S2o) S2o[pix] = s2m;
        return;
    }
    if (w_from_cube) {
        // sum w = remainder outside the V range (slot 0 of every replica) + the V-sum of the cube
        const A* m = reinterpret_cast<const A*>(acc_maps);
        const int64_t stride = (int64_t)npix * PNM_MAP_SLOTS;
        A tot = 0;
        for (int r = 0; r < replicas; ++r) tot += m[r * stride + (int64_t)pix * PNM_MAP_SLOTS + kSlotS0];
        for (int k = 0; k < 8; ++k) tot += part[k][threadIdx.x];
        s0 = (ACC == PNM_ACC_F64) ? (double)tot : __dmul_rn((double)tot, is0);
    } else {
        s0 = maps_slot<ACC>(acc_maps, npix, replicas, pix, kSlotS0, is0);
    }
    const double s1 = maps_slot<ACC>(acc_maps, npix, replicas, pix, kSlotS1, is1);
    const double s2 = maps_slot<ACC>(acc_maps, npix, replicas, pix, kSlotS2, is2);
    double v = 0.0, s = 0.0;
    if (s0 != 0.0) {
        v = __ddiv_rn(s1, s0);
        s = __dsqrt_rn(__dsub_rn(__ddiv_rn(s2, s0), __dmul_rn(v, v)));
    }
    if (M) M[pix] = s0;
    if (V) V[pix] = v;
    if (S) S[pix] = s;
    if (S1o) S1o[pix] = s1;
    if (S2o) S2o[pix] = s2;
}

// grid.py:203-205
template <int ACC>
__global__ void __launch_bounds__(256) k_finalize_map(const void* acc_maps, int npix, int replicas,
                                                      double is0, double is1, double* W, double* D, double* DW) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const double s0 = maps_slot<ACC>(acc_maps, npix, replicas, pix, kSlotS0, is0);
    const double s1 = maps_slot<ACC>(acc_maps, npix, replicas, pix, kSlotS1, is1);
    if (W) W[pix] = s0;
    if (D) D[pix] = (s0 != 0.0) ? __ddiv_rn(s1, s0) : 0.0;
    if (DW) DW[pix] = s1;
}

// fold the map replicas into replica 0 (and, if CLEAR, zero the others) before a multi-GPU reduce.
// Replicas are summed in index order; loads are issued eight at a time.
template <int ACC, bool CLEAR>
__global__ void __launch_bounds__(256) k_fold_maps(void* acc_maps, int64_t per_replica, int replicas) {
    using A = typename AccT<ACC>::type;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_replica) return;
    A* a = reinterpret_cast<A*>(acc_maps) + i;
    A t = a[0];
    for (int r = 1; r < replicas; r += 8) {
        A v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (r + k < replicas) ? a[(int64_t)(r + k) * per_replica] : A(0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (ACC == PNM_ACC_F64) t = (A)__dadd_rn((double)t, (double)v[k]);
            else t += v[k];
            if (CLEAR && r + k < replicas) a[(int64_t)(r + k) * per_replica] = A(0);
        }
    }
    a[0] = t;
}

// spread_amr (misc_io.py:184-224): every cell is replicated nsample times.
//   coordinates: out[k] = c[k / nsample] + ((u[k] - 0.5) * stretch) * scale[k / nsample]   (float64, misc_io.py:209-212)
//   values     : out[k] = v[k / nsample] / nsample  or  v[k / nsample]                     (input dtype, :216-222)
template <typename T>
__global__ void __launch_bounds__(256) k_amr_jitter(const T* __restrict__ c, const double* __restrict__ scale,
                                                    const double* __restrict__ u, int64_t total, int nsample,
                                                    double stretch, double* __restrict__ out) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = k / nsample;
        const double r = __dmul_rn(__dmul_rn(__dsub_rn(u[k], 0.5), stretch), scale[i]);
        out[k] = __dadd_rn((double)c[i], r);
    }
}

__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

template <typename T>
__global__ void __launch_bounds__(256) k_amr_repeat(const T* __restrict__ v, int64_t total, int nsample, int divide,
                                                    T* __restrict__ out) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        const T x = v[k / nsample];
        out[k] = divide ? div_rn(x, (T)nsample) : x;
    }
}

// cube accumulator [iv][q] -> reference layout [q][iv] = cube[ix][iy][iv] (grid.py:259) as float64:
// 32x32 shared-memory tile transpose, both sides coalesced.
template <int ACC>
__global__ void __launch_bounds__(256) k_cube_out(const void* acc_cube, double* out, int npix, int nv, double inv_scale,
                                                  int moments) {
    __shared__ double tile[32][33];
    const int q0 = blockIdx.x * 32, v0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int iv = v0 + r, q = q0 + threadIdx.x;
        double val = 0.0;
        if (iv < nv && q < npix) {
            // plain: [iv][q]; interleaved (PNM_CUBE_MOMENTS): slot iv & 1 of record [iv / 2][q]
            const int64_t at = moments ? ((int64_t)(iv >> 1) * npix + q) * 4 + (iv & 1) : (int64_t)iv * npix + q;
            val = (ACC == PNM_ACC_F64) ? reinterpret_cast<const double*>(acc_cube)[at]
                                       : __dmul_rn((double)reinterpret_cast<const long long*>(acc_cube)[at], inv_scale);
        }
        tile[r][threadIdx.x] = val;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int q = q0 + r, iv = v0 + threadIdx.x;
        if (q < npix && iv < nv) out[(int64_t)q * nv + iv] = tile[threadIdx.x][r];
    }
}

// ---------------------------------------------------------------- slab-sharded finalisation (multi-GPU)
// After a reduce-scatter over ranks along the record-slab axis of the interleaved cube (records [g][q][4]), a rank
// holds the slabs g0 <= g < g1 of the TOTAL.  k_slab_sums gives that rank's share of the three map sums per pixel
// (q order, accumulator type: the shares are summed over ranks by an all-reduce, exactly for int64), k_finalize_sums
// turns the totals into M, V, S (grid.py:133-138, same roundings as k_finalize_vmaps), k_cube_out_slabs writes the
// rank's velocity bins [iv0, iv0 + nvl) of the LOSVD cube as float64 [q][nvl] (grid.py:259).
// out[i] = chunks[0][i] + chunks[1][i] + ... in index order (the same order on every rank: deterministic float64 sums);
// `chunks` = nchunks arrays of `per` elements back to back.  128-bit loads, grid-stride.
template <int ACC>
__global__ void __launch_bounds__(256) k_sum_chunks(const void* chunks, int nchunks, int64_t per, void* out) {
    using A = typename AccT<ACC>::type;
    const A* c = reinterpret_cast<const A*>(chunks);
    A* o = reinterpret_cast<A*>(out);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 2;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < per; i += stride) {
        if (i + 1 < per && (per & 1) == 0) {
            A a0 = 0, a1 = 0;
            for (int r = 0; r < nchunks; ++r) {
                const longlong2 v = *reinterpret_cast<const longlong2*>(c + (int64_t)r * per + i);
                if (ACC == PNM_ACC_F64) { a0 = (A)__dadd_rn((double)a0, __longlong_as_double(v.x)); a1 = (A)__dadd_rn((double)a1, __longlong_as_double(v.y)); }
                else { a0 += (A)v.x; a1 += (A)v.y; }
            }
            o[i] = a0; o[i + 1] = a1;
        } else {
            for (int64_t j = i; j < per && j < i + 2; ++j) {
                A a = 0;
                for (int r = 0; r < nchunks; ++r) a += c[(int64_t)r * per + j];
                o[j] = a;
            }
        }
    }
}

// out[i] = src[0][i] + src[1][i] + ... (pointers into this device's or peer devices' memory; index order)
struct PeerPtrs { const void* p[PNM_MAX_PEERS]; int n; };
template <int ACC>
__global__ void __launch_bounds__(256) k_sum_peers(const PeerPtrs src, int64_t n, void* out) {
    using A = typename AccT<ACC>::type;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    A v[PNM_MAX_PEERS];
#pragma unroll
    for (int r = 0; r < PNM_MAX_PEERS; ++r) v[r] = r < src.n ? reinterpret_cast<const A*>(src.p[r])[i] : A(0);   // all loads in flight
    A t = 0;
#pragma unroll
    for (int r = 0; r < PNM_MAX_PEERS; ++r)
        if (r < src.n) { if (ACC == PNM_ACC_F64) t = (A)__dadd_rn((double)t, (double)v[r]); else t += v[r]; }
    reinterpret_cast<A*>(out)[i] = t;
}

template <int ACC>
__global__ void __launch_bounds__(256) k_slab_sums(const void* slabs, int npix, int g0, int g1, int G, void* sums3) {
    using A = typename AccT<ACC>::type;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= npix) return;
    const A* c = reinterpret_cast<const A*>(slabs);
    A t0 = 0, t1 = 0, t2 = 0;
    for (int g = g0; g < g1 && g <= G; ++g) {
        const A* rec = c + ((int64_t)(g - g0) * npix + q) * 4;
        t0 += rec[0] + rec[1]; t1 += rec[2]; t2 += rec[3];
    }
    A* o = reinterpret_cast<A*>(sums3);
    o[q] = t0; o[npix + q] = t1; o[2 * npix + q] = t2;
}

template <int ACC>
__global__ void __launch_bounds__(256) k_finalize_sums(const void* sums3, int nx, int ny, double is0, double is1, double is2,
                                                       double* M, double* V, double* S) {
    using A = typename AccT<ACC>::type;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;     // = ix*ny + iy
    const int npix = nx * ny;
    if (q >= npix) return;
    const A* a = reinterpret_cast<const A*>(sums3);
    const double s0 = (ACC == PNM_ACC_F64) ? (double)a[q] : __dmul_rn((double)a[q], is0);
    const double s1 = (ACC == PNM_ACC_F64) ? (double)a[npix + q] : __dmul_rn((double)a[npix + q], is1);
    const double s2 = (ACC == PNM_ACC_F64) ? (double)a[2 * npix + q] : __dmul_rn((double)a[2 * npix + q], is2);
    double v = 0.0, sg = 0.0;
    if (s0 != 0.0) {
        v = __ddiv_rn(s1, s0);
        sg = __dsqrt_rn(__dsub_rn(__ddiv_rn(s2, s0), __dmul_rn(v, v)));
    }
    const int ix = q / ny, iy = q - ix * ny;
    const int pix = iy * nx + ix;                            // map layout [iy][ix] (the reference's .T)
    if (M) M[pix] = s0;
    if (V) V[pix] = v;
    if (S) S[pix] = sg;
}

template <int ACC>
__global__ void __launch_bounds__(256) k_cube_out_slabs(const void* slabs, double* out, int npix, int g0, int iv0, int nvl,
                                                        double inv_scale) {
    __shared__ double tile[32][33];
    const int q0 = blockIdx.x * 32, v0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int il = v0 + r, q = q0 + threadIdx.x;          // il: velocity bin relative to iv0
        double val = 0.0;
        if (il < nvl && q < npix) {
            const int iv = iv0 + il;
            const int64_t at = ((int64_t)((iv >> 1) - g0) * npix + q) * 4 + (iv & 1);
            val = (ACC == PNM_ACC_F64) ? reinterpret_cast<const double*>(slabs)[at]
                                       : __dmul_rn((double)reinterpret_cast<const long long*>(slabs)[at], inv_scale);
        }
        tile[r][threadIdx.x] = val;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int q = q0 + r, il = v0 + threadIdx.x;
        if (q < npix && il < nvl) out[(int64_t)q * nvl + il] = tile[threadIdx.x][r];
    }
}

// ---------------------------------------------------------------- K4: separable Gaussian
// cube [nx][ny][nv], iv contiguous.  One pass convolves along AXIS (0 or 1) with zero fill:
//   out[i] = sum_t taps[t] * in[i + (t - r) * stride_axis]
// Threads run along iv so every tap is a coalesced row read served by L2 (the cube is resident).
struct Taps { double t[PNM_MAX_TAPS]; int n; };

// Each thread produces kSmoothOut consecutive outputs along AXIS from one sliding window of inputs:
// ntaps + 3 loads for 4 outputs instead of 4 * ntaps (the pass is bound by L2 -> SM traffic), one
// load + 4 FMAs per tap.  Every output is the same ordered sum_t fma(taps[t], in[..], acc) as a
// one-output-per-thread loop would compute; zero fill = out-of-range loads return 0.
constexpr int kSmoothOut = 4;

// astropy's nan_treatment='interpolate' (the default of convolve, losvd.py:130-131) for ONE output voxel: NaN samples are
// left out and the result is divided by the sum of the kernel values actually used (samples outside the array are the
// fill value 0 and do count); an output whose every sample is NaN keeps the input value.  The 2-D kernel is the outer
// product of the two normalised 1-D factors.  Only outputs that came out NaN from the separable passes get here.
__device__ __noinline__ double smooth_nan2d(const double* __restrict__ orig, int nx, int ny, int nv, int ix, int iy, int iv,
                                            const Taps& t0, const Taps& t1) {
    const int r0 = t0.n / 2, r1 = t1.n / 2;
    double top = 0.0, bot = 0.0;
    for (int a = 0; a < t0.n; ++a) {
        const int sx = ix + a - r0;
        for (int b = 0; b < t1.n; ++b) {
            const int sy = iy + b - r1;
            const double ker = __dmul_rn(t0.t[a], t1.t[b]);
            double val = 0.0;
            if (sx >= 0 && sx < nx && sy >= 0 && sy < ny) val = orig[((int64_t)sx * ny + sy) * nv + iv];
            if (val == val) { top = __fma_rn(val, ker, top); bot += ker; }
        }
    }
    return bot == 0.0 ? orig[((int64_t)ix * ny + iy) * nv + iv] : __ddiv_rn(top, bot);
}

// AXIS 1 is the second pass: with ``orig`` (the cube the first pass read) it repairs the outputs that have a NaN in
// their 2-D footprint (they come out NaN here) with smooth_nan2d.
template <int AXIS>
__global__ void __launch_bounds__(256) k_smooth_axis(const double* __restrict__ in, double* __restrict__ out,
                                                     int nx, int ny, int nv, const __grid_constant__ Taps taps,
                                                     const double* __restrict__ orig, const __grid_constant__ Taps taps_first) {
    // (the cube has < 2^31 voxels: checked at grid creation, so 32-bit indices are enough)
    const int len = AXIS == 0 ? nx : ny;
    const int nblk = (len + kSmoothOut - 1) / kSmoothOut;
    const int items = (AXIS == 0 ? nblk * ny : nx * nblk) * nv;
    const int step = AXIS == 0 ? ny * nv : nv;
    const int r = taps.n / 2;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < items; i += stride) {
        const int iv = i % nv, rest = i / nv;
        int pos0, base, ixo = 0;                        // first output position along AXIS, its flat index
        if (AXIS == 0) { const int iy = rest % ny; pos0 = (rest / ny) * kSmoothOut; base = (pos0 * ny + iy) * nv + iv; }
        else { const int b = rest % nblk; const int ix = rest / nblk; pos0 = b * kSmoothOut; base = (ix * ny + pos0) * nv + iv; ixo = ix; }
        auto at = [&](int p) -> double { return (p >= 0 && p < len) ? in[base + (p - pos0) * step] : 0.0; };
        double acc[kSmoothOut] = {0.0, 0.0, 0.0, 0.0};
        double w0 = at(pos0 - r), w1 = at(pos0 - r + 1), w2 = at(pos0 - r + 2), w3 = at(pos0 - r + 3);
        for (int t = 0; t < taps.n; ++t) {
            const double c = taps.t[t];
            acc[0] = __fma_rn(c, w0, acc[0]);
            acc[1] = __fma_rn(c, w1, acc[1]);
            acc[2] = __fma_rn(c, w2, acc[2]);
            acc[3] = __fma_rn(c, w3, acc[3]);
            w0 = w1; w1 = w2; w2 = w3;
            w3 = at(pos0 - r + t + 4);
        }
#pragma unroll
        for (int j = 0; j < kSmoothOut; ++j)
            if (pos0 + j < len) {
                double v = acc[j];
                if (AXIS == 1 && orig != nullptr && v != v) v = smooth_nan2d(orig, nx, ny, nv, ixo, pos0 + j, iv, taps_first, taps);
                out[base + j * step] = v;
            }
    }
}

// ---------------------------------------------------------------- f3: integer-factor spatial rebin
// losvd.py:69-92 via misc_io.py:383-394 (rebin_2darray): a.reshape(n0, fx, n1, fy).sum(-1).sum(1)
// per velocity slice, i.e. first the fy neighbours along Y, then the fx partial sums along X
// ('mean' divides by fy, then by fx, like the two chained .mean calls).
__global__ void __launch_bounds__(256) k_rebin_cube(const double* __restrict__ in, double* __restrict__ out,
                                                    int ny, int nv, int n0, int n1, int fx, int fy, int mean) {
    const int total = n0 * n1 * nv;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int k = i % nv, rest = i / nv;
        const int j = rest % n1, io = rest / n1;
        double acc = 0.0;
        for (int a = 0; a < fx; ++a) {
            const double* row = in + ((int64_t)(io * fx + a) * ny + (int64_t)j * fy) * nv + k;
            double s = 0.0;
            for (int b = 0; b < fy; ++b) s = __dadd_rn(s, row[(int64_t)b * nv]);
            if (mean) s = __ddiv_rn(s, (double)fy);
            acc = __dadd_rn(acc, s);
        }
        out[i] = mean ? __ddiv_rn(acc, (double)fx) : acc;
    }
}

// ---------------------------------------------------------------- f1: per-spaxel LOSVD moments
// losvd.py:41-51 + misc_functions.py:70-90: one warp per spaxel, shuffle reduction over iv.
// An all-zero LOSVD gives (0, 0, min(diff(binv))/3) (misc_functions.py:81-84) = empty_m2.
__global__ void __launch_bounds__(256) k_cube_moments(const double* __restrict__ cube, int nspax, int nv,
                                                      const double* __restrict__ binv, double empty_m2,
                                                      double* mom0, double* mom1, double* mom2) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nspax) return;
    const double* c = cube + (int64_t)warp * nv;
    double a0 = 0, a1 = 0, a2 = 0;
    bool any = false;
    for (int k = lane; k < nv; k += 32) {
        const double v = binv[k], f = c[k];
        any = any || (f != 0.0);
        a0 += f; a1 += v * f; a2 += v * v * f;
    }
    any = __any_sync(0xffffffffu, any);
    for (int o = 16; o; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) {
        double m0 = 0, m1 = 0, m2 = empty_m2;
        if (any) { m0 = a0; m1 = a1 / a0; m2 = sqrt(a2 / a0 - m1 * m1); }
        mom0[warp] = m0; mom1[warp] = m1; mom2[warp] = m2;
    }
}

// ---------------------------------------------------------------- f2: scattered cubic interpolation
// scipy.interpolate.griddata(method='cubic') + nearest-node fill (misc_functions.py:195-202), per
// particle: uniform cell -> candidate triangles -> barycentric test (scipy's eps = 100 * DBL_EPSILON)
// -> the 19-term Clough-Tocher cubic whose Bezier ordinates the host precomputed per triangle.
// Points outside the convex hull take the value of the nearest table node (nodes cached in shared
// memory: 636 x 24 B).
constexpr int kCtMaxNodes = 2048;

template <typename T>
__global__ void __launch_bounds__(256) k_ct_interp(const T* __restrict__ qx, const T* __restrict__ qy, int64_t n,
                                                   const double* __restrict__ xform, const double* __restrict__ coef,
                                                   const int* __restrict__ cell_start, const int* __restrict__ cell_tris,
                                                   int ncell, double lo0, double lo1, double inv0, double inv1,
                                                   const double* __restrict__ nodes, int nnode, int fill,
                                                   double* __restrict__ out) {
    extern __shared__ double s_nodes[];                  // [nnode][3] = x, y, value
    for (int i = threadIdx.x; i < nnode * 3; i += blockDim.x) s_nodes[i] = nodes[i];
    __syncthreads();
    const double eps = 100.0 * 2.220446049250313e-16;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double px = (double)qx[i], py = (double)qy[i];
        double res = __longlong_as_double(0x7ff8000000000000ll);   // NaN
        const double gx = floor((px - lo0) * inv0), gy = floor((py - lo1) * inv1);
        if (gx >= 0.0 && gx <= (double)ncell && gy >= 0.0 && gy <= (double)ncell) {   // false for NaN
            const int cx = min((int)gx, ncell - 1), cy = min((int)gy, ncell - 1);
            const int c = cx * ncell + cy;
            for (int k = cell_start[c]; k < cell_start[c + 1]; ++k) {
                const int t = cell_tris[k];
                const double* X = xform + (int64_t)t * 6;
                const double dx = px - X[4], dy = py - X[5];
                const double b0 = X[0] * dx + X[1] * dy, b1 = X[2] * dx + X[3] * dy, b2 = 1.0 - b0 - b1;
                if (fmin(b0, fmin(b1, b2)) >= -eps) {
                    const double* C = coef + (int64_t)t * 19;
                    const double mn = fmin(b0, fmin(b1, b2));
                    const double a1 = b0 - mn, a2 = b1 - mn, a3 = b2 - mn, a4 = 3.0 * mn;
                    res = a1 * a1 * a1 * C[0] + 3 * a1 * a1 * a2 * C[1] + 3 * a1 * a1 * a3 * C[2] + 3 * a1 * a1 * a4 * C[3]
                        + 3 * a1 * a2 * a2 * C[4] + 6 * a1 * a2 * a4 * C[5] + 3 * a1 * a3 * a3 * C[6] + 6 * a1 * a3 * a4 * C[7]
                        + 3 * a1 * a4 * a4 * C[8] + a2 * a2 * a2 * C[9] + 3 * a2 * a2 * a3 * C[10] + 3 * a2 * a2 * a4 * C[11]
                        + 3 * a2 * a3 * a3 * C[12] + 6 * a2 * a3 * a4 * C[13] + 3 * a2 * a4 * a4 * C[14] + a3 * a3 * a3 * C[15]
                        + 3 * a3 * a3 * a4 * C[16] + 3 * a3 * a4 * a4 * C[17] + a4 * a4 * a4 * C[18];
                    break;
                }
            }
        }
        if (fill && res != res && px == px && py == py) {   // outside the hull: nearest node
            double best = 1.7976931348623157e308;
            int arg = 0;
            for (int j = 0; j < nnode; ++j) {
                const double dx = px - s_nodes[3 * j], dy = py - s_nodes[3 * j + 1];
                const double d2 = dx * dx + dy * dy;
                if (d2 < best) { best = d2; arg = j; }
            }
            res = s_nodes[3 * arg + 2];
        }
        out[i] = res;
    }
}

// ---------------------------------------------------------------- memory-order probe
// Is the snapshot stored in a spatial order (tree / space-filling curve)?  Sample pairs of CONSECUTIVE particles
// (i, i + 1), i = k * stride: out[0] counts the pairs closer than `limit`, out[1] the pairs looked at.  "The median
// distance is below one pixel" <=> more than half of the pairs are closer than a pixel -- a count, no sort.
template <typename T>
__global__ void __launch_bounds__(256) k_order_probe(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ z,
                                                     int64_t n, int64_t stride, int64_t nsamples, double limit2,
                                                     unsigned long long* out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool close = false, looked = false;
    if (k < nsamples) {
        const int64_t i = k * stride;
        if (i + 1 < n) {
            const double dx = (double)x[i + 1] - (double)x[i], dy = (double)y[i + 1] - (double)y[i], dz = (double)z[i + 1] - (double)z[i];
            const double d2 = dx * dx + dy * dy + dz * dz;
            looked = true;
            close = d2 < limit2;                      // NaN distances are not close
        }
    }
    const unsigned c = __popc(__ballot_sync(0xffffffffu, close)), l = __popc(__ballot_sync(0xffffffffu, looked));
    if ((threadIdx.x & 31) == 0 && l) { atomicAdd(out, (unsigned long long)c); atomicAdd(out + 1, (unsigned long long)l); }
}

// ---------------------------------------------------------------- abs-max (fixed-point scale)
// Non-negative doubles order like their bit patterns, so one integer atomicMax per block suffices.
template <typename T>
__global__ void __launch_bounds__(256) k_absmax(const T* __restrict__ a, int64_t n, unsigned long long* out) {
    double m = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = fabs((double)a[i]);
        if (v > m && is_finite(v)) m = v;
    }
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) m = fmax(m, sm[k]);
        atomicMax(out, (unsigned long long)__double_as_longlong(m));
    }
}

}  // namespace pnm
