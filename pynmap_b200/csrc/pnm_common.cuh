// pnm_common.cuh -- shared device helpers for libpnm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pnm_b200.h"

namespace pnm {

constexpr int kNumSMs = 148;       // B200: 2 dies x 74 SMs
constexpr int kMaxMats = 64;       // nviews * nchain matrices carried in kernel parameters

// Slot order inside a 32-byte map record.  (S1, S2) share the first, 16-byte aligned half so that
// one 16-byte TMA bulk reduction updates both; S0 is touched rarely when the cube carries sum w.
constexpr int kSlotS1 = 0;         // sum w*d
constexpr int kSlotS2 = 1;         // sum w*d*d
constexpr int kSlotS0 = 2;         // sum w (or its out-of-V-range remainder)

// --------------------------------------------------------------------------------------------
// Rotation arithmetic.  np.dot(mat(3x3, f32), pos.T(3xN)) in the reference (rotation.py:115-117)
// rounds each output as a forward FMA chain over k = 0,1,2 (SURVEY.md 8c KAT 7).  Explicit
// intrinsics so that nvcc can neither contract differently nor re-associate.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ float rot_row(float m0, float m1, float m2, float x, float y, float z) {
    return __fmaf_rn(m2, z, __fmaf_rn(m1, y, __fmul_rn(m0, x)));
}
__device__ __forceinline__ double rot_row(float m0, float m1, float m2, double x, double y, double z) {
    return __fma_rn((double)m2, z, __fma_rn((double)m1, y, __dmul_rn((double)m0, x)));
}

// products of the weighted moments are rounded in the particle dtype before widening
// (grid.py:128 `wm * vm`, grid.py:131 `wm * vm**2`)
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

__device__ __forceinline__ bool is_finite(float v) { return fabsf(v) <= 3.402823466e+38f; }
__device__ __forceinline__ bool is_finite(double v) { return fabs(v) <= 1.7976931348623157e+308; }

// --------------------------------------------------------------------------------------------
// Bin tables.  `thr` holds n+1 thresholds in the particle dtype such that bin i is exactly
// thr[i] <= x < thr[i+1]  <=>  edges[i] <= (double)x < edges[i+1]  (numpy searchsorted 'right'),
// with thr[n] nudged one ulp up so that the closed last edge of np.histogramdd
// (_histograms_impl.py:1049-1053) is covered by the same test; NaN fails every test.  The lookup
// itself (uniform guess + exact table search near edges) lives in pnm_bin.cuh.
// --------------------------------------------------------------------------------------------

// --------------------------------------------------------------------------------------------
// Accumulation.  8-byte accumulators, two arithmetics:
//   F64   : RED.E.ADD.F64.RN at L2 (round-to-nearest, no flush-to-zero, unlike the f32 RED)
//   FIXED : RED.E.ADD.64 on rint(value * 2^k): integer addition is associative, so the result
//           is independent of atomic ordering and of how particles are split over GPUs.
// atomicAdd with an unused result compiles to REDG (no return trip).  See red_add / to_acc in
// pnm_bin.cuh.
// --------------------------------------------------------------------------------------------

}  // namespace pnm
