// abi_demo.cpp -- a compiled host that uses nothing but include/pnm_b200.h and the CUDA runtime:
// no Python, no PyTorch.  It runs the hot path (rotate + project + bin -> maps + LOSVD cube ->
// finalise -> smooth) on synthetic particles and checks the result against a plain double-precision
// loop written here with the reference's rules (forward FMA chain of np.dot, rotation.py:115-117;
// searchsorted 'right' with a closed last edge, grid.py:126-131 / :259).
//
//   nvcc -std=c++17 -I include examples/abi_demo.cpp -o examples/abi_demo -L pynmap_b200 -lpnm_b200 \
//        -Xlinker -rpath -Xlinker '$ORIGIN/../pynmap_b200'
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pnm_b200.h"

#define CK(call)                                                                        \
    do {                                                                                \
        if ((call) != 0) {                                                              \
            std::fprintf(stderr, "FAIL %s: %s\n", #call, pnm_last_error());             \
            return 1;                                                                   \
        }                                                                               \
    } while (0)
#define CU(call)                                                                        \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            std::fprintf(stderr, "FAIL %s: %s\n", #call, cudaGetErrorString(e_));       \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

static std::vector<double> edges(double lo, double hi, int n) {      // grid.py:119-124 (np.linspace)
    const double half = (hi - lo) / (n - 1) / 2.0, a = lo - half, b = hi + half, step = (b - a) / n;
    std::vector<double> e(n + 1);
    for (int i = 0; i <= n; ++i) e[i] = a + i * step;
    e[n] = b;
    return e;
}

static int bin_of(const std::vector<double>& e, double x) {          // searchsorted(e, x, 'right') - 1, last edge closed
    const int n = (int)e.size() - 1;
    if (!(x >= e[0]) || !(x <= e[n])) return -1;
    if (x == e[n]) return n - 1;
    return (int)(std::upper_bound(e.begin(), e.end(), x) - e.begin()) - 1;
}

static float row(const float* m, float x, float y, float z) { return fmaf(m[2], z, fmaf(m[1], y, m[0] * x)); }

int main() {
    if (pnm_device_count() < 1) { std::fprintf(stderr, "no sm_100 device\n"); return 2; }
    const int64_t n = 2000003;
    const int nx = 64, ny = 48, nv = 40;
    std::vector<float> h[7];
    for (auto& a : h) a.resize(n);
    uint64_t s = 88172645463325252ull;
    auto uni = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; };
    for (int64_t i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) h[k][i] = (float)((uni() + uni() + uni() - 1.5) * 8.0);
        for (int k = 3; k < 6; ++k) h[k][i] = (float)((uni() + uni() + uni() - 1.5) * 400.0);
        h[6][i] = (float)(0.5 + 1.5 * uni());
    }
    const float c = std::cos(0.5f), sn = std::sin(0.5f), ci = std::cos(1.0f), si = std::sin(1.0f);
    const float mat[9] = {c, -sn * ci, sn * si, sn, c * ci, -c * si, 0.f, si, ci};   // some rotation; any float32 matrix will do

    const auto ex = edges(-6.0, 6.0, nx), ey = edges(-5.0, 5.0, ny), ev = edges(-300.0, 300.0, nv);
    pnm_grid* g = nullptr;
    CK(pnm_grid_create(nx, ex.data(), ny, ey.data(), nv, ev.data(), &g));

    float* d[7];
    for (int k = 0; k < 7; ++k) {
        CU(cudaMalloc(&d[k], n * sizeof(float)));
        CU(cudaMemcpy(d[k], h[k].data(), n * sizeof(float), cudaMemcpyHostToDevice));
    }
    const int64_t me = pnm_maps_acc_elems(g), ce = pnm_cube_acc_elems(g), cme = pnm_cube_moments_acc_elems(g);
    double *acc_maps, *acc_cube, *M, *V, *S, *cube, *smooth, *work;
    CU(cudaMalloc(&acc_maps, me * 8)); CU(cudaMemset(acc_maps, 0, me * 8));
    CU(cudaMalloc(&acc_cube, cme * 8)); CU(cudaMemset(acc_cube, 0, cme * 8));   // interleaved cube: w per voxel + the moments
    CU(cudaMalloc(&M, nx * ny * 8)); CU(cudaMalloc(&V, nx * ny * 8)); CU(cudaMalloc(&S, nx * ny * 8));
    CU(cudaMalloc(&cube, ce * 8)); CU(cudaMalloc(&smooth, ce * 8)); CU(cudaMalloc(&work, ce * 8));

    const int sums = PNM_SUM_W | PNM_SUM_WD | PNM_SUM_WD2 | PNM_SUM_CUBE | PNM_W_FROM_CUBE | PNM_CUBE_MOMENTS;
    CK(pnm_project_bin(g, n, PNM_F32, d[0], d[1], d[2], d[3], d[4], d[5], d[6], nullptr, PNM_MASK_NONE, mat, 1, 1, sums,
                       PNM_ACC_F64, nullptr, acc_maps, acc_cube, nullptr));
    CK(pnm_finalize_vmaps(g, acc_maps, acc_cube, sums, PNM_ACC_F64, nullptr, M, V, S, nullptr, nullptr, nullptr));
    CK(pnm_finalize_cube(g, acc_cube, sums, PNM_ACC_F64, 1.0, cube, nullptr));
    const double taps[5] = {0.06136, 0.24477, 0.38774, 0.24477, 0.06136};
    CK(pnm_smooth_cube(cube, smooth, work, nx, ny, nv, taps, 5, taps, 5, nullptr));
    CU(cudaDeviceSynchronize());

    std::vector<double> hM(nx * ny), hV(nx * ny), hS(nx * ny), hC(ce), hSm(ce);
    CU(cudaMemcpy(hM.data(), M, nx * ny * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hV.data(), V, nx * ny * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hS.data(), S, nx * ny * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hC.data(), cube, ce * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hSm.data(), smooth, ce * 8, cudaMemcpyDeviceToHost));

    // the same thing in a plain loop
    std::vector<double> s0(nx * ny, 0.0), s1(nx * ny, 0.0), s2(nx * ny, 0.0), rc(ce, 0.0);
    for (int64_t i = 0; i < n; ++i) {
        const float px = row(mat, h[0][i], h[1][i], h[2][i]), py = row(mat + 3, h[0][i], h[1][i], h[2][i]);
        const float vz = row(mat + 6, h[3][i], h[4][i], h[5][i]), w = h[6][i];
        const int ix = bin_of(ex, px), iy = bin_of(ey, py);
        if (ix < 0 || iy < 0) continue;
        const int p = iy * nx + ix;                              // maps are (ny, nx): the reference's .T
        s0[p] += w; s1[p] += (double)(w * vz); s2[p] += (double)(w * (vz * vz));
        const int iv = bin_of(ev, vz);
        if (iv >= 0) rc[((int64_t)ix * ny + iy) * nv + iv] += w; // cube[ix][iy][iv]
    }
    double worst = 0.0, worst_c = 0.0, tot = 0.0, tot_s = 0.0;
    for (int p = 0; p < nx * ny; ++p) {
        worst = std::max(worst, std::fabs(hM[p] - s0[p]) / std::max(s0[p], 1e-300));
        if (s0[p] > 0) {
            const double v = s1[p] / s0[p];
            worst = std::max(worst, std::fabs(hV[p] - v) / 400.0);
            worst = std::max(worst, std::fabs(hS[p] * hS[p] - (s2[p] / s0[p] - v * v)) / 160000.0);
        }
    }
    for (int64_t q = 0; q < ce; ++q) {
        worst_c = std::max(worst_c, std::fabs(hC[q] - rc[q]) / std::max(rc[q], 1.0));
        tot += hC[q]; tot_s += hSm[q];
    }
    std::printf("maps: max deviation %.2e   cube: max deviation %.2e   sum(cube) %.6f   sum(smoothed) %.6f\n",
                worst, worst_c, tot, tot_s);
    const bool ok = worst < 1e-9 && worst_c < 1e-9 && tot > 0 && tot_s <= tot * (1 + 1e-12) && tot_s > 0.5 * tot;
    for (auto p : d) cudaFree(p);
    cudaFree(acc_maps); cudaFree(acc_cube); cudaFree(M); cudaFree(V); cudaFree(S); cudaFree(cube); cudaFree(smooth); cudaFree(work);
    pnm_grid_destroy(g);
    std::printf(ok ? "PASS\n" : "FAIL\n");
    return ok ? 0 : 1;
}
