/*
 * pnm_oracle.c -- TEST INFRASTRUCTURE ONLY.  Plain-C restatement of pynmap's particle-projection
 * path, used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER for
 * libpnm_b200.so.  The product never links or calls this file.
 *
 * Each function cites the reference lines it restates (paths relative to emsellem/pynmap).  The
 * reference's arithmetic lives in numpy (unpinned by the reference; numpy 2.3.5 in the build
 * container): np.dot (rotation.py:115-117), np.histogram2d / np.histogramdd (grid.py:126-131,
 * :200-201, :259; numpy/lib/_histograms_impl.py:1040-1071) and astropy.convolution.convolve
 * (losvd.py:127-135; astropy is absent here, its documented behaviour is restated -> smoothing
 * parity is "unpinned").  tests/test_oracle_golden.py pins every function below against outputs
 * of the reference's own code (tests/golden/*.npz, made by oracle/make_golden.py).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -shared -fPIC (see oracle/build_oracle.py); contraction
 * must stay off so that only the explicit fma()/fmaf() calls fuse.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PNMO_F32 0
#define PNMO_F64 1
#define PNMO_ACC_F64 0
#define PNMO_ACC_FIXED 1
#define PNMO_SUM_W 1
#define PNMO_SUM_WD 2
#define PNMO_SUM_WD2 4
#define PNMO_SUM_CUBE 8
#define PNMO_W_FROM_CUBE 16
#define PNMO_MASK_AND_NONFINITE 1
#define PNMO_MASK_PLAIN 2

/* ---- rotation.py:115-117: np.dot(mat, pos.T) rounds as a forward FMA chain over k = 0, 1, 2 ---- */
static float row32(const float* m, float x, float y, float z) { return fmaf(m[2], z, fmaf(m[1], y, m[0] * x)); }
static double row64(const float* m, double x, double y, double z) {
    return fma((double)m[2], z, fma((double)m[1], y, (double)m[0] * x));
}

void pnmo_rotate(int dtype, const void* x, const void* y, const void* z, void* ox, void* oy, void* oz,
                 int64_t n, const float* m) {
    for (int64_t i = 0; i < n; ++i) {
        if (dtype == PNMO_F32) {
            float a = ((const float*)x)[i], b = ((const float*)y)[i], c = ((const float*)z)[i];
            ((float*)ox)[i] = row32(m, a, b, c);
            ((float*)oy)[i] = row32(m + 3, a, b, c);
            ((float*)oz)[i] = row32(m + 6, a, b, c);
        } else {
            double a = ((const double*)x)[i], b = ((const double*)y)[i], c = ((const double*)z)[i];
            ((double*)ox)[i] = row64(m, a, b, c);
            ((double*)oy)[i] = row64(m + 3, a, b, c);
            ((double*)oz)[i] = row64(m + 6, a, b, c);
        }
    }
}

/* ---- _histograms_impl.py:1040-1053: searchsorted(edges, v, 'right'); v == edges[-1] joins the
 * last bin; index 0 / n+1 (and NaN, which sorts past the end) are dropped.  Returns -1 if dropped. */
static int bin_of(double v, const double* e, int n) {
    if (isnan(v)) return -1;
    int lo = 0, hi = n + 1;                 /* first index with e[idx] > v */
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (e[mid] <= v) lo = mid + 1; else hi = mid;
    }
    if (v == e[n]) lo -= 1;
    if (lo == 0 || lo == n + 1) return -1;
    return lo - 1;
}

void pnmo_bin_index(int dtype, const void* v, int64_t n, const double* edges, int nbins, int32_t* out) {
    for (int64_t i = 0; i < n; ++i) {
        double d = dtype == PNMO_F32 ? (double)((const float*)v)[i] : ((const double*)v)[i];
        out[i] = bin_of(d, edges, nbins);
    }
}

/* int64 fixed-point image of a term: rint(value * 2^k) with round-half-even; non-finite -> 0
 * (same convention as pnm_common.cuh acc_add<FIXED>) */
static int64_t quantise(double value, double scale) {
    double s = value * scale;
    if (!isfinite(s)) return 0;
    return (int64_t)llrint(s);
}

static void add_term(void* acc, int64_t idx, double value, int acc_mode, double scale) {
    if (acc_mode == PNMO_ACC_F64) ((double*)acc)[idx] += value;                    /* np.bincount: float64, particle order */
    else ((int64_t*)acc)[idx] = (int64_t)((uint64_t)((int64_t*)acc)[idx] + (uint64_t)quantise(value, scale));
}

/* ---- fused path: rotate_mat chain (rotation.py:73-119) -> x' y' v_los -> grid.py:126-131 sums and
 * grid.py:252-259 cube.  fused == 0: x, y, d(=vx) are already projected (points_2* entry points).
 *   acc_maps [view][iy][ix][4], acc_cube [view][ix][iy][iv]  (8-byte elements)                     */
void pnmo_project_bin(int fused, int64_t n, int dtype,
                      const void* x, const void* y, const void* z, const void* vx, const void* vy, const void* vz,
                      const void* w, const uint8_t* mask, int mask_mode,
                      const float* mats, int nchain, int nviews,
                      int nx, const double* ex, int ny, const double* ey, int nv, const double* ev,
                      int sums, int acc_mode, const double* scales, void* acc_maps, void* acc_cube) {
    const int64_t maps_stride = (int64_t)nx * ny * 4, cube_stride = (int64_t)nx * ny * nv;
    const double sc0 = scales ? scales[0] : 1.0, sc1 = scales ? scales[1] : 1.0, sc2 = scales ? scales[2] : 1.0;
    for (int64_t i = 0; i < n; ++i) {
        for (int view = 0; view < nviews; ++view) {
            double px, py, d, wd, wd2, wt;
            if (dtype == PNMO_F32) {
                float ax = ((const float*)x)[i], ay = ((const float*)y)[i], az = fused ? ((const float*)z)[i] : 0.f;
                float bx = ((const float*)vx)[i], by = fused ? ((const float*)vy)[i] : 0.f, bz = fused ? ((const float*)vz)[i] : 0.f;
                float dd = bx;
                if (fused) {
                    const float* m = mats + (size_t)view * nchain * 9;
                    for (int c = 0; c < nchain; ++c, m += 9) {
                        float t0 = row32(m, ax, ay, az), t1 = row32(m + 3, ax, ay, az), t2 = row32(m + 6, ax, ay, az);
                        float u0 = row32(m, bx, by, bz), u1 = row32(m + 3, bx, by, bz), u2 = row32(m + 6, bx, by, bz);
                        ax = t0; ay = t1; az = t2; bx = u0; by = u1; bz = u2;
                    }
                    dd = bz;
                }
                float ww = w ? ((const float*)w)[i] : 1.f;
                float p1 = ww * dd;              /* grid.py:128  wm * vm      (input dtype)  */
                float sq = dd * dd;              /* grid.py:131  vm**2                        */
                float p2 = ww * sq;              /*              wm * vm**2                   */
                px = ax; py = ay; d = dd; wt = ww; wd = p1; wd2 = p2;
            } else {
                double ax = ((const double*)x)[i], ay = ((const double*)y)[i], az = fused ? ((const double*)z)[i] : 0.;
                double bx = ((const double*)vx)[i], by = fused ? ((const double*)vy)[i] : 0., bz = fused ? ((const double*)vz)[i] : 0.;
                double dd = bx;
                if (fused) {
                    const float* m = mats + (size_t)view * nchain * 9;
                    for (int c = 0; c < nchain; ++c, m += 9) {
                        double t0 = row64(m, ax, ay, az), t1 = row64(m + 3, ax, ay, az), t2 = row64(m + 6, ax, ay, az);
                        double u0 = row64(m, bx, by, bz), u1 = row64(m + 3, bx, by, bz), u2 = row64(m + 6, bx, by, bz);
                        ax = t0; ay = t1; az = t2; bx = u0; by = u1; bz = u2;
                    }
                    dd = bz;
                }
                double ww = w ? ((const double*)w)[i] : 1.;
                px = ax; py = ay; d = dd; wt = ww; wd = ww * dd; wd2 = ww * (dd * dd);
            }
            int masked = mask && mask_mode && mask[i];
            if (mask_mode == PNMO_MASK_AND_NONFINITE) masked = masked && !isfinite(d);   /* grid.py:114-115 */
            if (masked) continue;
            int ix = bin_of(px, ex, nx), iy = bin_of(py, ey, ny);
            if (ix < 0 || iy < 0) continue;
            int iv = (sums & PNMO_SUM_CUBE) ? bin_of(d, ev, nv) : -1;
            int64_t pix = ((int64_t)iy * nx + ix) * 4 + view * maps_stride;              /* .T -> [iy][ix] */
            if ((sums & PNMO_SUM_W) && !((sums & PNMO_W_FROM_CUBE) && iv >= 0)) add_term(acc_maps, pix, wt, acc_mode, sc0);
            if (sums & PNMO_SUM_WD) add_term(acc_maps, pix + 1, wd, acc_mode, sc1);
            if (sums & PNMO_SUM_WD2) add_term(acc_maps, pix + 2, wd2, acc_mode, sc2);
            if (iv >= 0) add_term(acc_cube, ((int64_t)ix * ny + iy) * nv + iv + view * cube_stride, wt, acc_mode, sc0);
        }
    }
}

/* ---- grid.py:133-138 (and the W-from-cube variant of pnm_finalize_vmaps) ---- */
void pnmo_finalize_vmaps(const void* acc_maps, const void* acc_cube, int nx, int ny, int nv, int sums, int acc_mode,
                         const double* scales, double* M, double* V, double* S) {
    for (int pix = 0; pix < nx * ny; ++pix) {
        double s0, s1, s2;
        int iy = pix / nx, ix = pix - iy * nx;
        if (acc_mode == PNMO_ACC_F64) {
            const double* a = (const double*)acc_maps + (int64_t)pix * 4;
            s0 = a[0]; s1 = a[1]; s2 = a[2];
            if (sums & PNMO_W_FROM_CUBE)
                for (int k = 0; k < nv; ++k) s0 += ((const double*)acc_cube)[((int64_t)ix * ny + iy) * nv + k];
        } else {
            const int64_t* a = (const int64_t*)acc_maps + (int64_t)pix * 4;
            int64_t t = a[0];
            if (sums & PNMO_W_FROM_CUBE)
                for (int k = 0; k < nv; ++k) t += ((const int64_t*)acc_cube)[((int64_t)ix * ny + iy) * nv + k];
            s0 = (double)t * (1.0 / scales[0]); s1 = (double)a[1] * (1.0 / scales[1]); s2 = (double)a[2] * (1.0 / scales[2]);
        }
        double v = 0.0, s = 0.0;
        if (s0 != 0.0) {
            v = s1 / s0;
            double q = s2 / s0, vv = v * v;
            s = sqrt(q - vv);
        }
        M[pix] = s0; V[pix] = v; S[pix] = s;
    }
}

/* ---- losvd.py:127-135: astropy convolve(slice, Gaussian2DKernel) with its defaults boundary='fill', fill_value=0,
 * nan_treatment='interpolate', normalize_kernel=True (astropy/convolution/convolve.py and its C loop, restated from
 * the documentation: astropy is absent here): direct 'same' convolution over the zero-padded slice; NaN samples are
 * skipped and every output is divided by the sum of the kernel values it used (padding counts: 0 is a number); an
 * output all of whose samples are NaN keeps the input value.  ker[ka][kb] (ka along cube axis 0), already normalised. */
void pnmo_smooth_cube(const double* in, double* out, int nx, int ny, int nv, const double* ker, int ka, int kb) {
    const int ha = ka / 2, hb = kb / 2;
    for (int ix = 0; ix < nx; ++ix)
        for (int iy = 0; iy < ny; ++iy)
            for (int k = 0; k < nv; ++k) {
                double top = 0.0, bot = 0.0;
                for (int a = 0; a < ka; ++a) {
                    int sx = ix + ha - a;                 /* true convolution: kernel flipped */
                    for (int b = 0; b < kb; ++b) {
                        int sy = iy + hb - b;
                        double val = 0.0;
                        if (sx >= 0 && sx < nx && sy >= 0 && sy < ny) val = in[((int64_t)sx * ny + sy) * nv + k];
                        if (!isnan(val)) { top += ker[a * kb + b] * val; bot += ker[a * kb + b]; }
                    }
                }
                out[((int64_t)ix * ny + iy) * nv + k] = bot == 0.0 ? in[((int64_t)ix * ny + iy) * nv + k] : top / bot;
            }
}

/* ---- losvd.py:69-92 + misc_io.py:383-394: reshape(n0, fx, n1, fy).sum(-1).sum(1) per slice ---- */
void pnmo_rebin_cube(const double* in, double* out, int nx, int ny, int nv, int fx, int fy, int mean) {
    const int n0 = nx / fx, n1 = ny / fy;
    (void)nx;
    for (int i = 0; i < n0; ++i)
        for (int j = 0; j < n1; ++j)
            for (int k = 0; k < nv; ++k) {
                double acc = 0.0;
                for (int a = 0; a < fx; ++a) {
                    double s = 0.0;
                    for (int b = 0; b < fy; ++b) s += in[((int64_t)(i * fx + a) * ny + (j * fy + b)) * nv + k];
                    if (mean) s /= (double)fy;
                    acc += s;
                }
                out[((int64_t)i * n1 + j) * nv + k] = mean ? acc / (double)fx : acc;
            }
}

/* ---- losvd.py:41-51 + misc_functions.py:70-90 ---- */
void pnmo_cube_moments(const double* cube, int nx, int ny, int nv, const double* binv, double* m0, double* m1, double* m2) {
    double mind = INFINITY;
    for (int k = 0; k + 1 < nv; ++k) if (binv[k + 1] - binv[k] < mind) mind = binv[k + 1] - binv[k];
    for (int s = 0; s < nx * ny; ++s) {
        const double* c = cube + (int64_t)s * nv;
        double a0 = 0, a1 = 0, a2 = 0;
        int any = 0;
        for (int k = 0; k < nv; ++k) { any |= (c[k] != 0.0); a0 += c[k]; a1 += binv[k] * c[k]; a2 += binv[k] * binv[k] * c[k]; }
        if (any) { m0[s] = a0; m1[s] = a1 / a0; m2[s] = sqrt(a2 / a0 - m1[s] * m1[s]); }
        else { m0[s] = 0; m1[s] = 0; m2[s] = mind / 3.0; }        /* misc_functions.py:81-84 */
    }
}
