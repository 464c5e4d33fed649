// pnm_fit.cuh -- row f4: batched Gauss-Hermite fit of every spaxel's LOSVD (sm_100a only).
//
// Reference: LOSVD.fitlosvd (pynmap/losvd.py:53-67) loops fitgh (pynmap/fit_losvd.py:57-235) over
// the spaxels; each call minimises, over p = (V, S, h3, .. h_deg),
//     chi2(p) = sum_c ((d_c - A(p) u_c(p)) / e_c)^2,   u = gausshermite(x, [1, p])   (misc_functions.py:15-68)
//     A(p)    = sum (d/e)^2 / sum (d/e)(u/e)                                          (fit_losvd.py:25-54)
// inside the box fit_losvd.py:157-165, starting from the moments of the profile, with a vendored
// MINPACK port (utils/mpfit.py).  The optimiser is tolerance-defined; what is reproduced here is
// the objective, the start, the limits, mpfit's pegging rule at the limits (mpfit.py:1074-1094,
// :1171-1196), its convergence tests (mpfit.py:1303-1322) and status codes.  The Jacobian is
// analytic instead of mpfit's one-sided differences.
//
// One warp per spaxel: the nv channels are strided over the lanes, every reduction is a butterfly
// (all lanes hold bit-identical scalars, so control flow stays warp-uniform) and the NP x NP
// damped normal equations are solved redundantly in registers.  FP64 throughout.
#pragma once
#include "pnm_common.cuh"

namespace pnm {

constexpr int kFitMaxPar = 6;

struct FitParams {
    const double* cube;     // [nspax][nv]
    const double* err;      // [nspax][nv] or null (fit_losvd.py:180-185)
    const double* binv;     // [nv]
    int32_t nv;
    int64_t nspax;
    double start[kFitMaxPar];        // NaN: the default (moments for V, S; guess_h for h_n) -- fit_losvd.py:157-158, :169
    double lo[kFitMaxPar];           // -inf / +inf when not limited (fit_losvd.py:159-165, :170-174)
    double hi[kFitMaxPar];
    uint32_t fixed_mask;             // bit k: parameter k is held at its start value (fit_losvd.py:166, :170)
    double guess_h;                  // 0.02 (fit_losvd.py:157)
    double empty_m2;                 // min(diff(binv)) / 3 (misc_functions.py:84)
    double ftol, xtol, gtol;
    int32_t maxiter;
    double* gh;             // [NP + 1][nspax]
    double* bestfit;        // [nspax][nv] or null
    int32_t* status;        // [nspax]
    int32_t* niter;         // [nspax]
    double* chi2;           // [nspax]
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// u(x; p) with unit amplitude and its derivatives with respect to every parameter.
template <int NP, bool DERIV>
__device__ __forceinline__ double gh_basis(double x, const double (&p)[NP], double (&du)[NP]) {
    const double kR2 = 1.4142135623730951, kR3 = 1.7320508075688772;
    const double w = (x - p[0]) / p[1];
    const double w2 = w * w;
    double h_prev = (2.0 * w2 - 1.0) / kR2, d_prev = 4.0 * w / kR2;
    double h_cur = (2.0 * w2 - 3.0) * w / kR3, d_cur = (6.0 * w2 - 3.0) / kR3;
    double var = 1.0, dvar = 0.0;
    const double e = exp(-0.5 * w2);
#pragma unroll
    for (int k = 2; k < NP; ++k) {            // p[k] multiplies H_{k+1}
        var += p[k] * h_cur;
        if (DERIV) {
            dvar += p[k] * d_cur;
            du[k] = h_cur * e;
        }
        const double s = sqrt((double)(k + 2));
        const double h_next = (kR2 * h_cur * w - h_prev) / s;   // misc_functions.py:62 (no sqrt(i) on h_prev)
        const double d_next = (kR2 * (d_cur * w + h_cur) - d_prev) / s;
        h_prev = h_cur; d_prev = d_cur;
        h_cur = h_next; d_cur = d_next;
    }
    if (DERIV) {
        const double dudw = (dvar - w * var) * e;
        du[0] = -dudw / p[1];
        du[1] = -dudw * w / p[1];
    }
    return var * e;
}

template <int NP>
struct FitState {
    double f;                          // chi2
    double amp;                        // A(p)
    double g[NP];                      // J^T r
    double h[NP][NP];                  // J^T J (full, symmetric)
};

// chi2 and the amplitude only.
template <int NP>
__device__ __forceinline__ void fit_chi2(const FitParams& P, const double* d, const double* er, int lane,
                                         const double (&p)[NP], double d2, double& amp, double& f) {
    double dummy[NP];
    double sdu = 0.0;
    for (int c = lane; c < P.nv; c += 32) {
        double wgt = 1.0;
        if (er) { const double e = er[c]; wgt = e == 0.0 ? 1.0 : 1.0 / e; }
        sdu += d[c] * gh_basis<NP, false>(P.binv[c], p, dummy) * wgt * wgt;
    }
    sdu = warp_sum(sdu);
    amp = d2 / sdu;
    double acc = 0.0;
    for (int c = lane; c < P.nv; c += 32) {
        double wgt = 1.0;
        if (er) { const double e = er[c]; wgt = e == 0.0 ? 1.0 : 1.0 / e; }
        const double r = (d[c] - amp * gh_basis<NP, false>(P.binv[c], p, dummy)) * wgt;
        acc += r * r;
    }
    f = warp_sum(acc);
}

// chi2, gradient and Gauss-Newton matrix.
template <int NP>
__device__ __forceinline__ void fit_eval(const FitParams& P, const double* d, const double* er, int lane,
                                         const double (&p)[NP], double d2, FitState<NP>& st) {
    double du[NP];
    double sdu = 0.0, sddu[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) sddu[k] = 0.0;
    for (int c = lane; c < P.nv; c += 32) {
        double wgt = 1.0;
        if (er) { const double e = er[c]; wgt = e == 0.0 ? 1.0 : 1.0 / e; }
        const double dw = d[c] * wgt * wgt;
        const double u = gh_basis<NP, true>(P.binv[c], p, du);
        sdu += dw * u;
#pragma unroll
        for (int k = 0; k < NP; ++k) sddu[k] += dw * du[k];
    }
    sdu = warp_sum(sdu);
    const double amp = d2 / sdu;
    double ak[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) ak[k] = -amp * warp_sum(sddu[k]) / sdu;    // dA/dp_k

    double f = 0.0;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        st.g[k] = 0.0;
#pragma unroll
        for (int l = 0; l <= k; ++l) st.h[k][l] = 0.0;
    }
    for (int c = lane; c < P.nv; c += 32) {
        double wgt = 1.0;
        if (er) { const double e = er[c]; wgt = e == 0.0 ? 1.0 : 1.0 / e; }
        const double u = gh_basis<NP, true>(P.binv[c], p, du);
        const double r = (d[c] - amp * u) * wgt;
        f += r * r;
        double j[NP];
#pragma unroll
        for (int k = 0; k < NP; ++k) j[k] = -(ak[k] * u + amp * du[k]) * wgt;
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            st.g[k] += j[k] * r;
#pragma unroll
            for (int l = 0; l <= k; ++l) st.h[k][l] += j[k] * j[l];
        }
    }
    st.f = warp_sum(f);
    st.amp = amp;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        st.g[k] = warp_sum(st.g[k]);
#pragma unroll
        for (int l = 0; l <= k; ++l) {
            st.h[k][l] = warp_sum(st.h[k][l]);
            st.h[l][k] = st.h[k][l];
        }
    }
}

// Solve (H + lam diag^2) s = -g on the free parameters (Cholesky); false if not positive definite.
template <int NP>
__device__ __forceinline__ bool fit_solve(const FitState<NP>& st, const double (&diag)[NP], const bool (&peg)[NP],
                                          double lam, double (&s)[NP]) {
    double m[NP][NP], b[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        b[k] = peg[k] ? 0.0 : -st.g[k];
#pragma unroll
        for (int l = 0; l < NP; ++l) m[k][l] = (peg[k] || peg[l]) ? 0.0 : st.h[k][l];
        m[k][k] = peg[k] ? 1.0 : st.h[k][k] + lam * diag[k] * diag[k];
    }
    bool ok = true;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        double piv = m[k][k];
#pragma unroll
        for (int q = 0; q < k; ++q) piv -= m[k][q] * m[k][q];
        if (!(piv > 0.0) || !isfinite(piv)) { ok = false; piv = 1.0; }
        piv = sqrt(piv);
        m[k][k] = piv;
#pragma unroll
        for (int l = k + 1; l < NP; ++l) {
            double v = m[l][k];
#pragma unroll
            for (int q = 0; q < k; ++q) v -= m[l][q] * m[k][q];
            m[l][k] = v / piv;
        }
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) {            // L y = b
        double v = b[k];
#pragma unroll
        for (int q = 0; q < k; ++q) v -= m[k][q] * s[q];
        s[k] = v / m[k][k];
    }
#pragma unroll
    for (int k = NP - 1; k >= 0; --k) {       // L^T s = y
        double v = s[k];
#pragma unroll
        for (int q = k + 1; q < NP; ++q) v -= m[q][k] * s[q];
        s[k] = v / m[k][k];
    }
    return ok;
}

constexpr int kFitWarps = 4;

template <int NP>
__global__ void __launch_bounds__(kFitWarps * 32) k_fit_gh(const FitParams P) {
    const int lane = threadIdx.x & 31;
    const int64_t spax = (int64_t)blockIdx.x * kFitWarps + (threadIdx.x >> 5);
    if (spax >= P.nspax) return;
    const double* d = P.cube + spax * P.nv;
    const double* er = P.err ? P.err + spax * P.nv : nullptr;

    // ---- start: moment012 (misc_functions.py:70-90) and the limits --------------------------
    double m0 = 0.0, mx = 0.0, mxx = 0.0, d2 = 0.0;
    bool any = false;
    for (int c = lane; c < P.nv; c += 32) {
        const double v = d[c], x = P.binv[c];
        double wgt = 1.0;
        if (er) { const double e = er[c]; wgt = e == 0.0 ? 1.0 : 1.0 / e; }
        any |= (v != 0.0);
        m0 += v; mx += x * v; mxx += x * x * v;
        d2 += (v * wgt) * (v * wgt);
    }
    any = __any_sync(0xffffffffu, any);
    m0 = warp_sum(m0); mx = warp_sum(mx); mxx = warp_sum(mxx); d2 = warp_sum(d2);

    double p[NP], lo[NP], hi[NP];
    bool fixed[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        p[k] = P.guess_h; lo[k] = P.lo[k]; hi[k] = P.hi[k];
        fixed[k] = (P.fixed_mask >> k) & 1u;
    }
    if (any) {
        p[0] = mx / m0;
        p[1] = sqrt(mxx / m0 - p[0] * p[0]);
    } else {
        p[0] = 0.0;
        p[1] = P.empty_m2;
    }
#pragma unroll
    for (int k = 0; k < NP; ++k)
        if (P.start[k] == P.start[k]) p[k] = P.start[k];

    int status = 0, niter = 1;
    double amp = 0.0, f = 0.0;
    bool outside = false, bad = !any || !(d2 > 0.0);
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        bad |= !isfinite(p[k]);
        outside |= (p[k] < lo[k]) || (p[k] > hi[k]);
    }
    FitState<NP> st;
    if (bad) {
        status = -16;                   // 0/0 amplitude: mpfit ends with -16 at the start values (mpfit.py:1330-1336)
        amp = nan("");
        f = nan("");
    } else if (outside) {
        status = 0;                     // 'parameters are not within PARINFO limits' (mpfit.py:955-958)
        amp = nan("");
        f = nan("");
#pragma unroll
        for (int k = 0; k < NP; ++k) p[k] = nan("");
    } else {
        fit_eval<NP>(P, d, er, lane, p, d2, st);
        if (!isfinite(st.f)) status = -16;
    }

    // ---- Levenberg-Marquardt ------------------------------------------------------------------
    if (!bad && !outside && status == 0) {
        double diag[NP];
#pragma unroll
        for (int k = 0; k < NP; ++k) diag[k] = 0.0;
        double lam = 1e-3, nu = 2.0;
        int evals = 0;
        const int max_evals = 8 * P.maxiter + 64;
        while (status == 0) {
            bool peg[NP];
            double gnorm = 0.0;
            const double fnorm = sqrt(st.f);
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                peg[k] = fixed[k] || (p[k] == lo[k] && st.g[k] > 0.0) || (p[k] == hi[k] && st.g[k] < 0.0);
                const double hd = sqrt(st.h[k][k]);
                if (!peg[k] && hd > 0.0 && st.f > 0.0) gnorm = fmax(gnorm, fabs(st.g[k]) / (hd * fnorm));
                diag[k] = fmax(diag[k], hd > 0.0 ? hd : 1.0);
            }
            if (gnorm <= P.gtol) { status = 4; break; }
            if (P.maxiter == 0) { status = 5; break; }

            while (true) {                              // inner loop: until a step is accepted
                double s[NP];
                const bool ok = fit_solve<NP>(st, diag, peg, lam, s);
                ++evals;
                if (!ok) {
                    lam *= nu; nu *= 2.0;
                    if (evals >= max_evals || !isfinite(lam)) { status = 5; break; }
                    continue;
                }
                double alpha = 1.0;
#pragma unroll
                for (int k = 0; k < NP; ++k) {          // mpfit.py:1171-1196
                    if (p[k] == lo[k]) s[k] = fmax(s[k], 0.0);
                    if (p[k] == hi[k]) s[k] = fmin(s[k], 0.0);
                    if (s[k] != 0.0) {
                        if (p[k] + s[k] < lo[k]) alpha = fmin(alpha, (lo[k] - p[k]) / s[k]);
                        if (p[k] + s[k] > hi[k]) alpha = fmin(alpha, (hi[k] - p[k]) / s[k]);
                    }
                }
                double pn[NP], pnorm = 0.0, xnorm = 0.0, gs = 0.0, shs = 0.0;
                const double eps = 2.220446049250313e-16;
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                    s[k] *= alpha;
                    pn[k] = p[k] + s[k];
                    const double sgu = hi[k] >= 0.0 ? 1.0 : -1.0, sgl = lo[k] >= 0.0 ? 1.0 : -1.0;
                    if (pn[k] >= hi[k] * (1.0 - sgu * eps) - (hi[k] == 0.0 ? eps : 0.0)) pn[k] = hi[k];   // mpfit.py:1218-1230
                    if (pn[k] <= lo[k] * (1.0 + sgl * eps) + (lo[k] == 0.0 ? eps : 0.0)) pn[k] = lo[k];
                    if (!fixed[k]) {                    // mpfit's norms run over the free parameters
                        pnorm += diag[k] * s[k] * diag[k] * s[k];
                        xnorm += diag[k] * p[k] * diag[k] * p[k];
                    }
                    gs += st.g[k] * s[k];
                }
#pragma unroll
                for (int k = 0; k < NP; ++k)
#pragma unroll
                    for (int l = 0; l < NP; ++l) shs += s[k] * st.h[k][l] * s[l];
                pnorm = sqrt(pnorm); xnorm = sqrt(xnorm);

                double amp_n, f_n;
                fit_chi2<NP>(P, d, er, lane, pn, d2, amp_n, f_n);
                if (!isfinite(f_n)) { status = -16; break; }
                const double actred = (0.01 * f_n < st.f) ? 1.0 - f_n / st.f : -1.0;
                const double prered = -(2.0 * gs + shs) / st.f;
                const double ratio = prered != 0.0 ? actred / prered : 0.0;
                const bool accept = ratio >= 1e-4;
                if (accept) {
#pragma unroll
                    for (int k = 0; k < NP; ++k) p[k] = pn[k];
                    ++niter;
                    const double t = 2.0 * ratio - 1.0;
                    lam *= fmax(1.0 / 3.0, 1.0 - t * t * t);
                    nu = 2.0;
                    fit_eval<NP>(P, d, er, lane, p, d2, st);
                    xnorm = 0.0;
#pragma unroll
                    for (int k = 0; k < NP; ++k)
                        if (!fixed[k]) xnorm += diag[k] * p[k] * diag[k] * p[k];
                    xnorm = sqrt(xnorm);
                } else {
                    lam *= nu; nu *= 2.0;
                }
                const bool ftest = fabs(actred) <= P.ftol && prered <= P.ftol && 0.5 * ratio <= 1.0;   // mpfit.py:1303-1310
                if (ftest) status = 1;
                if (pnorm <= P.xtol * xnorm) status = ftest ? 3 : 2;
                if (status != 0) break;
                if (niter >= P.maxiter || evals >= max_evals) { status = 5; break; }
                if (accept) break;
            }
        }
        fit_chi2<NP>(P, d, er, lane, p, d2, amp, f);
    }

    // ---- outputs ------------------------------------------------------------------------------
    if (lane == 0) {
        P.gh[spax] = amp;
#pragma unroll
        for (int k = 0; k < NP; ++k) P.gh[(int64_t)(k + 1) * P.nspax + spax] = p[k];
        P.status[spax] = status;
        P.niter[spax] = niter;
        P.chi2[spax] = f;
    }
    if (P.bestfit) {
        double dummy[NP];
        for (int c = lane; c < P.nv; c += 32)
            P.bestfit[spax * P.nv + c] = amp * gh_basis<NP, false>(P.binv[c], p, dummy);
    }
}

}  // namespace pnm
