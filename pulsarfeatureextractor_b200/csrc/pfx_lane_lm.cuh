// pfx_lane_lm.cuh -- Levenberg-Marquardt for ONE LANE PER FIT with a streaming QR.
//
// The first version of the Bates kernels gave a whole warp to one candidate and kept the m x n
// Jacobian in shared memory (pfx_lm.cuh).  ncu showed that design issue-starved: n is 2..8, so
// most of the time one lane worked on the n x n sub-problem while 31 waited, the code ran once per
// iteration in long straight lines (instruction-cache misses were the top stall) and only a dozen
// warps fitted per SM.  Here every lane owns a fit, so the n x n work of 32 fits runs in SIMD
// lock-step, and the Jacobian is never stored: MINPACK's lmdif needs only R, the pivoting and the
// first n components of Q^T f, and those are accumulated row by row with Givens rotations while
// the residual rows are evaluated (one pass over the data per Jacobian, state in registers).
//
//   jac_begin()                     steps h_j = eps |x_j|  (fdjac2)
//   jac_row(a[N], b)                rotate row (J_i, f_i) into the triangular factor
//   jac_end()                       column-pivoted QR of the N x N factor with MINPACK's own qrfac
//                                   rules (same pivot order as qrfac on the full Jacobian in exact
//                                   arithmetic), Q^T f, scaled-gradient test, diag update
//   trial_begin(xnew)               lmpar -> step p, xnew = x + p
//   trial_end(|f(xnew)|)            gain ratio, trust-region update, convergence tests
//
// The control flow, constants and tests are lmdif's (see pfx_lm.cuh for the line-by-line
// correspondence); only the way R and Q^T f are obtained differs, which perturbs results at the
// rounding level exactly like the tree-ordered sums of the warp version did.
#pragma once

#include "pfx_lm.cuh"

namespace pfx {

// MINPACK qrfac (pivoting) on an n x n matrix a (column major, leading dimension NPMAX), applied
// to the triangular factor accumulated by the row rotations.  acnorm/rdiag/wa are preset by the
// caller to the streamed column norms of the full Jacobian.
template <class RV, class IV>
PFX_HD PFX_NOINLINE void qrfac_small(int n, int ld, RV a, IV ipvt, double* rdiag, double* wa) {
    for (int j = 0; j < n; ++j) ipvt[j] = j;
    for (int j = 0; j < n; ++j) {
        int kmax = j;
        for (int k = j; k < n; ++k)
            if (rdiag[k] > rdiag[kmax]) kmax = k;
        if (kmax != j) {
            for (int i = 0; i < n; ++i) {
                double t = a[i + ld * j];
                a[i + ld * j] = a[i + ld * kmax];
                a[i + ld * kmax] = t;
            }
            rdiag[kmax] = rdiag[j];
            wa[kmax] = wa[j];
            int t = ipvt[j];
            ipvt[j] = ipvt[kmax];
            ipvt[kmax] = t;
        }
        double s = 0.0;
        for (int i = j; i < n; ++i) s += a[i + ld * j] * a[i + ld * j];
        double ajnorm = sqrt(s);
        if (ajnorm != 0.0) {
            if (a[j + ld * j] < 0.0) ajnorm = -ajnorm;
            for (int i = j; i < n; ++i) a[i + ld * j] /= ajnorm;
            a[j + ld * j] += 1.0;
            for (int k = j + 1; k < n; ++k) {
                double sum = 0.0;
                for (int i = j; i < n; ++i) sum += a[i + ld * j] * a[i + ld * k];
                double temp = sum / a[j + ld * j];
                for (int i = j; i < n; ++i) a[i + ld * k] -= temp * a[i + ld * j];
                if (rdiag[k] != 0.0) {
                    temp = a[j + ld * k] / rdiag[k];
                    rdiag[k] *= sqrt(fmax(0.0, 1.0 - temp * temp));
                    double q = rdiag[k] / wa[k];
                    if (0.05 * (q * q) <= LM_EPSMCH) {
                        double s2 = 0.0;
                        for (int i = j + 1; i < n; ++i) s2 += a[i + ld * k] * a[i + ld * k];
                        rdiag[k] = sqrt(s2);
                        wa[k] = rdiag[k];
                    }
                }
            }
        }
        rdiag[j] = -ajnorm;
    }
}

enum LanePhase { LANE_JAC = 0, LANE_TRIAL = 1, LANE_DONE = 2 };

// per-lane workspace of the n x n sub-problem: r (n x n, column major), six n-vectors, ipvt
PFX_HD inline int lane_ws_doubles(int n) { return n * n + 6 * n + (n + 1) / 2; }
struct LaneWs {
    SPtr<double> r, diag, qtf, step, sdiag, wa1, wa2;
    SPtr<int> ipvt;
    // base points at this lane's first element; consecutive elements are `stride` doubles apart
    PFX_HD void carve(double* base, int stride, int n) {
        auto at = [&](int k) { SPtr<double> v; v.p = base + (size_t)k * stride; v.stride = stride; return v; };
        r = at(0);
        diag = at(n * n);
        qtf = at(n * n + n);
        step = at(n * n + 2 * n);
        sdiag = at(n * n + 3 * n);
        wa1 = at(n * n + 4 * n);
        wa2 = at(n * n + 5 * n);
        // ints are packed two per double slot, still strided per lane
        ipvt.p = (int*)(base + (size_t)(n * n + 6 * n) * stride);
        ipvt.stride = 2 * stride;
    }
};

template <int N>
struct LaneLM {
    // persistent over the whole solve
    double x[N], diag[N];
    double par, delta, xnorm, fnorm, gnorm;
    int iter, nfev, info;
    // Jacobian pass accumulators
    double h[N];
    double R[N * (N + 1) / 2];   // packed upper triangle, row i starts at i*N - i*(i-1)/2
    double q[N];
    double acn2[N];
    double f2;
    // n x n sub-problem: dynamically indexed small vectors in the lane's strided workspace
    // (shared memory on the device), see LaneWs
    double acnorm[N];
    // trial point
    double xnew[N];
    double pnorm, rpnorm;

    PFX_HD static constexpr int ridx(int i, int j) { return i * N - (i * (i - 1)) / 2 + (j - i); }

    PFX_HD void init(const double (&x0)[N]) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            x[j] = x0[j];
            diag[j] = 0.0;
        }
        par = 0.0;
        delta = 0.0;
        xnorm = 0.0;
        fnorm = 0.0;
        gnorm = 0.0;
        iter = 1;
        nfev = 1;          // the initial evaluation of lmdif; its values are the f_i of the first pass
        info = 0;
    }

    PFX_HD void jac_begin() {
        const double eps = 1.4901161193847656e-08;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double t = eps * fabs(x[j]);
            h[j] = (t == 0.0) ? eps : t;
            q[j] = 0.0;
            acn2[j] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < N * (N + 1) / 2; ++k) R[k] = 0.0;
        f2 = 0.0;
    }

    // a[j] = d f_i / d x_j (forward difference), b = f_i
    PFX_HD void jac_row(double (&a)[N], double b) {
        f2 = PFX_FMA(b, b, f2);
#pragma unroll
        for (int j = 0; j < N; ++j) acn2[j] = PFX_FMA(a[j], a[j], acn2[j]);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double aj = a[j];
            if (aj != 0.0) {
                const double rjj = R[ridx(j, j)];
                const double rr = PFX_FMA(rjj, rjj, aj * aj);
#if defined(__CUDA_ARCH__)
                const double inv = rsqrt(rr);
#else
                const double inv = 1.0 / sqrt(rr);
#endif
                const double c = rjj * inv, sn = aj * inv;
                R[ridx(j, j)] = rr * inv;
#pragma unroll
                for (int k = j + 1; k < N; ++k) {
                    const double rjk = R[ridx(j, k)], ak = a[k];
                    R[ridx(j, k)] = PFX_FMA(c, rjk, sn * ak);
                    a[k] = PFX_FMA(c, ak, -sn * rjk);
                }
                const double qj = q[j];
                q[j] = PFX_FMA(c, qj, sn * b);
                b = PFX_FMA(c, b, -sn * qj);
            }
        }
    }

    // returns LANE_TRIAL, or LANE_DONE with info set
    PFX_HD int jac_end(const LaneWs& s, int m) {
        nfev += N;
        if (iter == 1) fnorm = sqrt(f2);
        double rdiag[N], wa[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            acnorm[j] = sqrt(acn2[j]);
            rdiag[j] = acnorm[j];
            wa[j] = acnorm[j];
        }
        // unpack the triangular factor; rows beyond m do not exist (m >= N is guaranteed upstream)
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int i = 0; i < N; ++i) s.r[i + N * j] = (i <= j) ? R[ridx(i < j ? i : j, j)] : 0.0;
        (void)m;
        qrfac_small(N, N, s.r, s.ipvt, rdiag, wa);
        if (iter == 1) {
            double sq = 0.0;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                diag[j] = acnorm[j];
                if (acnorm[j] == 0.0) diag[j] = 1.0;
                double t = diag[j] * x[j];
                sq += t * t;
            }
            xnorm = sqrt(sq);
            delta = 100.0 * xnorm;
            if (delta == 0.0) delta = 100.0;
        }
        // Q^T f: apply the N Householder vectors stored in the lower part of s.r to q
        double w4[N];
#pragma unroll
        for (int j = 0; j < N; ++j) w4[j] = q[j];
        for (int j = 0; j < N; ++j) {
            const double ajj = s.r[j + N * j];
            if (ajj != 0.0) {
                double sum = 0.0;
                for (int i = j; i < N; ++i) sum += s.r[i + N * j] * w4[i];
                double temp = -sum / ajj;
                for (int i = j; i < N; ++i) w4[i] += s.r[i + N * j] * temp;
            }
            s.r[j + N * j] = rdiag[j];
            s.qtf[j] = w4[j];
        }
        gnorm = 0.0;
        if (fnorm != 0.0) {
            for (int j = 0; j < N; ++j) {
                const int l = s.ipvt[j];
                const double al = acnorm[l];
                if (al != 0.0) {
                    double sum = 0.0;
                    for (int i = 0; i <= j; ++i) sum += s.r[i + N * j] * (s.qtf[i] / fnorm);
                    gnorm = fmax(gnorm, fabs(sum / al));
                }
            }
        }
        if (gnorm <= 0.0) {     // gtol = 0
            info = 4;
            return LANE_DONE;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            diag[j] = fmax(diag[j], acnorm[j]);
            s.diag[j] = diag[j];
        }
        return LANE_TRIAL;
    }

    PFX_HD void trial_begin(const LaneWs& s) {
        lmpar_seq(N, N, s.r, s.ipvt, s.diag, s.qtf, delta, par, s.step, s.sdiag, s.wa1, s.wa2);
        for (int j = 0; j < N; ++j) s.wa1[j] = 0.0;
        for (int j = 0; j < N; ++j) {
            double temp = -s.step[s.ipvt[j]];
            for (int i = 0; i <= j; ++i) s.wa1[i] += s.r[i + N * j] * temp;
        }
        rpnorm = enorm_seq(N, s.wa1);
        double sq = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double p = -s.step[j];
            xnew[j] = x[j] + p;
            double t = diag[j] * p;
            sq += t * t;
        }
        pnorm = sqrt(sq);
        if (iter == 1) delta = fmin(delta, pnorm);
    }

    // fnorm1 = |f(xnew)|.  Returns LANE_JAC (step accepted), LANE_TRIAL (retry with a smaller
    // region) or LANE_DONE.
    PFX_HD int trial_end(double fnorm1) {
        const double p1 = 0.1, p5 = 0.5, p25 = 0.25, p75 = 0.75, p0001 = 1e-4;
        const double ftol = LM_TOL, xtol = LM_TOL, epsmch = LM_EPSMCH;
        const int maxfev = 200 * (N + 1);
        ++nfev;
        double actred = -1.0;
        if (p1 * fnorm1 < fnorm) {
            double t = fnorm1 / fnorm;
            actred = 1.0 - t * t;
        }
        const double temp1 = rpnorm / fnorm;
        const double temp2 = (sqrt(par) * pnorm) / fnorm;
        const double prered = temp1 * temp1 + temp2 * temp2 / p5;
        const double dirder = -(temp1 * temp1 + temp2 * temp2);
        double ratio = 0.0;
        if (prered != 0.0) ratio = actred / prered;
        if (ratio <= p25) {
            double temp = p5;
            if (actred >= 0.0) temp = p5;
            if (actred < 0.0) temp = p5 * dirder / (dirder + p5 * actred);
            if (p1 * fnorm1 >= fnorm || temp < p1) temp = p1;
            delta = temp * fmin(delta, pnorm / p1);
            par = par / temp;
        } else if (par == 0.0 || ratio >= p75) {
            delta = pnorm / p5;
            par = p5 * par;
        }
        if (ratio >= p0001) {
            double sq = 0.0;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                x[j] = xnew[j];
                double t = diag[j] * x[j];
                sq += t * t;
            }
            xnorm = sqrt(sq);
            fnorm = fnorm1;
            ++iter;
        }
        const bool c1 = fabs(actred) <= ftol && prered <= ftol && p5 * ratio <= 1.0;
        if (c1) info = 1;
        if (delta <= xtol * xnorm) info = 2;
        if (c1 && info == 2) info = 3;
        if (info != 0) return LANE_DONE;
        if (nfev >= maxfev) info = 5;
        if (fabs(actred) <= epsmch && prered <= epsmch && p5 * ratio <= 1.0) info = 6;
        if (delta <= epsmch * xnorm) info = 7;
        if (gnorm <= epsmch) info = 8;
        if (info != 0) return LANE_DONE;
        return (ratio >= p0001) ? LANE_JAC : LANE_TRIAL;
    }
};

}  // namespace pfx
