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

#include <string.h>

#include "pfx_lm.cuh"

namespace pfx {

// MINPACK qrfac (pivoting) on an n x n matrix a (column major, leading dimension NPMAX), applied
// to the triangular factor accumulated by the row rotations.  acnorm/rdiag/wa are preset by the
// caller to the streamed column norms of the full Jacobian.
// (rdiag, wa and the scratch vector acn0 may live in registers/local memory or in the lane's workspace)
template <class RV, class IV, class DV>
PFX_HD PFX_NOINLINE void qrfac_small(int n, int ld, RV a, IV ipvt, DV rdiag, DV wa, DV acn0) {
    // acn0: column norms of the full Jacobian (permuted along with rdiag)
    PFX_ROLL
    for (int j = 0; j < n; ++j) {
        ipvt[j] = j;
        acn0[j] = rdiag[j];
    }
    PFX_ROLL
    for (int j = 0; j < n; ++j) {
        int kmax = j;
        PFX_ROLL
        for (int k = j; k < n; ++k)
            if (rdiag[k] > rdiag[kmax]) kmax = k;
        if (kmax != j) {
            PFX_ROLL
            for (int i = 0; i < n; ++i) {
                double t = a[i + ld * j];
                a[i + ld * j] = a[i + ld * kmax];
                a[i + ld * kmax] = t;
            }
            rdiag[kmax] = rdiag[j];
            wa[kmax] = wa[j];
            acn0[kmax] = acn0[j];
            int t = ipvt[j];
            ipvt[j] = ipvt[kmax];
            ipvt[kmax] = t;
        }
        double s = 0.0;
        PFX_ROLL
        for (int i = j; i < n; ++i) s += a[i + ld * j] * a[i + ld * j];
        double ajnorm = sqrt(s);
        if (ajnorm != 0.0) {
            if (a[j + ld * j] < 0.0) ajnorm = -ajnorm;
            PFX_ROLL
            for (int i = j; i < n; ++i) a[i + ld * j] /= ajnorm;
            a[j + ld * j] += 1.0;
            PFX_ROLL
            for (int k = j + 1; k < n; ++k) {
                double sum = 0.0;
                PFX_ROLL
                for (int i = j; i < n; ++i) sum += a[i + ld * j] * a[i + ld * k];
                double temp = sum / a[j + ld * j];
                PFX_ROLL
                for (int i = j; i < n; ++i) a[i + ld * k] -= temp * a[i + ld * j];
                if (rdiag[k] != 0.0) {
                    temp = a[j + ld * k] / rdiag[k];
                    rdiag[k] *= sqrt(fmax(0.0, 1.0 - temp * temp));
                    double q = rdiag[k] / wa[k];
                    if (0.05 * (q * q) <= LM_EPSMCH) {
                        double s2 = 0.0;
                        PFX_ROLL
                        for (int i = j + 1; i < n; ++i) s2 += a[i + ld * k] * a[i + ld * k];
                        rdiag[k] = sqrt(s2);
                        // Exact rank deficiency.  On the full Jacobian MINPACK leaves EXACT zeros behind
                        // when the columns j, k, ... are non-zero in one and the same row only (a model
                        // that touches a single data point: every other DM trial clipped, a Gaussian
                        // narrower than a bin), and lmpar then sees r[k][k] == 0.0 and does not move
                        // those parameters.  The triangular factor accumulated block by block carries
                        // rounding residues of a few ulp of the column norm instead, which would turn
                        // into a step along a noise direction: a remainder below 1e-13 of the column's
                        // original norm is such a residue and is flushed to the exact zero.
                        if (rdiag[k] <= 1e-13 * acn0[k]) {
                            PFX_ROLL
                            for (int i = j + 1; i < n; ++i) a[i + ld * k] = 0.0;
                            rdiag[k] = 0.0;
                        }
                        wa[k] = rdiag[k];
                    }
                }
            }
        }
        rdiag[j] = -ajnorm;
    }
}

// The same factorisation with everything in registers (n <= 4): every loop unrolled, the pivot column
// chosen by predicated swaps.  Statement for statement qrfac_small (same sums in the same order, the same
// divisions), so the results are bit-identical; what it removes is the local-memory traffic of the
// dynamically indexed rdiag / wa / acn0 (ncu on k_gauss_rounds_s: 12.5 % of all samples in qrfac_small,
// two thirds of them waiting for local loads that miss the small L1 left beside 222 KB of shared memory).
// a[j][i] = element (row i, column j).
template <int N>
PFX_HD inline void qrfac_reg(double (&a)[N][N], int (&ipvt)[N], double (&rdiag)[N], double (&wa)[N]) {
    double acn0[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        ipvt[j] = j;
        acn0[j] = rdiag[j];
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
        int kmax = j;
        double vmax = rdiag[j];
#pragma unroll
        for (int k = j + 1; k < N; ++k)
            if (rdiag[k] > vmax) {
                kmax = k;
                vmax = rdiag[k];
            }
#pragma unroll
        for (int k = j + 1; k < N; ++k)
            if (kmax == k) {
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double t = a[j][i];
                    a[j][i] = a[k][i];
                    a[k][i] = t;
                }
                rdiag[k] = rdiag[j];
                wa[k] = wa[j];
                acn0[k] = acn0[j];
                const int t = ipvt[j];
                ipvt[j] = ipvt[k];
                ipvt[k] = t;
            }
        double s = 0.0;
#pragma unroll
        for (int i = j; i < N; ++i) s = PFX_FMA(a[j][i], a[j][i], s);
        double ajnorm = sqrt(s);
        if (ajnorm != 0.0) {
            if (a[j][j] < 0.0) ajnorm = -ajnorm;
#pragma unroll
            for (int i = j; i < N; ++i) a[j][i] /= ajnorm;
            a[j][j] = PFX_ADD(a[j][j], 1.0);
#pragma unroll
            for (int k = j + 1; k < N; ++k) {
                double sum = 0.0;
#pragma unroll
                for (int i = j; i < N; ++i) sum = PFX_FMA(a[j][i], a[k][i], sum);
                double temp = sum / a[j][j];
#pragma unroll
                for (int i = j; i < N; ++i) a[k][i] = PFX_FMA(-temp, a[j][i], a[k][i]);
                if (rdiag[k] != 0.0) {
                    temp = a[k][j] / rdiag[k];
                    rdiag[k] = PFX_MUL(rdiag[k], sqrt(fmax(0.0, PFX_FMA(-temp, temp, 1.0))));
                    double q = rdiag[k] / wa[k];
                    if (PFX_MUL(0.05, PFX_MUL(q, q)) <= LM_EPSMCH) {
                        double s2 = 0.0;
#pragma unroll
                        for (int i = j + 1; i < N; ++i) s2 = PFX_FMA(a[k][i], a[k][i], s2);
                        rdiag[k] = sqrt(s2);
                        if (rdiag[k] <= PFX_MUL(1e-13, acn0[k])) {      // exact rank deficiency, see qrfac_small
#pragma unroll
                            for (int i = j + 1; i < N; ++i) a[k][i] = 0.0;
                            rdiag[k] = 0.0;
                        }
                        wa[k] = rdiag[k];
                    }
                }
            }
        }
        rdiag[j] = -ajnorm;
    }
}

// ---------------------------------------------------------------------------------------------
// C lanes of a warp share one fit (C = 1, 2, 4, ... 32).  The rows of the residual vector are
// dealt to the C lanes, every lane accumulates the triangular factor of its own rows, and the C
// factors are merged by a butterfly (LaneLM::jac_combine) after which all lanes of the group hold
// bit-identical state and run the n x n sub-problem redundantly.  A slow fit then finishes C
// times sooner (the kernels end in a tail of slow candidates) and C times more warps are
// resident for the same shared-memory footprint.  On the host (logic check) C is 1.
template <int C>
struct LaneGroup {
    static constexpr int size = C;
#if defined(__CUDA_ARCH__)
    PFX_HD static int rank() { return (C == 1) ? 0 : (int)(threadIdx.x & (C - 1)); }
    PFX_HD static unsigned mask() {
        return (C == 32) ? 0xffffffffu : (((1u << (C & 31)) - 1u) << (threadIdx.x & 31 & ~(C - 1)));
    }
    PFX_HD static double xchg(double v, int s) { return (C == 1) ? v : __shfl_xor_sync(mask(), v, s); }
    PFX_HD static double bcast(double v) { return (C == 1) ? v : __shfl_sync(mask(), v, 0, C); }
    PFX_HD static long long bcast(long long v) { return (C == 1) ? v : __shfl_sync(mask(), v, 0, C); }
    PFX_HD static int bcast(int v) { return (C == 1) ? v : __shfl_sync(mask(), v, 0, C); }
    PFX_HD static void sync() { if (C > 1) __syncwarp(mask()); }
#else
    // host build (logic check): groups of one lane only; nvcc's host pass merely has to parse this
    PFX_HD static int rank() { return 0; }
    PFX_HD static double xchg(double v, int) { return v; }
    PFX_HD static double bcast(double v) { return v; }
    PFX_HD static long long bcast(long long v) { return v; }
    PFX_HD static int bcast(int v) { return v; }
    PFX_HD static void sync() {}
#endif
    // IEEE addition and max are commutative, so both partners of an exchange get the same bits
    PFX_HD static double sum(double v) {
#pragma unroll
        for (int s = 1; s < C; s <<= 1) v = PFX_ADD(v, xchg(v, s));
        return v;
    }
    PFX_HD static double max_gt(double v) {     // NaN-free inputs
#pragma unroll
        for (int s = 1; s < C; s <<= 1) {
            const double o = xchg(v, s);
            v = (o > v) ? o : v;
        }
        return v;
    }
    PFX_HD static int sum(int v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
        for (int s = 1; s < C; s <<= 1) v += __shfl_xor_sync(mask(), v, s);
#endif
        return v;
    }
};

enum LanePhase { LANE_JAC = 0, LANE_TRIAL = 1, LANE_DONE = 2 };

// reciprocal / reciprocal square root for the orthogonal transformations (not for the model or
// for MINPACK's own formulas): a few ulp are irrelevant there and the short dependency chain is
// what bounds the latency of a Jacobian pass
PFX_HD inline double fast_rcp(double x) {
#if defined(__CUDA_ARCH__)
    return __drcp_rn(x);
#else
    return 1.0 / x;
#endif
}
// exp(a) for a <= 0 (or NaN), branch free: the lane kernels spend most of their time in
// exp(-t^2/2), and the branches of the library exp()/division fragment the instruction stream
// (instruction-fetch stalls were the top stall reason in ncu).  Cody-Waite reduction, degree-13
// polynomial, two-step scaling so that denormal results come out right.  Measured against glibc
// on 2e7 arguments in [-760, 0]: max error 1 ulp (scripts/micro/exp_check.cpp).
PFX_HD inline double exp_neg(double a) {
    const double L2E = 1.4426950408889634074, LN2_HI = 6.93147180369123816490e-01,
                 LN2_LO = 1.90821492927058770002e-10;
    const double ac = (a < -745.5) ? -745.5 : a;
    const double nf = rint(ac * L2E);
    double r = fma(-nf, LN2_HI, ac);
    r = fma(-nf, LN2_LO, r);
    double p = 1.6059043836821613e-10;
    p = fma(p, r, 2.08767569878681e-09);
    p = fma(p, r, 2.505210838544172e-08);
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 1.984126984126984e-04);
    p = fma(p, r, 1.388888888888889e-03);
    p = fma(p, r, 8.333333333333333e-03);
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int n = (int)nf;
    const int n1 = n / 2, n2 = n - n1;
#if defined(__CUDA_ARCH__)
    const double s1 = __hiloint2double((n1 + 1023) << 20, 0), s2 = __hiloint2double((n2 + 1023) << 20, 0);
#else
    double s1, s2;
    {
        unsigned long long b1 = (unsigned long long)(n1 + 1023) << 52, b2 = (unsigned long long)(n2 + 1023) << 52;
        memcpy(&s1, &b1, 8);
        memcpy(&s2, &b2, 8);
    }
#endif
    const double res = (p * s1) * s2;
    return (a != a) ? a : res;
}

PFX_HD inline double fast_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

// ---------------------------------------------------------------------------------------------
// Register-resident lmpar / qrsolv for small N (fully unrolled, static indexing; the permutation
// is applied with select chains).  Same algorithm, branches and sign conventions as MINPACK
// (pfx_lm.cuh: lmpar_seq / qrsolv_seq); the Givens rotations are formed from one reciprocal
// square root instead of two divisions and a square root, which shortens the dependency chain
// that bounds the latency of the trust-region solve (measured: the shared-memory version took
// ~3e5 cycles per call, more than a whole Jacobian pass).
#if defined(PFX_COUNT_LMPAR) && !defined(__CUDA_ARCH__)
static long long g_lmpar_hist[2][12];
#endif
template <int N>
struct SmallLM {
    double r[N][N];      // r[i][j]: upper triangle = R, strictly lower = S^T (qrsolv work)
    int ipvt[N];
    double diag[N], qtb[N];

    PFX_HD static double sel(const double (&a)[N], int idx) {
        double v = a[0];
#pragma unroll
        for (int k = 1; k < N; ++k) v = (k == idx) ? a[k] : v;
        return v;
    }
    PFX_HD static void put(double (&a)[N], int idx, double v) {
#pragma unroll
        for (int k = 0; k < N; ++k) a[k] = (k == idx) ? v : a[k];
    }
    PFX_HD static double nrm(const double (&a)[N]) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) s += a[k] * a[k];
        return sqrt(s);
    }

    // qrsolv: d = diagonal in pivoted order is looked up through ipvt
    PFX_HD void qrsolv(const double (&d)[N], double (&x)[N], double (&sdiag)[N]) {
        double wa[N], rd[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
#pragma unroll
            for (int i = j + 1; i < N; ++i) r[i][j] = r[j][i];
            rd[j] = r[j][j];
            wa[j] = qtb[j];
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double dl = sel(d, ipvt[j]);
            if (dl != 0.0) {
#pragma unroll
                for (int k = j; k < N; ++k) sdiag[k] = 0.0;
                sdiag[j] = dl;
                double qtbpj = 0.0;
#pragma unroll
                for (int k = j; k < N; ++k) {
                    const double sk = sdiag[k];
                    if (sk != 0.0) {
                        const double rkk = r[k][k];
                        const double h2 = PFX_FMA(rkk, rkk, sk * sk);
                        double c, sn;
                        if (h2 > 1e-280 && h2 < 1e280) {
                            const double inv = fast_rsqrt(h2);
                            if (fabs(rkk) >= fabs(sk)) {          // tan branch: c = |a|/h, s = sign(a) b/h
                                c = fabs(rkk) * inv;
                                sn = (rkk < 0.0 ? -sk : sk) * inv;
                            } else {                               // cotan branch: s = |b|/h, c = sign(b) a/h
                                sn = fabs(sk) * inv;
                                c = (sk < 0.0 ? -rkk : rkk) * inv;
                            }
                        } else if (fabs(rkk) >= fabs(sk)) {
                            const double t = sk / rkk;
                            c = 0.5 / sqrt(0.25 + 0.25 * (t * t));
                            sn = c * t;
                        } else {
                            const double ct = rkk / sk;
                            sn = 0.5 / sqrt(0.25 + 0.25 * (ct * ct));
                            c = sn * ct;
                        }
                        r[k][k] = c * rkk + sn * sk;
                        const double temp = c * wa[k] + sn * qtbpj;
                        qtbpj = -sn * wa[k] + c * qtbpj;
                        wa[k] = temp;
#pragma unroll
                        for (int i = k + 1; i < N; ++i) {
                            const double rik = r[i][k];
                            const double t2 = c * rik + sn * sdiag[i];
                            sdiag[i] = -sn * rik + c * sdiag[i];
                            r[i][k] = t2;
                        }
                    }
                }
            }
            sdiag[j] = r[j][j];
            r[j][j] = rd[j];
        }
        int nsing = N;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            if (sdiag[j] == 0.0 && nsing == N) nsing = j;
            if (nsing < N) wa[j] = 0.0;
        }
#pragma unroll
        for (int kk = 0; kk < N; ++kk) {
            const int j = N - 1 - kk;
            if (j < nsing) {
                double sum = 0.0;
#pragma unroll
                for (int i = j + 1; i < N; ++i)
                    if (i < nsing) sum += r[i][j] * wa[i];
                wa[j] = (wa[j] - sum) / sdiag[j];
            }
        }
#pragma unroll
        for (int j = 0; j < N; ++j) put(x, ipvt[j], wa[j]);
    }

    // lmpar in two pieces (see lmpar_begin / lmpar_iter in pfx_lm.cuh); x = step (un-negated), returns through par
    PFX_HD bool lmpar_begin(double delta, double& par, double (&x)[N], LmparState& st) {
        const double p1 = 0.1;
        double wa1[N], wa2[N];
        int nsing = N;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            wa1[j] = qtb[j];
            if (r[j][j] == 0.0 && nsing == N) nsing = j;
            if (nsing < N) wa1[j] = 0.0;
        }
#pragma unroll
        for (int kk = 0; kk < N; ++kk) {
            const int j = N - 1 - kk;
            if (j < nsing) {
                wa1[j] = wa1[j] / r[j][j];
                const double temp = wa1[j];
#pragma unroll
                for (int i = 0; i < j; ++i) wa1[i] -= r[i][j] * temp;
            }
        }
#pragma unroll
        for (int j = 0; j < N; ++j) put(x, ipvt[j], wa1[j]);
        st.iter = 0;
#pragma unroll
        for (int j = 0; j < N; ++j) wa2[j] = diag[j] * x[j];
        const double dxnorm = nrm(wa2);
        const double fp = dxnorm - delta;
        if (fp <= p1 * delta) {
#if defined(PFX_COUNT_LMPAR) && !defined(__CUDA_ARCH__)
            g_lmpar_hist[0][0]++;
#endif
            par = 0.0;
            return true;
        }
        double parl = 0.0;
        if (nsing >= N) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const int l = ipvt[j];
                wa1[j] = sel(diag, l) * (sel(wa2, l) / dxnorm);
            }
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double sum = 0.0;
#pragma unroll
                for (int i = 0; i < j; ++i) sum += r[i][j] * wa1[i];
                wa1[j] = (wa1[j] - sum) / r[j][j];
            }
            const double temp = nrm(wa1);
            parl = ((fp / delta) / temp) / temp;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double sum = 0.0;
#pragma unroll
            for (int i = 0; i <= j; ++i) sum += r[i][j] * qtb[i];
            wa1[j] = sum / sel(diag, ipvt[j]);
        }
        const double gnorm = nrm(wa1);
        double paru = gnorm / delta;
        if (paru == 0.0) paru = LM_DWARF / fmin(delta, p1);
        par = fmax(par, parl);
        par = fmin(par, paru);
        if (par == 0.0) par = gnorm / dxnorm;
        st.parl = parl;
        st.paru = paru;
        st.fp = fp;
        return false;
    }
    // one iteration on a freshly loaded sub-problem (qrsolv leaves its factor below the diagonal of r)
    PFX_HD bool lmpar_iter(double delta, double& par, double (&x)[N], double (&sdiag)[N], LmparState& st) {
        const double p1 = 0.1, p001 = 0.001;
        double wa1[N], wa2[N];
        ++st.iter;
        if (par == 0.0) par = fmax(LM_DWARF, p001 * st.paru);
        double temp = sqrt(par);
#pragma unroll
        for (int j = 0; j < N; ++j) wa1[j] = temp * diag[j];
        qrsolv(wa1, x, sdiag);
#pragma unroll
        for (int j = 0; j < N; ++j) wa2[j] = diag[j] * x[j];
        const double dxnorm = nrm(wa2);
        temp = st.fp;
        const double fp = dxnorm - delta;
        st.fp = fp;
        if (fabs(fp) <= p1 * delta || (st.parl == 0.0 && fp <= temp && temp < 0.0) || st.iter == 10) {
#if defined(PFX_COUNT_LMPAR) && !defined(__CUDA_ARCH__)
            g_lmpar_hist[0][st.iter]++;
#endif
            return true;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int l = ipvt[j];
            wa1[j] = sel(diag, l) * (sel(wa2, l) / dxnorm);
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            wa1[j] = wa1[j] / sdiag[j];
            const double t = wa1[j];
#pragma unroll
            for (int i = j + 1; i < N; ++i) wa1[i] -= r[i][j] * t;
        }
        temp = nrm(wa1);
        const double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0) st.parl = fmax(st.parl, par);
        if (fp < 0.0) st.paru = fmin(st.paru, par);
        par = fmax(st.parl, par + parc);
        return false;
    }
    PFX_HD void lmpar(double delta, double& par, double (&x)[N], double (&sdiag)[N]) {
        LmparState st;
        if (lmpar_begin(delta, par, x, st)) return;
        while (!lmpar_iter(delta, par, x, sdiag, st)) {
        }
    }
};

// per-lane workspace of the n x n sub-problem: r (n x n, column major), six n-vectors, ipvt
#ifndef PFX_BLOCK_ROWS
#define PFX_BLOCK_ROWS 8
#endif
constexpr int LANE_BLOCK_ROWS = PFX_BLOCK_ROWS;
// r, six vectors, ipvt (one int per float64 slot), and the staged row block of the Jacobian pass
// The staged block is only live during the row loop of a Jacobian pass and the n x n vectors only
// after it, so they share the same storage.
PFX_HD inline int lane_ws_doubles(int n) {
    const int a = n * n + 7 * n, b = LANE_BLOCK_ROWS * (n + 1);
    return a > b ? a : b;
}
struct LaneWs {
    SPtr<double> r, diag, qtf, step, sdiag, wa1, wa2, blk, stg;
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
        // one int per float64 slot (low half), strided like the rest
        ipvt.p = (int*)(base + (size_t)(n * n + 6 * n) * stride);
        ipvt.stride = 2 * stride;
        blk = at(0);      // aliases r .. ipvt (see lane_ws_doubles)
        // hybrid Gaussian kernel (n = 4): a 4-row staging block in step .. ipvt (20 slots), which the
        // register-resident n x n solve leaves free, so a lane whose own fit is between a Jacobian pass
        // and its trial points (r and qtf live) can serve the row passes of other candidates
        stg = at(n * n + 2 * n);
    }
};

template <int N>
struct LaneLM {
    // persistent over the whole solve
    double x[N], diag[N];
    double par, delta, xnorm, fnorm, gnorm;
    int iter, nfev, info;
    int ipv;             // pivot order of the last factorisation, one byte per column (N <= 4)
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
        rotate_row(a, b);
    }

    // Givens rotations of one row (a, b) into the triangular factor
    PFX_HD void rotate_row(double (&a)[N], double b) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double aj = a[j];
            const double rjj = R[ridx(j, j)];
            const double rr = PFX_FMA(rjj, rjj, aj * aj);
            // (entries whose squares underflow are dropped: see jac_block_staged)
            if (aj != 0.0 && !(rr <= 1e-290)) {
                const double inv = fast_rsqrt(rr);
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

    // Merge the factors of the C lanes of a group (each built from its own rows): log2(C)
    // exchange steps; in each, both partners stack the factor of the lower lane on top of the
    // factor of the upper lane and re-triangularise, so they end with the same bits.
    template <int C>
    PFX_HD void jac_combine() {
        typedef LaneGroup<C> G;
        if (C == 1) return;
#pragma unroll 1
        for (int s = 1; s < C; s <<= 1) {
            const bool upper = (G::rank() & s) != 0;
            double Ro[N * (N + 1) / 2], qo[N];
#pragma unroll
            for (int k = 0; k < N * (N + 1) / 2; ++k) {
                const double o = G::xchg(R[k], s);
                Ro[k] = upper ? R[k] : o;       // factor of the upper lane: rows to insert
                R[k] = upper ? o : R[k];        // factor of the lower lane: the base
            }
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const double o = G::xchg(q[k], s);
                qo[k] = upper ? q[k] : o;
                q[k] = upper ? o : q[k];
                acn2[k] = PFX_ADD(acn2[k], G::xchg(acn2[k], s));
            }
            f2 = PFX_ADD(f2, G::xchg(f2, s));
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double a[N];
#pragma unroll
                for (int j = 0; j < N; ++j) a[j] = (j >= i) ? Ro[ridx(i < j ? i : j, j)] : 0.0;
                rotate_row(a, qo[i]);
            }
        }
    }

    // B rows at once: Householder reflections eliminate the block against the triangular factor.
    // One sqrt and one division per column per block instead of one rsqrt per column per row,
    // and the B rows (their exponentials, their dot products) are independent instruction
    // streams, which is what a single resident warp per scheduler needs to keep the FP64 pipe
    // busy.  Rows beyond the data are passed as zeros (they change nothing).
    template <int B>
    PFX_HD void jac_block(double (&A)[B][N], double (&bv)[B]) {
#pragma unroll
        for (int r = 0; r < B; ++r) {
            f2 = PFX_FMA(bv[r], bv[r], f2);
#pragma unroll
            for (int j = 0; j < N; ++j) acn2[j] = PFX_FMA(A[r][j], A[r][j], acn2[j]);
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double ss = 0.0;
#pragma unroll
            for (int r = 0; r < B; ++r) ss = PFX_FMA(A[r][j], A[r][j], ss);
            const double rjj = R[ridx(j, j)];
            const double t2 = PFX_FMA(rjj, rjj, ss);
            if (ss != 0.0 && !(t2 <= 1e-290)) {          // see jac_block_staged
                // reciprocal square root and reciprocal instead of sqrt and a division: two short
                // dependency chains per column instead of two long ones
                const double inv = fast_rsqrt(t2);
                const double nrm = t2 * inv;
                const double alpha = (rjj > 0.0) ? -nrm : nrm;
                const double v0 = rjj - alpha;
                const double beta = inv * fast_rcp(fabs(v0));
                R[ridx(j, j)] = alpha;
#pragma unroll
                for (int k = j + 1; k < N; ++k) {
                    double d = v0 * R[ridx(j, k)];
#pragma unroll
                    for (int r = 0; r < B; ++r) d = PFX_FMA(A[r][j], A[r][k], d);
                    const double t = beta * d;
                    R[ridx(j, k)] = PFX_FMA(-t, v0, R[ridx(j, k)]);
#pragma unroll
                    for (int r = 0; r < B; ++r) A[r][k] = PFX_FMA(-t, A[r][j], A[r][k]);
                }
                double d = v0 * q[j];
#pragma unroll
                for (int r = 0; r < B; ++r) d = PFX_FMA(A[r][j], bv[r], d);
                const double t = beta * d;
                q[j] = PFX_FMA(-t, v0, q[j]);
#pragma unroll
                for (int r = 0; r < B; ++r) bv[r] = PFX_FMA(-t, A[r][j], bv[r]);
            }
        }
    }

    // Same elimination with the block staged in the lane's strided workspace (shared memory):
    // row r, column k at blk[r * (N + 1) + k], column N holding f.  The rows are produced by a
    // compact loop (the exponentials are not replicated B times in the instruction stream) and
    // consumed here column by column.
    template <int B>
    PFX_HD void jac_block_staged(const SPtr<double>& blk, int rows) {
        double c0[B];      // column j of the block
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double ss = 0.0;
#pragma unroll
            for (int r = 0; r < B; ++r) {
                c0[r] = (r < rows) ? blk[r * (N + 1) + j] : 0.0;
                ss = PFX_FMA(c0[r], c0[r], ss);
            }
            // (column norms are accumulated by the producer loop)
            // A column whose entries are so small that rjj^2 + ss leaves the normal range (rows far
            // out in the tail of a narrow Gaussian give derivatives of 1e-150 and less) is dropped:
            // the reciprocals below would overflow and poison the factor with NaN (round 1: "gnorm = 0
            // -> info 4" on the first pass of such fits), and against entries of order one elsewhere
            // in the column its weight is below 1e-290.
            const double rjj = R[ridx(j, j)];
            const double t2 = PFX_FMA(rjj, rjj, ss);
            if (ss != 0.0 && !(t2 <= 1e-290)) {
                const double inv = fast_rsqrt(t2);
                const double nrm = t2 * inv;
                const double alpha = (rjj > 0.0) ? -nrm : nrm;
                const double v0 = rjj - alpha;
                const double beta = inv * fast_rcp(fabs(v0));
                R[ridx(j, j)] = alpha;
#pragma unroll
                for (int k = j + 1; k <= N; ++k) {
                    double ck[B];
                    double d = v0 * ((k < N) ? R[ridx(j, k < N ? k : j)] : q[j]);
#pragma unroll
                    for (int r = 0; r < B; ++r) {
                        ck[r] = (r < rows) ? blk[r * (N + 1) + k] : 0.0;
                        d = PFX_FMA(c0[r], ck[r], d);
                    }
                    const double t = beta * d;
                    if (k < N) R[ridx(j, k < N ? k : j)] = PFX_FMA(-t, v0, R[ridx(j, k < N ? k : j)]);
                    else q[j] = PFX_FMA(-t, v0, q[j]);
#pragma unroll
                    for (int r = 0; r < B; ++r)
                        if (r < rows) blk[r * (N + 1) + k] = PFX_FMA(-t, c0[r], ck[r]);
                }
            }
        }
    }

    // returns LANE_TRIAL, or LANE_DONE with info set
    // pre_unpacked: s.r already holds the triangular factor (upper triangle, zeros below), put there by the
    // lanes that ran the row pass; R[] of this object is not read
    PFX_HD int jac_end(const LaneWs& s, int m, bool pre_unpacked = false) {
        nfev += N;
        if (iter == 1) fnorm = sqrt(f2);
        double rdiag[N], wa[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            acnorm[j] = sqrt(acn2[j]);
            rdiag[j] = acnorm[j];
            wa[j] = acnorm[j];
        }
#ifndef PFX_QRFAC_REG
#define PFX_QRFAC_REG 1
#endif
        (void)m;
        if constexpr (PFX_QRFAC_REG && N <= 4) {
            // the n x n factorisation, Q^T f and the gradient test in registers (qrfac_reg); the workspace
            // receives what the trial points read: the upper triangle, Q^T f and diag
            double a[N][N];
            int ip[N];
#pragma unroll
            for (int j = 0; j < N; ++j)
#pragma unroll
                for (int i = 0; i < N; ++i) a[j][i] = (i <= j) ? R[ridx(i < j ? i : j, j)] : 0.0;
            qrfac_reg<N>(a, ip, rdiag, wa);
            ipv = 0;
#pragma unroll
            for (int j = 0; j < N; ++j) ipv |= ip[j] << (8 * j);
            if (iter == 1) {
                double sq = 0.0;
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    diag[j] = acnorm[j];
                    if (acnorm[j] == 0.0) diag[j] = 1.0;
                    const double t = PFX_MUL(diag[j], x[j]);
                    sq = PFX_FMA(t, t, sq);
                }
                xnorm = sqrt(sq);
                delta = PFX_MUL(100.0, xnorm);
                if (delta == 0.0) delta = 100.0;
            }
            double w4[N], qtf[N];
#pragma unroll
            for (int j = 0; j < N; ++j) w4[j] = q[j];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double ajj = a[j][j];
                if (ajj != 0.0) {
                    double sum = 0.0;
#pragma unroll
                    for (int i = j; i < N; ++i) sum = PFX_FMA(a[j][i], w4[i], sum);
                    double temp = -sum / ajj;
#pragma unroll
                    for (int i = j; i < N; ++i) w4[i] = PFX_FMA(a[j][i], temp, w4[i]);
                }
                a[j][j] = rdiag[j];
                qtf[j] = w4[j];
            }
            gnorm = 0.0;
            if (fnorm != 0.0) {
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    double al = acnorm[0];
#pragma unroll
                    for (int k = 1; k < N; ++k) al = (ip[j] == k) ? acnorm[k] : al;
                    if (al != 0.0) {
                        double sum = 0.0;
#pragma unroll
                        for (int i = 0; i <= j; ++i) sum = PFX_FMA(a[j][i], qtf[i] / fnorm, sum);
                        gnorm = fmax(gnorm, fabs(sum / al));
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < N; ++j) {
#pragma unroll
                for (int i = 0; i <= j; ++i) s.r[i + N * j] = a[j][i];
                s.qtf[j] = qtf[j];
                s.ipvt[j] = ip[j];      // (read by the rolled lmpar of the histogram and DM kernels)
            }
        } else {
        // unpack the triangular factor; rows beyond m do not exist (m >= N is guaranteed upstream)
        if (!pre_unpacked) {
#pragma unroll
            for (int j = 0; j < N; ++j)
#pragma unroll
                for (int i = 0; i < N; ++i) s.r[i + N * j] = (i <= j) ? R[ridx(i < j ? i : j, j)] : 0.0;
        }
        if constexpr (N > 4) {
            // the vectors of the factorisation in the lane's workspace (the lmpar scratch rows, idle here): as local
            // arrays they are indexed dynamically and every access is a local-memory round trip (ncu on the final-fit
            // kernels: two thirds of the samples of qrfac_small waiting on them)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                s.step[j] = rdiag[j];
                s.sdiag[j] = wa[j];
                s.wa2[j] = acnorm[j];
            }
            qrfac_small(N, N, s.r, s.ipvt, s.step, s.sdiag, s.wa1);
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
            // Q^T f (unrolled: w4 stays in registers)
            double w4[N];
#pragma unroll
            for (int j = 0; j < N; ++j) w4[j] = q[j];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double ajj = s.r[j + N * j];
                if (ajj != 0.0) {
                    double sum = 0.0;
#pragma unroll
                    for (int i = j; i < N; ++i) sum = PFX_FMA(s.r[i + N * j], w4[i], sum);
                    double temp = -sum / ajj;
#pragma unroll
                    for (int i = j; i < N; ++i) w4[i] = PFX_FMA(s.r[i + N * j], temp, w4[i]);
                }
                s.r[j + N * j] = s.step[j];
                s.qtf[j] = w4[j];
            }
            gnorm = 0.0;
            if (fnorm != 0.0) {
                for (int j = 0; j < N; ++j) {
                    const int l = s.ipvt[j];
                    const double al = s.wa2[l];
                    if (al != 0.0) {
                        double sum = 0.0;
                        for (int i = 0; i <= j; ++i) sum += s.r[i + N * j] * (s.qtf[i] / fnorm);
                        gnorm = fmax(gnorm, fabs(sum / al));
                    }
                }
            }
        } else {
        double acn0[N];
        qrfac_small(N, N, s.r, s.ipvt, &rdiag[0], &wa[0], &acn0[0]);
        if (N <= 4) {
            ipv = 0;
#pragma unroll
            for (int j = 0; j < (N <= 4 ? N : 0); ++j) ipv |= s.ipvt[j] << (8 * j);
        }
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
        double stepv[N];
#ifndef PFX_SMALL_LM
#define PFX_SMALL_LM 1
#endif
        if (PFX_SMALL_LM && N <= 4) {
            // load the sub-problem into registers, solve there
            SmallLM<N> q;
#pragma unroll
            for (int j = 0; j < N; ++j) {
#pragma unroll
                for (int i = 0; i < N; ++i) q.r[i][j] = (i <= j) ? s.r[i + N * j] : 0.0;
                q.ipvt[j] = (ipv >> (8 * j)) & 0xff;
                q.diag[j] = diag[j];
                q.qtb[j] = s.qtf[j];
            }
            double sd[N];
            q.lmpar(delta, par, stepv, sd);
            double w3[N];
#pragma unroll
            for (int j = 0; j < N; ++j) w3[j] = 0.0;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double temp = -SmallLM<N>::sel(stepv, q.ipvt[j]);
#pragma unroll
                for (int i = 0; i <= j; ++i) w3[i] += q.r[i][j] * temp;
            }
            rpnorm = SmallLM<N>::nrm(w3);
        } else {
            lmpar_seq(N, N, s.r, s.ipvt, s.diag, s.qtf, delta, par, s.step, s.sdiag, s.wa1, s.wa2);
            trial_point_seq(s);
            return;
        }
        double sq = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double p = -stepv[j];
            xnew[j] = x[j] + p;
            double t = diag[j] * p;
            sq += t * t;
        }
        pnorm = sqrt(sq);
        if (iter == 1) delta = fmin(delta, pnorm);
    }
    // trial_begin with the register-resident sub-problem (N <= 4) and at most `budget` iterations of lmpar per
    // call: returns false when lmpar is not finished (call again with pending = true; st carries the iteration)
    PFX_HD bool trial_begin_budget(const LaneWs& s, LmparState& st, bool pending, int budget) {
        SmallLM<N> q;
#pragma unroll
        for (int j = 0; j < N; ++j) {
#pragma unroll
            for (int i = 0; i < N; ++i) q.r[i][j] = (i <= j) ? s.r[i + N * j] : 0.0;
            q.ipvt[j] = (ipv >> (8 * j)) & 0xff;
            q.diag[j] = diag[j];
            q.qtb[j] = s.qtf[j];
        }
        double stepv[N], sd[N];
#pragma unroll
        for (int j = 0; j < N; ++j) stepv[j] = sd[j] = 0.0;
        bool done = pending ? false : q.lmpar_begin(delta, par, stepv, st);
        for (int it = 0; !done && it < budget; ++it) done = q.lmpar_iter(delta, par, stepv, sd, st);
        if (!done) return false;
        double w3[N];
#pragma unroll
        for (int j = 0; j < N; ++j) w3[j] = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double temp = -SmallLM<N>::sel(stepv, q.ipvt[j]);
#pragma unroll
            for (int i = 0; i <= j; ++i) w3[i] += q.r[i][j] * temp;
        }
        rpnorm = SmallLM<N>::nrm(w3);
        double sq = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double p = -stepv[j];
            xnew[j] = x[j] + p;
            double t = diag[j] * p;
            sq += t * t;
        }
        pnorm = sqrt(sq);
        if (iter == 1) delta = fmin(delta, pnorm);
        return true;
    }
    // the rolled (workspace) form of trial_begin in pieces, for kernels that spread the iterations of lmpar
    // over several trips: lmpar_start, lmpar_step until done, then trial_point_seq
    PFX_HD bool lmpar_start(const LaneWs& s, LmparState& st) {
        return lmpar_begin(N, N, s.r, s.ipvt, s.diag, s.qtf, delta, par, s.step, s.wa1, s.wa2, st);
    }
    PFX_HD bool lmpar_step(const LaneWs& s, LmparState& st) {
        return lmpar_iter(N, N, s.r, s.ipvt, s.diag, s.qtf, delta, par, s.step, s.sdiag, s.wa1, s.wa2, st);
    }
    // predicted reduction and the trial point from the step in s.step
    PFX_HD void trial_point_seq(const LaneWs& s) {
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

    // ---- one n x n workspace per group --------------------------------------------------------------
    // With the workspace shared by the C lanes of a group (it is the larger part of the shared memory
    // a fit with n = 8 needs), lane 0 runs the n x n steps and hands the state they change to the
    // others; everything else stays redundant and bit-identical.
    template <int C>
    PFX_HD int jac_end_shared(const LaneWs& s, int m) {
        typedef LaneGroup<C> G;
        int r = 0;
        if (G::rank() == 0) r = jac_end(s, m);
        if (C > 1) {
            r = G::bcast(r);
#pragma unroll
            for (int j = 0; j < N; ++j) diag[j] = G::bcast(diag[j]);
            xnorm = G::bcast(xnorm);
            delta = G::bcast(delta);
            fnorm = G::bcast(fnorm);
            gnorm = G::bcast(gnorm);
            info = G::bcast(info);
            nfev = G::bcast(nfev);
        }
        return r;
    }
    template <int C>
    PFX_HD void trial_begin_shared(const LaneWs& s) {
        typedef LaneGroup<C> G;
        if (G::rank() == 0) trial_begin(s);
        if (C > 1) {
#pragma unroll
            for (int j = 0; j < N; ++j) xnew[j] = G::bcast(xnew[j]);
            pnorm = G::bcast(pnorm);
            rpnorm = G::bcast(rpnorm);
            par = G::bcast(par);
            delta = G::bcast(delta);
        }
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
