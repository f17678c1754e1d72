// pfx_lm.cuh -- group-parallel Levenberg-Marquardt: MINPACK lmdif as scipy.optimize.leastsq
// drives it (default ftol = xtol = 1.49012e-8, gtol = 0, maxfev = 200(n+1), epsfcn = eps,
// factor = 100, mode 1 scaling), which is what every fit of the reference calls
// (ProfileOperations.py:449,549,949,1045,1246,1379,1473; PHCXOperations.py:212;
// PFDOperations.py:369).  The algorithm (fdjac2 forward differences, column-pivoted Householder
// qrfac, lmpar/qrsolv, the 0.25/0.75 trust-region update and the four convergence tests) is
// MINPACK's; the mapping is not: the m residual rows are striped over the lanes of a group,
// the m x n Jacobian lives in shared memory column by column, inner products are lane-partial
// sums joined by a butterfly, and the n x n trust-region sub-problem is solved by the group
// leader on a copy of R in shared memory while the other lanes wait at a __syncwarp.
#pragma once

#include "pfx_group.cuh"

// Optional phase timers (development builds with -DPFX_PROFILE): cycles spent per LM phase,
// accumulated by the group leader into pfx_prof[] and dumped by pfx_prof_dump().
#if defined(PFX_PROFILE_LM) && defined(__CUDA_ARCH__)
extern __device__ unsigned long long pfx_prof[16];
#define PFX_T0() long long t_prof_ = clock64()
#define PFX_TACC(slot)                                                        \
    do {                                                                      \
        long long t_now_ = clock64();                                         \
        if (g.leader()) atomicAdd(&pfx_prof[slot], (unsigned long long)(t_now_ - t_prof_)); \
        t_prof_ = t_now_;                                                     \
    } while (0)
#define PFX_TCOUNT(slot) do { if (g.leader()) atomicAdd(&pfx_prof[slot], 1ULL); } while (0)
#else
#define PFX_T0() do {} while (0)
#define PFX_TACC(slot) do {} while (0)
#define PFX_TCOUNT(slot) do {} while (0)
#endif

namespace pfx {

constexpr int NPMAX = 8;
constexpr double LM_EPSMCH = 2.220446049250313e-16;
constexpr double LM_DWARF = 2.2250738585072014e-308;
constexpr double LM_TOL = 1.49012e-8;

// small per-group scratch in shared memory (leader writes, everyone reads after a sync)
struct LmShared {
    double r[NPMAX * NPMAX];  // column major: r[i + NPMAX*j] = R(i,j) upper, S^T strictly lower
    double diag[NPMAX];
    double qtf[NPMAX];
    double step[NPMAX];
    double sdiag[NPMAX];
    double wa1[NPMAX];
    double wa2[NPMAX];
    double bc[4];
    int ipvt[NPMAX];
    int pad[NPMAX];
};

// Loops of the n x n sub-problem stay rolled: the lane kernels are bound by instruction fetch (a
// few warps per scheduler, each at its own program counter; ncu: 37% of the instruction-cache
// requests of k_gauss_rounds missed), so a compact body that stays in the instruction cache beats
// an unrolled one.
#define PFX_ROLL _Pragma("unroll 1")

// Strided view: element k of a small per-lane vector lives at p[k * stride].  With stride = block
// size and the lane index folded into p, the 32 lanes of a warp touch 32 consecutive words, i.e.
// shared-memory accesses of the lane-per-fit kernels are bank-conflict free.
template <class T>
struct SPtr {
    T* p;
    int stride;
    PFX_HD T& operator[](int k) const { return p[k * stride]; }
};

template <class V>
PFX_HD inline double enorm_seq(int n, V x) {
    double s = 0.0;
    PFX_ROLL
    for (int i = 0; i < n; ++i) s += x[i] * x[i];
    return sqrt(s);
}

// MINPACK qrsolv on the n x n system, r column major with leading dimension NPMAX.
template <class RV, class IV>
PFX_HD PFX_NOINLINE void qrsolv_seq(int n, int ld, RV r, IV ipvt, RV diag, RV qtb, RV x, RV sdiag, RV wa) {
    PFX_ROLL
    for (int j = 0; j < n; ++j) {
        PFX_ROLL
        for (int i = j; i < n; ++i) r[i + ld * j] = r[j + ld * i];
        x[j] = r[j + ld * j];
        wa[j] = qtb[j];
    }
    PFX_ROLL
    for (int j = 0; j < n; ++j) {
        int l = ipvt[j];
        if (diag[l] != 0.0) {
            PFX_ROLL
            for (int k = j; k < n; ++k) sdiag[k] = 0.0;
            sdiag[j] = diag[l];
            double qtbpj = 0.0;
            PFX_ROLL
            for (int k = j; k < n; ++k) {
                if (sdiag[k] == 0.0) continue;
                double rkk = r[k + ld * k];
                double c, s;
                if (fabs(rkk) >= fabs(sdiag[k])) {
                    double t = sdiag[k] / rkk;
                    c = 0.5 / sqrt(0.25 + 0.25 * (t * t));
                    s = c * t;
                } else {
                    double ct = rkk / sdiag[k];
                    s = 0.5 / sqrt(0.25 + 0.25 * (ct * ct));
                    c = s * ct;
                }
                r[k + ld * k] = c * rkk + s * sdiag[k];
                double temp = c * wa[k] + s * qtbpj;
                qtbpj = -s * wa[k] + c * qtbpj;
                wa[k] = temp;
                PFX_ROLL
                for (int i = k + 1; i < n; ++i) {
                    double rik = r[i + ld * k];
                    temp = c * rik + s * sdiag[i];
                    sdiag[i] = -s * rik + c * sdiag[i];
                    r[i + ld * k] = temp;
                }
            }
        }
        sdiag[j] = r[j + ld * j];
        r[j + ld * j] = x[j];
    }
    int nsing = n;
    PFX_ROLL
    for (int j = 0; j < n; ++j) {
        if (sdiag[j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa[j] = 0.0;
    }
    PFX_ROLL
    for (int k = 0; k < nsing; ++k) {
        int j = nsing - 1 - k;
        double sum = 0.0;
        PFX_ROLL
        for (int i = j + 1; i < nsing; ++i) sum += r[i + ld * j] * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
    PFX_ROLL
    for (int j = 0; j < n; ++j) x[ipvt[j]] = wa[j];
}

#if defined(PFX_COUNT_LMPAR) && !defined(__CUDA_ARCH__)
static long long g_lmpar_hist8[12];
#endif
// MINPACK lmpar, split so that a kernel can spread the iterations of a hard sub-problem over several
// trips (lmpar_begin, then lmpar_iter until it returns true); lmpar_seq runs them back to back.  What an
// iteration needs from the previous one are four scalars (LmparState) and par; the vectors are rebuilt.
struct LmparState {
    double parl, paru, fp;
    int iter;
};
// returns true when the Gauss-Newton step is accepted (par = 0): no iterations needed
template <class RV, class IV>
PFX_HD PFX_NOINLINE bool lmpar_begin(int n, int ld, RV r, IV ipvt, RV diag, RV qtb, double delta, double& par,
                                     RV x, RV wa1, RV wa2, LmparState& st) {
    const double p1 = 0.1;
    int nsing = n;
    PFX_ROLL
    for (int j = 0; j < n; ++j) {
        wa1[j] = qtb[j];
        if (r[j + ld * j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa1[j] = 0.0;
    }
    PFX_ROLL
    for (int k = 0; k < nsing; ++k) {
        int j = nsing - 1 - k;
        wa1[j] = wa1[j] / r[j + ld * j];
        double temp = wa1[j];
        PFX_ROLL
        for (int i = 0; i < j; ++i) wa1[i] -= r[i + ld * j] * temp;
    }
    PFX_ROLL
    for (int j = 0; j < n; ++j) x[ipvt[j]] = wa1[j];

    st.iter = 0;
    PFX_ROLL
    for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
    double dxnorm = enorm_seq(n, wa2);
    double fp = dxnorm - delta;
    if (fp <= p1 * delta) {
#if defined(PFX_COUNT_LMPAR) && !defined(__CUDA_ARCH__)
        if (n == 8) g_lmpar_hist8[0]++;
#endif
        par = 0.0;
        return true;
    }
    double parl = 0.0;
    if (nsing >= n) {
        PFX_ROLL
        for (int j = 0; j < n; ++j) {
            int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
        }
        PFX_ROLL
        for (int j = 0; j < n; ++j) {
            double sum = 0.0;
            PFX_ROLL
            for (int i = 0; i < j; ++i) sum += r[i + ld * j] * wa1[i];
            wa1[j] = (wa1[j] - sum) / r[j + ld * j];
        }
        double temp = enorm_seq(n, wa1);
        parl = ((fp / delta) / temp) / temp;
    }
    PFX_ROLL
    for (int j = 0; j < n; ++j) {
        double sum = 0.0;
        PFX_ROLL
        for (int i = 0; i <= j; ++i) sum += r[i + ld * j] * qtb[i];
        wa1[j] = sum / diag[ipvt[j]];
    }
    double gnorm = enorm_seq(n, wa1);
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
// one iteration; returns true when par is final (x and sdiag then hold the step and the diagonal of S)
template <class RV, class IV>
PFX_HD PFX_NOINLINE bool lmpar_iter(int n, int ld, RV r, IV ipvt, RV diag, RV qtb, double delta, double& par,
                                    RV x, RV sdiag, RV wa1, RV wa2, LmparState& st) {
    const double p1 = 0.1, p001 = 0.001;
    ++st.iter;
    if (par == 0.0) par = fmax(LM_DWARF, p001 * st.paru);
    double temp = sqrt(par);
    PFX_ROLL
    for (int j = 0; j < n; ++j) wa1[j] = temp * diag[j];
    qrsolv_seq(n, ld, r, ipvt, wa1, qtb, x, sdiag, wa2);
    PFX_ROLL
    for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
    const double dxnorm = enorm_seq(n, wa2);
    temp = st.fp;
    const double fp = dxnorm - delta;
    st.fp = fp;
    if (fabs(fp) <= p1 * delta || (st.parl == 0.0 && fp <= temp && temp < 0.0) || st.iter == 10) {
#if defined(PFX_COUNT_LMPAR) && !defined(__CUDA_ARCH__)
        if (n == 8) g_lmpar_hist8[st.iter]++;
#endif
        return true;
    }
    PFX_ROLL
    for (int j = 0; j < n; ++j) {
        int l = ipvt[j];
        wa1[j] = diag[l] * (wa2[l] / dxnorm);
    }
    PFX_ROLL
    for (int j = 0; j < n; ++j) {
        wa1[j] = wa1[j] / sdiag[j];
        double t = wa1[j];
        PFX_ROLL
        for (int i = j + 1; i < n; ++i) wa1[i] -= r[i + ld * j] * t;
    }
    temp = enorm_seq(n, wa1);
    double parc = ((fp / delta) / temp) / temp;
    if (fp > 0.0) st.parl = fmax(st.parl, par);
    if (fp < 0.0) st.paru = fmin(st.paru, par);
    par = fmax(st.parl, par + parc);
    return false;
}
template <class RV, class IV>
PFX_HD inline void lmpar_seq(int n, int ld, RV r, IV ipvt, RV diag, RV qtb, double delta, double& par,
                             RV x, RV sdiag, RV wa1, RV wa2) {
    LmparState st;
    if (lmpar_begin(n, ld, r, ipvt, diag, qtb, delta, par, x, wa1, wa2, st)) return;
    PFX_ROLL
    while (!lmpar_iter(n, ld, r, ipvt, diag, qtb, delta, par, x, sdiag, wa1, wa2, st)) {
    }
}

template <int NP>
PFX_HD inline double vget(const double (&a)[NP], int idx) {
    double v = a[0];
#pragma unroll
    for (int k = 1; k < NP; ++k)
        if (k == idx) v = a[k];
    return v;
}
template <int NP>
PFX_HD inline void vset(double (&a)[NP], int idx, double v) {
#pragma unroll
    for (int k = 0; k < NP; ++k)
        if (k == idx) a[k] = v;
}
template <int NP>
PFX_HD inline int iget(const int (&a)[NP], int idx) {
    int v = a[0];
#pragma unroll
    for (int k = 1; k < NP; ++k)
        if (k == idx) v = a[k];
    return v;
}
template <int NP>
PFX_HD inline void iset(int (&a)[NP], int idx, int v) {
#pragma unroll
    for (int k = 0; k < NP; ++k)
        if (k == idx) a[k] = v;
}

// Large per-fit work vectors (shared memory, or global scratch for oversized histograms).
struct LmVec {
    double* fvec;  // [m]
    double* wa4;   // [m]
    double* fjac;  // [NP][ld], column j at fjac + j*ld
    int ld;
};

// Residual evaluation is kept out of line: lmdif calls it from three places and the models
// carry exp / sin, so inlining would triple the hot code (instruction-cache misses dominated
// the first version of these kernels).
template <int NP, class G, class Model>
PFX_HD PFX_NOINLINE void lm_eval(const G& g, const Model& f, const double (&p)[NP], double* out, int m) {
    f.eval(g, p, out, m);
}

// Returns MINPACK's info (scipy's ier).  x: start on entry, solution on exit (replicated in
// every lane).  Model::eval(g, p, out, m) writes the residual rows owned by the calling lane.
template <int NP, class G, class Model>
PFX_HD inline int lmdif(const G& g, const Model& f, int m, double (&x)[NP], const LmVec& v,
                        LmShared& s, int& nfev_out) {
    const double p1 = 0.1, p5 = 0.5, p25 = 0.25, p75 = 0.75, p0001 = 1e-4;
    const double ftol = LM_TOL, xtol = LM_TOL, gtol = 0.0, factor = 100.0;
    const double epsmch = LM_EPSMCH;
    const int maxfev = 200 * (NP + 1);
    const double eps = 1.4901161193847656e-08;  // sqrt(max(epsfcn, epsmch)), epsfcn = eps
    double* fvec = v.fvec;
    double* wa4 = v.wa4;
    const int ld = v.ld;
    int info = 0;
    int nfev = 0;

    PFX_T0();
    PFX_TCOUNT(8);
    lm_eval<NP>(g, f, x, fvec, m);
    nfev = 1;
    double fnorm = sqrt(rows_sumsq(g, fvec, 0, m));
    PFX_TACC(0);
    double par = 0.0;
    int iter = 1;
    double diag[NP], qtf[NP], acnorm[NP], rdiag[NP], wa[NP], xnew[NP], p[NP];
    int ipvt[NP];
    double xnorm = 0.0, delta = 0.0, gnorm = 0.0;
#pragma unroll
    for (int j = 0; j < NP; ++j) diag[j] = 0.0;

    for (;;) {
        // ---- fdjac2: forward-difference Jacobian ------------------------------------------
#pragma unroll 1
        for (int j = 0; j < NP; ++j) {
            double temp = vget<NP>(x, j);
            double h = eps * fabs(temp);
            if (h == 0.0) h = eps;
            vset<NP>(x, j, temp + h);
            lm_eval<NP>(g, f, x, wa4, m);
            vset<NP>(x, j, temp);
            double* col = v.fjac + j * ld;
            for (int i = g.lane(); i < m; i += G::W) col[i] = (wa4[i] - fvec[i]) / h;
        }
        nfev += NP;
        PFX_TACC(1);
        PFX_TCOUNT(9);

        // ---- qrfac with column pivoting ---------------------------------------------------
        // (loops are deliberately not unrolled; the small vectors live in registers and are
        // indexed with select chains)
#pragma unroll 1
        for (int j = 0; j < NP; ++j) {
            double a = sqrt(rows_sumsq(g, v.fjac + j * ld, 0, m));
            vset<NP>(acnorm, j, a);
            vset<NP>(rdiag, j, a);
            vset<NP>(wa, j, a);
            iset<NP>(ipvt, j, j);
        }
        const int minmn = (m < NP) ? m : NP;
#pragma unroll 1
        for (int j = 0; j < minmn; ++j) {
            int kmax = j;
            double best = vget<NP>(rdiag, j);
#pragma unroll 1
            for (int k = j + 1; k < NP; ++k) {
                double rk = vget<NP>(rdiag, k);
                if (rk > best) {
                    best = rk;
                    kmax = k;
                }
            }
            double* cj = v.fjac + j * ld;
            if (kmax != j) {
                double* ck = v.fjac + kmax * ld;
                for (int i = g.lane(); i < m; i += G::W) {
                    double t = cj[i];
                    cj[i] = ck[i];
                    ck[i] = t;
                }
                vset<NP>(rdiag, kmax, vget<NP>(rdiag, j));
                vset<NP>(wa, kmax, vget<NP>(wa, j));
                int t = iget<NP>(ipvt, j);
                iset<NP>(ipvt, j, iget<NP>(ipvt, kmax));
                iset<NP>(ipvt, kmax, t);
            }
            double ajnorm = sqrt(rows_sumsq(g, cj, j, m));
            if (ajnorm != 0.0) {
                g.sync();
                if (cj[j] < 0.0) ajnorm = -ajnorm;
                g.sync();
                for (int i = g.lane(); i < m; i += G::W)
                    if (i >= j) {
                        double t = cj[i] / ajnorm;
                        if (i == j) t += 1.0;
                        cj[i] = t;
                    }
                g.sync();
                const double ajj = cj[j];
#pragma unroll 1
                for (int k = j + 1; k < NP; ++k) {
                    double* ck = v.fjac + k * ld;
                    double sum = rows_dot(g, cj, ck, j, m);
                    double temp = sum / ajj;
                    for (int i = g.lane(); i < m; i += G::W)
                        if (i >= j) ck[i] -= temp * cj[i];
                }
                g.sync();
#pragma unroll 1
                for (int k = j + 1; k < NP; ++k) {
                    double rk = vget<NP>(rdiag, k);
                    if (rk != 0.0) {
                        const double* ck = v.fjac + k * ld;
                        double temp = ck[j] / rk;
                        rk *= sqrt(fmax(0.0, 1.0 - temp * temp));
                        double q = rk / vget<NP>(wa, k);
                        if (0.05 * (q * q) <= epsmch) {
                            rk = sqrt(rows_sumsq(g, ck, j + 1, m));
                            vset<NP>(wa, k, rk);
                        }
                        vset<NP>(rdiag, k, rk);
                    }
                }
            }
            vset<NP>(rdiag, j, -ajnorm);
        }

        PFX_TACC(2);
        if (iter == 1) {
            double sq = 0.0;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                diag[j] = acnorm[j];
                if (acnorm[j] == 0.0) diag[j] = 1.0;
                double t = diag[j] * x[j];
                sq += t * t;
            }
            xnorm = sqrt(sq);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }

        // ---- (Q^T fvec)[0..n) ---------------------------------------------------------------
        for (int i = g.lane(); i < m; i += G::W) wa4[i] = fvec[i];
        g.sync();
#pragma unroll 1
        for (int j = 0; j < NP; ++j) {
            double qj = 0.0;
            if (j < m) {
                const double* cj = v.fjac + j * ld;
                const double ajj = cj[j];
                if (ajj != 0.0) {
                    double sum = rows_dot(g, cj, wa4, j, m);
                    double temp = -sum / ajj;
                    for (int i = g.lane(); i < m; i += G::W)
                        if (i >= j) wa4[i] += cj[i] * temp;
                }
                g.sync();
                qj = wa4[j];
            }
            vset<NP>(qtf, j, qj);
        }

        PFX_TACC(3);
        // ---- norm of the scaled gradient ----------------------------------------------------
        gnorm = 0.0;
        if (fnorm != 0.0) {
#pragma unroll 1
            for (int j = 0; j < NP; ++j) {
                int l = iget<NP>(ipvt, j);
                double al = vget<NP>(acnorm, l);
                if (al != 0.0) {
                    const double* cj = v.fjac + j * ld;
                    double sum = 0.0;
#pragma unroll 1
                    for (int i = 0; i < j; ++i) sum += cj[i] * (vget<NP>(qtf, i) / fnorm);
                    sum += vget<NP>(rdiag, j) * (vget<NP>(qtf, j) / fnorm);
                    gnorm = fmax(gnorm, fabs(sum / al));   // fmax: a NaN never replaces gnorm
                }
            }
        }
        if (gnorm <= gtol) info = 4;
        if (info != 0) break;
#pragma unroll
        for (int j = 0; j < NP; ++j) diag[j] = fmax(diag[j], acnorm[j]);

        // leader: publish R, permutation, Q^T f and the scaling for the n x n sub-problem
        if (g.leader()) {
#pragma unroll 1
            for (int j = 0; j < NP; ++j) {
                const double* cj = v.fjac + j * ld;
                for (int i = 0; i < j; ++i) s.r[i + NPMAX * j] = cj[i];
                s.r[j + NPMAX * j] = vget<NP>(rdiag, j);
                s.ipvt[j] = iget<NP>(ipvt, j);
                s.qtf[j] = vget<NP>(qtf, j);
                s.diag[j] = vget<NP>(diag, j);
            }
        }
        g.sync();

        PFX_TACC(4);
        // ---- inner loop ---------------------------------------------------------------------
        for (;;) {
            PFX_TCOUNT(10);
            if (g.leader()) {
                double lpar = par;
                lmpar_seq<double*, int*>(NP, NPMAX, s.r, s.ipvt, s.diag, s.qtf, delta, lpar, s.step, s.sdiag, s.wa1, s.wa2);
                s.bc[0] = lpar;
                // || R P^T p ||
                for (int j = 0; j < NP; ++j) s.wa1[j] = 0.0;
                for (int j = 0; j < NP; ++j) {
                    double temp = -s.step[s.ipvt[j]];
                    for (int i = 0; i <= j; ++i) s.wa1[i] += s.r[i + NPMAX * j] * temp;
                }
                s.bc[1] = enorm_seq<double*>(NP, s.wa1);
            }
            g.sync();
            par = s.bc[0];
            const double rpnorm = s.bc[1];
            double sq = 0.0;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                p[j] = -s.step[j];
                xnew[j] = x[j] + p[j];
                double t = diag[j] * p[j];
                sq += t * t;
            }
            g.sync();
            PFX_TACC(5);
            const double pnorm = sqrt(sq);
            if (iter == 1) delta = fmin(delta, pnorm);

            lm_eval<NP>(g, f, xnew, wa4, m);
            ++nfev;
            const double fnorm1 = sqrt(rows_sumsq(g, wa4, 0, m));
            PFX_TACC(6);

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
                double sq2 = 0.0;
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    x[j] = xnew[j];
                    double t = diag[j] * x[j];
                    sq2 += t * t;
                }
                for (int i = g.lane(); i < m; i += G::W) fvec[i] = wa4[i];
                xnorm = sqrt(sq2);
                fnorm = fnorm1;
                ++iter;
            }

            const bool c1 = fabs(actred) <= ftol && prered <= ftol && p5 * ratio <= 1.0;
            if (c1) info = 1;
            if (delta <= xtol * xnorm) info = 2;
            if (c1 && info == 2) info = 3;
            if (info != 0) break;
            if (nfev >= maxfev) info = 5;
            if (fabs(actred) <= epsmch && prered <= epsmch && p5 * ratio <= 1.0) info = 6;
            if (delta <= epsmch * xnorm) info = 7;
            if (gnorm <= epsmch) info = 8;
            if (info != 0) break;
            PFX_TACC(7);
            if (ratio >= p0001) break;   // successful step: new Jacobian
        }
        if (info != 0) break;
    }
    g.sync();
    nfev_out = nfev;
    return info;
}

}  // namespace pfx
