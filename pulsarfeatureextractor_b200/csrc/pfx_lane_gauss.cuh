// pfx_lane_gauss.cuh -- scores 8-11 with one lane per candidate (see pfx_lane_lm.cuh for why).
//
// A candidate goes through ten least-squares fits (ProfileOperations.py:724-768,1061-1483):
//   stage 0      Gaussian + background on the background-subtracted, rotated profile  (PO:1194-1264)
//   stage 1..8   the eight 4-parameter rounds of the double-Gaussian procedure        (PO:1358-1405)
//   final        the 8-parameter fit and the choice between it and the "combi" curve  (PO:1407-1483)
// Stages 0..8 all fit |max| exp(-((x-mu)/s)^2/2) + b with four parameters, so one kernel runs them
// as a per-lane state machine: every trip of the kernel loop is "one Jacobian pass and one trial
// point" for all 32 lanes of a warp, whatever stage or candidate each lane is in.  The final fit
// has eight parameters and runs in a second kernel; the few numbers it needs travel through a
// small per-candidate record in global memory (GaussCarry).
//
// Per-lane working data: dat[m] (float64 data vector of the current fit) and xs[m] (uint16 bin
// numbers) in shared memory, strided so that the lanes of a warp never collide on a bank.
#pragma once

#include "pfx_driver.cuh"
#include "pfx_lane_lm.cuh"

// -DPFX_PROFILE: per-phase cycle and active-lane accounting of the lane kernels (development
// builds only; read back with pfx_debug_profile, scripts/gpu_phase_profile.py).  Slot k holds
// [2k] warp cycles spent in the phase and [2k+1] the same weighted by the lanes that executed it.
#if defined(PFX_PROFILE) && defined(__CUDACC__)
static __device__ unsigned long long g_pfx_prof[32];
#endif
#if defined(PFX_PROFILE) && defined(__CUDA_ARCH__)
#define PFX_PROF_T0(v) const long long v = clock64()
#define PFX_PROF_ADD(slot, v)                                                            \
    do {                                                                                 \
        const unsigned am_ = __activemask();                                             \
        const long long dt_ = clock64() - (v);                                           \
        if ((int)(threadIdx.x & 31) == __ffs(am_) - 1) {                                 \
            atomicAdd(&g_pfx_prof[2 * (slot)], (unsigned long long)dt_);                 \
            atomicAdd(&g_pfx_prof[2 * (slot) + 1], (unsigned long long)dt_ * __popc(am_)); \
        }                                                                                \
    } while (0)
#else
#define PFX_PROF_T0(v)
#define PFX_PROF_ADD(slot, v)
#endif
enum { PROF_SETUP = 0, PROF_JAC_ROWS = 1, PROF_JAC_END = 2, PROF_LMPAR = 3, PROF_TRIAL_ROWS = 4, PROF_TRIAL_END = 5,
       PROF_FINISH = 6, PROF_QUEUE = 7, PROF_TOTAL = 8, PROF_TAIL = 9,
       PROF8_START = 10, PROF8_JAC_ROWS = 11, PROF8_JAC_END = 12, PROF8_TRIAL = 13, PROF8_FINISH = 14, PROF8_TOTAL = 15 };

#ifndef PFX_HYBRID
#define PFX_HYBRID 0    // host build: 1 = follow the hybrid kernel's arithmetic, 0 = the lane kernel's (the default kernel)
#endif
#ifndef PFX_COOP
#define PFX_COOP 1      // 0: every lane sets up and closes its own fits (round-1 behaviour, for comparison)
#endif

namespace pfx {

struct LaneBuf {
    double* dat;           // element i of this lane at dat[i * stride]
    unsigned short* xs;    // element i of this lane at xs[i * stride]
    int stride;
    int pitch = 1;         // distance between the vectors of neighbouring lanes (1: lane index fastest;
                           // m + pad with stride 1: one contiguous vector per candidate, hybrid kernel)
    PFX_HD double& d(int i) const { return dat[(size_t)i * stride]; }
    PFX_HD unsigned short& x(int i) const { return xs[(size_t)i * stride]; }
};

// what the final-fit kernel needs from the rounds kernel
struct GaussCarry {
    double p_t1[4];
    double p1[4];
    double p2[4];
    double minbg, sd, t1_fwhm;
    int skip;      // 1: candidate finished in the rounds kernel (failure or sentinel)
    int pad;
};

enum { GP_IDLE = 0, GP_SETUP = 1, GP_JAC = 2, GP_TRIAL = 3, GP_FIN = 4, GP_WAIT = 5 };

#if !defined(__CUDA_ARCH__)
// host build (tests/hostsim): > 0 = run lmpar of the 4-parameter fits in resumable pieces of that many iterations
inline int g_host_lmpar_budget = 0;
#endif

// rows of the 4-parameter family: value = |p2| exp(-((x-p1)/s)^2/2) + b,
// s = |p0| or p0, b = p3 or |p3|
struct G4Flags {
    bool abs_sigma, abs_bg;
};

// exp(-((x-mu)/sigma)^2 / 2) with the division replaced by a multiplication with 1/sigma
// (computed once per pass with an IEEE division) and the branch-free exponential: the two
// deviations from the reference expression are below 2 ulp, the size of the differences between
// any two exp() implementations.
PFX_HD inline double gauss_fast(double xv, double mu, double inv_sigma) {
    const double t = PFX_MUL(PFX_SUB(xv, mu), inv_sigma);
    return exp_neg(PFX_MUL(-PFX_MUL(t, t), 0.5));
}

// one evaluation point of the 4-parameter family, prepared once per pass
struct G4Eval {
    double mu, inv, amx, bg;
    PFX_HD void set(const double (&p)[4], G4Flags fl) {
        mu = p[1];
        inv = 1.0 / (fl.abs_sigma ? fabs(p[0]) : p[0]);
        amx = fabs(p[2]);
        bg = fl.abs_bg ? fabs(p[3]) : p[3];
    }
    PFX_HD double value(double xv) const { return PFX_ADD(PFX_MUL(amx, gauss_fast(xv, mu, inv)), bg); }
};

PFX_HD inline double g4_value(const double (&p)[4], G4Flags fl, double xv) {
    G4Eval e;
    e.set(p, fl);
    return e.value(xv);
}

// one Jacobian row of the 4-parameter family: residual f0 and the four forward differences.  Three
// exponentials per row (the max and background columns reuse the base exponential, which gives
// bit-identical values to re-evaluating the model).  The constants are prepared once per pass.
struct G4Jac {
    double isg, isgh, mu, muh, amx, amxh, bg, bgh, ih0, ih1, ih2, ih3;
    PFX_HD void set(const double* x, const double* h, G4Flags fl) {
        isg = 1.0 / (fl.abs_sigma ? fabs(x[0]) : x[0]);
        isgh = 1.0 / (fl.abs_sigma ? fabs(x[0] + h[0]) : (x[0] + h[0]));
        mu = x[1];
        muh = x[1] + h[1];
        amx = fabs(x[2]);
        amxh = fabs(x[2] + h[2]);
        bg = fl.abs_bg ? fabs(x[3]) : x[3];
        bgh = fl.abs_bg ? fabs(x[3] + h[3]) : (x[3] + h[3]);
        ih0 = 1.0 / h[0];
        ih1 = 1.0 / h[1];
        ih2 = 1.0 / h[2];
        ih3 = 1.0 / h[3];
    }
    PFX_HD void row(double xv, double yv, double& a0, double& a1, double& a2, double& a3, double& f0) const {
        const double e0 = gauss_fast(xv, mu, isg);
        const double e1 = gauss_fast(xv, mu, isgh);
        const double e2 = gauss_fast(xv, muh, isg);
        const double base = PFX_MUL(amx, e0);
        f0 = PFX_SUB(yv, PFX_ADD(base, bg));
        a0 = (PFX_SUB(yv, PFX_ADD(PFX_MUL(amx, e1), bg)) - f0) * ih0;
        a1 = (PFX_SUB(yv, PFX_ADD(PFX_MUL(amx, e2), bg)) - f0) * ih1;
        a2 = (PFX_SUB(yv, PFX_ADD(PFX_MUL(amxh, e0), bg)) - f0) * ih2;
        a3 = (PFX_SUB(yv, PFX_ADD(base, bgh)) - f0) * ih3;
    }
};

// ---- hybrid rows: EIGHT LANES PER CANDIDATE for the passes over the data, ONE for the n x n work ----
// The row loops of a fit (Jacobian pass, trial point) are the part of a trip that is proportional to
// the number of bins, the n x n work (pivoting, lmpar, the convergence tests) is not.  Run by the
// candidate's own lane, the row loops cost a warp the same whether 32 or 3 of its candidates need
// them -- the lanes whose fit is retrying a step, is being set up or has run out of candidates idle
// through them (20 of 32 lanes active on average, profiles/r02_phase_profile_coop.txt; the tail of
// slow candidates is 28 % of a warp's life).  In the hybrid kernel the warp instead serves the
// candidates that need a pass four at a time with eight lanes each: lane s of a group takes the rows
// s, s + 8, s + 16, ... in blocks of four, eliminates them against its own triangular factor in
// registers, and the eight factors are merged by a butterfly (both partners stack the lower lane's
// factor on the upper lane's, so they end with the same bits); the group leader hands R, Q^T f, the
// column norms and |f|^2 to the owner through its n x n workspace.  The cost of a pass is now
// proportional to the candidates that need it, every lane of the warp works in it, and four rows
// (twelve exponentials) are in flight per lane.  The arithmetic (which rows meet in which factor,
// the merge tree) depends only on the number of rows, never on scheduling; the host build walks the
// eight lanes of a group in turn.
struct RowFactor4 {
    double R[10], q[4], acn2[4], f2;      // packed upper triangle (row i at i*4 - i(i-1)/2), Q^T f, sums
    PFX_HD static constexpr int ridx(int i, int j) { return i * 4 - (i * (i - 1)) / 2 + (j - i); }
    PFX_HD void zero() {
#pragma unroll
        for (int k = 0; k < 10; ++k) R[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            q[k] = 0.0;
            acn2[k] = 0.0;
        }
        f2 = 0.0;
    }
    // Householder elimination of four rows (A[r][0..3] | A[r][4] = f) against the factor; with TRI the
    // block is itself upper triangular (row r empty left of column r: the factor of a partner lane).
    // Same formulas as LaneLM::jac_block (reciprocal square root and reciprocal; columns whose norm
    // leaves the normal range are dropped).
    template <bool TRI>
    PFX_HD void eliminate(double (&A)[4][5]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double ss = 0.0;
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (!TRI || r <= j) ss = PFX_FMA(A[r][j], A[r][j], ss);
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
                for (int k = j + 1; k <= 4; ++k) {
                    double d = v0 * ((k < 4) ? R[ridx(j, k < 4 ? k : j)] : q[j]);
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (!TRI || r <= j) d = PFX_FMA(A[r][j], A[r][k], d);
                    const double t = beta * d;
                    if (k < 4) R[ridx(j, k < 4 ? k : j)] = PFX_FMA(-t, v0, R[ridx(j, k < 4 ? k : j)]);
                    else q[j] = PFX_FMA(-t, v0, q[j]);
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (!TRI || r <= j) A[r][k] = PFX_FMA(-t, A[r][j], A[r][k]);
                }
            }
        }
    }
    // The same elimination with the block staged in the lane's workspace (row r, column k at
    // blk[r * 5 + k], column 4 = f; rows beyond `rows` count as zeros): the rows are produced by a compact
    // rolled loop -- unrolling four rows with their twelve exponentials made the row pass four times the
    // size of the instruction cache (ncu: instruction fetch 57 % of the stalls there).
    PFX_HD void eliminate_staged(const SPtr<double>& blk, int rows) {
        double c0[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double ss = 0.0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                c0[r] = (r < rows) ? blk[r * 5 + j] : 0.0;
                ss = PFX_FMA(c0[r], c0[r], ss);
            }
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
                for (int k = j + 1; k <= 4; ++k) {
                    double ck[4];
                    double d = v0 * ((k < 4) ? R[ridx(j, k < 4 ? k : j)] : q[j]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        ck[r] = (r < rows) ? blk[r * 5 + k] : 0.0;
                        d = PFX_FMA(c0[r], ck[r], d);
                    }
                    const double t = beta * d;
                    if (k < 4) R[ridx(j, k < 4 ? k : j)] = PFX_FMA(-t, v0, R[ridx(j, k < 4 ? k : j)]);
                    else q[j] = PFX_FMA(-t, v0, q[j]);
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (r < rows) blk[r * 5 + k] = PFX_FMA(-t, c0[r], ck[r]);
                }
            }
        }
    }
    // this := factor of (rows of this stacked on the rows of up); the sums add (commutative: both
    // partners of an exchange get the same bits)
    PFX_HD void absorb(const RowFactor4& up) {
        double A[4][5];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k) A[r][k] = (k >= r) ? up.R[ridx(r < k ? r : k, k)] : 0.0;
            A[r][4] = up.q[r];
        }
        eliminate<true>(A);
#pragma unroll
        for (int k = 0; k < 4; ++k) acn2[k] = PFX_ADD(acn2[k], up.acn2[k]);
        f2 = PFX_ADD(f2, up.f2);
    }
};
constexpr int HY_GROUP = 8;                 // lanes per candidate in the row passes
// Slots of the owner's n x n workspace used for the hand-over (LaneWs::blk[k]).  Slots 24 .. 43 of every
// lane are its staging block during Jacobian passes (LaneWs::stg).  A lane that owns a Jacobian pass has
// a dead workspace (its factor is rebuilt by jac_close): constants in 0 .. 11, the result over them in
// 0 .. 18 once the group is done with the constants.  Around a trial point r and qtf (0 .. 23) are live:
// constants and result go to the staging area, which is idle during trial passes.
enum { HY_JC = 0, HY_JR = 0, HY_TC = 24, HY_TS = 28 };
// doubles per candidate slot in the packed layout of the hybrid kernel (0..23 r/diag/qtf, 24..28 trial constants
// and result, ipvt over 24..27); also the number of workspace slots the bit set of the peak stripping may use
constexpr int HY_WSN = 30;

// ---- warp-cooperative helpers --------------------------------------------------------------
// The set-up and the close of a fit (copying / rebuilding the data vector, its statistics, the
// chi-square) are passes over the bins of ONE candidate.  Run by the candidate's own lane they were
// executed with 1.6-1.8 of 32 lanes active and took a fifth of k_gauss_rounds (phase profile,
// profiles/r02_phase_profile_baseline.txt), because the lanes of a warp reach them at different
// times.  Here the whole warp serves one candidate at a time: the bins are dealt to the 32 lanes
// and sums are closed by a butterfly.  The summation order (32 strided partial sums, then the
// butterfly) is fixed and reproduced on the host by one lane emulating the 32, so host and device
// builds of the kernel logic stay comparable.
struct WarpCoop {
    static constexpr int W = 32;
#if defined(__CUDA_ARCH__)
    PFX_HD static int lane() { return (int)(threadIdx.x & 31); }
    PFX_HD static double bc(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
    PFX_HD static int bc(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
    PFX_HD static long long bc(long long v, int src) { return __shfl_sync(0xffffffffu, v, src); }
    PFX_HD static void sync() { __syncwarp(); }
    template <class F>
    PFX_HD static void rows(int n, F f) {
        for (int i = lane(); i < n; i += W) f(i);
    }
    // sum of val(i); with SQ the terms are val(i)^2 accumulated with one fused multiply-add each
    template <bool SQ, class F>
    PFX_HD static double sum(int n, F val) {
        double p = 0.0;
        for (int i = lane(); i < n; i += W) {
            const double v = val(i);
            p = SQ ? PFX_FMA(v, v, p) : PFX_ADD(p, v);
        }
#pragma unroll
        for (int s = W / 2; s > 0; s >>= 1) p = PFX_ADD(p, __shfl_xor_sync(0xffffffffu, p, s));
        return p;
    }
    // numpy.argmax (argmax_better is a total order, so the combination order is irrelevant) and
    // Python's builtin max (a NaN survives only in first position)
    template <class F>
    PFX_HD static void maxima(int n, F val, int& amax, double& vmax, double& bmax) {
        int bi = -1;
        double bv = 0.0, mx = -INFINITY;
        for (int i = lane(); i < n; i += W) {
            const double v = val(i);
            if (bi < 0 || argmax_better(v, i, bv, bi)) {
                bv = v;
                bi = i;
            }
            if (v > mx) mx = v;
        }
#pragma unroll
        for (int s = W / 2; s > 0; s >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, s), om = __shfl_xor_sync(0xffffffffu, mx, s);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
            if (oi >= 0 && (bi < 0 || argmax_better(ov, oi, bv, bi))) {
                bv = ov;
                bi = oi;
            }
            if (om > mx) mx = om;
        }
        const double d0 = val(0);
        amax = bi;
        vmax = bv;
        bmax = (d0 != d0) ? d0 : mx;
    }
#else
    PFX_HD static int lane() { return 0; }
    PFX_HD static double bc(double v, int) { return v; }
    PFX_HD static int bc(int v, int) { return v; }
    PFX_HD static long long bc(long long v, int) { return v; }
    PFX_HD static void sync() {}
    template <class F>
    PFX_HD static void rows(int n, F f) {
        for (int i = 0; i < n; ++i) f(i);
    }
    template <bool SQ, class F>
    PFX_HD static double sum(int n, F val) {
        double part[W];
        for (int l = 0; l < W; ++l) {
            double p = 0.0;
            for (int i = l; i < n; i += W) {
                const double v = val(i);
                p = SQ ? PFX_FMA(v, v, p) : PFX_ADD(p, v);
            }
            part[l] = p;
        }
        for (int s = W / 2; s > 0; s >>= 1) {
            double t[W];
            for (int l = 0; l < W; ++l) t[l] = PFX_ADD(part[l], part[l ^ s]);
            for (int l = 0; l < W; ++l) part[l] = t[l];
        }
        return part[0];
    }
    template <class F>
    PFX_HD static void maxima(int n, F val, int& amax, double& vmax, double& bmax) {
        int bi = -1;
        double bv = 0.0, mx = -INFINITY;
        for (int i = 0; i < n; ++i) {
            const double v = val(i);
            if (bi < 0 || argmax_better(v, i, bv, bi)) {
                bv = v;
                bi = i;
            }
            if (v > mx) mx = v;
        }
        const double d0 = val(0);
        amax = bi;
        vmax = bv;
        bmax = (d0 != d0) ? d0 : mx;
    }
#endif
};

// The index list of the peak stripping (PO:1305-1354) as a bit set over the bins: the list is always
// an increasing subsequence of 0..m-1, so "delete the element at list position k" is "clear the k-th
// set bit".  The reference deletes one or two elements per step of its loop, and a list kept as an
// array pays a shift of the whole tail for each (O(m^2) shared-memory moves per candidate by a single
// lane, 5 % of the kernel in round 1).  Words live in the lane's n x n workspace, which is idle while
// a fit is set up.
struct BinSet {
    SPtr<double> w;      // 64 bins per slot
    int nwords;
    PFX_HD static int capacity(int slots) { return slots * 64; }
    PFX_HD unsigned long long get(int k) const {
        unsigned long long v;
        const double d = w[k];
        memcpy(&v, &d, 8);
        return v;
    }
    PFX_HD void put(int k, unsigned long long v) const {
        double d;
        memcpy(&d, &v, 8);
        w[k] = d;
    }
    PFX_HD void fill(int m) {
        nwords = (m + 63) / 64;
        for (int k = 0; k < nwords; ++k) {
            const int left = m - 64 * k;
            put(k, left >= 64 ? ~0ULL : ((1ULL << left) - 1ULL));
        }
    }
    PFX_HD static int popc(unsigned long long v) {
#if defined(__CUDA_ARCH__)
        return __popcll(v);
#else
        return __builtin_popcountll(v);
#endif
    }
    PFX_HD static int lowest(unsigned long long v) {
#if defined(__CUDA_ARCH__)
        return __ffsll((long long)v) - 1;
#else
        return __builtin_ctzll(v);
#endif
    }
    // clear the k-th set bit (0 <= k < number of set bits)
    PFX_HD void erase(int k) const {
        for (int wd = 0; wd < nwords; ++wd) {
            unsigned long long v = get(wd);
            const int c = popc(v);
            if (k < c) {
                unsigned long long t = v;
                for (int q = 0; q < k; ++q) t &= t - 1;
                put(wd, v & ~(1ULL << lowest(t)));
                return;
            }
            k -= c;
        }
    }
    PFX_HD bool has(int i) const { return (get(i >> 6) >> (i & 63)) & 1ULL; }
};

// C lanes of a warp share the candidate (LaneGroup): one copy of the data vector per group.
// Element-wise passes over the data are dealt to the lanes (element i to lane i % C) and closed
// by a group barrier; order-sensitive scans (sequential sums, argmax, the index-list surgery of
// the peak stripping) are done by every lane redundantly on the shared copy (reads) or by lane 0
// (writes), so that all lanes of a group always hold bit-identical state.
template <int C = 1>
struct GaussLaneT {
    typedef LaneGroup<C> G;
    // candidate
    const void* profile;
    int dtype;
    long long cand;
    int m, cut;
    double ph_expect, mean, sd, minbg;
    // state
    int phase, stage, nrows;
    G4Flags fl;
    LaneLM<4> lm;
    double p_t1[4], p1[4], p2[4];
    double t1_fwhm, t1_chi;
    unsigned status;
    int8_t ier[10];      // slots PFX_FIT_T1 .. PFX_FIT_DG_ROUND0+7 (index = slot - PFX_FIT_T1)
    bool finished, sentinel;

    PFX_HD double yrot(int i) const {
        int src = i + cut;
        if (src >= m) src -= m;
        return load_elem(profile, dtype, cand * (long long)m + src);
    }

    PFX_HD void start(const void* prof, int dt, long long c, int m_, double ph) {
        profile = prof;
        dtype = dt;
        cand = c;
        m = m_;
        cut = m_ / 2;
        ph_expect = ph;
        stage = 0;
        phase = GP_SETUP;
        status = 0;
        finished = false;
        sentinel = false;
        for (int k = 0; k < 10; ++k) ier[k] = 0;
        t1_fwhm = NAN;
        t1_chi = NAN;
    }

    // mean / std / argmax / builtin-max of dat[0..n)
    PFX_HD void stats(const LaneBuf& b, int n, double& mu, double& sdev, int& amax, double& vmax, double& bmax) const {
        double s = 0.0;
        amax = 0;
        vmax = b.d(0);
        bmax = b.d(0);
        for (int i = 0; i < n; ++i) {
            const double v = b.d(i);
            s += v;
            if (i > 0) {
                if (argmax_better(v, i, vmax, amax)) {
                    vmax = v;
                    amax = i;
                }
                if (v > bmax) bmax = v;      // builtin max: sequential '>'
            }
        }
        mu = s / (double)n;
        double q = 0.0;
        for (int i = 0; i < n; ++i) {
            const double dlt = b.d(i) - mu;
            q = PFX_FMA(dlt, dlt, q);
        }
        sdev = sqrt(q / (double)n);
    }

    PFX_HD void setup(const LaneBuf& b) {
        double x0[4];
        if (stage == 0) {
            // profile statistics and the T1 data vector (PO:724-739, 1098-1110)
            for (int i = G::rank(); i < m; i += C) {
                b.d(i) = yrot(i);
                b.x(i) = (unsigned short)i;
            }
            G::sync();
            double s = 0.0;
            for (int i = 0; i < m; ++i) s += b.d(i);
            mean = s / (double)m;
            double q = 0.0;
            for (int i = 0; i < m; ++i) {
                const double dlt = b.d(i) - mean;
                q = PFX_FMA(dlt, dlt, q);
            }
            sd = sqrt(q / (double)m);
            minbg = (mean < ph_expect) ? mean : ph_expect;     // builtin min(ph_expect, mean)
            G::sync();
            if (minbg > 0.0) {
                for (int i = G::rank(); i < m; i += C) {
                    double v = PFX_ADD(PFX_SUB(b.d(i), minbg), sd);
                    if (v < 0.0) v = 0.0;
                    b.d(i) = v;
                }
            }
            G::sync();
            double mu, sdev, vmax, bmax;
            int amax;
            stats(b, m, mu, sdev, amax, vmax, bmax);
            x0[0] = sdev;
            x0[1] = (double)amax;
            x0[2] = vmax;
            x0[3] = 1.0;
            fl.abs_sigma = true;
            fl.abs_bg = false;
            nrows = m;
        } else if (stage == 1) {
            // rotated raw profile, peak stripping (PO:1305-1354)
            for (int i = G::rank(); i < m; i += C) b.d(i) = yrot(i);
            G::sync();
            int amax = 0;
            double vmax = 0.0;
            for (int i = 0; i < m; ++i) {
                const double v = b.d(i);
                if (i == 0 || argmax_better(v, i, vmax, amax)) {
                    vmax = v;
                    amax = i;
                }
            }
            const int nlen = strip_peak(b, amax);
            G::sync();
            if (nlen < 0) {                 // IndexError -> 1e6 sentinels (PO:762-764)
                sentinel = true;
                status |= PFX_ST_DG_SENTINEL;
                finished = true;
                phase = GP_IDLE;
                return;
            }
            if (nlen == 0) {
                status |= PFX_ST_EMPTY_SEQUENCE;
                finished = true;
                phase = GP_IDLE;
                return;
            }
            if (G::rank() == 0)
                for (int k = 0; k < nlen; ++k) b.d(k) = b.d(b.x(k));   // x(k) >= k, so no overwrite hazard
            G::sync();
            double mu, sdev, bmax;
            int am2;
            stats(b, nlen, mu, sdev, am2, vmax, bmax);
            x0[0] = sdev;
            x0[1] = (double)b.x(am2);
            x0[2] = bmax;
            x0[3] = mu;
            nrows = nlen;
            for (int k = nlen; k < 4; ++k) {     // PO:1373-1377 zero padding
                if (G::rank() == 0) {
                    b.d(k) = 0.0;
                    b.x(k) = 0;
                }
                nrows = 4;
            }
            G::sync();
            fl.abs_sigma = false;
            fl.abs_bg = true;
        } else {
            // rounds 2..8: data were rebuilt by the previous round's finish()
            double mu, sdev, vmax, bmax;
            int amax;
            stats(b, m, mu, sdev, amax, vmax, bmax);
            x0[0] = sdev;
            x0[1] = (double)b.x(amax);
            x0[2] = bmax;
            x0[3] = mu;
            nrows = m;
            fl.abs_sigma = false;
            fl.abs_bg = true;
        }
        lm.init(x0);
        phase = GP_JAC;
    }

    // PO:1305-1354 on the index list kept in xs; returns the survivor count or -1 (IndexError)
    PFX_HD int strip_peak(const LaneBuf& b, int pos) const { return strip_peak_list(b, m, pos); }
    PFX_HD static int strip_peak_list(const LaneBuf& b, int n, int pos) {
        int len = n;
        const bool writer = (G::rank() == 0);     // every lane follows the control flow, lane 0 edits the list
        if (writer)
            for (int i = 0; i < n; ++i) b.x(i) = (unsigned short)i;
#define PFX_DEL(k_)                                                  \
    do {                                                             \
        int k = (k_);                                                \
        if (k < -len || k >= len) return -1;                         \
        if (k < 0) k += len;                                         \
        if (writer)                                                  \
            for (int q_ = k; q_ < len - 1; ++q_) b.x(q_) = b.x(q_ + 1);  \
        --len;                                                       \
    } while (0)
        PFX_DEL(pos);
        int tol = 0;
        const int limit = 5;
        for (int i = 1; i < n; ++i) {
            if ((pos - i) > 0 && (pos + i) < n) {
                const double yl = b.d(pos - i), yl1 = b.d(pos - i + 1), yr_ = b.d(pos + i), yr1 = b.d(pos + i - 1);
                if ((yl < yl1) && (yr_ < yr1)) {
                    PFX_DEL(pos - i);
                    PFX_DEL(pos - i);
                } else if ((yl >= yl1) || ((yr_ >= yr1) && (tol < limit))) {
                    PFX_DEL(pos - i);
                    PFX_DEL(pos - i);
                    ++tol;
                } else {
                    break;
                }
            } else if ((pos - i) < 0) {
                if (pos + i >= n) return -1;
                const double yr_ = b.d(pos + i), yr1 = b.d(pos + i - 1);
                if (yr_ < yr1) {
                    PFX_DEL(pos - i + 1);
                } else if ((yr_ >= yr1) && (tol < limit)) {
                    PFX_DEL(pos - i + 1);
                    ++tol;
                } else {
                    break;
                }
            } else if ((pos + i) > n) {
                const double yl = b.d(pos - i), yl1 = b.d(pos - i + 1);
                if (yl < yl1) {
                    PFX_DEL(pos - i + 1);
                } else if ((yl >= yl1) && (tol < limit)) {
                    PFX_DEL(pos - i);
                    ++tol;
                } else {
                    break;
                }
            }
        }
#undef PFX_DEL
        return len;
    }

    // one Jacobian pass; rows are produced into the staged block and eliminated eight at a time
    PFX_HD void jac_pass(const LaneBuf& b, const LaneWs& ws) {
        PFX_PROF_T0(pt0);
        lm.jac_begin();
        G4Jac J;
        J.set(lm.x, lm.h, fl);
        constexpr int B = LANE_BLOCK_ROWS;
        for (int i0 = G::rank() * B; i0 < nrows; i0 += C * B) {     // blocks of rows dealt to the lanes
            const int rows = (nrows - i0 < B) ? nrows - i0 : B;
#pragma unroll 1
            for (int r = 0; r < rows; ++r) {
                const int i = i0 + r;
                double a0, a1, a2, a3, f0;
                J.row((double)b.x(i), b.d(i), a0, a1, a2, a3, f0);
                lm.acn2[0] = PFX_FMA(a0, a0, lm.acn2[0]);
                lm.acn2[1] = PFX_FMA(a1, a1, lm.acn2[1]);
                lm.acn2[2] = PFX_FMA(a2, a2, lm.acn2[2]);
                lm.acn2[3] = PFX_FMA(a3, a3, lm.acn2[3]);
                lm.f2 = PFX_FMA(f0, f0, lm.f2);
                ws.blk[r * 5 + 0] = a0;
                ws.blk[r * 5 + 1] = a1;
                ws.blk[r * 5 + 2] = a2;
                ws.blk[r * 5 + 3] = a3;
                ws.blk[r * 5 + 4] = f0;
            }
            lm.template jac_block_staged<B>(ws.blk, rows);
        }
        lm.template jac_combine<C>();
        PFX_PROF_ADD(PROF_JAC_ROWS, pt0);
        PFX_PROF_T0(pt1);
        const int r = lm.jac_end(ws, nrows);
        phase = (r == LANE_TRIAL) ? GP_TRIAL : GP_FIN;
        PFX_PROF_ADD(PROF_JAC_END, pt1);
    }

    PFX_HD void trial_pass(const LaneBuf& b, const LaneWs& ws) {
        PFX_PROF_T0(pt0);
        lm.trial_begin(ws);
        PFX_PROF_ADD(PROF_LMPAR, pt0);
        PFX_PROF_T0(pt1);
        G4Eval ev;
        ev.set(lm.xnew, fl);
        double s = 0.0;
#pragma unroll 4
        for (int i = G::rank(); i < nrows; i += C) {
            const double f = PFX_SUB(b.d(i), ev.value((double)b.x(i)));
            s = PFX_FMA(f, f, s);
        }
        s = G::sum(s);
        PFX_PROF_ADD(PROF_TRIAL_ROWS, pt1);
        PFX_PROF_T0(pt2);
        const int r = lm.trial_end(sqrt(s));
        phase = (r == LANE_JAC) ? GP_JAC : (r == LANE_TRIAL ? GP_TRIAL : GP_FIN);
        PFX_PROF_ADD(PROF_TRIAL_END, pt2);
    }

    // ---- hybrid rows (see RowFactor4): lane-mode halves of a pass -------------------------------
    PFX_HD void jac_prep(const LaneWs& ws) {
        lm.jac_begin();
        G4Jac J;
        J.set(lm.x, lm.h, fl);
        const double c[12] = {J.isg, J.isgh, J.mu, J.muh, J.amx, J.amxh, J.bg, J.bgh, J.ih0, J.ih1, J.ih2, J.ih3};
#pragma unroll
        for (int k = 0; k < 12; ++k) ws.blk[HY_JC + k] = c[k];
    }
    PFX_HD void jac_close(const LaneWs& ws) {
        PFX_PROF_T0(pt1);
#pragma unroll
        for (int k = 0; k < 10; ++k) lm.R[k] = ws.blk[HY_JR + k];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            lm.q[k] = ws.blk[HY_JR + 10 + k];
            lm.acn2[k] = ws.blk[HY_JR + 14 + k];
        }
        lm.f2 = ws.blk[HY_JR + 18];
        const int r = lm.jac_end(ws, nrows);
        phase = (r == LANE_TRIAL) ? GP_TRIAL : GP_FIN;
        PFX_PROF_ADD(PROF_JAC_END, pt1);
    }
    PFX_HD void trial_prep(const LaneWs& ws) {
        PFX_PROF_T0(pt0);
        lm.trial_begin(ws);
        G4Eval ev;
        ev.set(lm.xnew, fl);
        ws.blk[HY_TC + 0] = ev.mu;
        ws.blk[HY_TC + 1] = ev.inv;
        ws.blk[HY_TC + 2] = ev.amx;
        ws.blk[HY_TC + 3] = ev.bg;
        PFX_PROF_ADD(PROF_LMPAR, pt0);
    }
    // trial_prep with at most `budget` lmpar iterations (block-synchronous kernel): false = not finished yet
    PFX_HD bool trial_prep_budget(const LaneWs& ws, LmparState& st, bool pending, int budget) {
        if (!lm.trial_begin_budget(ws, st, pending, budget)) return false;
        G4Eval ev;
        ev.set(lm.xnew, fl);
        ws.blk[HY_TC + 0] = ev.mu;
        ws.blk[HY_TC + 1] = ev.inv;
        ws.blk[HY_TC + 2] = ev.amx;
        ws.blk[HY_TC + 3] = ev.bg;
        return true;
    }
    PFX_HD void trial_close(const LaneWs& ws) {
        PFX_PROF_T0(pt2);
        const int r = lm.trial_end(sqrt(ws.blk[HY_TS]));
        phase = (r == LANE_JAC) ? GP_JAC : (r == LANE_TRIAL ? GP_TRIAL : GP_FIN);
        PFX_PROF_ADD(PROF_TRIAL_END, pt2);
    }
    // the rows of lane s of a group: s, s + 8, s + 16, ... in blocks of four
    PFX_HD static void rows_factor(const LaneBuf& bs, int nr, int s, const G4Jac& J, RowFactor4& F,
                                   const SPtr<double>& stg) {
        F.zero();
#pragma unroll 1
        for (int i0 = s; i0 < nr; i0 += 4 * HY_GROUP) {
            int rows = 0;
#pragma unroll 1
            for (int i = i0; i < nr && rows < 4; i += HY_GROUP, ++rows) {
                double a0, a1, a2, a3, f0;
                J.row((double)bs.x(i), bs.d(i), a0, a1, a2, a3, f0);
                F.acn2[0] = PFX_FMA(a0, a0, F.acn2[0]);
                F.acn2[1] = PFX_FMA(a1, a1, F.acn2[1]);
                F.acn2[2] = PFX_FMA(a2, a2, F.acn2[2]);
                F.acn2[3] = PFX_FMA(a3, a3, F.acn2[3]);
                F.f2 = PFX_FMA(f0, f0, F.f2);
                stg[rows * 5 + 0] = a0;
                stg[rows * 5 + 1] = a1;
                stg[rows * 5 + 2] = a2;
                stg[rows * 5 + 3] = a3;
                stg[rows * 5 + 4] = f0;
            }
            F.eliminate_staged(stg, rows);
        }
    }
    PFX_HD static double rows_trial(const LaneBuf& bs, int nr, int s, const G4Eval& ev) {
        double acc = 0.0;
#pragma unroll 4
        for (int i = s; i < nr; i += HY_GROUP) {
            const double f = PFX_SUB(bs.d(i), ev.value((double)bs.x(i)));
            acc = PFX_FMA(f, f, acc);
        }
        return acc;
    }
    PFX_HD static void load_jac_consts(const double* w, int st, G4Jac& J) {
        J.isg = w[(HY_JC + 0) * st];
        J.isgh = w[(HY_JC + 1) * st];
        J.mu = w[(HY_JC + 2) * st];
        J.muh = w[(HY_JC + 3) * st];
        J.amx = w[(HY_JC + 4) * st];
        J.amxh = w[(HY_JC + 5) * st];
        J.bg = w[(HY_JC + 6) * st];
        J.bgh = w[(HY_JC + 7) * st];
        J.ih0 = w[(HY_JC + 8) * st];
        J.ih1 = w[(HY_JC + 9) * st];
        J.ih2 = w[(HY_JC + 10) * st];
        J.ih3 = w[(HY_JC + 11) * st];
    }
    PFX_HD static void store_factor(double* w, int st, const RowFactor4& F) {
#pragma unroll
        for (int k = 0; k < 10; ++k) w[(HY_JR + k) * st] = F.R[k];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            w[(HY_JR + 10 + k) * st] = F.q[k];
            w[(HY_JR + 14 + k) * st] = F.acn2[k];
        }
        w[(HY_JR + 18) * st] = F.f2;
    }
#if defined(__CUDACC__)
    // One Jacobian pass of one candidate by the eight lanes of a group (s = lane of the group, gmask = its
    // lanes): constants from the owner's workspace w (element k at w[k * st]), rows of lane s eliminated
    // into its own factor (staging block stg), butterfly merge, result into the owner's workspace.
    __device__ static void group_jac_pass(const LaneBuf& bs, int nr, int s, unsigned gmask, double* w, int st,
                                          const SPtr<double>& stg) {
        RowFactor4 F;
        {
            G4Jac J;
            load_jac_consts(w, st, J);
            rows_factor(bs, nr, s, J, F, stg);
        }
#pragma unroll 1
        for (int sh = 1; sh < HY_GROUP; sh <<= 1) {
            const bool upper = (s & sh) != 0;
            RowFactor4 lo, up;
#pragma unroll
            for (int k = 0; k < 10; ++k) {
                const double o = __shfl_xor_sync(gmask, F.R[k], sh);
                lo.R[k] = upper ? o : F.R[k];
                up.R[k] = upper ? F.R[k] : o;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double o = __shfl_xor_sync(gmask, F.q[k], sh);
                lo.q[k] = upper ? o : F.q[k];
                up.q[k] = upper ? F.q[k] : o;
                const double oa = __shfl_xor_sync(gmask, F.acn2[k], sh);
                lo.acn2[k] = upper ? oa : F.acn2[k];
                up.acn2[k] = upper ? F.acn2[k] : oa;
            }
            {
                const double o = __shfl_xor_sync(gmask, F.f2, sh);
                lo.f2 = upper ? o : F.f2;
                up.f2 = upper ? F.f2 : o;
            }
            lo.absorb(up);
            F = lo;
        }
        if (s == 0) store_factor(w, st, F);
    }
    __device__ static void group_trial_pass(const LaneBuf& bs, int nr, int s, unsigned gmask, double* w, int st) {
        G4Eval ev;
        ev.mu = w[(HY_TC + 0) * st];
        ev.inv = w[(HY_TC + 1) * st];
        ev.amx = w[(HY_TC + 2) * st];
        ev.bg = w[(HY_TC + 3) * st];
        double acc = rows_trial(bs, nr, s, ev);
#pragma unroll
        for (int sh = 1; sh < HY_GROUP; sh <<= 1) acc = PFX_ADD(acc, __shfl_xor_sync(gmask, acc, sh));
        if (s == 0) w[HY_TS * st] = acc;
    }
    // Called by all 32 lanes.  mask = the owners (lanes of this warp) whose fit needs a Jacobian pass and
    // have run jac_prep; my_nrows = the caller's own row count.  Four candidates per turn, eight lanes each.
    __device__ static void hybrid_jac_rows(unsigned mask, int my_nrows, const LaneBuf& b, const LaneWs& ws) {
        const int lane = (int)(threadIdx.x & 31), g = lane >> 3, s = lane & (HY_GROUP - 1);
        const unsigned gmask = 0xffu << (8 * g);
        while (mask) {
            const unsigned pos = __fns(mask, 0, g + 1);          // lane of the g-th candidate of this turn
            const int src = (pos < 32u) ? (int)pos : -1;
            const int nr = __shfl_sync(0xffffffffu, my_nrows, src < 0 ? 0 : src);
#pragma unroll
            for (int t = 0; t < 32 / HY_GROUP; ++t) mask &= mask - 1;
            if (src >= 0) group_jac_pass(buf_of(b, src), nr, s, gmask, ws.blk.p + (src - lane), ws.blk.stride, ws.stg);
            __syncwarp();
        }
    }
    __device__ static void hybrid_trial_rows(unsigned mask, int my_nrows, const LaneBuf& b, const LaneWs& ws) {
        const int lane = (int)(threadIdx.x & 31), g = lane >> 3, s = lane & (HY_GROUP - 1);
        const unsigned gmask = 0xffu << (8 * g);
        while (mask) {
            const unsigned pos = __fns(mask, 0, g + 1);
            const int src = (pos < 32u) ? (int)pos : -1;
            const int nr = __shfl_sync(0xffffffffu, my_nrows, src < 0 ? 0 : src);
#pragma unroll
            for (int t = 0; t < 32 / HY_GROUP; ++t) mask &= mask - 1;
            if (src >= 0) group_trial_pass(buf_of(b, src), nr, s, gmask, ws.blk.p + (src - lane), ws.blk.stride);
            __syncwarp();
        }
    }
#endif
#if !defined(__CUDA_ARCH__)
    // host build: one lane walks the eight lanes of its group in turn (same rows, same merge tree)
    PFX_HD void host_group_jac(const LaneBuf& b, const LaneWs& ws) {
        RowFactor4 F[HY_GROUP];
        G4Jac J;
        load_jac_consts(ws.blk.p, ws.blk.stride, J);
        for (int s = 0; s < HY_GROUP; ++s) rows_factor(b, nrows, s, J, F[s], ws.stg);
        for (int sh = 1; sh < HY_GROUP; sh <<= 1)
            for (int s = 0; s < HY_GROUP; ++s)
                if (!(s & sh)) {
                    RowFactor4 lo = F[s];
                    lo.absorb(F[s | sh]);
                    F[s] = lo;
                    F[s | sh] = lo;
                }
        store_factor(ws.blk.p, ws.blk.stride, F[0]);
    }
    PFX_HD void host_group_trial(const LaneBuf& b, const LaneWs& ws) {
        G4Eval ev;
        ev.mu = ws.blk[HY_TC + 0];
        ev.inv = ws.blk[HY_TC + 1];
        ev.amx = ws.blk[HY_TC + 2];
        ev.bg = ws.blk[HY_TC + 3];
        double p[HY_GROUP];
        for (int s = 0; s < HY_GROUP; ++s) p[s] = rows_trial(b, nrows, s, ev);
        for (int sh = 1; sh < HY_GROUP; sh <<= 1) {
            double t[HY_GROUP];
            for (int s = 0; s < HY_GROUP; ++s) t[s] = PFX_ADD(p[s], p[s ^ sh]);
            for (int s = 0; s < HY_GROUP; ++s) p[s] = t[s];
        }
        ws.blk[HY_TS] = p[0];
    }
#endif

    PFX_HD void finish(const LaneBuf& b) {
        double p[4] = {lm.x[0], lm.x[1], lm.x[2], lm.x[3]};
        ier[stage] = (int8_t)lm.info;
        if (stage == 0) {
            for (int k = 0; k < 4; ++k) p_t1[k] = p[k];
            t1_fwhm = fabs(FWHM_K * p[0]);
            G4Eval ev;
            ev.set(p, fl);
            double c = 0.0;
#pragma unroll 4
            for (int i = G::rank(); i < m; i += C) {
                const double dlt = b.d(i) - ev.value((double)i);
                c = PFX_FMA(dlt, dlt, c);
            }
            c = G::sum(c);
            t1_chi = c / (double)m;
        } else {
            // rebuild the data for the next round over the full profile (PO:1388-1399)
            const double nfwhm = fabs(FWHM_K * p[0]);
            const double lo = p[2] - (1.5 * nfwhm) / 2.0;
            const double hi = p[2] + (1.5 * nfwhm) / 2.0;
            G4Eval fe;
            fe.set(p, fl);
            G::sync();
#pragma unroll 4
            for (int i = G::rank(); i < m && stage < 8; i += C) {
                const double xv = (double)i;
                const double ev = fe.value(xv);
                const double yv = yrot(i);
                double nv;
                if (ev <= yv) nv = PFX_ADD(PFX_SUB(yv, ev), p[3]);
                else if ((ev > yv) && (xv > lo) && (xv < hi)) nv = p[3];
                else nv = yv;
                b.d(i) = nv;
                b.x(i) = (unsigned short)i;
            }
            G::sync();
            if (stage == 7)
                for (int k = 0; k < 4; ++k) p2[k] = p[k];
            if (stage == 8)
                for (int k = 0; k < 4; ++k) p1[k] = p[k];
        }
        ++stage;
        if (stage > 8) {
            finished = true;
            phase = GP_IDLE;
        } else {
            phase = GP_SETUP;
        }
    }

    // ---- warp-cooperative set-up and close of a fit (one lane per candidate, C == 1) -------------
    // Called by ALL lanes of the warp with the same src; the state of lane src is broadcast, the
    // passes over its bins are dealt to the 32 lanes, and lane src alone commits the results.
    PFX_HD static LaneBuf buf_of(const LaneBuf& b, int src) {
        LaneBuf o = b;
        if (b.pitch < 0) {      // buffers of the lanes anywhere in memory (wave kernel): fetch the owner's pointers
            o.dat = (double*)WarpCoop::bc((long long)b.dat, src);
            o.xs = (unsigned short*)WarpCoop::bc((long long)b.xs, src);
            return o;
        }
        const int d = (src - WarpCoop::lane()) * b.pitch;
        o.dat = b.dat + d;
        o.xs = b.xs + d;
        return o;
    }
    PFX_HD static double yrot_of(const void* prof, int dt, long long cd, int mm, int ct, int i) {
        int sidx = i + ct;
        if (sidx >= mm) sidx -= mm;
        return load_elem(prof, dt, cd * (long long)mm + sidx);
    }
    PFX_HD static void coop_stats(const LaneBuf& bs, int n, double& mu, double& sdev, int& amax, double& vmax,
                                  double& bmax) {
        typedef WarpCoop WC;
        const double sm = WC::template sum<false>(n, [&](int i) { return bs.d(i); });
        mu = sm / (double)n;
        const double mu_ = mu;
        const double q = WC::template sum<true>(n, [&](int i) { return bs.d(i) - mu_; });
        sdev = sqrt(q / (double)n);
        WC::maxima(n, [&](int i) { return bs.d(i); }, amax, vmax, bmax);
    }

    // PO:1305-1354 with the index list kept as a bit set (BinSet); same control flow as strip_peak
    PFX_HD static int strip_peak_bits(const LaneBuf& b, int n, BinSet& set, int pos) {
        int len = n;
        set.fill(n);
#define PFX_DEL(k_)                              \
    do {                                         \
        int k = (k_);                            \
        if (k < -len || k >= len) return -1;     \
        if (k < 0) k += len;                     \
        set.erase(k);                            \
        --len;                                   \
    } while (0)
        PFX_DEL(pos);
        int tol = 0;
        const int limit = 5;
        for (int i = 1; i < n; ++i) {
            if ((pos - i) > 0 && (pos + i) < n) {
                const double yl = b.d(pos - i), yl1 = b.d(pos - i + 1), yr_ = b.d(pos + i), yr1 = b.d(pos + i - 1);
                if ((yl < yl1) && (yr_ < yr1)) {
                    PFX_DEL(pos - i);
                    PFX_DEL(pos - i);
                } else if ((yl >= yl1) || ((yr_ >= yr1) && (tol < limit))) {
                    PFX_DEL(pos - i);
                    PFX_DEL(pos - i);
                    ++tol;
                } else {
                    break;
                }
            } else if ((pos - i) < 0) {
                if (pos + i >= n) return -1;
                const double yr_ = b.d(pos + i), yr1 = b.d(pos + i - 1);
                if (yr_ < yr1) {
                    PFX_DEL(pos - i + 1);
                } else if ((yr_ >= yr1) && (tol < limit)) {
                    PFX_DEL(pos - i + 1);
                    ++tol;
                } else {
                    break;
                }
            } else if ((pos + i) > n) {
                const double yl = b.d(pos - i), yl1 = b.d(pos - i + 1);
                if (yl < yl1) {
                    PFX_DEL(pos - i + 1);
                } else if ((yl >= yl1) && (tol < limit)) {
                    PFX_DEL(pos - i);
                    ++tol;
                } else {
                    break;
                }
            }
        }
#undef PFX_DEL
        return len;
    }

    // The passes of a set-up / close as functions of plain values, called by ALL lanes of a warp with the
    // same arguments (bs = the candidate's data vector, lead = the lane that runs the serial list surgery
    // in bits[] and writes single elements): the owner-lane forms below broadcast the owner's state and
    // commit the results to it; the block-synchronous kernel runs them as tasks on any warp.
    struct SetupOut {
        double x0[4], mean, sd, minbg;
        int nr;          // rows of the fit; <= 0: no fit (-1: IndexError -> sentinels, 0: empty sequence)
    };
    PFX_HD static void work_setup(int lead, const LaneBuf& bs, const SPtr<double>& bits, int bit_slots, int st, double ph,
                                  const void* prof, int dt, long long cd, int mm, int ct, SetupOut& o) {
        typedef WarpCoop WC;
        const bool me = WC::lane() == lead;
        o.nr = mm;
        o.mean = o.sd = o.minbg = 0.0;
        if (st == 0) {
            // profile statistics and the T1 data vector (PO:724-739, 1098-1110)
            WC::rows(mm, [&](int i) {
                bs.d(i) = yrot_of(prof, dt, cd, mm, ct, i);
                bs.x(i) = (unsigned short)i;
            });
            WC::sync();
            const double sm = WC::template sum<false>(mm, [&](int i) { return bs.d(i); });
            const double mean_ = sm / (double)mm;
            const double q = WC::template sum<true>(mm, [&](int i) { return bs.d(i) - mean_; });
            const double sd_ = sqrt(q / (double)mm);
            const double minbg_ = (mean_ < ph) ? mean_ : ph;     // builtin min(ph_expect, mean)
            WC::sync();
            if (minbg_ > 0.0) {
                WC::rows(mm, [&](int i) {
                    double v = PFX_ADD(PFX_SUB(bs.d(i), minbg_), sd_);
                    if (v < 0.0) v = 0.0;
                    bs.d(i) = v;
                });
            }
            WC::sync();
            double mu, sdev, vmax, bmax;
            int amax;
            coop_stats(bs, mm, mu, sdev, amax, vmax, bmax);
            o.x0[0] = sdev;
            o.x0[1] = (double)amax;
            o.x0[2] = vmax;
            o.x0[3] = 1.0;
            o.mean = mean_;
            o.sd = sd_;
            o.minbg = minbg_;
        } else if (st == 1) {
            // rotated raw profile, peak stripping (PO:1305-1354)
            WC::rows(mm, [&](int i) { bs.d(i) = yrot_of(prof, dt, cd, mm, ct, i); });
            WC::sync();
            int amax;
            double vmax, bmax;
            WC::maxima(mm, [&](int i) { return bs.d(i); }, amax, vmax, bmax);
            int nlen = 0;
            if (me) {
                if (mm <= BinSet::capacity(bit_slots)) {
                    BinSet set;
                    set.w = bits;
                    nlen = strip_peak_bits(bs, mm, set, amax);
                    if (nlen > 0) {
                        int cnt = 0;
                        for (int i = 0; i < mm; ++i)
                            if (set.has(i)) {       // i >= cnt: no overwrite hazard
                                bs.d(cnt) = bs.d(i);
                                bs.x(cnt) = (unsigned short)i;
                                ++cnt;
                            }
                    }
                } else {
                    nlen = strip_peak_list(bs, mm, amax);
                    for (int k = 0; k < nlen; ++k) bs.d(k) = bs.d(bs.x(k));
                }
            }
            nlen = WC::bc(nlen, lead);
            WC::sync();
            if (nlen <= 0) {
                o.nr = nlen;
                return;
            }
            double mu, sdev;
            int am2;
            coop_stats(bs, nlen, mu, sdev, am2, vmax, bmax);
            o.x0[0] = sdev;
            o.x0[1] = (double)bs.x(am2);
            o.x0[2] = bmax;
            o.x0[3] = mu;
            o.nr = nlen;
            WC::sync();
            if (nlen < 4) {                          // PO:1373-1377 zero padding
                if (me)
                    for (int k = nlen; k < 4; ++k) {
                        bs.d(k) = 0.0;
                        bs.x(k) = 0;
                    }
                o.nr = 4;
            }
            WC::sync();
        } else {
            // rounds 2..8: data were rebuilt by the previous round's close
            double mu, sdev, vmax, bmax;
            int amax;
            coop_stats(bs, mm, mu, sdev, amax, vmax, bmax);
            o.x0[0] = sdev;
            o.x0[1] = (double)bs.x(amax);
            o.x0[2] = bmax;
            o.x0[3] = mu;
        }
    }
    // close of fit `st` with parameters p: the chi-square of the T1 fit (returned), or the data of the next round
    PFX_HD static double work_finish(const LaneBuf& bs, int st, const double (&p)[4], const void* prof, int dt,
                                     long long cd, int mm, int ct) {
        typedef WarpCoop WC;
        G4Flags f;
        f.abs_sigma = (st == 0);
        f.abs_bg = (st != 0);
        if (st == 0) {
            G4Eval ev;
            ev.set(p, f);
            const double c = WC::template sum<true>(mm, [&](int i) { return bs.d(i) - ev.value((double)i); });
            return c / (double)mm;
        }
        // rebuild the data for the next round over the full profile (PO:1388-1399)
        const double nfwhm = fabs(FWHM_K * p[0]);
        const double lo = p[2] - (1.5 * nfwhm) / 2.0;
        const double hi = p[2] + (1.5 * nfwhm) / 2.0;
        G4Eval fe;
        fe.set(p, f);
        WC::sync();
        if (st < 8) {
            WC::rows(mm, [&](int i) {
                const double xv = (double)i;
                const double ev = fe.value(xv);
                const double yv = yrot_of(prof, dt, cd, mm, ct, i);
                double nv;
                if (ev <= yv) nv = PFX_ADD(PFX_SUB(yv, ev), p[3]);
                else if ((ev > yv) && (xv > lo) && (xv < hi)) nv = p[3];
                else nv = yv;
                bs.d(i) = nv;
                bs.x(i) = (unsigned short)i;
            });
        }
        WC::sync();
        return 0.0;
    }
    // owner-side halves: what a set-up / close changes in the state of the candidate's lane
    PFX_HD void commit_setup(int st, const SetupOut& o) {
        if (st == 0) {
            mean = o.mean;
            sd = o.sd;
            minbg = o.minbg;
        }
        if (o.nr <= 0) {
            if (o.nr < 0) {                 // IndexError -> 1e6 sentinels (PO:762-764)
                sentinel = true;
                status |= PFX_ST_DG_SENTINEL;
            } else {
                status |= PFX_ST_EMPTY_SEQUENCE;
            }
            finished = true;
            phase = GP_IDLE;
            return;
        }
        fl.abs_sigma = (st == 0);
        fl.abs_bg = (st != 0);
        nrows = o.nr;
        lm.init(o.x0);
        phase = GP_JAC;
    }
    // (t1_chi of a stage-0 close is set by the caller from work_finish)
    PFX_HD void commit_finish() {
        const int st = stage;
        if (st == 0) {
            for (int k = 0; k < 4; ++k) p_t1[k] = lm.x[k];
            t1_fwhm = fabs(FWHM_K * lm.x[0]);
        }
        if (st == 7)
            for (int k = 0; k < 4; ++k) p2[k] = lm.x[k];
        if (st == 8)
            for (int k = 0; k < 4; ++k) p1[k] = lm.x[k];
        ier[st] = (int8_t)lm.info;
        ++stage;
        if (stage > 8) {
            finished = true;
            phase = GP_IDLE;
        } else {
            phase = GP_SETUP;
        }
    }

    PFX_HD void coop_setup(int src, const LaneBuf& b, const LaneWs& ws) {
        typedef WarpCoop WC;
        const LaneBuf bs = buf_of(b, src);
        const int st = WC::bc(stage, src), mm = WC::bc(m, src), ct = WC::bc(cut, src), dt = WC::bc(dtype, src);
        const long long cd = WC::bc(cand, src);
        const void* prof = (const void*)WC::bc((long long)profile, src);
        const double ph = WC::bc(ph_expect, src);
        SetupOut o;
        work_setup(src, bs, ws.blk, HY_WSN, st, ph, prof, dt, cd, mm, ct, o);
        if (WC::lane() == src) commit_setup(st, o);
    }

    PFX_HD void coop_finish(int src, const LaneBuf& b) {
        typedef WarpCoop WC;
        const LaneBuf bs = buf_of(b, src);
        const int st = WC::bc(stage, src), mm = WC::bc(m, src), ct = WC::bc(cut, src), dt = WC::bc(dtype, src);
        const long long cd = WC::bc(cand, src);
        const void* prof = (const void*)WC::bc((long long)profile, src);
        double p[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) p[k] = WC::bc(lm.x[k], src);
        const double chi = work_finish(bs, st, p, prof, dt, cd, mm, ct);
        if (WC::lane() == src) {
            if (st == 0) t1_chi = chi;
            commit_finish();
        }
    }

    // one trip of the kernel loop.  With one lane per candidate the kernel serves the SETUP and FIN
    // phases of the warp cooperatively before it calls step(); the host build (one lane) serves
    // itself here.
    PFX_HD void step(const LaneBuf& b, const LaneWs& ws) {
#if !defined(__CUDA_ARCH__)
        if (C == 1 && PFX_HYBRID) {
            // host build of the hybrid kernel's trip (the device kernel drives these pieces itself)
            if (phase == GP_SETUP) coop_setup(0, b, ws);
            if (phase == GP_JAC) {
                jac_prep(ws);
                host_group_jac(b, ws);
                jac_close(ws);
            }
            if (phase == GP_TRIAL) {
                if (g_host_lmpar_budget > 0) {
                    // as the block-synchronous kernels do: lmpar in pieces of at most `budget` iterations, the
                    // sub-problem reloaded from the workspace before every piece
                    LmparState lst;
                    bool pending = false;
                    while (!trial_prep_budget(ws, lst, pending, g_host_lmpar_budget)) pending = true;
                } else {
                    trial_prep(ws);
                }
                host_group_trial(b, ws);
                trial_close(ws);
            }
            if (phase == GP_FIN) coop_finish(0, b);
            return;
        }
#endif
        if (C == 1 && PFX_COOP) {
#if !defined(__CUDA_ARCH__)
            if (phase == GP_SETUP) coop_setup(0, b, ws);
#endif
        } else if (phase == GP_SETUP) {
            setup(b);
        }
        if (phase == GP_JAC) jac_pass(b, ws);
        if (phase == GP_TRIAL) trial_pass(b, ws);
        if (C == 1 && PFX_COOP) {
#if !defined(__CUDA_ARCH__)
            if (phase == GP_FIN) coop_finish(0, b);
#endif
        } else if (phase == GP_FIN) {
            finish(b);
        }
    }

};

typedef GaussLaneT<1> GaussLane;

// ---- final 8-parameter fit -----------------------------------------------------------------
template <int C = 1>
struct Gauss8LaneT {
    typedef LaneGroup<C> G;
    const void* profile;
    int dtype;
    long long cand;
    int m, cut, phase;
    LaneLM<8> lm;
    GaussCarry carry;
    double s10, s11;
    int8_t ier;
    bool finished;

    PFX_HD void start(const void* prof, int dt, long long c, int m_, const GaussCarry& cy, const LaneBuf& b) {
        profile = prof;
        dtype = dt;
        cand = c;
        m = m_;
        cut = m_ / 2;
        carry = cy;
        finished = false;
        for (int i = G::rank(); i < m; i += C) {
            int src = i + cut;
            if (src >= m) src -= m;
            b.d(i) = load_elem(profile, dtype, cand * (long long)m + src);
        }
        G::sync();
        double x0[8];
        for (int k = 0; k < 4; ++k) {
            x0[k] = cy.p1[k];
            x0[4 + k] = cy.p2[k];
        }
        lm.init(x0);
        phase = GP_JAC;
    }

    PFX_HD static double value(const double* p, double xv) {
        const double a = PFX_MUL(fabs(p[2]), gauss_fast(xv, p[1], 1.0 / fabs(p[0])));
        const double c = PFX_MUL(fabs(p[6]), gauss_fast(xv, p[5], 1.0 / fabs(p[4])));
        const double bg = PFX_MUL(PFX_ADD(fabs(p[3]), fabs(p[7])), 0.5);
        return PFX_ADD(PFX_ADD(a, c), bg);
    }

    // rows of one Jacobian pass dealt to the C lanes of a group (L.x set, L.jac_begin() done): every lane
    // eliminates its blocks of rows against its own factor in L
    PFX_HD static void rows_jac(LaneLM<8>& lm, const LaneBuf& b, int m) {
        const double* x = lm.x;
        const double* h = lm.h;
        const double s1 = 1.0 / fabs(x[0]), s1h = 1.0 / fabs(x[0] + h[0]), m1 = x[1], m1h = x[1] + h[1];
        const double a1 = fabs(x[2]), a1h = fabs(x[2] + h[2]);
        const double s2 = 1.0 / fabs(x[4]), s2h = 1.0 / fabs(x[4] + h[4]), m2 = x[5], m2h = x[5] + h[5];
        const double a2 = fabs(x[6]), a2h = fabs(x[6] + h[6]);
        const double bg = PFX_MUL(PFX_ADD(fabs(x[3]), fabs(x[7])), 0.5);
        const double bg3 = PFX_MUL(PFX_ADD(fabs(x[3] + h[3]), fabs(x[7])), 0.5);
        const double bg7 = PFX_MUL(PFX_ADD(fabs(x[3]), fabs(x[7] + h[7])), 0.5);
        double ih[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) ih[k] = 1.0 / h[k];
#ifndef PFX_G8_BLOCK
#define PFX_G8_BLOCK 2
#endif
        constexpr int B = PFX_G8_BLOCK;
#pragma unroll 1
        for (int i0 = G::rank() * B; i0 < m; i0 += C * B) {
            double A[B][8], bv[B];
#pragma unroll
            for (int r = 0; r < B; ++r) {
                const int i = i0 + r;
                if (i < m) {
                    const double xv = (double)i, yv = b.d(i);
                    const double e1 = gauss_fast(xv, m1, s1), e2 = gauss_fast(xv, m2, s2);
                    const double g1 = PFX_MUL(a1, e1), g2 = PFX_MUL(a2, e2);
                    const double f0 = PFX_SUB(yv, PFX_ADD(PFX_ADD(g1, g2), bg));
                    A[r][0] = (PFX_SUB(yv, PFX_ADD(PFX_ADD(PFX_MUL(a1, gauss_fast(xv, m1, s1h)), g2), bg)) - f0) * ih[0];
                    A[r][1] = (PFX_SUB(yv, PFX_ADD(PFX_ADD(PFX_MUL(a1, gauss_fast(xv, m1h, s1)), g2), bg)) - f0) * ih[1];
                    A[r][2] = (PFX_SUB(yv, PFX_ADD(PFX_ADD(PFX_MUL(a1h, e1), g2), bg)) - f0) * ih[2];
                    A[r][3] = (PFX_SUB(yv, PFX_ADD(PFX_ADD(g1, g2), bg3)) - f0) * ih[3];
                    A[r][4] = (PFX_SUB(yv, PFX_ADD(PFX_ADD(g1, PFX_MUL(a2, gauss_fast(xv, m2, s2h))), bg)) - f0) * ih[4];
                    A[r][5] = (PFX_SUB(yv, PFX_ADD(PFX_ADD(g1, PFX_MUL(a2, gauss_fast(xv, m2h, s2))), bg)) - f0) * ih[5];
                    A[r][6] = (PFX_SUB(yv, PFX_ADD(PFX_ADD(g1, PFX_MUL(a2h, e2)), bg)) - f0) * ih[6];
                    A[r][7] = (PFX_SUB(yv, PFX_ADD(PFX_ADD(g1, g2), bg7)) - f0) * ih[7];
                    bv[r] = f0;
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) A[r][k] = 0.0;
                    bv[r] = 0.0;
                }
            }
            lm.template jac_block<B>(A, bv);
        }
    }
    PFX_HD void jac_pass(const LaneBuf& b, const LaneWs& ws) {
        PFX_PROF_T0(pt0);
        lm.jac_begin();
        rows_jac(lm, b, m);
        lm.template jac_combine<C>();
        PFX_PROF_ADD(PROF8_JAC_ROWS, pt0);
        PFX_PROF_T0(pt1);
        const int r = lm.template jac_end_shared<C>(ws, m);      // one n x n workspace per group
        phase = (r == LANE_TRIAL) ? GP_TRIAL : GP_FIN;
        PFX_PROF_ADD(PROF8_JAC_END, pt1);
    }

    PFX_HD void trial_pass(const LaneBuf& b, const LaneWs& ws) {
        PFX_PROF_T0(pt0);
        lm.template trial_begin_shared<C>(ws);
        double s = 0.0;
        for (int i = G::rank(); i < m; i += C) {
            const double f = PFX_SUB(b.d(i), value(lm.xnew, (double)i));
            s = PFX_FMA(f, f, s);
        }
        s = G::sum(s);
        const int r = lm.trial_end(sqrt(s));
        phase = (r == LANE_JAC) ? GP_JAC : (r == LANE_TRIAL ? GP_TRIAL : GP_FIN);
        PFX_PROF_ADD(PROF8_TRIAL, pt0);
    }

    // PO:1407-1428 and PO:747-761
    PFX_HD void finish(const LaneBuf& b) {
        ier = (int8_t)lm.info;
        const double* q = lm.x;
        const double fwhm1 = fabs(FWHM_K * q[0]), fwhm2 = fabs(FWHM_K * q[4]);
        const G4Flags rf = {false, true};
        const G4Flags tf = {true, false};
        const double bgmix = PFX_MUL(PFX_ADD(carry.p1[3], carry.p2[3]), 0.5);
        double c = 0.0, cc = 0.0;
        // pass 1: both chi-squares
        for (int i = G::rank(); i < m; i += C) {
            const double xv = (double)i, yv = b.d(i);
            const double fi = value(q, xv);
            if (fi >= 1.0) {
                const double dlt = yv - fi;
                c += (dlt * dlt) / (double)m;
            }
            const double cf = PFX_ADD(PFX_SUB(PFX_SUB(PFX_ADD(g4_value(carry.p1, rf, xv), g4_value(carry.p2, rf, xv)),
                                                      carry.p1[3]), carry.p2[3]), bgmix);
            if (cf >= 1.0) {
                const double dlt = yv - cf;
                cc += (dlt * dlt) / (double)m;
            }
        }
        c = G::sum(c);
        cc = G::sum(cc);
        const bool use_final = (c <= cc);
        double fw_a, fw_b, dg_chi;
        if (use_final) {
            fw_a = fwhm1;
            fw_b = fwhm2;
            dg_chi = c;
        } else {
            fw_a = fabs(FWHM_K * carry.p2[0]);
            fw_b = fabs(FWHM_K * carry.p1[0]);
            dg_chi = cc;
        }
        // pass 2 and 3: std of (chosen fit) - (T1 fit + minbg - std)   (PO:755-756)
        double ds = 0.0;
        G::sync();
        for (int i = G::rank(); i < m; i += C) {
            const double xv = (double)i;
            const double dg = use_final ? value(q, xv)
                                        : PFX_ADD(PFX_SUB(PFX_SUB(PFX_ADD(g4_value(carry.p1, rf, xv), g4_value(carry.p2, rf, xv)),
                                                                  carry.p1[3]), carry.p2[3]), bgmix);
            const double dlt = dg - PFX_SUB(PFX_ADD(g4_value(carry.p_t1, tf, xv), carry.minbg), carry.sd);
            b.d(i) = dlt;       // the data vector is no longer needed
            ds += dlt;
        }
        ds = G::sum(ds);
        G::sync();
        const double dmean = ds / (double)m;
        double dv = 0.0;
        for (int i = G::rank(); i < m; i += C) {
            const double dlt = b.d(i) - dmean;
            dv = PFX_FMA(dlt, dlt, dv);
        }
        dv = G::sum(dv);
        const double dstd = fabs(sqrt(dv / (double)m));
        if (dstd < 3.0) s10 = carry.t1_fwhm;
        else s10 = (fw_b < fw_a) ? fw_b : fw_a;
        s11 = dg_chi;
        finished = true;
        phase = GP_IDLE;
    }

    // ---- block-synchronous kernel (k_gauss_final_s): row passes by a group of C lanes for a slot whose fit
    // state lives in its owner lane; hand-over through the slot's workspace w (element k at w[k * st]) ----
    //   G8_X  x of the pass (owner)        G8_XN  trial point (owner) / parameters of the close
    //   r rows 0..63  merged factor, unpacked (group leader)    G8_Q, G8_ACN2, G8_F2  Q^T f, column norms^2, |f|^2
    //   G8_TS  sum of squares at the trial point
    enum { G8_Q = 80, G8_ACN2 = 88, G8_F2 = 96, G8_TS = 97, G8_OP = 98, G8_CFIN = 99, G8_CNEW = 100,
           G8_X = 120, G8_XN = 128, G8_WSN = 136 };
#if defined(__CUDACC__)
    __device__ static void group_jac(const LaneBuf& b, int m, double* w, int st) {
        LaneLM<8> L;
#pragma unroll
        for (int k = 0; k < 8; ++k) L.x[k] = w[(size_t)(G8_X + k) * st];
        L.jac_begin();
        rows_jac(L, b, m);
        L.template jac_combine<C>();
        if (G::rank() == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    w[(size_t)(i + 8 * j) * st] = (i <= j) ? L.R[LaneLM<8>::ridx(i < j ? i : j, j)] : 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                w[(size_t)(G8_Q + k) * st] = L.q[k];
                w[(size_t)(G8_ACN2 + k) * st] = L.acn2[k];
            }
            w[(size_t)G8_F2 * st] = L.f2;
        }
    }
    __device__ static void group_trial(const LaneBuf& b, int m, double* w, int st) {
        double xn[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) xn[k] = w[(size_t)(G8_XN + k) * st];
        double s = 0.0;
        for (int i = G::rank(); i < m; i += C) {
            const double f = PFX_SUB(b.d(i), value(xn, (double)i));
            s = PFX_FMA(f, f, s);
        }
        s = G::sum(s);
        if (G::rank() == 0) w[(size_t)G8_TS * st] = s;
    }
#endif
    // tasks of the block-synchronous kernel, called by all lanes of a warp with the same arguments:
    // the rotated profile of a new candidate ...
    PFX_HD static void task_start(const LaneBuf& bs, const void* prof, int dt, long long cd, int mm) {
        typedef WarpCoop WC;
        WC::rows(mm, [&](int i) { bs.d(i) = GaussLaneT<1>::yrot_of(prof, dt, cd, mm, mm / 2, i); });
        WC::sync();
    }
    // ... and the close of a fit (PO:1407-1428, PO:747-761; finish() with the bins dealt to the warp)
    PFX_HD static void task_finish(const LaneBuf& bs, const double (&q)[8], const GaussCarry& cy, int mm, double& s10,
                                   double& s11) {
        typedef WarpCoop WC;
        const double fwhm1 = fabs(FWHM_K * q[0]), fwhm2 = fabs(FWHM_K * q[4]);
        const G4Flags rf = {false, true};
        const G4Flags tf = {true, false};
        const double bgmix = PFX_MUL(PFX_ADD(cy.p1[3], cy.p2[3]), 0.5);
        auto combi = [&](double xv) {
            return PFX_ADD(PFX_SUB(PFX_SUB(PFX_ADD(g4_value(cy.p1, rf, xv), g4_value(cy.p2, rf, xv)), cy.p1[3]), cy.p2[3]), bgmix);
        };
        const double c = WC::template sum<false>(mm, [&](int i) {
            const double fi = value(q, (double)i);
            const double dlt = bs.d(i) - fi;
            return (fi >= 1.0) ? (dlt * dlt) / (double)mm : 0.0;
        });
        const double cc = WC::template sum<false>(mm, [&](int i) {
            const double cf = combi((double)i);
            const double dlt = bs.d(i) - cf;
            return (cf >= 1.0) ? (dlt * dlt) / (double)mm : 0.0;
        });
        const bool use_final = (c <= cc);
        double fw_a, fw_b, dg_chi;
        if (use_final) {
            fw_a = fwhm1;
            fw_b = fwhm2;
            dg_chi = c;
        } else {
            fw_a = fabs(FWHM_K * cy.p2[0]);
            fw_b = fabs(FWHM_K * cy.p1[0]);
            dg_chi = cc;
        }
        // std of (chosen fit) - (T1 fit + minbg - std)   (PO:755-756)
        WC::sync();
        WC::rows(mm, [&](int i) {
            const double xv = (double)i;
            const double dg = use_final ? value(q, xv) : combi(xv);
            bs.d(i) = dg - PFX_SUB(PFX_ADD(g4_value(cy.p_t1, tf, xv), cy.minbg), cy.sd);      // the data vector is no longer needed
        });
        WC::sync();
        const double ds = WC::template sum<false>(mm, [&](int i) { return bs.d(i); });
        const double dmean = ds / (double)mm;
        const double dv = WC::template sum<true>(mm, [&](int i) { return bs.d(i) - dmean; });
        const double dstd = fabs(sqrt(dv / (double)mm));
        if (dstd < 3.0) s10 = cy.t1_fwhm;
        else s10 = (fw_b < fw_a) ? fw_b : fw_a;
        s11 = dg_chi;
        WC::sync();
    }

    PFX_HD void step(const LaneBuf& b, const LaneWs& ws) {
        if (phase == GP_JAC) jac_pass(b, ws);
        if (phase == GP_TRIAL) trial_pass(b, ws);
        if (phase == GP_FIN) {
            PFX_PROF_T0(pt);
            finish(b);
            PFX_PROF_ADD(PROF8_FINISH, pt);
        }
    }
};

typedef Gauss8LaneT<1> Gauss8Lane;

}  // namespace pfx
