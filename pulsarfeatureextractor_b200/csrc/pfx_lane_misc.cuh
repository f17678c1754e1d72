// pfx_lane_misc.cuh -- lane state machines for scores 1-4 (two sinusoid fits) and scores 12-19
// (DM-curve fit).  Same execution model as pfx_lane_gauss.cuh: a trip is "one Jacobian pass + one
// trial point"; a candidate belongs to a group of C lanes (LaneGroup, pfx_lane_lm.cuh) that share
// its rows; the data row is read straight from the input arrays (it stays in L1), only the small
// trust-region workspace lives in shared memory.
#pragma once

#include "pfx_driver.cuh"
#include "pfx_lane_gauss.cuh"

namespace pfx {

// =============================================================================================
// scores 1-4                                                                   PO:190-587
// =============================================================================================
// C lanes share the candidate (LaneGroup): the profile statistics of setup() are computed by every
// lane redundantly in the reference's order (scores 3 and 4 are exact), the rows of the two fits
// and of the chi-squares are dealt to the lanes.
template <int C = 1>
struct SinLaneT {
    typedef LaneGroup<C> G;
    const void* profile;
    int dtype;
    long long cand;
    int m, phase, stage, maxima;
    double half, s4, chi[2];
    LaneLM<2> lm;
    int8_t ier[2];

    PFX_HD double y(int i) const { return load_elem(profile, dtype, cand * (long long)m + i); }

    PFX_HD void start(const void* prof, int dt, long long c, int m_) {
        profile = prof;
        dtype = dt;
        cand = c;
        m = m_;
        stage = 0;
        phase = GP_SETUP;
        ier[0] = ier[1] = 0;
    }

    PFX_HD void setup() {
        if (stage == 0) {
            double s = 0.0, vmax = -INFINITY, vmin = INFINITY;
            for (int i = 0; i < m; ++i) {
                const double v = y(i);
                s += v;
                vmax = fmax(vmax, v);
                vmin = fmin(vmin, v);
            }
            const double mean = s / (double)m;
            double q = 0.0;
            for (int i = 0; i < m; ++i) {
                const double d = y(i) - mean;
                q = PFX_FMA(d, d, q);
            }
            const double sd = sqrt(q / (double)m);
            half = fabs(vmax - vmin) / 2.0;
            double acc = 0.0;
            for (int i = 0; i < m; ++i) acc += half - y(i);        // score 4, PO:246-247
            s4 = acc;
            // block scan of PO:294-326 (see count_peaks in pfx_scores.cuh)
            int peaks = 0, zeros = 0;
            bool nonzero = false;
            for (int i = 0; i < m; ++i) {
                double v = (y(i) - mean) - sd;
                if (v < 0.0) v = 0.0;
                if (v != 0.0) nonzero = true;
                else if (zeros < 4) ++zeros;
                else {
                    if (nonzero) ++peaks;
                    nonzero = false;
                    zeros = 0;
                }
            }
            if (nonzero) ++peaks;
            maxima = peaks;
        }
        double x0[2];
        const double y0 = y(0);
        if (stage == 0) {
            const double freq = (double)maxima / ((double)m - 1.0);
            double phi0;
            if (y0 == half) phi0 = 0.0;
            else if (y0 < half) phi0 = (freq != 0.0) ? -1.0 / (4.0 * freq) : -1.0 / (4.0 * 0.00000000001);
            else phi0 = (freq != 0.0) ? 1.0 / (4.0 * freq) : 1.0 / (4.0 * 0.00000000001);
            x0[0] = freq;
            x0[1] = phi0;
        } else {
            const double freq = (double)maxima / ((double)m - 1.0) / 2.0;
            double phi0;
            if (y0 == 0.0) phi0 = 0.0;
            else phi0 = (freq != 0.0) ? -1.0 / (4.0 * freq) : -1.0 / (4.0 * 0.00000000001);
            x0[0] = freq;
            x0[1] = phi0;
        }
        lm.init(x0);
        phase = GP_JAC;
    }

    // residual of row i: stage 0  y - (A sin t + bg)      (PO:418)
    //                    stage 1  (y - A sin^2 t) + bg    (PO:522; the +bg sits outside the bracket)
    PFX_HD double resid(double w, double phi, int i, double yv) const {
        const double sn = sin(PFX_ADD(PFX_MUL(w, (double)i), phi));
        if (stage == 0) return PFX_SUB(yv, PFX_ADD(PFX_MUL(half, sn), half));
        return PFX_ADD(PFX_SUB(yv, PFX_MUL(half, PFX_MUL(sn, sn))), half);
    }

    PFX_HD void jac_pass(const LaneWs& ws) {
        lm.jac_begin();
        const double w0 = PFX_MUL(TWO_PI, lm.x[0]), w1 = PFX_MUL(TWO_PI, lm.x[0] + lm.h[0]);
        const double p0 = lm.x[1], p1 = lm.x[1] + lm.h[1];
        const double ih0 = 1.0 / lm.h[0], ih1 = 1.0 / lm.h[1];
        constexpr int B = LANE_BLOCK_ROWS;
        for (int i0 = G::rank() * B; i0 < m; i0 += C * B) {
            const int rows = (m - i0 < B) ? m - i0 : B;
#pragma unroll 1
            for (int r = 0; r < rows; ++r) {
                const int i = i0 + r;
                const double yv = y(i);
                const double f0 = resid(w0, p0, i, yv);
                const double a0 = (resid(w1, p0, i, yv) - f0) * ih0;
                const double a1 = (resid(w0, p1, i, yv) - f0) * ih1;
                lm.acn2[0] = PFX_FMA(a0, a0, lm.acn2[0]);
                lm.acn2[1] = PFX_FMA(a1, a1, lm.acn2[1]);
                lm.f2 = PFX_FMA(f0, f0, lm.f2);
                ws.blk[r * 3 + 0] = a0;
                ws.blk[r * 3 + 1] = a1;
                ws.blk[r * 3 + 2] = f0;
            }
            lm.template jac_block_staged<B>(ws.blk, rows);
        }
        lm.template jac_combine<C>();
        const int r = lm.jac_end(ws, m);
        phase = (r == LANE_TRIAL) ? GP_TRIAL : GP_FIN;
    }

    PFX_HD void trial_pass(const LaneWs& ws) {
        lm.trial_begin(ws);
        const double w = PFX_MUL(TWO_PI, lm.xnew[0]), ph = lm.xnew[1];
        double s = 0.0;
        for (int i = G::rank(); i < m; i += C) {
            const double f = resid(w, ph, i, y(i));
            s = PFX_FMA(f, f, s);
        }
        s = G::sum(s);
        const int r = lm.trial_end(sqrt(s));
        phase = (r == LANE_JAC) ? GP_JAC : (r == LANE_TRIAL ? GP_TRIAL : GP_FIN);
    }

    PFX_HD void finish() {
        ier[stage] = (int8_t)lm.info;
        // chi-square against the evaluated curve A sin(..) + bg resp. A sin^2(..) + bg (PO:425,529)
        const double w = PFX_MUL(TWO_PI, lm.x[0]), ph = lm.x[1];
        double c = 0.0;
        for (int i = G::rank(); i < m; i += C) {
            const double sn = sin(PFX_ADD(PFX_MUL(w, (double)i), ph));
            const double fit = (stage == 0) ? PFX_ADD(PFX_MUL(half, sn), half)
                                            : PFX_ADD(PFX_MUL(half, PFX_MUL(sn, sn)), half);
            const double d = y(i) - fit;
            c = PFX_FMA(d, d, c);
        }
        c = G::sum(c);
        chi[stage] = c / (double)m;
        ++stage;
        phase = (stage > 1) ? GP_IDLE : GP_SETUP;
    }

    PFX_HD void step(const LaneWs& ws) {
        if (phase == GP_SETUP) setup();
        if (phase == GP_JAC) jac_pass(ws);
        if (phase == GP_TRIAL) trial_pass(ws);
        if (phase == GP_FIN) finish();
    }

    PFX_HD void write(double* feat) const {
        feat[0] = chi[0] / (double)maxima;      // maxima == 0 -> inf / nan as numpy (PO:373-374)
        feat[1] = chi[1] / (double)maxima;
        feat[2] = (double)(maxima > 0 ? maxima - 1 : 0);
        feat[3] = s4;
    }
};

typedef SinLaneT<1> SinLane;

// =============================================================================================
// scores 12-19                                              PHXO:81-233, PFDO:274-393, CFI:84-112
// =============================================================================================
template <int N, int C = 1>    // N = 3: PHCX family, 4: PFD family; C lanes share the candidate
struct DmLane {
    typedef LaneGroup<C> G;
    const void* dmcurve;
    int dm_dtype;
    long long cand;
    int n_dm, phase;
    bool pfd;
    double snr, dm, period, width, dm_start, dm_step, wint, df, f3, amp0, yscale;
    double s16, s17, s18, s19;
    unsigned status;
    LaneLM<N> lm;
    int8_t ier;

    PFX_HD double xval(int i) const {
        const double xi = pfd ? (double)i : (double)(128 * i - 1);      // PHXO:251-257
        return PFX_ADD(dm_start, PFX_MUL(xi, dm_step));
    }
    PFX_HD double yval(int i) const {
        if (pfd && dm_dtype == PFX_F32)   // float32 arithmetic of PFDO:332 on the float32 curve
            return (double)((float)yscale * ((const float*)dmcurve)[cand * (long long)n_dm + i]);
        const double v = load_elem(dmcurve, dm_dtype, cand * (long long)n_dm + i);
        return pfd ? yscale * v : v;
    }
    // sqrt((P - weff)/weff), weff = sqrt(wint + (prop kdm |dm + shift - x| df / f^3)^2)
    PFX_HD double shape(double prop, double shift, double xv) const {
        const double a = fabs(PFX_SUB(PFX_ADD(dm, shift), xv));
        const double q = PFX_DIV(PFX_MUL(PFX_MUL(PFX_MUL(prop, 8300000.000000001), a), df), f3);
        double weff = sqrt(PFX_ADD(wint, PFX_MUL(q, q)));
        if (pfd && weff > period) weff = period;
        return sqrt(PFX_DIV(PFX_SUB(period, weff), weff));
    }
    PFX_HD double theo_shape(double xv) const {
        const double q = PFX_DIV(PFX_MUL(PFX_MUL(8300000.000000001, fabs(PFX_SUB(dm, xv))), df), f3);
        double weff = sqrt(PFX_ADD(wint, PFX_MUL(q, q)));
        if (pfd && weff > period) weff = period;
        return sqrt(PFX_DIV(PFX_SUB(period, weff), weff));
    }

    PFX_HD void start(const void* curve, int dt, long long c, int n, int family, const double* prm) {
        dmcurve = curve;
        dm_dtype = dt;
        cand = c;
        n_dm = n;
        pfd = (family == PFX_FAMILY_PFD);
        snr = prm[0];
        dm = prm[1];
        period = prm[2] * 1000.0;
        width = prm[3];
        dm_start = prm[4];
        double length_all = prm[6];
        if (!(length_all > 0.0)) length_all = pfd ? (double)n : 128.0 * (double)n;
        dm_step = fabs(prm[4] - prm[5]) / length_all;
        wint = (width * period) * (width * period);
        df = pfd ? 32.0 : 400.0;
        f3 = pfd ? 2460375.0 : 2593941624.0;
        status = 0;
        ier = 0;
        phase = GP_SETUP;
    }

    PFX_HD void setup() {
        const double sw = sqrt(wint);
        s16 = snr / sqrt((period - sw) / sw);
        yscale = 1.0;
        if (pfd) {
            // yData = 255./max(yData)*yData  (PFDO:332), builtin max
            // builtin max: a NaN survives only when it is first, later ones lose every comparison;
            // the lanes of the group scan interleaved elements and merge (max of NaN-free values)
            if (dm_dtype == PFX_F32) {
                const float* p = (const float*)dmcurve + cand * (long long)n_dm;
                float mx = p[0];
                for (int i = 1 + G::rank(); i < n_dm; i += C)
                    if (p[i] > mx) mx = p[i];
                if (mx == mx) mx = (float)G::max_gt((double)mx);
                yscale = (double)(255.0f / mx);
            } else {
                double mx = load_elem(dmcurve, dm_dtype, cand * (long long)n_dm);
                for (int i = 1 + G::rank(); i < n_dm; i += C) {
                    const double v = load_elem(dmcurve, dm_dtype, cand * (long long)n_dm + i);
                    if (v > mx) mx = v;
                }
                if (mx == mx) mx = G::max_gt(mx);
                yscale = 255.0 / mx;
            }
        }
        // max(_help) with Python's builtin max: a NaN survives only when it is first
        double hmax = theo_shape(xval(0));
        for (int i = 1 + G::rank(); i < n_dm; i += C) {
            const double h = theo_shape(xval(i));
            if (h > hmax) hmax = h;
        }
        if (hmax == hmax) hmax = G::max_gt(hmax);
        amp0 = 255.0 / hmax;
        double x0[N];
        x0[0] = amp0;
        x0[1] = 1.0;
        x0[2] = 0.0;
        if (N == 4) x0[N - 1] = 0.0;
        lm.init(x0);
        phase = GP_JAC;
    }

    PFX_HD void jac_pass(const LaneWs& ws) {
        lm.jac_begin();
        const double* x = lm.x;
        const double* h = lm.h;
        double ih[N];
#pragma unroll
        for (int k = 0; k < N; ++k) ih[k] = 1.0 / h[k];
        const double up = (N == 4) ? x[N - 1] : 0.0;
        constexpr int B = LANE_BLOCK_ROWS;
        for (int i0 = G::rank() * B; i0 < n_dm; i0 += C * B) {     // blocks of rows dealt to the lanes
            const int rows = (n_dm - i0 < B) ? n_dm - i0 : B;
#pragma unroll 1
            for (int r = 0; r < rows; ++r) {
                const int i = i0 + r;
                const double xv = xval(i), yv = yval(i);
                const double s0 = shape(x[1], x[2], xv);
                const double s1 = shape(x[1] + h[1], x[2], xv);
                const double s2 = shape(x[1], x[2] + h[2], xv);
                double a[N], f0;
                if (N == 3) {
                    f0 = PFX_SUB(yv, PFX_MUL(x[0], s0));
                    a[0] = (PFX_SUB(yv, PFX_MUL(x[0] + h[0], s0)) - f0) * ih[0];
                    a[1] = (PFX_SUB(yv, PFX_MUL(x[0], s1)) - f0) * ih[1];
                    a[2] = (PFX_SUB(yv, PFX_MUL(x[0], s2)) - f0) * ih[2];
                } else {
                    f0 = PFX_SUB(yv, PFX_ADD(up, PFX_MUL(x[0], s0)));
                    a[0] = (PFX_SUB(yv, PFX_ADD(up, PFX_MUL(x[0] + h[0], s0))) - f0) * ih[0];
                    a[1] = (PFX_SUB(yv, PFX_ADD(up, PFX_MUL(x[0], s1))) - f0) * ih[1];
                    a[2] = (PFX_SUB(yv, PFX_ADD(up, PFX_MUL(x[0], s2))) - f0) * ih[2];
                    a[N - 1] = (PFX_SUB(yv, PFX_ADD(up + h[N - 1], PFX_MUL(x[0], s0))) - f0) * ih[N - 1];
                }
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    lm.acn2[k] = PFX_FMA(a[k], a[k], lm.acn2[k]);
                    ws.blk[r * (N + 1) + k] = a[k];
                }
                lm.f2 = PFX_FMA(f0, f0, lm.f2);
                ws.blk[r * (N + 1) + N] = f0;
            }
            lm.template jac_block_staged<B>(ws.blk, rows);
        }
        lm.template jac_combine<C>();
        const int r = lm.jac_end(ws, n_dm);
        phase = (r == LANE_TRIAL) ? GP_TRIAL : GP_FIN;
    }

    PFX_HD double model(const double* p, double xv) const {
        const double v = PFX_MUL(p[0], shape(p[1], p[2], xv));
        return (N == 4) ? PFX_ADD(p[N - 1], v) : v;
    }

    PFX_HD void trial_pass(const LaneWs& ws) {
        lm.trial_begin(ws);
        double s = 0.0;
        for (int i = G::rank(); i < n_dm; i += C) {
            const double f = PFX_SUB(yval(i), model(lm.xnew, xval(i)));
            s = PFX_FMA(f, f, s);
        }
        s = G::sum(s);
        const int r = lm.trial_end(sqrt(s));
        phase = (r == LANE_JAC) ? GP_JAC : (r == LANE_TRIAL ? GP_TRIAL : GP_FIN);
    }

    PFX_HD void finish() {
        ier = (int8_t)lm.info;
        s17 = fabs(1.0 - lm.x[1]);
        s18 = fabs(lm.x[2]);                          // filterScore(18, .), CFI:110
        if (!pfd) {
            // PHXO:222-230: chi-square of the THEORETICAL curve where the FIT is >= 1
            double c = 0.0;
            for (int i = G::rank(); i < n_dm; i += C) {
                const double xv = xval(i);
                if (model(lm.x, xv) >= 1.0) {
                    const double d = yval(i) - PFX_MUL(amp0, theo_shape(xv));
                    c = PFX_FMA(d, d, c);
                }
            }
            c = G::sum(c);
            s19 = c / (double)n_dm;
        } else {
            double c = 0.0;
            int ndeg = 0;
            for (int i = G::rank(); i < n_dm; i += C) {
                const double th = PFX_MUL(amp0, theo_shape(xval(i)));
                if (th > 0.0) {
                    const double d = yval(i) - th;
                    c += (d * d) / th;
                    ++ndeg;
                }
            }
            c = G::sum(c);
            ndeg = G::sum(ndeg);
            if (ndeg == 0) {
                status |= PFX_ST_DM_NDEG_ZERO;      // ZeroDivisionError, PFDO:385
                s19 = NAN;
            } else {
                s19 = c / (double)ndeg;
            }
        }
        phase = GP_IDLE;
    }

    PFX_HD void step(const LaneWs& ws) {
        if (phase == GP_SETUP) setup();
        if (phase == GP_JAC) jac_pass(ws);
        if (phase == GP_TRIAL) trial_pass(ws);
        if (phase == GP_FIN) finish();
    }

    PFX_HD void write(double* feat) const {
        feat[11] = period;
        feat[12] = filter_13_14(snr);
        feat[13] = filter_13_14(dm);
        feat[14] = width;
        if (!(status & PFX_ST_FAIL_MASK)) {
            feat[15] = s16;
            feat[16] = s17;
            feat[17] = s18;
            feat[18] = s19;
        }
    }
};

// =============================================================================================
// PFD family: S/N and width derived from the profile              PFDOperations.py:93-233
// =============================================================================================
// snr   = sum((p - avg') / sqrt(var')) with avg', var' over the bins inside avg +- 3 sigma; < 0 -> 0.1
// width = (right - left - 1) / n from the half-maximum walk on the profile rotated so that its peak
//         sits at n//2, minus its minimum.  fft_rotate (PFDO:490-500) by an integer number of bins
//         is a circular shift; the reference's FFT round trip adds ~1e-13 of noise, which can only
//         matter when a bin equals the half maximum exactly.
// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src: pairwise_sum): fewer than
// 8 terms sequentially; up to 128 terms with eight interleaved accumulators combined as
// ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)) and a sequential remainder; longer inputs split in halves
// (the first a multiple of 8).  Reproduced because the PFD S/N below is a sum that cancels to
// rounding noise whenever no bin is clipped, and its SIGN decides between ~1e-15 and 0.1.
template <class F>
PFX_HD double np_pairwise_sum(const F& f, int lo, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r = PFX_ADD(r, f(lo + i));      // PFX_ADD: never fused with a product
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = f(lo + k);
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] = PFX_ADD(r[k], f(lo + i + k));
        double res = PFX_ADD(PFX_ADD(PFX_ADD(r[0], r[1]), PFX_ADD(r[2], r[3])),
                             PFX_ADD(PFX_ADD(r[4], r[5]), PFX_ADD(r[6], r[7])));
        for (; i < n; ++i) res = PFX_ADD(res, f(lo + i));
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return PFX_ADD(np_pairwise_sum(f, lo, n2), np_pairwise_sum(f, lo + n2, n - n2));
}

PFX_HD inline void pfd_params(const void* profile, int dtype, long long cand, int n, double& snr, double& width) {
    const long long off = cand * (long long)n;
    auto p = [&](int i) { return load_elem(profile, dtype, off + i); };
    // profile.mean(), profile.var() (numpy: pairwise sums)
    const double avg = np_pairwise_sum(p, 0, n) / (double)n;
    const double var = np_pairwise_sum([&](int i) { const double d = PFX_SUB(p(i), avg); return PFX_MUL(d, d); }, 0, n) / (double)n;
    const double sigma = sqrt(var);
    // the clipped sub-array is compacted by the reference; its sums run over the kept bins in order
    const double lo3 = PFX_SUB(avg, PFX_MUL(3.0, sigma)), hi3 = PFX_ADD(avg, PFX_MUL(3.0, sigma));
    int cnt = 0;
    for (int i = 0; i < n; ++i) cnt += (p(i) > lo3 && p(i) < hi3) ? 1 : 0;
    auto kept = [&](int k) {            // k-th kept bin (n is small: 64..1024)
        int seen = -1;
        for (int i = 0; i < n; ++i) {
            const double v = p(i);
            if (v > lo3 && v < hi3) {
                if (++seen == k) return v;
            }
        }
        return (double)NAN;
    };
    const bool all_kept = (cnt == n);
    double avg2, var2;
    if (all_kept) {
        avg2 = avg;
        var2 = var;
    } else {
        avg2 = np_pairwise_sum(kept, 0, cnt) / (double)cnt;      // cnt == 0 -> NaN like numpy
        var2 = np_pairwise_sum([&](int k) { const double d = PFX_SUB(kept(k), avg2); return PFX_MUL(d, d); }, 0, cnt) / (double)cnt;
    }
    const double sd2 = sqrt(var2);
    const double acc = np_pairwise_sum([&](int i) { return PFX_DIV(PFX_SUB(p(i), avg2), sd2); }, 0, n);
    snr = (acc < 0.0) ? 0.1 : acc;
    int peak = 0;
    double vmax = p(0), vmin = vmax;
    for (int i = 1; i < n; ++i) {
        const double v = p(i);
        if (argmax_better(v, i, vmax, peak)) {
            vmax = v;
            peak = i;
        }
        if (v < vmin) vmin = v;
    }
    const int shift = peak - n / 2;
    auto rot = [&](int k) {
        int src = (k + shift) % n;
        if (src < 0) src += n;
        return load_elem(profile, dtype, off + src) - vmin;
    };
    int rpeak = 0;
    double rmax = rot(0);
    for (int k = 1; k < n; ++k) {
        const double v = rot(k);
        if (argmax_better(v, k, rmax, rpeak)) {
            rmax = v;
            rpeak = k;
        }
    }
    const double half = rmax / 2.0;
    int left = rpeak;
    while (left > 0) {
        if (rot(left) < half) break;
        --left;
    }
    int right = rpeak;
    while (right < n) {
        if (rot(right) < half) break;
        ++right;
    }
    width = (1.0 * ((double)(right - left) - 1.0)) / (double)n;
}

}  // namespace pfx
