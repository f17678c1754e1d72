// pfx_lane_hist.cuh -- scores 5-7 with one lane per candidate: Freedman-Diaconis histograms of the
// profile and of its derivative and the three Gaussian fits on them
// (ProfileOperations.py:654-722,774-1057, ProfileOperationsInterface.py:138-186).
//
// Per-lane working arrays (profile, derivative, sort buffer, bin edges, counts) live in a global
// scratch area laid out [element][lane], so the 32 lanes of a warp read consecutive float64.
// Stages of the state machine:  0 derivative-histogram fit (with the reference's retry loop),
// 1 profile-histogram fit (retry loop), 2 fixed-expectation fit.  A histogram with more bins than
// the per-lane capacity is not truncated: the candidate is handed to the warp-form kernel
// (pfx_k_hist.cu), which has a 4096-bin scratch slot per warp.
#pragma once

#include "pfx_lane_gauss.cuh"

namespace pfx {

constexpr int HIST_LANE_CAP = 256;

PFX_HD inline int hist_lane_doubles(int m) { return 2 * m + next_pow2(m) + 2 * HIST_LANE_CAP; }

struct HistLaneBuf {
    SPtr<double> y, dy, srt, hx, hy;
    PFX_HD void carve(double* base, int stride, int m) {
        auto at = [&](int k) { SPtr<double> v; v.p = base + (size_t)k * stride; v.stride = stride; return v; };
        y = at(0);
        dy = at(m);
        srt = at(2 * m);
        hx = at(2 * m + next_pow2(m));
        hy = at(2 * m + next_pow2(m) + HIST_LANE_CAP);
    }
    PFX_HD void pad(int i) const {      // PO:943-947 zero padding to three points
        hx[i] = 0.0;
        hy[i] = 0.0;
    }
};

// The histogram of a uint8 profile of at most 255 bins as k_hist_lane_u8 keeps it: the counts are
// bytes in shared memory (a count never exceeds the number of bins of the profile) and the left
// edges are not stored at all -- numpy's edge i is i * step + first, two operations.  Reads only; the
// kernel builds the table with the whole warp (pfx_k_hist.cu).  Element i of the table of a lane is
// stored at (i + rot) & 255 with rot = 4 * lane, which spreads the 32 lanes of a warp over the 32
// banks when every lane reads its own bin i, and keeps the bins of one lane consecutive when the
// warp fills one table together.
struct HistEdgeView {
    double first, step;
    int nb;
    PFX_HD double operator[](int i) const { return (i < nb) ? PFX_ADD(PFX_MUL((double)i, step), first) : 0.0; }
};
struct HistCountView {
    const unsigned char* p;
    int rot, nb;
    PFX_HD double operator[](int i) const { return (i < nb) ? (double)p[(i + rot) & 255] : 0.0; }
};
struct HistBufU8 {
    HistEdgeView hx;
    HistCountView hy;
    PFX_HD void pad(int) const {}       // entries beyond the histogram read as zeros
};

template <class BUF>
struct HistLaneT {
    long long cand;
    int m, phase, stage;
    bool int_data;
    double pow_m, pow_m1;
    // histogram in hx/hy
    int nb, nfit, first_max, counter, nfits;
    double sigma0, max0, meansq, expect;
    int nbp, nbd;
    double pmin, pmax, dmin, dmax;
    // results
    double dy_expect, pp[3], pf[2], xmax;
    unsigned status;
    bool overflow;
    int8_t ier[3], counts[2];
    LaneLM<3> lm3;
    LaneLM<2> lm2;

    PFX_HD void start(long long c, int m_, bool isint, double pm, double pm1) {
        cand = c;
        m = m_;
        int_data = isint;
        pow_m = pm;
        pow_m1 = pm1;
        stage = 0;
        phase = GP_SETUP;
        status = 0;
        overflow = false;
        ier[0] = ier[1] = ier[2] = 0;
        counts[0] = counts[1] = 0;
    }

    // in-place bitonic sort of srt[0..n2): data independent control flow, all lanes in step
    PFX_HD static void sort(const SPtr<double>& a, int n2) {
        for (int k = 2; k <= n2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1)
                for (int i = 0; i < n2; ++i) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const double x = a[i], z = a[ixj];
                        const bool up = ((i & k) == 0);
                        if ((x > z) == up) {
                            a[i] = z;
                            a[ixj] = x;
                        }
                    }
                }
    }
    PFX_HD static double percentile(const SPtr<double>& srt, int n, double per) {
        const double idx = per / 100.0 * (double)(n - 1);
        const int i = (int)idx;
        if ((double)i == idx) return srt[i] * 1.0 / 1.0;
        const int j = i + 1;
        const double w0 = (double)j - idx, w1 = idx - (double)i;
        return (srt[i] * w0 + srt[i + 1] * w1) / (w0 + w1);
    }
    // POI:138-166
    PFX_HD long long fd_bins(const SPtr<double>& d, int n, const SPtr<double>& srt, double pow_n, double& lo, double& hi) const {
        int n2 = 1;
        while (n2 < n) n2 <<= 1;
        for (int i = 0; i < n2; ++i) srt[i] = (i < n) ? d[i] : INFINITY;
        sort(srt, n2);
        const double iqr = percentile(srt, n, 75.0) - percentile(srt, n, 25.0);
        lo = srt[0];
        hi = srt[n - 1];
        const double binwidth = 2.0 * iqr * pow_n;
        const double rnge = hi - lo;
        double q;
        if (binwidth <= 0.0) q = int_data ? floor(rnge / 60.0) : rnge / 60.0;
        else q = rnge / binwidth;
        const double c = ceil(q);
        if (!(c < 9.0e15)) return 9000000000000000LL;
        return (long long)c;
    }
    // numpy.histogram uniform-bin path; fills hx (left edges) and hy (counts)
    PFX_HD static void histogram(const SPtr<double>& d, int n, int nbins, double first, double last,
                                 const SPtr<double>& hx, const SPtr<double>& hy) {
        if (first == last) {
            first -= 0.5;
            last += 0.5;
        }
        const double step = (last - first) / (double)nbins;
        for (int i = 0; i < nbins; ++i) {
            hx[i] = PFX_ADD(PFX_MUL((double)i, step), first);
            hy[i] = 0.0;
        }
        const double denom = last - first;
        for (int r = 0; r < n; ++r) {
            const double a = d[r];
            if (!(a >= first && a <= last)) continue;
            const double fi = PFX_MUL(PFX_DIV(PFX_SUB(a, first), denom), (double)nbins);
            int idx = (int)fi;
            if (idx == nbins) idx -= 1;
            if (a < hx[idx]) idx -= 1;
            else {
                const double e_hi = (idx + 1 == nbins) ? last : hx[idx + 1];
                if (a >= e_hi && idx != nbins - 1) idx += 1;
            }
            hy[idx] += 1.0;
        }
    }

    // set up the fit of the histogram currently in hx/hy (PO:829-857, 943-947)
    PFX_HD void begin_hist_fit(const BUF& b, int nbins) {
        nb = nbins;
        int am = 0;
        double vmax = b.hy[0], bmax = b.hy[0], s = 0.0;
        for (int i = 0; i < nbins; ++i) {
            const double v = b.hy[i];
            s += v;
            if (i > 0) {
                if (argmax_better(v, i, vmax, am)) {
                    vmax = v;
                    am = i;
                }
                if (v > bmax) bmax = v;
            }
        }
        const double mean = s / (double)nbins;
        double q = 0.0;
        for (int i = 0; i < nbins; ++i) {
            const double dlt = b.hy[i] - mean;
            q = PFX_FMA(dlt, dlt, q);
        }
        sigma0 = sqrt(q / (double)nbins);
        max0 = bmax;
        meansq = mean * mean;
        first_max = am;
        expect = b.hx[am];
        nfit = nbins;
        for (int i = nbins; i < 3; ++i) {
            b.pad(i);
            nfit = 3;
        }
        counter = 0;
        nfits = 0;
    }

    // histogram currently described by the lane (k_hist_lane_u8 builds its views from these)
    double h_first, h_step;
    int h_nb;

    // heavy part of a set-up: order statistics, bin counts, the histogram (stages 0 and 1)
    PFX_HD void setup_data(const BUF& b, const void* profile, int dtype) {
        if (stage == 0) {
            for (int i = 0; i < m; ++i) b.y[i] = load_elem(profile, dtype, cand * (long long)m + i);
            for (int i = 0; i < m - 1; ++i) b.dy[i] = b.y[i] - b.y[i + 1];
            const long long np_ = fd_bins(b.y, m, b.srt, pow_m, pmin, pmax);
            const long long nd_ = fd_bins(b.dy, m - 1, b.srt, pow_m1, dmin, dmax);
            if (nd_ < 1 || np_ < 1) {
                status |= PFX_ST_HIST_ZERO_BINS;
                phase = GP_IDLE;
                return;
            }
            if (nd_ > HIST_LANE_CAP || np_ > HIST_LANE_CAP) {
                overflow = true;
                phase = GP_IDLE;
                return;
            }
            nbp = (int)np_;
            nbd = (int)nd_;
            histogram(b.dy, m - 1, nbd, dmin, dmax, b.hx, b.hy);
        } else if (stage == 1) {
            histogram(b.y, m, nbp, pmin, pmax, b.hx, b.hy);
        }
    }
    PFX_HD void setup(const BUF& b, const void* profile, int dtype) {
        setup_data(b, profile, dtype);
        if (phase == GP_SETUP) setup_fit(b);
    }
    // the rest: statistics of the counts, starting values
    PFX_HD void setup_fit(const BUF& b) {
        if (stage == 0) begin_hist_fit(b, nbd);
        else if (stage == 1) begin_hist_fit(b, nbp);
        if (stage <= 1) {
            double x0[3] = {sigma0, expect, max0};
            lm3.init(x0);
        } else {
            // fixed-expectation fit on the (unpadded) profile histogram, PO:988-1057
            if (nbp < 2) {
                status |= PFX_ST_LSQ_IMPROPER;
                phase = GP_IDLE;
                return;
            }
            int xi = nbp / 2 - 1;
            if (xi < 0) xi += nbp;
            xmax = b.hx[xi];
            double s = 0.0, bmax = b.hy[0];
            for (int i = 0; i < nbp; ++i) {
                s += b.hy[i];
                if (i > 0 && b.hy[i] > bmax) bmax = b.hy[i];
            }
            const double mean = s / (double)nbp;
            double q = 0.0;
            for (int i = 0; i < nbp; ++i) {
                const double dlt = b.hy[i] - mean;
                q = PFX_FMA(dlt, dlt, q);
            }
            double x0[2] = {sqrt(q / (double)nbp), bmax};
            lm2.init(x0);
        }
        phase = GP_JAC;
    }

    PFX_HD void jac_pass(const BUF& b, const LaneWs& ws) {
        constexpr int B = LANE_BLOCK_ROWS;
        if (stage <= 1) {
            lm3.jac_begin();
            const double* x = lm3.x;
            const double* h = lm3.h;
            const double isg = 1.0 / x[0], isgh = 1.0 / (x[0] + h[0]);
            const double mu = x[1], muh = x[1] + h[1], amx = fabs(x[2]), amxh = fabs(x[2] + h[2]);
            const double ih0 = 1.0 / h[0], ih1 = 1.0 / h[1], ih2 = 1.0 / h[2];
            for (int i0 = 0; i0 < nfit; i0 += B) {
                const int rows = (nfit - i0 < B) ? nfit - i0 : B;
#pragma unroll 1
                for (int r = 0; r < rows; ++r) {
                    const double xv = b.hx[i0 + r], yv = b.hy[i0 + r];
                    const double e0 = gauss_fast(xv, mu, isg);
                    const double f0 = PFX_SUB(yv, PFX_MUL(amx, e0));
                    const double a0 = (PFX_SUB(yv, PFX_MUL(amx, gauss_fast(xv, mu, isgh))) - f0) * ih0;
                    const double a1 = (PFX_SUB(yv, PFX_MUL(amx, gauss_fast(xv, muh, isg))) - f0) * ih1;
                    const double a2 = (PFX_SUB(yv, PFX_MUL(amxh, e0)) - f0) * ih2;
                    lm3.acn2[0] = PFX_FMA(a0, a0, lm3.acn2[0]);
                    lm3.acn2[1] = PFX_FMA(a1, a1, lm3.acn2[1]);
                    lm3.acn2[2] = PFX_FMA(a2, a2, lm3.acn2[2]);
                    lm3.f2 = PFX_FMA(f0, f0, lm3.f2);
                    ws.blk[r * 4 + 0] = a0;
                    ws.blk[r * 4 + 1] = a1;
                    ws.blk[r * 4 + 2] = a2;
                    ws.blk[r * 4 + 3] = f0;
                }
                lm3.jac_block_staged<B>(ws.blk, rows);
            }
            const int r = lm3.jac_end(ws, nfit);
            phase = (r == LANE_TRIAL) ? GP_TRIAL : GP_FIN;
        } else {
            lm2.jac_begin();
            const double* x = lm2.x;
            const double* h = lm2.h;
            const double isg = 1.0 / x[0], isgh = 1.0 / (x[0] + h[0]);
            const double amx = fabs(x[1]), amxh = fabs(x[1] + h[1]);
            const double ih0 = 1.0 / h[0], ih1 = 1.0 / h[1];
            for (int i0 = 0; i0 < nbp; i0 += B) {
                const int rows = (nbp - i0 < B) ? nbp - i0 : B;
#pragma unroll 1
                for (int r = 0; r < rows; ++r) {
                    const double xv = b.hx[i0 + r], yv = b.hy[i0 + r];
                    const double e0 = gauss_fast(xv, xmax, isg);
                    const double f0 = PFX_SUB(yv, PFX_MUL(amx, e0));
                    const double a0 = (PFX_SUB(yv, PFX_MUL(amx, gauss_fast(xv, xmax, isgh))) - f0) * ih0;
                    const double a1 = (PFX_SUB(yv, PFX_MUL(amxh, e0)) - f0) * ih1;
                    lm2.acn2[0] = PFX_FMA(a0, a0, lm2.acn2[0]);
                    lm2.acn2[1] = PFX_FMA(a1, a1, lm2.acn2[1]);
                    lm2.f2 = PFX_FMA(f0, f0, lm2.f2);
                    ws.blk[r * 3 + 0] = a0;
                    ws.blk[r * 3 + 1] = a1;
                    ws.blk[r * 3 + 2] = f0;
                }
                lm2.jac_block_staged<B>(ws.blk, rows);
            }
            const int r = lm2.jac_end(ws, nbp);
            phase = (r == LANE_TRIAL) ? GP_TRIAL : GP_FIN;
        }
    }

    PFX_HD void trial_pass(const BUF& b, const LaneWs& ws) {
        int r;
        if (stage <= 1) {
            lm3.trial_begin(ws);
            const double isg = 1.0 / lm3.xnew[0], mu = lm3.xnew[1], amx = fabs(lm3.xnew[2]);
            double s = 0.0;
            for (int i = 0; i < nfit; ++i) {
                const double f = PFX_SUB(b.hy[i], PFX_MUL(amx, gauss_fast(b.hx[i], mu, isg)));
                s = PFX_FMA(f, f, s);
            }
            r = lm3.trial_end(sqrt(s));
        } else {
            lm2.trial_begin(ws);
            const double isg = 1.0 / lm2.xnew[0], amx = fabs(lm2.xnew[1]);
            double s = 0.0;
            for (int i = 0; i < nbp; ++i) {
                const double f = PFX_SUB(b.hy[i], PFX_MUL(amx, gauss_fast(b.hx[i], xmax, isg)));
                s = PFX_FMA(f, f, s);
            }
            r = lm2.trial_end(sqrt(s));
        }
        phase = (r == LANE_JAC) ? GP_JAC : (r == LANE_TRIAL ? GP_TRIAL : GP_FIN);
    }

    PFX_HD void finish(const BUF& b) {
        if (stage <= 1) {
            // PO:957-978: chi-square over the original bins, retry test, moved start position
            ++nfits;
            ier[stage] = (int8_t)lm3.info;
            const double p0 = lm3.x[0], p1 = lm3.x[1], p2 = lm3.x[2];
            const double isg = 1.0 / p0, amx = fabs(p2);
            double c = 0.0;
            for (int i = 0; i < nb; ++i) {
                const double dlt = b.hy[i] - PFX_MUL(amx, gauss_fast(b.hx[i], p1, isg));
                c = PFX_FMA(dlt, dlt, c);
            }
            const double chisq = c / (double)nfit;
            bool again = false;
            if ((chisq > meansq * (double)nb) && (p0 < 0.2 * (double)nb)) {
                ++counter;
                const int len_before = nb - (counter - 1);
                if (first_max >= len_before) {
                    status |= PFX_ST_INDEX_OOB;
                } else if (len_before - 1 == 0) {
                    status |= PFX_ST_EMPTY_SEQUENCE;
                } else {
                    // argmax of the counts with [first_max, first_max + counter) removed
                    int best = -1;
                    double bv = 0.0;
                    for (int i = 0; i < nb; ++i) {
                        if (i >= first_max && i < first_max + counter) continue;
                        const int ti = (i < first_max) ? i : i - counter;
                        if (best < 0 || argmax_better(b.hy[i], ti, bv, best)) {
                            bv = b.hy[i];
                            best = ti;
                        }
                    }
                    if (best + counter >= nfit) status |= PFX_ST_INDEX_OOB;
                    else {
                        expect = b.hx[best + counter];
                        again = (counter <= 5);
                    }
                }
            }
            if (status & PFX_ST_FAIL_MASK) {
                counts[stage] = (int8_t)nfits;
                phase = GP_IDLE;
                return;
            }
            if (again) {
                double x0[3] = {sigma0, expect, max0};
                lm3.init(x0);
                phase = GP_JAC;
                return;
            }
            counts[stage] = (int8_t)nfits;
            if (stage == 0) dy_expect = p1;
            else {
                pp[0] = p0;
                pp[1] = p1;
                pp[2] = p2;
            }
            ++stage;
            phase = GP_SETUP;
        } else {
            ier[2] = (int8_t)lm2.info;
            pf[0] = lm2.x[0];
            pf[1] = lm2.x[1];
            ++stage;
            phase = GP_IDLE;
        }
    }

    PFX_HD void step(const BUF& b, const LaneWs& ws, const void* profile, int dtype) {
        if (phase == GP_SETUP) setup(b, profile, dtype);
        if (phase == GP_JAC) jac_pass(b, ws);
        if (phase == GP_TRIAL) trial_pass(b, ws);
        if (phase == GP_FIN) finish(b);
    }

    // scores 5-7 and the value parked for the Gaussian kernels; only for a completed candidate
    PFX_HD void write(double* feat, double* ph_expect) const {
        feat[4] = fabs(xmax - pp[1]);
        feat[5] = fabs(pf[1] / pp[2]);
        feat[6] = fabs(dy_expect - pp[1]);
        *ph_expect = pp[1];
    }
    PFX_HD bool complete() const { return stage == 3 && !(status & PFX_ST_FAIL_MASK) && !overflow; }
};

typedef HistLaneT<HistLaneBuf> HistLane;

}  // namespace pfx
