// pfx_scores.cuh -- the scoring routines, one per reference operator method, written once for
// any lane-group width (see pfx_group.cuh).  Reference citations are file:line under
// PulsarFeatureExtractor/src/ (PO = ProfileOperations.py, POI = ProfileOperationsInterface.py,
// PHXO = PHCXOperations.py, PFDO = PFDOperations.py, CFI = CandidateFileInterface.py).
#pragma once

#include "../../include/pfx.h"
#include "pfx_lm.cuh"

namespace pfx {

constexpr double TWO_PI = 6.283185307179586;            // 2 * numpy.pi
constexpr double FWHM_K = 2.3548200450309493;           // 2 * sqrt(2 * log(2))

// =============================================================================================
// residual models
// =============================================================================================

// PO:405-419: y - (|A| sin(2 pi f x + phi) + |bg|), x = bin index
struct SineModel {
    const double* y;
    double aamp, abg;
    template <class G>
    PFX_HD void eval(const G& g, const double (&p)[2], double* out, int m) const {
        const double w = PFX_MUL(TWO_PI, p[0]);
        for (int i = g.lane(); i < m; i += G::W) {
            double t = PFX_ADD(PFX_MUL(w, (double)i), p[1]);
            out[i] = PFX_SUB(y[i], PFX_ADD(PFX_MUL(aamp, sin(t)), abg));
        }
    }
    PFX_HD double fit(const double (&p)[2], int i) const {
        double t = PFX_ADD(PFX_MUL(PFX_MUL(TWO_PI, p[0]), (double)i), p[1]);
        return PFX_ADD(PFX_MUL(aamp, sin(t)), abg);
    }
};

// PO:509-523: y - |A| sin^2(2 pi f x + phi) + |bg|   (the +|bg| sits OUTSIDE the bracket);
// the evaluated curve PO:526-529 is |A| sin^2(...) + |bg|.
struct SineSqModel {
    const double* y;
    double aamp, abg;
    template <class G>
    PFX_HD void eval(const G& g, const double (&p)[2], double* out, int m) const {
        const double w = PFX_MUL(TWO_PI, p[0]);
        for (int i = g.lane(); i < m; i += G::W) {
            double s = sin(PFX_ADD(PFX_MUL(w, (double)i), p[1]));
            out[i] = PFX_ADD(PFX_SUB(y[i], PFX_MUL(aamp, PFX_MUL(s, s))), abg);
        }
    }
    PFX_HD double fit(const double (&p)[2], int i) const {
        double s = sin(PFX_ADD(PFX_MUL(PFX_MUL(TWO_PI, p[0]), (double)i), p[1]));
        return PFX_ADD(PFX_MUL(aamp, PFX_MUL(s, s)), abg);
    }
};

PFX_HD inline double gauss_core(double x, double mu, double sigma) {
    double t = PFX_DIV(PFX_SUB(x, mu), sigma);
    return exp(PFX_MUL(-PFX_MUL(t, t), 0.5));
}

// PO:800-816: y - |max| exp(-((x-mu)/sigma)^2 / 2), params (sigma, mu, max)
struct Gauss3Model {
    const double* x;
    const double* y;
    PFX_HD double fit(const double (&p)[3], int i) const {
        return PFX_MUL(fabs(p[2]), gauss_core(x[i], p[1], p[0]));
    }
    template <class G>
    PFX_HD void eval(const G& g, const double (&p)[3], double* out, int m) const {
        for (int i = g.lane(); i < m; i += G::W) out[i] = PFX_SUB(y[i], fit(p, i));
    }
};

// PO:1016-1030: expectation fixed to xmax, params (sigma, max)
struct Gauss2FixedModel {
    const double* x;
    const double* y;
    double xmax;
    PFX_HD double fit(const double (&p)[2], int i) const {
        return PFX_MUL(fabs(p[1]), gauss_core(x[i], xmax, p[0]));
    }
    template <class G>
    PFX_HD void eval(const G& g, const double (&p)[2], double* out, int m) const {
        for (int i = g.lane(); i < m; i += G::W) out[i] = PFX_SUB(y[i], fit(p, i));
    }
};

// PO:1219-1233: |max| exp(-((x-mu)/|sigma|)^2/2) + bg (signed), x = bin index
struct Gauss4BgModel {
    const double* y;
    PFX_HD double fit(const double (&p)[4], int i) const {
        return PFX_ADD(PFX_MUL(fabs(p[2]), gauss_core((double)i, p[1], fabs(p[0]))), p[3]);
    }
    template <class G>
    PFX_HD void eval(const G& g, const double (&p)[4], double* out, int m) const {
        for (int i = g.lane(); i < m; i += G::W) out[i] = PFX_SUB(y[i], fit(p, i));
    }
};

// PO:1288-1303: |max| exp(-((x-mu)/sigma)^2/2) + |bg|, x from a list (bins that survived)
struct Gauss4AbsModel {
    const double* x;
    const double* y;
    PFX_HD double at(const double (&p)[4], double xv) const {
        return PFX_ADD(PFX_MUL(fabs(p[2]), gauss_core(xv, p[1], p[0])), fabs(p[3]));
    }
    // curve on the full bin grid x = 0..m-1 (PO:1391, 1413-1414)
    PFX_HD double fit(const double (&p)[4], int i) const { return at(p, (double)i); }
    template <class G>
    PFX_HD void eval(const G& g, const double (&p)[4], double* out, int m) const {
        for (int i = g.lane(); i < m; i += G::W) out[i] = PFX_SUB(y[i], at(p, x[i]));
    }
};

// PO:1452-1468
struct Gauss8Model {
    const double* y;
    PFX_HD double fit(const double (&p)[8], int i) const {
        double a = PFX_MUL(fabs(p[2]), gauss_core((double)i, p[1], fabs(p[0])));
        double b = PFX_MUL(fabs(p[6]), gauss_core((double)i, p[5], fabs(p[4])));
        double bg = PFX_MUL(PFX_ADD(fabs(p[3]), fabs(p[7])), 0.5);
        return PFX_ADD(PFX_ADD(a, b), bg);
    }
    template <class G>
    PFX_HD void eval(const G& g, const double (&p)[8], double* out, int m) const {
        for (int i = g.lane(); i < m; i += G::W) out[i] = PFX_SUB(y[i], fit(p, i));
    }
};

// PHXO:150-162: Amp sqrt((P - weff)/weff), weff = sqrt(wint + (Prop kdm |DM+Shift-x| df / f^3)^2)
struct DmPhcxModel {
    const double* x;
    const double* y;
    double wint, period, dm, kdm, df, f3;
    PFX_HD double curve(double amp, double prop, double shift, double xv) const {
        double a = fabs(PFX_SUB(PFX_ADD(dm, shift), xv));
        double q = PFX_DIV(PFX_MUL(PFX_MUL(PFX_MUL(prop, kdm), a), df), f3);
        double weff = sqrt(PFX_ADD(wint, PFX_MUL(q, q)));
        return PFX_MUL(amp, sqrt(PFX_DIV(PFX_SUB(period, weff), weff)));
    }
    PFX_HD double fit(const double (&p)[3], int i) const { return curve(p[0], p[1], p[2], x[i]); }
    template <class G>
    PFX_HD void eval(const G& g, const double (&p)[3], double* out, int m) const {
        for (int i = g.lane(); i < m; i += G::W) out[i] = PFX_SUB(y[i], curve(p[0], p[1], p[2], x[i]));
    }
};

// PFDO:302-320: weff clipped to the period, plus an offset
struct DmPfdModel {
    const double* x;
    const double* y;
    double wint, period, dm, kdm, df, f3;
    PFX_HD double curve(double amp, double prop, double shift, double up, double xv) const {
        double a = fabs(PFX_SUB(PFX_ADD(dm, shift), xv));
        double q = PFX_DIV(PFX_MUL(PFX_MUL(PFX_MUL(prop, kdm), a), df), f3);
        double weff = sqrt(PFX_ADD(wint, PFX_MUL(q, q)));
        if (weff > period) weff = period;
        return PFX_ADD(up, PFX_MUL(amp, sqrt(PFX_DIV(PFX_SUB(period, weff), weff))));
    }
    template <class G>
    PFX_HD void eval(const G& g, const double (&p)[4], double* out, int m) const {
        for (int i = g.lane(); i < m; i += G::W)
            out[i] = PFX_SUB(y[i], curve(p[0], p[1], p[2], p[3], x[i]));
    }
};

// Evaluate the fitted curve on every row, out of line (one copy per model; see lm_eval).
template <int NP, class G, class Model>
PFX_HD PFX_NOINLINE void fit_rows(const G& g, const Model& f, const double (&p)[NP], double* out, int m) {
    for (int i = g.lane(); i < m; i += G::W) out[i] = f.fit(p, i);
}

// sum over own rows of (y - fit)^2, joined over the group
template <class G>
PFX_HD inline double chisq_rows(const G& g, const double* y, const double* fit, int m) {
    double c = 0.0;
    for (int i = g.lane(); i < m; i += G::W) {
        double d = y[i] - fit[i];
        c = PFX_FMA(d, d, c);
    }
    return g.sum(c);
}

// =============================================================================================
// scores 1-4                                                                   PO:190-587
// =============================================================================================
struct SinWork {
    double* y;       // [m] profile as float64
    LmVec v;         // fvec, wa4, fjac[2]
    LmShared* s;
};

// Block scan of PO:294-326 over max(0, p - mean - std); returns the number of peaks.
PFX_HD inline int count_peaks(const double* y, int m, double mean, double sd) {
    // A block collects bins until the fifth zero after its start; the block holds a peak when
    // any of its values is non-zero (values are >= 0, so max(block) != 0).  The zero that
    // triggers the flush is dropped and a block of only zeros is ignored.
    int peaks = 0, zeros = 0;
    bool nonzero = false;
    for (int i = 0; i < m; ++i) {
        double v = (y[i] - mean) - sd;
        if (v < 0.0) v = 0.0;
        if (v != 0.0) {
            nonzero = true;
        } else if (zeros < 4) {
            ++zeros;
        } else {
            if (nonzero) ++peaks;
            nonzero = false;
            zeros = 0;
        }
    }
    if (nonzero) ++peaks;
    return peaks;
}

template <class G>
PFX_HD inline void sinusoid_scores(const G& g, const SinWork& w, int m, double* feat, int8_t* ier) {
    const double* y = w.y;
    double mean, sd;
    rows_mean_std(g, y, m, mean, sd);
    double vmax = -INFINITY, vmin = INFINITY;
    for (int i = g.lane(); i < m; i += G::W) {
        vmax = fmax(vmax, y[i]);
        vmin = fmin(vmin, y[i]);
    }
    vmax = g.max(vmax);
    vmin = g.min(vmin);
    const double half = fabs(vmax - vmin) / 2.0;
    // score 4, PO:246-247
    double s4 = 0.0;
    for (int i = g.lane(); i < m; i += G::W) s4 += half - y[i];
    s4 = g.sum(s4);
    const int maxima = count_peaks(y, m, mean, sd);
    const double ndiff = (double)(maxima > 0 ? maxima - 1 : 0);

    // score 1, PO:380-491
    double chi1, chi2;
    {
        SineModel f;
        f.y = y;
        f.aamp = half;
        f.abg = half;
        double freq = (double)maxima / ((double)m - 1.0);
        double phi0;
        const double y0 = y[0];
        if (y0 == half) phi0 = 0.0;
        else if (y0 < half) phi0 = (freq != 0.0) ? -1.0 / (4.0 * freq) : -1.0 / (4.0 * 0.00000000001);
        else phi0 = (freq != 0.0) ? 1.0 / (4.0 * freq) : 1.0 / (4.0 * 0.00000000001);
        double p[2] = {freq, phi0};
        int nfev;
        int info = lmdif<2>(g, f, m, p, w.v, *w.s, nfev);
        if (ier && g.leader()) ier[PFX_FIT_SINE] = (int8_t)info;
        fit_rows<2>(g, f, p, w.v.wa4, m);
        chi1 = chisq_rows(g, y, w.v.wa4, m) / (double)m;
    }
    // score 2, PO:495-587
    {
        SineSqModel f;
        f.y = y;
        f.aamp = half;
        f.abg = half;
        double freq = (double)maxima / ((double)m - 1.0) / 2.0;
        double phi0;
        if (y[0] == 0.0) phi0 = 0.0;
        else phi0 = (freq != 0.0) ? -1.0 / (4.0 * freq) : -1.0 / (4.0 * 0.00000000001);
        double p[2] = {freq, phi0};
        int nfev;
        int info = lmdif<2>(g, f, m, p, w.v, *w.s, nfev);
        if (ier && g.leader()) ier[PFX_FIT_SINESQ] = (int8_t)info;
        fit_rows<2>(g, f, p, w.v.wa4, m);
        chi2 = chisq_rows(g, y, w.v.wa4, m) / (double)m;
    }
    if (g.leader()) {
        feat[0] = chi1 / (double)maxima;   // maxima == 0 -> inf / nan, as numpy does (PO:373-374)
        feat[1] = chi2 / (double)maxima;
        feat[2] = ndiff;
        feat[3] = s4;
    }
}

// =============================================================================================
// scores 5-7: histogram fits                                             PO:654-722, 774-1057
// =============================================================================================
struct HistWork {
    double* y;        // [m] profile
    double* srt;      // [pow2 >= m] sort buffer
    double* hx;       // [cap] left bin edges
    double* hy;       // [cap] counts
    int* cnt;         // [cap] integer counts
    LmVec v;          // fvec[cap], wa4[cap], fjac[3][cap]
    LmShared* s;
    int cap;          // histogram capacity of this workspace
    // outputs kept for the later stages
    double ph_expect; // mu of the profile-histogram fit
};

// in-place bitonic sort of a[0..n2), n2 a power of two
template <class G>
PFX_HD inline void bitonic_sort(const G& g, double* a, int n2) {
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            g.sync();
            for (int i = g.lane(); i < n2; i += G::W) {
                int ixj = i ^ j;
                if (ixj > i) {
                    double x = a[i], z = a[ixj];
                    bool up = ((i & k) == 0);
                    if ((x > z) == up) {
                        a[i] = z;
                        a[ixj] = x;
                    }
                }
            }
        }
    }
    g.sync();
}

// scipy.stats.scoreatpercentile on sorted data (_stats_py.py:2034-2074)
PFX_HD inline double percentile_sorted(const double* srt, int n, double per) {
    double idx = per / 100.0 * (double)(n - 1);
    int i = (int)idx;
    if ((double)i == idx) return srt[i] * 1.0 / 1.0;
    int j = i + 1;
    double w0 = (double)j - idx, w1 = idx - (double)i;
    double sumval = w0 + w1;
    return (srt[i] * w0 + srt[i + 1] * w1) / sumval;
}

// POI:138-166 on data d[0..n); returns nbins (may be 0 or huge).  pow_n = pow(n, -0.3333333)
// evaluated on the host with the C library (what Python's pow does).
template <class G>
PFX_HD PFX_NOINLINE long long fd_bins(const G& g, const double* d, int n, double* srt, double pow_n,
                               bool int_data, double& dmin, double& dmax) {
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int i = g.lane(); i < n2; i += G::W) srt[i] = (i < n) ? d[i] : INFINITY;
    bitonic_sort(g, srt, n2);
    double iqr = percentile_sorted(srt, n, 75.0) - percentile_sorted(srt, n, 25.0);
    dmin = srt[0];
    dmax = srt[n - 1];
    g.sync();
    double binwidth = 2.0 * iqr * pow_n;
    double rnge = dmax - dmin;
    double q;
    if (binwidth <= 0.0) {
        // Python 2: integer range / integer 60 floors (POI:153-158)
        q = int_data ? floor(rnge / 60.0) : rnge / 60.0;
    } else {
        q = rnge / binwidth;
    }
    double c = ceil(q);
    if (!(c < 9.0e15)) return (long long)9000000000000000LL;
    return (long long)c;
}

// numpy.histogram(d, nb) uniform-bin path (numpy/lib/_histograms_impl.py:814-873); fills
// hx[0..nb) with the left edges and hy[0..nb) with the counts.
template <class G>
PFX_HD PFX_NOINLINE void np_histogram(const G& g, const double* d, int n, int nb, double first, double last,
                                double* hx, double* hy, int* cnt) {
    if (first == last) {
        first -= 0.5;
        last += 0.5;
    }
    const double step = (last - first) / (double)nb;
    for (int i = g.lane(); i < nb; i += G::W) {
        hx[i] = PFX_ADD(PFX_MUL((double)i, step), first);
        cnt[i] = 0;
    }
    g.sync();
    const double denom = last - first;
    for (int r = g.lane(); r < n; r += G::W) {
        double a = d[r];
        if (!(a >= first && a <= last)) continue;
        double fi = PFX_MUL(PFX_DIV(PFX_SUB(a, first), denom), (double)nb);
        int idx = (int)fi;
        if (idx == nb) idx -= 1;
        double e_lo = hx[idx];
        if (a < e_lo) idx -= 1;
        else {
            double e_hi = (idx + 1 == nb) ? last : hx[idx + 1];
            if (a >= e_hi && idx != nb - 1) idx += 1;
        }
#if defined(__CUDA_ARCH__)
        atomicAdd(&cnt[idx], 1);
#else
        cnt[idx] += 1;
#endif
    }
    g.sync();
    for (int i = g.lane(); i < nb; i += G::W) hy[i] = (double)cnt[i];
    g.sync();
}

// PO:774-983.  On entry hx/hy hold nb bins; they are zero-padded to 3 points when shorter.
// Returns status bits; p = (sigma, mu, max) of the last fit done.
template <class G>
PFX_HD PFX_NOINLINE unsigned fit_gaussian_hist(const G& g, HistWork& w, int nb, double (&p)[3],
                                        int8_t* ier, int slot, int count_slot) {
    double* hx = w.hx;
    double* hy = w.hy;
    const int n0 = nb;
    ValIdx am = rows_argmax(g, hy, nb);
    const int first_max = am.i;
    double expect = hx[first_max];
    double cmean, sigma0;
    rows_mean_std(g, hy, nb, cmean, sigma0);
    const double max0 = rows_builtin_max(g, hy, nb);
    const double meansq = cmean * cmean;
    int mfit = nb;
    if (nb < 3) {
        g.sync();
        if (g.leader())
            for (int i = nb; i < 3; ++i) {
                hx[i] = 0.0;
                hy[i] = 0.0;
            }
        mfit = 3;
        g.sync();
    }
    Gauss3Model f;
    f.x = hx;
    f.y = hy;
    int counter = 0, nfits = 0;
    unsigned st = 0;
    for (;;) {
        p[0] = sigma0;
        p[1] = expect;
        p[2] = max0;
        int nfev;
        int info = lmdif<3>(g, f, mfit, p, w.v, *w.s, nfev);
        ++nfits;
        if (ier && g.leader()) ier[slot] = (int8_t)info;
        fit_rows<3>(g, f, p, w.v.wa4, n0);
        const double chisq = chisq_rows(g, hy, w.v.wa4, n0) / (double)mfit;
        if ((chisq > meansq * (double)n0) && (p[0] < 0.2 * (double)n0)) {
            ++counter;
            // temp = delete(temp, first_max): counts with [first_max, first_max+counter) removed
            const int len_before = nb - (counter - 1);
            if (first_max >= len_before) {
                st |= PFX_ST_INDEX_OOB;
                break;
            }
            const int len_after = len_before - 1;
            if (len_after == 0) {
                st |= PFX_ST_EMPTY_SEQUENCE;
                break;
            }
            ValIdx best;
            best.v = -INFINITY;
            best.i = 0x7fffffff;
            bool have = false;
            for (int i = g.lane(); i < nb; i += G::W) {
                if (i >= first_max && i < first_max + counter) continue;
                int ti = (i < first_max) ? i : i - counter;
                if (!have || argmax_better(hy[i], ti, best.v, best.i)) {
                    best.v = hy[i];
                    best.i = ti;
                    have = true;
                }
            }
            best = g.argmax(best);
            const int pos = best.i;
            if (pos + counter >= mfit) {
                st |= PFX_ST_INDEX_OOB;
                break;
            }
            expect = hx[pos + counter];
            if (counter > 5) break;
        } else {
            break;
        }
    }
    if (ier && g.leader()) ier[count_slot] = (int8_t)nfits;
    return st;
}

// scores 5, 6, 7 and the profile-histogram expectation needed by scores 8-11.
// `w` arrives carved for shared memory (capacity w.cap); when a histogram has more bins than
// that, the histogram and fit vectors are re-pointed at the group's global scratch slot
// (gscratch, capacity cap_global) -- candidates are flagged, never truncated.
template <class G>
PFX_HD inline unsigned hist_scores(const G& g, HistWork& w, int m, double pow_m, double pow_m1,
                                   bool int_data, double* gscratch, int cap_global, double* feat,
                                   int8_t* ier) {
    const double* y = w.y;
    unsigned st = 0;
    double dmin, dmax;
    // profile bins (POI:138-166 on the profile)
    long long nbp = fd_bins(g, y, m, w.srt, pow_m, int_data, dmin, dmax);
    const double pmin = dmin, pmax = dmax;
    // derivative dy[i] = y[i] - y[i+1] (POI:170-186), staged in the shared-memory Jacobian area
    double* dy = w.v.fjac;
    for (int i = g.lane(); i < m - 1; i += G::W) dy[i] = y[i] - y[i + 1];
    g.sync();
    long long nbd = fd_bins(g, dy, m - 1, w.srt, pow_m1, int_data, dmin, dmax);
    if (nbd < 1 || nbp < 1) return st | PFX_ST_HIST_ZERO_BINS;
    const long long nbmax = nbd > nbp ? nbd : nbp;
    if (nbmax > w.cap) {
        if (gscratch == nullptr || nbmax > cap_global) return st | PFX_ST_HIST_OVERFLOW;
        w.hx = gscratch;
        w.hy = gscratch + cap_global;
        w.cnt = (int*)(gscratch + 2 * cap_global);
        w.v.fvec = gscratch + 3 * cap_global;
        w.v.wa4 = gscratch + 4 * cap_global;
        w.v.fjac = gscratch + 5 * cap_global;
        w.v.ld = cap_global;
        w.cap = cap_global;
    }

    // derivative histogram + fit (PO:657-661); the histogram is finished before the fit
    // overwrites the staging area
    np_histogram(g, dy, m - 1, (int)nbd, dmin, dmax, w.hx, w.hy, w.cnt);
    double pd[3];
    st |= fit_gaussian_hist(g, w, (int)nbd, pd, ier, PFX_FIT_DYHIST, PFX_FIT_DYHIST_COUNT);
    if (st & PFX_ST_FAIL_MASK) return st;
    const double dy_expect = pd[1];

    // profile histogram + fit (PO:678-682)
    np_histogram(g, y, m, (int)nbp, pmin, pmax, w.hx, w.hy, w.cnt);
    double pp[3];
    st |= fit_gaussian_hist(g, w, (int)nbp, pp, ier, PFX_FIT_PROFHIST, PFX_FIT_PROFHIST_COUNT);
    if (st & PFX_ST_FAIL_MASK) return st;
    // the retry loop may have padded hx/hy (nb < 3); the fixed fit uses the unpadded histogram
    const int nb = (int)nbp;

    // fixed-expectation fit (PO:988-1057); xmax = x[bins//2 - 1], negative index wraps
    if (nb < 2) return st | PFX_ST_LSQ_IMPROPER;
    int xi = nb / 2 - 1;
    if (xi < 0) xi += nb;
    Gauss2FixedModel f;
    f.x = w.hx;
    f.y = w.hy;
    f.xmax = w.hx[xi];
    double cmean, sigma0;
    rows_mean_std(g, w.hy, nb, cmean, sigma0);
    double pf[2] = {sigma0, rows_builtin_max(g, w.hy, nb)};
    int nfev;
    int info = lmdif<2>(g, f, nb, pf, w.v, *w.s, nfev);
    if (ier && g.leader()) ier[PFX_FIT_FIXED] = (int8_t)info;

    w.ph_expect = pp[1];
    if (g.leader()) {
        feat[4] = fabs(f.xmax - pp[1]);       // score 5, PO:715
        feat[5] = fabs(pf[1] / pp[2]);        // score 6, PO:716
        feat[6] = fabs(dy_expect - pp[1]);    // score 7, PO:717
    }
    return st;
}

// =============================================================================================
// scores 8-11: Gaussian and double-Gaussian fits to the profile     PO:724-768, 1061-1483
// =============================================================================================
struct GaussWork {
    double* y;       // [m] profile
    double* yr;      // [m] rotated (background-subtracted) profile for T1, then rotated raw profile
    double* t1fit;   // [m]
    double* nx;      // [m] newx
    double* ny;      // [m] newy
    int* idx;        // [m] index list for the peak-stripping loop
    LmVec v;         // fvec[m], wa4[m], fjac[8][m]
    LmShared* s;
};

// Peak stripping of PO:1305-1354 on the rotated profile y[0..n): fills idx with the surviving
// bin numbers and returns their count, or -1 when the reference would raise IndexError.
PFX_HD inline int dg_strip_peak(const double* y, int n, int pos, int* idx) {
    int len = n;
    for (int i = 0; i < n; ++i) idx[i] = i;
#define PFX_DEL(k_)                                        \
    do {                                                   \
        int k = (k_);                                      \
        if (k < -len || k >= len) return -1;               \
        if (k < 0) k += len;                               \
        for (int q = k; q < len - 1; ++q) idx[q] = idx[q + 1]; \
        --len;                                             \
    } while (0)
#define PFX_Y(k_) y[((k_) < 0) ? (k_) + n : (k_)]
    PFX_DEL(pos);
    int tol = 0;
    const int limit = 5;
    for (int i = 1; i < n; ++i) {
        if ((pos - i) > 0 && (pos + i) < n) {
            if ((PFX_Y(pos - i) < PFX_Y(pos - i + 1)) && (y[pos + i] < y[pos + i - 1])) {
                PFX_DEL(pos - i);
                PFX_DEL(pos - i);
            } else if ((PFX_Y(pos - i) >= PFX_Y(pos - i + 1)) ||
                       ((y[pos + i] >= y[pos + i - 1]) && (tol < limit))) {
                PFX_DEL(pos - i);
                PFX_DEL(pos - i);
                ++tol;
            } else {
                break;
            }
        } else if ((pos - i) < 0) {
            if (pos + i >= n) return -1;                     // y[pos+i] IndexError
            if (y[pos + i] < y[pos + i - 1]) {
                PFX_DEL(pos - i + 1);
            } else if ((y[pos + i] >= y[pos + i - 1]) && (tol < limit)) {
                PFX_DEL(pos - i + 1);
                ++tol;
            } else {
                break;
            }
        } else if ((pos + i) > n) {
            if (PFX_Y(pos - i) < PFX_Y(pos - i + 1)) {
                PFX_DEL(pos - i + 1);
            } else if ((PFX_Y(pos - i) >= PFX_Y(pos - i + 1)) && (tol < limit)) {
                PFX_DEL(pos - i);
                ++tol;
            } else {
                break;
            }
        }
    }
#undef PFX_DEL
#undef PFX_Y
    return len;
}

template <class G>
PFX_HD inline unsigned gauss_scores(const G& g, GaussWork& w, int m, double ph_expect,
                                    double* feat, int8_t* ier) {
    const double* y = w.y;
    unsigned st = 0;
    double mean, sd;
    rows_mean_std(g, y, m, mean, sd);
    // minbg = min(ph_expect, mean) with Python's builtin min (PO:724)
    const double minbg = (mean < ph_expect) ? mean : ph_expect;
    const int cut = m / 2;

    // ---- T1: Gaussian + background on the background-subtracted, rotated profile ------------
    for (int i = g.lane(); i < m; i += G::W) {
        int src = i + cut;
        if (src >= m) src -= m;
        double v = y[src];
        if (minbg > 0.0) {
            v = PFX_ADD(PFX_SUB(v, minbg), sd);
            if (v < 0.0) v = 0.0;
        }
        w.yr[i] = v;
    }
    g.sync();
    double t1_fwhm, t1_chi;
    {
        Gauss4BgModel f;
        f.y = w.yr;
        ValIdx am = rows_argmax(g, w.yr, m);
        double rm, rs;
        rows_mean_std(g, w.yr, m, rm, rs);
        double p[4] = {rs, (double)am.i, am.v, 1.0};
        int nfev;
        int info = lmdif<4>(g, f, m, p, w.v, *w.s, nfev);
        if (ier && g.leader()) ier[PFX_FIT_T1] = (int8_t)info;
        t1_fwhm = fabs(FWHM_K * p[0]);
        fit_rows<4>(g, f, p, w.t1fit, m);
        t1_chi = chisq_rows(g, w.yr, w.t1fit, m) / (double)m;
    }
    g.sync();
    if (g.leader()) {
        feat[7] = t1_fwhm;   // score 8
        feat[8] = t1_chi;    // score 9
    }

    // ---- T2: double Gaussian on the rotated raw profile ---------------------------------------
    for (int i = g.lane(); i < m; i += G::W) {
        int src = i + cut;
        if (src >= m) src -= m;
        w.yr[i] = y[src];
    }
    g.sync();
    const double* yr = w.yr;
    ValIdx am = rows_argmax(g, yr, m);
    // peak stripping: sequential, done by the leader on the shared index list
    int nlen;
    {
        int len = 0;
        if (g.leader()) {
            len = dg_strip_peak(yr, m, am.i, w.idx);
            w.s->pad[0] = len;
        }
        g.sync();
        nlen = w.s->pad[0];
        g.sync();
    }
    if (nlen < 0) {
        if (g.leader()) {
            feat[9] = 1000000.0;    // PO:762-764
            feat[10] = 1000000.0;
        }
        return st | PFX_ST_DG_SENTINEL;
    }
    if (nlen == 0) return st | PFX_ST_EMPTY_SEQUENCE;
    for (int i = g.lane(); i < nlen; i += G::W) {
        int b = w.idx[i];
        w.nx[i] = (double)b;
        w.ny[i] = yr[b];
    }
    g.sync();
    int nm = nlen;
    double p1[4], p2[4];
    Gauss4AbsModel f4;
    f4.x = w.nx;
    f4.y = w.ny;
    for (int rnd = 0; rnd < 8; ++rnd) {
        ValIdx nmx = rows_argmax(g, w.ny, nm);
        double nmean, nsd;
        rows_mean_std(g, w.ny, nm, nmean, nsd);
        double p[4] = {nsd, w.nx[nmx.i], rows_builtin_max(g, w.ny, nm), nmean};
        if (nm < 4) {
            g.sync();
            if (g.leader())
                for (int i = nm; i < 4; ++i) {
                    w.nx[i] = 0.0;
                    w.ny[i] = 0.0;
                }
            nm = 4;
            g.sync();
        }
        int nfev;
        int info = lmdif<4>(g, f4, nm, p, w.v, *w.s, nfev);
        if (ier && g.leader()) ier[PFX_FIT_DG_ROUND0 + rnd] = (int8_t)info;
        const double nfwhm = fabs(FWHM_K * p[0]);
        const double lo = p[2] - (1.5 * nfwhm) / 2.0;
        const double hi = p[2] + (1.5 * nfwhm) / 2.0;
        fit_rows<4>(g, f4, p, w.v.wa4, m);
        for (int i = g.lane(); i < m; i += G::W) {
            double xv = (double)i;
            double ev = w.v.wa4[i];
            double yv = yr[i];
            double nv;
            if (ev <= yv) nv = PFX_ADD(PFX_SUB(yv, ev), p[3]);
            else if ((ev > yv) && (xv > lo) && (xv < hi)) nv = p[3];
            else nv = yv;
            w.ny[i] = nv;
            w.nx[i] = xv;
        }
        nm = m;
        g.sync();
        if (rnd == 6) {
#pragma unroll
            for (int k = 0; k < 4; ++k) p2[k] = p[k];
        } else if (rnd == 7) {
#pragma unroll
            for (int k = 0; k < 4; ++k) p1[k] = p[k];
        }
    }
    // final 8-parameter fit (PO:1432-1483)
    Gauss8Model f8;
    f8.y = yr;
    double q[8] = {p1[0], p1[1], p1[2], p1[3], p2[0], p2[1], p2[2], p2[3]};
    int nfev;
    int info = lmdif<8>(g, f8, m, q, w.v, *w.s, nfev);
    if (ier && g.leader()) ier[PFX_FIT_DG_FINAL] = (int8_t)info;
    const double fwhm1 = fabs(FWHM_K * q[0]);
    const double fwhm2 = fabs(FWHM_K * q[4]);
    double c = 0.0, cc = 0.0;
    const double bgmix = PFX_MUL(PFX_ADD(p1[3], p2[3]), 0.5);
    fit_rows<8>(g, f8, q, w.ny, m);          // final fit
    fit_rows<4>(g, f4, p1, w.v.wa4, m);      // fit1
    fit_rows<4>(g, f4, p2, w.v.fvec, m);     // fit2
    for (int i = g.lane(); i < m; i += G::W) {
        double fi = w.ny[i];
        if (fi >= 1.0) {
            double d = yr[i] - fi;
            c += (d * d) / (double)m;
        }
        // combifit = fit1 + fit2 - bg1 - bg2 + (bg1 + bg2)/2   (PO:1413-1415)
        double cf = PFX_ADD(PFX_SUB(PFX_SUB(PFX_ADD(w.v.wa4[i], w.v.fvec[i]), p1[3]), p2[3]), bgmix);
        w.nx[i] = cf;
        if (cf >= 1.0) {
            double d = yr[i] - cf;
            cc += (d * d) / (double)m;
        }
    }
    const double chisq = g.sum(c);
    const double combi_chisq = g.sum(cc);
    double fw_a, fw_b, dg_chi;
    const double* dgfit;
    if (chisq <= combi_chisq) {
        fw_a = fwhm1;
        fw_b = fwhm2;
        dg_chi = chisq;
        dgfit = w.ny;
    } else {
        fw_a = fabs(FWHM_K * p2[0]);   // combi_fwhm2 sits in slot [1], combi_fwhm1 in slot [6]
        fw_b = fabs(FWHM_K * p1[0]);
        dg_chi = combi_chisq;
        dgfit = w.nx;
    }
    // std of (double-Gaussian fit) - (T1 fit + minbg - std)   (PO:755-756)
    double ds = 0.0;
    for (int i = g.lane(); i < m; i += G::W) {
        double d = dgfit[i] - PFX_SUB(PFX_ADD(w.t1fit[i], minbg), sd);
        ds += d;
    }
    const double dmean = g.sum(ds) / (double)m;
    double dv = 0.0;
    for (int i = g.lane(); i < m; i += G::W) {
        double d = (dgfit[i] - PFX_SUB(PFX_ADD(w.t1fit[i], minbg), sd)) - dmean;
        dv = PFX_FMA(d, d, dv);
    }
    const double dstd = fabs(sqrt(g.sum(dv) / (double)m));
    double s10;
    if (dstd < 3.0) s10 = t1_fwhm;
    else s10 = (fw_b < fw_a) ? fw_b : fw_a;   // builtin min(fw_a, fw_b)
    if (g.leader()) {
        feat[9] = s10;
        feat[10] = dg_chi;
    }
    g.sync();
    return st;
}

// =============================================================================================
// scores 12-19                                              PHXO:81-233, PFDO:274-393, CFI:84-112
// =============================================================================================
struct DmWork {
    double* x;      // [n_dm]
    double* y;      // [n_dm]
    double* theo;   // [n_dm]
    LmVec v;        // fvec, wa4, fjac[4]
    LmShared* s;
};

PFX_HD inline double filter_13_14(double v) {
    // CFI:97-108 with isEqual(value, 0, 5e-6) == -1
    if (fabs(v - 0.0) > 0.000005 && v < 0.0) return 0.0;
    return v;
}

// y must already hold the DM curve as float64 (PFD: scaled by 255/max, PFDO:332)
template <class G>
PFX_HD inline unsigned dm_scores(const G& g, DmWork& w, int n_dm, int family, const double* prm,
                                 double snr, double dm, double period_ms, double width,
                                 double* feat, int8_t* ier) {
    unsigned st = 0;
    const double dm_start = prm[4], dm_end = prm[5];
    double length_all = prm[6];
    const bool pfd = (family == PFX_FAMILY_PFD);
    if (!(length_all > 0.0)) length_all = pfd ? (double)n_dm : 128.0 * (double)n_dm;
    const double dm_step = fabs(dm_start - dm_end) / length_all;
    const double wint = (width * period_ms) * (width * period_ms);
    const double kdm = 8300000.000000001;   // 8.3*10**6 as Python evaluates it (PHXO:187, PFDO:342): one ulp above 8.3e6
    const double df = pfd ? 32.0 : 400.0;
    const double f3 = pfd ? 2460375.0 : 2593941624.0;   // 135^3, 1374^3
    const double sw = sqrt(wint);
    const double peak = snr / sqrt((period_ms - sw) / sw);   // score 16, PHXO:191
    for (int i = g.lane(); i < n_dm; i += G::W) {
        double xi = pfd ? (double)i : (double)(128 * i - 1);   // PHXO:251-257 x index 128k-1
        double xv = PFX_ADD(dm_start, PFX_MUL(xi, dm_step));
        w.x[i] = xv;
        double q = PFX_DIV(PFX_MUL(PFX_MUL(kdm, fabs(PFX_SUB(dm, xv))), df), f3);
        double weff = sqrt(PFX_ADD(wint, PFX_MUL(q, q)));
        if (pfd && weff > period_ms) weff = period_ms;
        w.theo[i] = sqrt(PFX_DIV(PFX_SUB(period_ms, weff), weff));
    }
    g.sync();
    const double hmax = rows_builtin_max(g, w.theo, n_dm);
    const double amp0 = 255.0 / hmax;
    g.sync();
    for (int i = g.lane(); i < n_dm; i += G::W) w.theo[i] = PFX_MUL(amp0, w.theo[i]);
    g.sync();
    double prop, shift, chi_theo;
    if (!pfd) {
        DmPhcxModel f;
        f.x = w.x; f.y = w.y; f.wint = wint; f.period = period_ms; f.dm = dm; f.kdm = kdm; f.df = df; f.f3 = f3;
        double p[3] = {amp0, 1.0, 0.0};
        int nfev;
        int info = lmdif<3>(g, f, n_dm, p, w.v, *w.s, nfev);
        if (ier && g.leader()) ier[PFX_FIT_DM] = (int8_t)info;
        prop = p[1];
        shift = p[2];
        double c = 0.0;
        fit_rows<3>(g, f, p, w.v.wa4, n_dm);
        if constexpr (G::ordered) {
            // (a masked term adds nothing, like the skipped iteration of the reference's loop)
            c = ordered_rows(g, 0, n_dm, [&](int i) {
                const double d = PFX_SUB(w.y[i], w.theo[i]);
                return (w.v.wa4[i] >= 1.0) ? PFX_MUL(d, d) : 0.0;
            });
        } else {
            for (int i = g.lane(); i < n_dm; i += G::W) {
                double fit = w.v.wa4[i];
                if (fit >= 1.0) {
                    double d = w.y[i] - w.theo[i];
                    c = PFX_FMA(d, d, c);
                }
            }
            c = g.sum(c);
        }
        chi_theo = c / (double)n_dm;    // PHXO:222-230: theoretical curve, masked by the fit
    } else {
        DmPfdModel f;
        f.x = w.x; f.y = w.y; f.wint = wint; f.period = period_ms; f.dm = dm; f.kdm = kdm; f.df = df; f.f3 = f3;
        double p[4] = {amp0, 1.0, 0.0, 0.0};
        int nfev;
        int info = lmdif<4>(g, f, n_dm, p, w.v, *w.s, nfev);
        if (ier && g.leader()) ier[PFX_FIT_DM] = (int8_t)info;
        prop = p[1];
        shift = p[2];
        double c = 0.0;
        int ndeg = 0;
        for (int i = g.lane(); i < n_dm; i += G::W)
            if (w.theo[i] > 0.0) ++ndeg;
        ndeg = g.sum(ndeg);
        if (ndeg == 0) return st | PFX_ST_DM_NDEG_ZERO;
        if constexpr (G::ordered) {
            c = ordered_rows(g, 0, n_dm, [&](int i) {
                const double th = w.theo[i];
                const double d = PFX_SUB(w.y[i], th);
                return (th > 0.0) ? PFX_DIV(PFX_MUL(d, d), th) : 0.0;
            });
        } else {
            for (int i = g.lane(); i < n_dm; i += G::W) {
                double th = w.theo[i];
                if (th > 0.0) {
                    double d = w.y[i] - th;
                    c += (d * d) / th;
                }
            }
            c = g.sum(c);
        }
        chi_theo = c / (double)ndeg;    // PFDO:378-387
    }
    if (g.leader()) {
        feat[15] = peak;
        feat[16] = fabs(1.0 - prop);
        feat[17] = fabs(shift);                // filterScore(18, .), CFI:110
        feat[18] = chi_theo;
    }
    g.sync();
    return st;
}

// =============================================================================================
// scores 20-22                                    PO:1585-1686, PHXO:387-415, PFDO:445-466
// =============================================================================================
struct SubWork {
    double* sums;    // [nsub][L] box-car sums, later centred in place
    double* prof;    // [nbin] profile, later centred in place
    double* sub;     // [nsub][nbin] sub-bands as float64 (may be NULL when read from global)
    double* maxbin;  // [nsub]
    double* rnorm;   // [nsub] sqrt(c_ii)
};

// sub(i, j) accessor is passed as a functor so that uint8 / float inputs need no staging copy
template <class G, class SubAt>
PFX_HD inline unsigned subband_scores(const G& g, SubWork& w, const SubAt& sub, int nbin, int nsub,
                                      double width, double* feat) {
    // width_bins = int(ceil(bestWidth * prof_bins))  (PO:1603)
    const double wb = ceil(width * (double)nbin);
    if (!(wb <= (double)nbin)) return PFX_ST_SUBBAND_BAD_WIDTH;   // also catches NaN
    if (!(wb >= 1.0)) return PFX_ST_SUBBAND_NO_PAIRS;              // w <= 0: every pair is NaN
    const int wbin = (int)wb;
    const int L = nbin - wbin + 1;
    // box-car sums, accumulated left to right as the reference does (PO:1608-1618)
    for (int t = g.lane(); t < nsub * L; t += G::W) {
        int i = t / L, j = t - i * L;
        double s = 0.0;
        for (int b = 0; b < wbin; ++b) s += sub(i, j + b);
        w.sums[t] = s;
    }
    g.sync();
    // first strict maximum per band (PO:1622-1631), max_bin = j + w//2
    for (int i = g.lane(); i < nsub; i += G::W) {
        double best = -10000.0;
        int bj = -1;
        const double* row = w.sums + i * L;
        for (int j = 0; j < L; ++j)
            if (row[j] > best) {
                best = row[j];
                bj = j;
            }
        w.maxbin[i] = (bj < 0) ? NAN : (double)(bj + wbin / 2);
    }
    g.sync();
    // a band whose sums never exceed -10000 keeps the previous band's max_bin (stale variable);
    // for the first band the reference raises UnboundLocalError.
    unsigned st = 0;
    {
        double prev = NAN;
        bool bad = false;
        if (g.leader()) {
            for (int i = 0; i < nsub; ++i) {
                double v = w.maxbin[i];
                if (v != v) {
                    if (i == 0) bad = true;
                    v = prev;
                    w.maxbin[i] = v;
                }
                prev = v;
            }
            w.rnorm[0] = bad ? 1.0 : 0.0;
        }
        g.sync();
        bad = w.rnorm[0] != 0.0;
        g.sync();
        if (bad) return PFX_ST_SUBBAND_BAD_WIDTH;
    }
    // med, count, variance: sequential over the (few) bands, identical in every lane
    double rms;
    {
        double med = rows_sum(g, w.maxbin, 0, nsub) / (double)nsub;
        int count = 0;
        double var_med = 0.0;
        for (int i = 0; i < nsub; ++i) {
            double d = w.maxbin[i] - med;
            if (fabs(d) <= (double)wbin) {
                ++count;
                var_med += d * d;
            }
        }
        double var;
        if (count > 1) {
            var = var_med / (double)(count - 1);
        } else {
            double mean = 0.0;
            for (int i = 0; i < nsub; ++i) mean += w.maxbin[i];
            mean /= (double)nsub;
            var = 0.0;
            for (int i = 0; i < nsub; ++i) {
                double d = w.maxbin[i] - mean;
                var += d * d;
            }
            if (nsub - 1 == 0) return PFX_ST_SUBBAND_NO_PAIRS;
            var /= (double)(nsub - 1);
        }
        rms = sqrt(var) / (double)wbin;
    }
    // ---- pair correlations of the box-car sums: numpy.corrcoef semantics ----------------------
    // centre every row, c_ii = dot/(L-1), r = c_ik / sqrt(c_ii) / sqrt(c_kk), clipped to [-1, 1]
    const double inv = 1.0 / (double)(L - 1);    // L == 1 -> inf -> NaN rows, as numpy
    for (int i = 0; i < nsub; ++i) {
        double* row = w.sums + i * L;
        double mu = rows_sum(g, row, 0, L) / (double)L;
        double ss = 0.0;
        for (int j = g.lane(); j < L; j += G::W) {
            double d = row[j] - mu;
            row[j] = d;
            ss = PFX_FMA(d, d, ss);
        }
        ss = g.sum(ss);
        if (g.leader()) w.rnorm[i] = sqrt(ss * inv);
    }
    g.sync();
    const int npairs = nsub * (nsub - 1) / 2;
    double csum = 0.0;
    int cm = 0;
    // pairs are distributed over lanes; each lane walks whole rows (shared-memory reads of
    // different rows by different lanes: L is odd-ish so bank conflicts stay low)
    for (int t = g.lane(); t < npairs; t += G::W) {
        // decode (i, k), i < k, from the row-major pair index
        int i = 0, rem = t;
        while (rem >= nsub - 1 - i) {
            rem -= nsub - 1 - i;
            ++i;
        }
        int k = i + 1 + rem;
        const double* a = w.sums + i * L;
        const double* b = w.sums + k * L;
        double d = 0.0;
        for (int j = 0; j < L; ++j) d = PFX_FMA(a[j], b[j], d);
        double r = ((d * inv) / w.rnorm[i]) / w.rnorm[k];
        if (r == r) {
            r = fmin(1.0, fmax(-1.0, r));
            csum += r;
            ++cm;
        }
    }
    csum = g.sum(csum);
    cm = g.sum(cm);
    if (cm == 0) return PFX_ST_SUBBAND_NO_PAIRS;    // ZeroDivisionError, PO:1681
    const double mean_corr = csum / (double)cm;

    // ---- score 22: |corrcoef(sub-band, profile)| above 0.0055 ----------------------------------
    double pmu = rows_sum(g, w.prof, 0, nbin) / (double)nbin;
    double pss = 0.0;
    g.sync();
    for (int j = g.lane(); j < nbin; j += G::W) {
        double d = w.prof[j] - pmu;
        w.prof[j] = d;
        pss = PFX_FMA(d, d, pss);
    }
    pss = g.sum(pss);
    g.sync();
    const double inv2 = 1.0 / (double)(nbin - 1);
    const double pnorm = sqrt(pss * inv2);
    double s22 = 0.0;
    for (int i = 0; i < nsub; ++i) {
        double s1 = 0.0;
        for (int j = g.lane(); j < nbin; j += G::W) s1 += sub(i, j);
        const double mu = g.sum(s1) / (double)nbin;
        double ss = 0.0, dp = 0.0;
        for (int j = g.lane(); j < nbin; j += G::W) {
            double d = sub(i, j) - mu;
            ss = PFX_FMA(d, d, ss);
            dp = PFX_FMA(d, w.prof[j], dp);
        }
        ss = g.sum(ss);
        dp = g.sum(dp);
        double r = ((dp * inv2) / sqrt(ss * inv2)) / pnorm;
        if (r == r) r = fmin(1.0, fmax(-1.0, r));
        r = fabs(r);
        if (r > 0.0055) s22 += r;     // NaN fails the test and is dropped (PHXO:411-412)
    }
    if (g.leader()) {
        feat[19] = rms;
        feat[20] = mean_corr;
        feat[21] = s22;
    }
    return st;
}

// =============================================================================================
// Lyon statistics                                                            PHX:320-379
// =============================================================================================
// mean, std, skew, kurtosis from the lane-partial central sums (scipy.stats semantics)
PFX_HD inline void lyon_finish(double n, double mean, double s2, double s3, double s4, double* out) {
    const double m2 = s2 / n, m3 = s3 / n, m4 = s4 / n;
    out[0] = mean;
    out[1] = sqrt(m2);
    const double thr = (LM_EPSMCH * mean) * (LM_EPSMCH * mean);
    if (m2 <= thr) {
        out[2] = NAN;
        out[3] = NAN;
    } else {
        out[2] = m3 / (m2 * sqrt(m2));      // m2 ** 1.5
        out[3] = m4 / (m2 * m2) - 3.0;
    }
}

}  // namespace pfx
