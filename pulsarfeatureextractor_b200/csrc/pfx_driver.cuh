// pfx_driver.cuh -- per-candidate drivers: workspace layout, typed loads, status handling.
// Shared by the CUDA kernels (pfx_kernels.cu) and the host logic check (tests/hostsim).
#pragma once

#include "pfx_scores.cuh"

namespace pfx {

PFX_HD inline int elem_size(int dtype) { return dtype == PFX_U8 ? 1 : (dtype == PFX_F32 ? 4 : 8); }

// Inputs are read-only for the whole call: on the device they are read through the non-coherent
// path (__ldg), which also tells the compiler that stores to working memory cannot alias them, so
// the loads of a loop can be issued ahead of the stores of earlier iterations.
PFX_HD inline double load_elem(const void* base, int dtype, long long idx) {
#if defined(__CUDA_ARCH__)
    if (dtype == PFX_U8) return (double)__ldg((const unsigned char*)base + idx);
    if (dtype == PFX_F32) return (double)__ldg((const float*)base + idx);
    return __ldg((const double*)base + idx);
#else
    if (dtype == PFX_U8) return (double)((const unsigned char*)base)[idx];
    if (dtype == PFX_F32) return (double)((const float*)base)[idx];
    return ((const double*)base)[idx];
#endif
}

// copy row `row` (len elements) of a typed [n][len] array into float64 work memory
template <class G>
PFX_HD inline void load_row(const G& g, const void* base, int dtype, long long row, int len, double* dst) {
    const long long off = row * (long long)len;
    if (dtype == PFX_U8) {
        const unsigned char* p = (const unsigned char*)base + off;
        for (int i = g.lane(); i < len; i += G::W) dst[i] = (double)p[i];
    } else if (dtype == PFX_F32) {
        const float* p = (const float*)base + off;
        for (int i = g.lane(); i < len; i += G::W) dst[i] = (double)p[i];
    } else {
        const double* p = (const double*)base + off;
        for (int i = g.lane(); i < len; i += G::W) dst[i] = p[i];
    }
}

PFX_HD inline int align_up(int v, int a) { return (v + a - 1) / a * a; }
PFX_HD inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

constexpr int LM_SHARED_DOUBLES = (int)((sizeof(LmShared) + 7) / 8);

// ---- scores 1-4 ---------------------------------------------------------------------------
PFX_HD inline int sin_work_doubles(int m) { return LM_SHARED_DOUBLES + 5 * align_up(m, 2); }

template <class G>
PFX_HD inline void sinusoid_candidate(const G& g, double* ws, const void* profile, int dtype,
                                      long long cand, int m, double* feat, int8_t* ier) {
    const int ma = align_up(m, 2);
    SinWork w;
    w.s = (LmShared*)ws;
    double* q = ws + LM_SHARED_DOUBLES;
    w.y = q;
    w.v.fvec = q + ma;
    w.v.wa4 = q + 2 * ma;
    w.v.fjac = q + 3 * ma;
    w.v.ld = ma;
    load_row(g, profile, dtype, cand, m, w.y);
    g.sync();
    sinusoid_scores(g, w, m, feat, ier);
    g.sync();
}

// ---- scores 5-7 (histogram fits) -------------------------------------------------------------
// shared-memory layout: LmShared | y[m] | srt[pow2(m)] | hx, hy, cnt, fvec, wa4, fjac[3] (8*cap)
PFX_HD inline int hist_work_doubles(int m, int cap) {
    return LM_SHARED_DOUBLES + align_up(m, 2) + next_pow2(m) + 8 * cap;
}
// capacity (in bins) of a global scratch slot used when a histogram does not fit in shared memory
PFX_HD inline int hist_scratch_doubles(int cap) { return 8 * cap; }

// Returns status bits; *ph_expect = mu of the profile-histogram fit (input of scores 8-11).
template <class G>
PFX_HD inline unsigned hist_candidate(const G& g, double* ws, int cap_smem, double* gscratch,
                                      int cap_global, const void* profile, int dtype, long long cand,
                                      int m, double pow_m, double pow_m1, double* feat, int8_t* ier,
                                      double* ph_expect) {
    const int ma = align_up(m, 2);
    HistWork h;
    h.s = (LmShared*)ws;
    h.y = ws + LM_SHARED_DOUBLES;
    h.srt = h.y + ma;
    double* q = h.srt + next_pow2(m);
    load_row(g, profile, dtype, cand, m, h.y);
    g.sync();
    h.hx = q;
    h.hy = q + cap_smem;
    h.cnt = (int*)(q + 2 * cap_smem);
    h.v.fvec = q + 3 * cap_smem;
    h.v.wa4 = q + 4 * cap_smem;
    h.v.fjac = q + 5 * cap_smem;
    h.v.ld = cap_smem;
    h.cap = cap_smem;
    h.ph_expect = NAN;
    unsigned st = hist_scores(g, h, m, pow_m, pow_m1, dtype == PFX_U8, gscratch, cap_global, feat, ier);
    g.sync();
    *ph_expect = h.ph_expect;
    return st;
}

// ---- scores 8-11 (Gaussian and double-Gaussian fits to the profile) -----------------------
// shared-memory layout: LmShared | y | yr | t1fit | nx | ny | idx | fvec | wa4 | fjac[8]
PFX_HD inline int gauss_work_doubles(int m) { return LM_SHARED_DOUBLES + 16 * align_up(m, 2); }

template <class G>
PFX_HD inline unsigned gauss_candidate(const G& g, double* ws, const void* profile, int dtype,
                                       long long cand, int m, double ph_expect, double* feat,
                                       int8_t* ier) {
    const int ma = align_up(m, 2);
    GaussWork w;
    w.s = (LmShared*)ws;
    double* q = ws + LM_SHARED_DOUBLES;
    w.y = q;
    w.yr = q + ma;
    w.t1fit = q + 2 * ma;
    w.nx = q + 3 * ma;
    w.ny = q + 4 * ma;
    w.idx = (int*)(q + 5 * ma);
    w.v.fvec = q + 6 * ma;
    w.v.wa4 = q + 7 * ma;
    w.v.fjac = q + 8 * ma;
    w.v.ld = ma;
    load_row(g, profile, dtype, cand, m, w.y);
    g.sync();
    unsigned st = gauss_scores(g, w, m, ph_expect, feat, ier);
    g.sync();
    return st;
}

// ---- scores 12-19 -------------------------------------------------------------------------
PFX_HD inline int dm_work_doubles(int n_dm) { return LM_SHARED_DOUBLES + 9 * align_up(n_dm, 2); }

// PFD family: period, S/N and width are derived from the profile (PFDO:93-233); this first
// version takes them from params like the PHCX family (SURVEY 8f2 lists the PFD reader side as
// "next"), so params = (snr, dm, period_s, width, dm_start, dm_end, len_all).
template <class G>
PFX_HD inline unsigned dmfit_candidate(const G& g, double* ws, const void* dmcurve, int dm_dtype,
                                       long long cand, int n_dm, int family, const double* prm,
                                       double* feat, int8_t* ier) {
    const int na = align_up(n_dm, 2);
    DmWork w;
    w.s = (LmShared*)ws;
    double* q = ws + LM_SHARED_DOUBLES;
    w.x = q;
    w.y = q + na;
    w.theo = q + 2 * na;
    w.v.fvec = q + 3 * na;
    w.v.wa4 = q + 4 * na;
    w.v.fjac = q + 5 * na;
    w.v.ld = na;
    const double snr = prm[0], dm = prm[1], period_ms = prm[2] * 1000.0, width = prm[3];
    if (g.leader()) {
        feat[11] = period_ms;            // score 12, PHXO:109
        feat[12] = filter_13_14(snr);    // score 13
        feat[13] = filter_13_14(dm);     // score 14
        feat[14] = width;                // score 15
    }
    if (family == PFX_FAMILY_PFD && dm_dtype == PFX_F32) {
        // yData = 255./max(yData)*yData in float32 arithmetic (PFDO:332 on the float32 curve of PFD:393)
        const float* p = (const float*)dmcurve + cand * (long long)n_dm;
        float mx = -INFINITY;
        for (int i = g.lane(); i < n_dm; i += G::W)
            if (p[i] == p[i]) mx = fmaxf(mx, p[i]);
        mx = (float)g.max((double)mx);
        if (p[0] != p[0]) mx = p[0];
        const float sc = 255.0f / mx;
        for (int i = g.lane(); i < n_dm; i += G::W) w.y[i] = (double)(sc * p[i]);
    } else {
        load_row(g, dmcurve, dm_dtype, cand, n_dm, w.y);
        if (family == PFX_FAMILY_PFD) {
            g.sync();
            const double mx = rows_builtin_max(g, w.y, n_dm);
            g.sync();
            const double sc = 255.0 / mx;
            for (int i = g.lane(); i < n_dm; i += G::W) w.y[i] = sc * w.y[i];
        }
    }
    g.sync();
    unsigned st = dm_scores(g, w, n_dm, family, prm, snr, dm, period_ms, width, feat, ier);
    g.sync();
    return st;
}

// ---- scores 20-22 -------------------------------------------------------------------------
PFX_HD inline int sub_work_doubles(int nbin, int nsub) {
    return nsub * nbin + align_up(nbin, 2) + 2 * align_up(nsub, 2);
}

struct SubAtTyped {
    const void* base;   // first element of this candidate's [nsub][nbin] block
    int dtype;
    int nbin;
    PFX_HD double operator()(int i, int j) const { return load_elem(base, dtype, (long long)i * nbin + j); }
};

template <class G>
PFX_HD inline unsigned subband_candidate(const G& g, double* ws, const void* profile,
                                         const void* subbands, int dtype, long long cand, int nbin,
                                         int nsub, double width, double* feat) {
    SubWork w;
    w.sums = ws;
    w.prof = ws + nsub * nbin;
    w.maxbin = w.prof + align_up(nbin, 2);
    w.rnorm = w.maxbin + align_up(nsub, 2);
    w.sub = nullptr;
    load_row(g, profile, dtype, cand, nbin, w.prof);
    g.sync();
    SubAtTyped at;
    at.base = (const char*)subbands + (size_t)cand * nsub * nbin * elem_size(dtype);
    at.dtype = dtype;
    at.nbin = nbin;
    unsigned st = subband_scores(g, w, at, nbin, nsub, width, feat);
    g.sync();
    return st;
}

// ---- Lyon statistics ----------------------------------------------------------------------
// generic (any dtype, any length) two-pass version on one row
template <class G>
PFX_HD inline void lyon4_row(const G& g, const void* base, int dtype, long long row, int len, double* out) {
    const long long off = row * (long long)len;
    double s = 0.0;
    for (int i = g.lane(); i < len; i += G::W) s += load_elem(base, dtype, off + i);
    const double mean = g.sum(s) / (double)len;
    double s2 = 0.0, s3 = 0.0, s4 = 0.0;
    for (int i = g.lane(); i < len; i += G::W) {
        double d = load_elem(base, dtype, off + i) - mean;
        double d2 = d * d;
        s2 += d2;
        s3 = PFX_FMA(d2, d, s3);
        s4 = PFX_FMA(d2, d2, s4);
    }
    s2 = g.sum(s2);
    s3 = g.sum(s3);
    s4 = g.sum(s4);
    if (g.leader()) lyon_finish((double)len, mean, s2, s3, s4, out);
}

}  // namespace pfx
