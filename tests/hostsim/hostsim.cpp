// hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the scoring routines of pulsarfeatureextractor_b200/csrc (the same templates the CUDA
// kernels instantiate with 32-lane groups) for the host with a one-lane group, so the kernel
// LOGIC (MINPACK port, histogram, peak stripping, status bits ...) can be checked against the
// oracle on machines without a GPU.  It is built by tests/hostsim/build.sh into
// tests/hostsim/_hostsim.so and loaded only by tests/test_hostsim_logic.py.  The product
// package never loads it: pulsarfeatureextractor_b200 has no CPU path.
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../pulsarfeatureextractor_b200/csrc/pfx_driver.cuh"

using namespace pfx;
typedef Group<1> G1;

extern "C" {

int hs_sinusoid(const void* profile, int dtype, int m, double* feat, int8_t* ier) {
    std::vector<double> ws(sin_work_doubles(m) + 16);
    G1 g;
    sinusoid_candidate(g, ws.data(), profile, dtype, 0, m, feat, ier);
    return 0;
}

unsigned hs_gaussian(const void* profile, int dtype, int m, int cap_smem, int cap_global,
                     double* feat, int8_t* ier) {
    std::vector<double> ws(hist_work_doubles(m, cap_smem) + 16);
    std::vector<double> gs(hist_scratch_doubles(cap_global) + 16);
    G1 g;
    double pow_m = pow((double)m, -0.3333333), pow_m1 = pow((double)(m - 1), -0.3333333);
    double ph_expect = NAN;
    unsigned st = hist_candidate(g, ws.data(), cap_smem, cap_global > 0 ? gs.data() : nullptr, cap_global,
                                 profile, dtype, 0, m, pow_m, pow_m1, feat, ier, &ph_expect);
    if (st & PFX_ST_FAIL_MASK) return st;
    std::vector<double> ws2(gauss_work_doubles(m) + 16);
    st |= gauss_candidate(g, ws2.data(), profile, dtype, 0, m, ph_expect, feat, ier);
    return st;
}

unsigned hs_dmfit(const void* dmcurve, int dm_dtype, int n_dm, int family, const double* prm,
                  double* feat, int8_t* ier) {
    std::vector<double> ws(dm_work_doubles(n_dm) + 16);
    G1 g;
    return dmfit_candidate(g, ws.data(), dmcurve, dm_dtype, 0, n_dm, family, prm, feat, ier);
}

unsigned hs_subband(const void* profile, const void* subbands, int dtype, int nbin, int nsub,
                    double width, double* feat) {
    std::vector<double> ws(sub_work_doubles(nbin, nsub) + 16);
    G1 g;
    return subband_candidate(g, ws.data(), profile, subbands, dtype, 0, nbin, nsub, width, feat);
}

void hs_lyon4(const void* data, int dtype, int len, double* out) {
    G1 g;
    lyon4_row(g, data, dtype, 0, len, out);
}

}  // extern "C"

// ---- generic LM with a caller-supplied residual (Python callback) -------------------------
// Lets tests run the SAME numpy residual through scipy.optimize.leastsq and through this port.
typedef void (*hs_resid_cb)(const double* p, double* out);
template <int NP>
struct CallbackModel {
    hs_resid_cb cb;
    template <class G>
    void eval(const G&, const double (&p)[NP], double* out, int) const { cb(p, out); }
};
template <int NP>
static int run_cb(int m, double* x, hs_resid_cb cb, int* nfev) {
    std::vector<double> buf((NP + 2) * (size_t)m + 16);
    LmShared s;
    LmVec v;
    v.fvec = buf.data();
    v.wa4 = buf.data() + m;
    v.fjac = buf.data() + 2 * (size_t)m;
    v.ld = m;
    CallbackModel<NP> f;
    f.cb = cb;
    double p[NP];
    for (int i = 0; i < NP; ++i) p[i] = x[i];
    G1 g;
    int info = lmdif<NP>(g, f, m, p, v, s, *nfev);
    for (int i = 0; i < NP; ++i) x[i] = p[i];
    return info;
}
extern "C" int hs_lmdif_cb(int n, int m, double* x, hs_resid_cb cb, int* nfev) {
    switch (n) {
        case 2: return run_cb<2>(m, x, cb, nfev);
        case 3: return run_cb<3>(m, x, cb, nfev);
        case 4: return run_cb<4>(m, x, cb, nfev);
        case 8: return run_cb<8>(m, x, cb, nfev);
    }
    return -100;
}

// ---- lane-per-fit LM (streaming QR) with a caller-supplied residual ---------------------------
#include "../../pulsarfeatureextractor_b200/csrc/pfx_lane_lm.cuh"
template <int NP>
static int run_lane_cb(int m, double* x, hs_resid_cb cb, int* nfev) {
    std::vector<double> f0(m), fj((size_t)NP * m), f1(m);
    double x0[NP];
    for (int i = 0; i < NP; ++i) x0[i] = x[i];
    LaneLM<NP> lm;
    lm.init(x0);
    std::vector<double> wsbuf(lane_ws_doubles(NP) + 8);
    LaneWs ws;
    ws.carve(wsbuf.data(), 1, NP);
    int phase = LANE_JAC;
    while (phase != LANE_DONE) {
        if (phase == LANE_JAC) {
            lm.jac_begin();
            cb(lm.x, f0.data());
            for (int j = 0; j < NP; ++j) {
                double p[NP];
                for (int k = 0; k < NP; ++k) p[k] = lm.x[k];
                p[j] = lm.x[j] + lm.h[j];
                cb(p, fj.data() + (size_t)j * m);
            }
            for (int i = 0; i < m; ++i) {
                double a[NP];
                for (int j = 0; j < NP; ++j) a[j] = (fj[(size_t)j * m + i] - f0[i]) / lm.h[j];
                lm.jac_row(a, f0[i]);
            }
            phase = lm.jac_end(ws, m);
        } else {
            lm.trial_begin(ws);
            cb(lm.xnew, f1.data());
            double s = 0.0;
            for (int i = 0; i < m; ++i) s += f1[i] * f1[i];
            phase = lm.trial_end(sqrt(s));
        }
    }
    for (int i = 0; i < NP; ++i) x[i] = lm.x[i];
    *nfev = lm.nfev;
    return lm.info;
}
extern "C" int hs_lanelm_cb(int n, int m, double* x, hs_resid_cb cb, int* nfev) {
    switch (n) {
        case 2: return run_lane_cb<2>(m, x, cb, nfev);
        case 3: return run_lane_cb<3>(m, x, cb, nfev);
        case 4: return run_lane_cb<4>(m, x, cb, nfev);
        case 8: return run_lane_cb<8>(m, x, cb, nfev);
    }
    return -100;
}

// ---- lane-per-candidate scores 8-11 (pfx_lane_gauss.cuh) ----------------------------------------
#include "../../pulsarfeatureextractor_b200/csrc/pfx_lane_gauss.cuh"
extern "C" unsigned hs_gauss_lane(const void* profile, int dtype, int m, double ph_expect, double* feat,
                                  int8_t* ier) {
    std::vector<double> dat(m + 8);
    std::vector<unsigned short> xs(m + 8);
    LaneBuf b;
    b.dat = dat.data();
    b.xs = xs.data();
    b.stride = 1;
    std::vector<double> wsbuf(lane_ws_doubles(8) + 8);
    LaneWs ws;
    ws.carve(wsbuf.data(), 1, 4);
    GaussLane g;
    g.start(profile, dtype, 0, m, ph_expect);
    while (g.phase != GP_IDLE) g.step(b, ws);
    for (int k = 0; k < 9; ++k) ier[PFX_FIT_T1 + k] = g.ier[k];
    feat[7] = g.t1_fwhm;
    feat[8] = g.t1_chi;
    if (g.sentinel) {
        feat[9] = 1000000.0;
        feat[10] = 1000000.0;
        return g.status;
    }
    if (g.status & PFX_ST_FAIL_MASK) return g.status;
    GaussCarry cy;
    for (int k = 0; k < 4; ++k) {
        cy.p_t1[k] = g.p_t1[k];
        cy.p1[k] = g.p1[k];
        cy.p2[k] = g.p2[k];
    }
    cy.minbg = g.minbg;
    cy.sd = g.sd;
    cy.t1_fwhm = g.t1_fwhm;
    cy.skip = 0;
    Gauss8Lane f;
    ws.carve(wsbuf.data(), 1, 8);
    f.start(profile, dtype, 0, m, cy, b);
    while (f.phase != GP_IDLE) f.step(b, ws);
    ier[PFX_FIT_DG_FINAL] = f.ier;
    feat[9] = f.s10;
    feat[10] = f.s11;
    return g.status;
}

// scores 5-7 only (histogram stage); returns status, *ph_expect = mu of the profile-histogram fit
extern "C" unsigned hs_hist(const void* profile, int dtype, int m, int cap_smem, int cap_global,
                            double* feat, int8_t* ier, double* ph_expect) {
    std::vector<double> ws(hist_work_doubles(m, cap_smem) + 16);
    std::vector<double> gs(hist_scratch_doubles(cap_global) + 16);
    G1 g;
    double pow_m = pow((double)m, -0.3333333), pow_m1 = pow((double)(m - 1), -0.3333333);
    return hist_candidate(g, ws.data(), cap_smem, cap_global > 0 ? gs.data() : nullptr, cap_global,
                          profile, dtype, 0, m, pow_m, pow_m1, feat, ier, ph_expect);
}

// ---- lane-per-candidate scores 1-4 and 12-19 (pfx_lane_misc.cuh) ---------------------------------
#include "../../pulsarfeatureextractor_b200/csrc/pfx_lane_misc.cuh"
extern "C" int hs_sin_lane(const void* profile, int dtype, int m, double* feat, int8_t* ier) {
    std::vector<double> wsbuf(lane_ws_doubles(4) + 8);
    LaneWs ws;
    ws.carve(wsbuf.data(), 1, 2);
    SinLane g;
    g.start(profile, dtype, 0, m);
    while (g.phase != GP_IDLE) g.step(ws);
    g.write(feat);
    ier[PFX_FIT_SINE] = g.ier[0];
    ier[PFX_FIT_SINESQ] = g.ier[1];
    return 0;
}
extern "C" unsigned hs_dm_lane(const void* dmcurve, int dm_dtype, int n_dm, int family, const double* prm,
                               double* feat, int8_t* ier) {
    std::vector<double> wsbuf(lane_ws_doubles(4) + 8);
    LaneWs ws;
    if (family == PFX_FAMILY_PFD) {
        ws.carve(wsbuf.data(), 1, 4);
        DmLane<4> g;
        g.start(dmcurve, dm_dtype, 0, n_dm, family, prm);
        while (g.phase != GP_IDLE) g.step(ws);
        g.write(feat);
        ier[PFX_FIT_DM] = g.ier;
        return g.status;
    }
    ws.carve(wsbuf.data(), 1, 3);
    DmLane<3> g;
    g.start(dmcurve, dm_dtype, 0, n_dm, family, prm);
    while (g.phase != GP_IDLE) g.step(ws);
    g.write(feat);
    ier[PFX_FIT_DM] = g.ier;
    return g.status;
}

// ---- lane-per-candidate scores 5-7 (pfx_lane_hist.cuh) --------------------------------------------
#include "../../pulsarfeatureextractor_b200/csrc/pfx_lane_hist.cuh"
// returns status; *overflow = 1 when the candidate must go to the warp-form kernel
extern "C" unsigned hs_hist_lane(const void* profile, int dtype, int m, double* feat, int8_t* ier,
                                 double* ph_expect, int* overflow) {
    std::vector<double> buf(hist_lane_doubles(m) + 8), wsbuf(lane_ws_doubles(4) + 8);
    HistLaneBuf b;
    b.carve(buf.data(), 1, m);
    LaneWs ws;
    ws.carve(wsbuf.data(), 1, 3);
    HistLane g;
    g.start(0, m, dtype == PFX_U8, pow((double)m, -0.3333333), pow((double)(m - 1), -0.3333333));
    while (g.phase != GP_IDLE) g.step(b, ws, profile, dtype);
    *overflow = g.overflow ? 1 : 0;
    ier[PFX_FIT_DYHIST] = g.ier[0];
    ier[PFX_FIT_PROFHIST] = g.ier[1];
    ier[PFX_FIT_FIXED] = g.ier[2];
    ier[PFX_FIT_DYHIST_COUNT] = g.counts[0];
    ier[PFX_FIT_PROFHIST_COUNT] = g.counts[1];
    if (g.complete()) g.write(feat, ph_expect);
    return g.status;
}

extern "C" void hs_pfd_params(const void* profile, int dtype, int n, double* snr, double* width) {
    pfd_params(profile, dtype, 0, n, *snr, *width);
}
#if defined(PFX_COUNT_LMPAR)
extern "C" void hs_lmpar_hist(long long* out24) {
    for (int k = 0; k < 12; ++k) { out24[k] = pfx::g_lmpar_hist[0][k]; out24[12 + k] = pfx::g_lmpar_hist8[k]; }
}
#endif

// lmpar of the Gaussian rounds in resumable pieces (what k_gauss_rounds_s does with PFX_LMPAR_BUDGET): 0 = in one go
extern "C" void hs_set_lmpar_budget(int budget) { pfx::g_host_lmpar_budget = budget; }
