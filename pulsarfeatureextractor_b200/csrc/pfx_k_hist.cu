// pfx_k_hist.cu -- scores 5-7: Freedman-Diaconis histograms of the profile and of its derivative and
// the three Gaussian fits on them (ProfileOperations.py:654-722,774-1057,
// ProfileOperationsInterface.py:138-186).  One warp per candidate; the histogram lives in shared
// memory, or in the warp's global scratch slot when it has more bins than fit there.
// The expectation of the profile-histogram fit, needed by scores 8-11, is parked in the score-8
// column of the feature row; the next kernel (pfx_k_gauss.cu) consumes it.
// the n x n trust-region solve runs from the rolled shared-memory routines here (smaller code;
// measured slightly faster than the register-resident SmallLM for this kernel)
#define PFX_SMALL_LM 0
#include "pfx_internal.cuh"
#include "pfx_lane_hist.cuh"

using namespace pfx;

namespace {

__global__ void __launch_bounds__(BLOCK_THREADS, 4)
k_hist(const void* __restrict__ profile, int dtype, long long n, int m, double pow_m, double pow_m1,
       double* feat, long long ld, int8_t* ier, int32_t* status, u64* counter, int work_doubles,
       int cap_smem, double* gscratch, int cap_global, const int* __restrict__ list,
       const u64* __restrict__ list_count) {
    extern __shared__ double smem[];
    Group<WARP> g;
    const int gid = threadIdx.x / WARP;
    double* ws = smem + gid * work_doubles;
    double* gs = gscratch
                     ? gscratch + ((size_t)blockIdx.x * GROUPS_PER_BLOCK + gid) * hist_scratch_doubles(cap_global)
                     : nullptr;
    // with an index list the kernel scores only those candidates (histograms that overflowed the
    // per-lane capacity of k_hist_lane)
    const long long total = list ? (long long)*list_count : n;
    for (;;) {
        const long long q = next_candidate(g, counter);
        if (q >= total) break;
        const long long c = list ? (long long)list[q] : q;
        double ph_expect;
        unsigned st = hist_candidate(g, ws, cap_smem, gs, cap_global, profile, dtype, c, m, pow_m, pow_m1,
                                     feat + c * ld, ier ? ier + c * PFX_NIER : nullptr, &ph_expect);
        if (g.leader()) {
            feat[c * ld + 7] = ph_expect;
            if (st) atomicOr(status + c, (int)st);
        }
    }
}

__global__ void __launch_bounds__(128, 3)
k_hist_lane(const void* __restrict__ profile, int dtype, long long n, int m, double pow_m, double pow_m1,
            double* feat, long long ld, int8_t* ier, int32_t* status, u64* counter, double* gbuf,
            int* ovf_list, u64* ovf_count) {
    extern __shared__ double smem[];
    LaneWs ws;
    ws.carve(smem + threadIdx.x, blockDim.x, 3);
    HistLaneBuf b;
    b.carve(gbuf + (size_t)blockIdx.x * blockDim.x + threadIdx.x, (int)(gridDim.x * blockDim.x), m);
    HistLane g;
    g.phase = GP_IDLE;
    g.cand = -1;
    bool more = true;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (;;) {
        if (g.phase == GP_IDLE) {
            if (g.cand >= 0) {
                double* row = feat + g.cand * ld;
                if (g.overflow) {
                    ovf_list[atomicAdd(ovf_count, 1ULL)] = (int)g.cand;
                } else {
                    double ph = qnan;
                    if (g.complete()) g.write(row, &ph);
                    row[7] = ph;                   // parked for the Gaussian kernels
                    if (ier) {
                        int8_t* ir = ier + g.cand * PFX_NIER;
                        ir[PFX_FIT_DYHIST] = g.ier[0];
                        ir[PFX_FIT_PROFHIST] = g.ier[1];
                        ir[PFX_FIT_FIXED] = g.ier[2];
                        ir[PFX_FIT_DYHIST_COUNT] = g.counts[0];
                        ir[PFX_FIT_PROFHIST_COUNT] = g.counts[1];
                    }
                    if (g.status) atomicOr(status + g.cand, (int)g.status);
                }
                g.cand = -1;
            }
            if (more) {
                const long long c = (long long)atomicAdd(counter, 1ULL);
                if (c < n) g.start(c, m, dtype == PFX_U8, pow_m, pow_m1);
                else more = false;
            }
        }
        if (!__any_sync(0xffffffffu, g.phase != GP_IDLE || more)) break;
        g.step(b, ws, profile, dtype);
    }
}

}  // namespace

namespace {
struct HistPlan {
    int grid, lgrid, wd, cap_smem;
    size_t smem, lsmem;
};
}  // namespace

static int hist_plan(pfx_ctx* ctx, pfx_scratch_set* set, int m, int64_t n, HistPlan* p) {
    p->cap_smem = 192;                              // typical histograms have 6..110 bins
    while (3 * p->cap_smem < m) p->cap_smem += 64;  // the derivative is staged in the Jacobian area
    p->wd = align_up(hist_work_doubles(m, p->cap_smem), 2);
    p->smem = (size_t)p->wd * 8 * GROUPS_PER_BLOCK;
    int rc = persistent_grid(ctx, k_hist, BLOCK_THREADS, p->smem, &p->grid);
    if (rc) return rc;
    // global histogram slots of the warp-form kernel: one per warp of its grid, per scratch set
    rc = ensure_buf(ctx, &set->hist_warp, (size_t)p->grid * GROUPS_PER_BLOCK * hist_scratch_doubles(HIST_CAP_GLOBAL) * 8);
    if (rc) return rc;
    const int threads = 128;
    p->lsmem = (size_t)lane_ws_doubles(3) * 8 * threads;
    rc = persistent_grid(ctx, k_hist_lane, threads, p->lsmem, &p->lgrid);
    if (rc) return rc;
#if defined(PFX_HIST_BLOCKS_PER_SM)
    if (p->lgrid > PFX_HIST_BLOCKS_PER_SM * ctx->sm_count) p->lgrid = PFX_HIST_BLOCKS_PER_SM * ctx->sm_count;
#endif
    if ((long long)p->lgrid * threads > n) p->lgrid = (int)((n + threads - 1) / threads);
    rc = ensure_buf(ctx, &set->hist_lane, (size_t)p->lgrid * threads * (size_t)hist_lane_doubles(m) * 8);
    if (rc) return rc;
    return ensure_buf(ctx, &set->ovf, (size_t)n * sizeof(int));
}

int pfx::reserve_hist(pfx_ctx* ctx, pfx_scratch_set* set, const pfx_shape* s, int64_t n) {
    HistPlan p;
    return hist_plan(ctx, set, s->nbin, n, &p);
}

int pfx::launch_hist(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* feat,
                     int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, pfx_scratch_set* set, int ck) {
    const int m = s->nbin;
    if (m < 8) return set_err(ctx, PFX_ERR_UNSUPPORTED, "nbin < 8");
    HistPlan p;
    int rc = hist_plan(ctx, set, m, n, &p);
    if (rc) return rc;
    const double pow_m = pow((double)m, -0.3333333);        // POI:151, evaluated by the C library
    const double pow_m1 = pow((double)(m - 1), -0.3333333);
    // pass 1: one lane per candidate, histograms up to HIST_LANE_CAP bins
    u64* ovf_count = take_counter(set, 7, st);
    k_hist_lane<<<p.lgrid, 128, p.lsmem, st>>>(profile, s->dtype, n, m, pow_m, pow_m1, feat, ld, ier, status,
                                                take_counter(set, ck, st), (double*)set->hist_lane.p,
                                                (int*)set->ovf.p, ovf_count);
    PFX_CUDA(ctx, cudaGetLastError());
    // pass 2: the few candidates whose histogram has more bins go through the warp-form kernel
    // (4096-bin scratch slot per warp); a few per cent of a synthetic batch end up here (1-2 k
    // candidates of 100 k): one warp each, so the whole persistent grid is launched; blocks without
    // work leave at once
    k_hist<<<p.grid, BLOCK_THREADS, p.smem, st>>>(profile, s->dtype, n, m, pow_m, pow_m1, feat, ld, ier, status,
                                                   take_counter(set, 6, st), p.wd, p.cap_smem,
                                                   (double*)set->hist_warp.p, HIST_CAP_GLOBAL, (int*)set->ovf.p, ovf_count);
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}
