// pfx_k_hist.cu -- scores 5-7: Freedman-Diaconis histograms of the profile and of its derivative and
// the three Gaussian fits on them (ProfileOperations.py:654-722,774-1057,
// ProfileOperationsInterface.py:138-186).  One warp per candidate; the histogram lives in shared
// memory, or in the warp's global scratch slot when it has more bins than fit there.
// The expectation of the profile-histogram fit, needed by scores 8-11, is parked in the score-8
// column of the feature row; the next kernel (pfx_k_gauss.cu) consumes it.
#include "pfx_internal.cuh"

using namespace pfx;

namespace {

__global__ void __launch_bounds__(BLOCK_THREADS, 4)
k_hist(const void* __restrict__ profile, int dtype, long long n, int m, double pow_m, double pow_m1,
       double* feat, long long ld, int8_t* ier, int32_t* status, u64* counter, int work_doubles,
       int cap_smem, double* gscratch, int cap_global) {
    extern __shared__ double smem[];
    Group<WARP> g;
    const int gid = threadIdx.x / WARP;
    double* ws = smem + gid * work_doubles;
    double* gs = gscratch
                     ? gscratch + ((size_t)blockIdx.x * GROUPS_PER_BLOCK + gid) * hist_scratch_doubles(cap_global)
                     : nullptr;
    for (;;) {
        const long long c = next_candidate(g, counter);
        if (c >= n) break;
        double ph_expect;
        unsigned st = hist_candidate(g, ws, cap_smem, gs, cap_global, profile, dtype, c, m, pow_m, pow_m1,
                                     feat + c * ld, ier ? ier + c * PFX_NIER : nullptr, &ph_expect);
        if (g.leader()) {
            feat[c * ld + 7] = ph_expect;
            if (st) atomicOr(status + c, (int)st);
        }
    }
}

}  // namespace

int pfx::launch_hist(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* feat,
                     int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, int ck) {
    const int m = s->nbin;
    if (m < 8) return set_err(ctx, PFX_ERR_UNSUPPORTED, "nbin < 8");
    int cap_smem = 192;                       // typical histograms have 6..110 bins
    while (3 * cap_smem < m) cap_smem += 64;  // the derivative is staged in the Jacobian area
    const int wd = align_up(hist_work_doubles(m, cap_smem), 2);
    const size_t smem = (size_t)wd * 8 * GROUPS_PER_BLOCK;
    int grid;
    int rc = persistent_grid(ctx, k_hist, BLOCK_THREADS, smem, &grid);
    if (rc) return rc;
    const int groups = grid * GROUPS_PER_BLOCK;
    const size_t need = (size_t)groups * hist_scratch_doubles(HIST_CAP_GLOBAL) * 8;
    if (ctx->hist_scratch_bytes < need) {
        if (ctx->hist_scratch) {
            cudaDeviceSynchronize();
            cudaFree(ctx->hist_scratch);
            ctx->scratch_bytes -= ctx->hist_scratch_bytes;
            ctx->hist_scratch = nullptr;
            ctx->hist_scratch_bytes = 0;
        }
        PFX_CUDA(ctx, cudaMalloc(&ctx->hist_scratch, need));
        ctx->hist_scratch_bytes = need;
        ctx->scratch_bytes += need;
    }
    const double pow_m = pow((double)m, -0.3333333);        // POI:151, evaluated by the C library
    const double pow_m1 = pow((double)(m - 1), -0.3333333);
    k_hist<<<grid, BLOCK_THREADS, smem, st>>>(profile, s->dtype, n, m, pow_m, pow_m1, feat, ld, ier, status,
                                               take_counter(ctx, ck, st), wd, cap_smem, ctx->hist_scratch,
                                               HIST_CAP_GLOBAL);
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}
