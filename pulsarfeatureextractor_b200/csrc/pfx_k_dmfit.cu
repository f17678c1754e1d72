// pfx_k_dmfit.cu -- scores 12-19 (PHCXOperations.py:81-233, PFDOperations.py:274-393).
#include "pfx_internal.cuh"

using namespace pfx;

namespace {
// ---- scores 12-19 --------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK_THREADS, 4)
k_dmfit(const void* __restrict__ dmcurve, int dm_dtype, long long n, int n_dm, int family,
        const double* __restrict__ params, double* feat, long long ld, int8_t* ier,
        int32_t* status, u64* counter, int work_doubles) {
    extern __shared__ double smem[];
    Group<WARP> g;
    double* ws = smem + (threadIdx.x / WARP) * work_doubles;
    for (;;) {
        const long long c = next_candidate(g, counter);
        if (c >= n) break;
        unsigned st = dmfit_candidate(g, ws, dmcurve, dm_dtype, c, n_dm, family, params + c * PFX_NPARAM,
                                      feat + c * ld, ier ? ier + c * PFX_NIER : nullptr);
        if (st && g.leader()) atomicOr(status + c, (int)st);
    }
}

}  // namespace

int pfx::launch_dmfit(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* dmcurve,
                        const double* params, double* feat, int64_t ld, int8_t* ier, int32_t* status,
                        cudaStream_t st, int ck) {
    if (s->n_dm < 4) return set_err(ctx, PFX_ERR_UNSUPPORTED, "n_dm < 4");
    const int wd = align_up(dm_work_doubles(s->n_dm), 2);
    const size_t smem = (size_t)wd * 8 * GROUPS_PER_BLOCK;
    int grid;
    int rc = persistent_grid(ctx, k_dmfit, BLOCK_THREADS, smem, &grid);
    if (rc) return rc;
    k_dmfit<<<grid, BLOCK_THREADS, smem, st>>>(dmcurve, s->dm_dtype, n, s->n_dm, s->family, params, feat, ld,
                                                ier, status, take_counter(ctx, ck, st), wd);
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}

