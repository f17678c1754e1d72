// pfx_k_sinusoid.cu -- scores 1-4 (ProfileOperations.py:190-587), one warp per candidate.
#include "pfx_internal.cuh"

using namespace pfx;

namespace {
// ---- scores 1-4 ----------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK_THREADS, 4)
k_sinusoid(const void* __restrict__ profile, int dtype, long long n, int m, double* feat,
           long long ld, int8_t* ier, u64* counter, int work_doubles) {
    extern __shared__ double smem[];
    Group<WARP> g;
    double* ws = smem + (threadIdx.x / WARP) * work_doubles;
    for (;;) {
        const long long c = next_candidate(g, counter);
        if (c >= n) break;
        sinusoid_candidate(g, ws, profile, dtype, c, m, feat + c * ld, ier ? ier + c * PFX_NIER : nullptr);
    }
}

}  // namespace

int pfx::launch_sinusoid(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* feat,
                           int64_t ld, int8_t* ier, cudaStream_t st, int ck) {
    if (s->nbin < 2) return set_err(ctx, PFX_ERR_UNSUPPORTED, "nbin < 2");
    const int wd = sin_work_doubles(s->nbin);
    const size_t smem = (size_t)wd * 8 * GROUPS_PER_BLOCK;
    int grid;
    int rc = persistent_grid(ctx, k_sinusoid, BLOCK_THREADS, smem, &grid);
    if (rc) return rc;
    k_sinusoid<<<grid, BLOCK_THREADS, smem, st>>>(profile, s->dtype, n, s->nbin, feat, ld, ier,
                                                   take_counter(ctx, ck, st), wd);
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}

