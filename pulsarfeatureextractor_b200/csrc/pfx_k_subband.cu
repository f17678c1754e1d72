// pfx_k_subband.cu -- scores 20-22 (ProfileOperations.py:1585-1686, PHCXOperations.py:387-415).
#include "pfx_internal.cuh"

using namespace pfx;

namespace {
// ---- scores 20-22 --------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK_THREADS, 4)
k_subband(const void* __restrict__ profile, const void* __restrict__ subbands, int dtype,
          long long n, int nbin, int nsub, const double* __restrict__ params, double* feat,
          long long ld, int32_t* status, u64* counter, int work_doubles) {
    extern __shared__ double smem[];
    Group<WARP> g;
    double* ws = smem + (threadIdx.x / WARP) * work_doubles;
    for (;;) {
        const long long c = next_candidate(g, counter);
        if (c >= n) break;
        unsigned st = subband_candidate(g, ws, profile, subbands, dtype, c, nbin, nsub,
                                        params[c * PFX_NPARAM + 3], feat + c * ld);
        if (st && g.leader()) atomicOr(status + c, (int)st);
    }
}

}  // namespace

int pfx::launch_subband(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile,
                          const void* subbands, const double* params, double* feat, int64_t ld,
                          int32_t* status, cudaStream_t st, pfx_scratch_set* set, int ck) {
    if (s->nsub < 1 || s->nbin < 2) return set_err(ctx, PFX_ERR_UNSUPPORTED, "bad sub-band shape");
    const int wd = align_up(sub_work_doubles(s->nbin, s->nsub), 2);
    // the box-car sums of one candidate (nsub x nbin float64) live in shared memory; as many warps
    // per block as fit (HTRU 16 x 128: four; survey 16 x 1024: one)
    int warps = GROUPS_PER_BLOCK;
    while (warps > 1 && (size_t)wd * 8 * warps > (size_t)ctx->smem_optin) warps >>= 1;
    const size_t smem = (size_t)wd * 8 * warps;
    if (smem > (size_t)ctx->smem_optin)
        return set_err(ctx, PFX_ERR_UNSUPPORTED, "sub-bands %d x %d need %zu bytes of shared memory per candidate, limit %d",
                       s->nsub, s->nbin, smem, ctx->smem_optin);
    const int threads = warps * WARP;
    int grid;
    int rc = persistent_grid(ctx, k_subband, threads, smem, &grid);
    if (rc) return rc;
    k_subband<<<grid, threads, smem, st>>>(profile, subbands, s->dtype, n, s->nbin, s->nsub, params, feat,
                                                  ld, status, take_counter(set, ck, st), wd);
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}

