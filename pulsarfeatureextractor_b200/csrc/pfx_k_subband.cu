// pfx_k_subband.cu -- scores 20-22 (ProfileOperations.py:1585-1686, PHCXOperations.py:387-415).
#include "pfx_internal.cuh"

using namespace pfx;

namespace {
// ---- scores 20-22 --------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK_THREADS, 4)
k_subband(const void* __restrict__ profile, const void* __restrict__ subbands, int dtype,
          long long n, int nbin, int nsub, const double* __restrict__ params, double* feat,
          long long ld, int32_t* status, u64* counter, int work_doubles, double* gws) {
    extern __shared__ double smem[];
    Group<WARP> g;
    // the box-car sums of the candidate a warp works on: shared memory, or (long profiles) the warp's slot of a
    // global scratch that stays in L2
    double* ws = gws ? gws + ((size_t)blockIdx.x * (blockDim.x / WARP) + threadIdx.x / WARP) * work_doubles
                     : smem + (threadIdx.x / WARP) * work_doubles;
    for (;;) {
        const long long c = next_candidate(g, counter);
        if (c >= n) break;
        unsigned st = subband_candidate(g, ws, profile, subbands, dtype, c, nbin, nsub,
                                        params[c * PFX_NPARAM + 3], feat + c * ld);
        if (st && g.leader()) atomicOr(status + c, (int)st);
    }
}

}  // namespace

namespace {
struct SubPlan {
    int wd, threads, grid;
    size_t smem;
    bool global;       // box-car sums in the per-warp global scratch
};
}  // namespace

// The box-car sums of one candidate (nsub x nbin float64) live in shared memory when the four warps of a block
// fit (HTRU 16 x 128: 17 KB per candidate, PFD 64 x 64: 33 KB); longer profiles (survey 16 x 1024: 131 KB, which
// left one warp per SM: 866 ms per 50 000 candidates) keep them in a per-warp global scratch instead, two blocks
// per SM, L2 resident.
static int subband_plan(pfx_ctx* ctx, pfx_scratch_set* set, const pfx_shape* s, SubPlan* p) {
    p->wd = align_up(sub_work_doubles(s->nbin, s->nsub), 2);
    p->threads = GROUPS_PER_BLOCK * WARP;
    p->smem = (size_t)p->wd * 8 * GROUPS_PER_BLOCK;
    const bool huge = p->smem > (size_t)ctx->smem_optin;
    // mid-size shapes (PFD 64 x 64: 133 KB per block of four warps, one block per SM) are faster from the global
    // scratch too, with four blocks per SM: 66 -> 30 ms per 100 000 candidates (PFX_SUB_GLOBAL=0: shared memory)
    p->global = huge || (ctx->sub_global > 0 && p->smem > (size_t)ctx->smem_optin / 3);
    if (p->global) p->smem = 0;
    int rc = persistent_grid(ctx, k_subband, p->threads, p->smem, &p->grid);
    if (rc) return rc;
    if (p->global) {
        const int cap = ctx->sub_blocks > 0 ? ctx->sub_blocks : (huge ? 2 : 4);
        if (p->grid > cap * ctx->sm_count) p->grid = cap * ctx->sm_count;
        rc = ensure_buf(ctx, &set->sub_scratch, (size_t)p->grid * GROUPS_PER_BLOCK * p->wd * 8);
    }
    return rc;
}

int pfx::reserve_subband(pfx_ctx* ctx, pfx_scratch_set* set, const pfx_shape* s) {
    if (s->nsub < 1 || s->nbin < 2) return PFX_OK;
    SubPlan p;
    return subband_plan(ctx, set, s, &p);
}

int pfx::launch_subband(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile,
                          const void* subbands, const double* params, double* feat, int64_t ld,
                          int32_t* status, cudaStream_t st, pfx_scratch_set* set, int ck) {
    if (s->nsub < 1 || s->nbin < 2) return set_err(ctx, PFX_ERR_UNSUPPORTED, "bad sub-band shape");
    SubPlan p;
    int rc = subband_plan(ctx, set, s, &p);
    if (rc) return rc;
    k_subband<<<p.grid, p.threads, p.smem, st>>>(profile, subbands, s->dtype, n, s->nbin, s->nsub, params, feat, ld, status,
                                                 take_counter(set, ck, st), p.wd, p.global ? (double*)set->sub_scratch.p : nullptr);
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}
