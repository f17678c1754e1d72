// pfx_k_dmfit_exact.cu -- scores 12-19 with the DM-curve fit reproducing scipy BIT FOR BIT
// (PHCXOperations.py:121-259, PFDOperations.py:274-393).
//
// The DM model is the one least-squares model of the path that needs only correctly rounded
// operations (+ - * / sqrt; no exp, no sin), so its residuals can be made bit-identical to numpy's.
// What is left is the order of every sum inside lmdif: this translation unit is compiled with
// -fmad=false -DPFX_NO_FMA (no contraction anywhere, like the Fortran MINPACK of the scipy wheels) and
// instantiates the stored-Jacobian form of pfx_lm.cuh on an OrderedGroup, whose inner products run in
// row order on the leader lane while the element-wise work (model evaluation, forward differences,
// Householder updates) stays striped over the 32 lanes.  Checked on the host build of the same
// templates (tests/test_hostsim_logic.py): scores 16-19 and ier equal to the oracle's scipy run on
// every PHCX-shaped candidate and on 99 % of the PFD-shaped ones, where the lane form with its
// streaming factorisation reaches 73 % / 23 % (the fits are ill conditioned: PFD curves clip most
// trial DMs, and what scipy returns is decided by rounding residues).
// Cost: an inner product is a dependent chain of n_dm additions on one lane (about 13 of them per
// LM iteration); a warp per candidate and eleven warps per SM hide most of it.
#include "pfx_internal.cuh"

using namespace pfx;

namespace {

__global__ void __launch_bounds__(WARP)
k_dmfit_exact(const void* __restrict__ dmcurve, int dm_dtype, long long n, int n_dm, int family,
              const double* __restrict__ params, double* feat, long long ld, int8_t* ier,
              int32_t* status, u64* counter) {
    extern __shared__ double smem[];
    OrderedGroup<WARP> g;
    for (;;) {
        const long long c = next_candidate(g, counter);
        if (c >= n) break;
        unsigned st = dmfit_candidate(g, smem, dmcurve, dm_dtype, c, n_dm, family, params + c * PFX_NPARAM,
                                      feat + c * ld, ier ? ier + c * PFX_NIER : nullptr);
        if (st && g.leader()) atomicOr(status + c, (int)st);
    }
}

}  // namespace

int pfx::launch_dmfit_exact(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* dmcurve, const double* params,
                            double* feat, int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st,
                            pfx_scratch_set* set, int ck) {
    const size_t smem = (size_t)align_up(dm_work_doubles(s->n_dm), 2) * 8;
    if (smem > (size_t)ctx->smem_optin)
        return set_err(ctx, PFX_ERR_UNSUPPORTED, "DM curve of %d trials needs %zu bytes of shared memory, limit %d",
                       s->n_dm, smem, ctx->smem_optin);
    int grid;
    int rc = persistent_grid(ctx, k_dmfit_exact, WARP, smem, &grid);
    if (rc) return rc;
    if (grid > n) grid = (int)(n > 0 ? n : 1);
    k_dmfit_exact<<<grid, WARP, smem, st>>>(dmcurve, s->dm_dtype, n, s->n_dm, s->family, params, feat, ld, ier, status,
                                             take_counter(set, ck, st));
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}
