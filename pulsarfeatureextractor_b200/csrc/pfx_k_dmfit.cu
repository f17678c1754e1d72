// pfx_k_dmfit.cu -- scores 12-19 (PHCXOperations.py:81-233, PFDOperations.py:274-393,
// CandidateFileInterface.py:84-112): pass-through parameters and the DM-curve fit, one lane per
// candidate group (state machine of pfx_lane_misc.cuh).  N = 3 parameters for the PHCX family, 4
// for PFD.  The fit is one LM solve over a few hundred rows whose cost varies by two orders of
// magnitude between candidates (most converge in a handful of iterations, some run to maxfev), so
// a whole warp shares each candidate (DM_LANES = 32): the rows are dealt to the lanes and the
// kernel no longer ends in a long tail of single lanes finishing slow fits.
// the n x n trust-region solve runs from the rolled shared-memory routines here (smaller code;
// measured slightly faster than the register-resident SmallLM for this kernel)
#define PFX_SMALL_LM 0
#include "pfx_internal.cuh"
#include "pfx_lane_misc.cuh"

using namespace pfx;

namespace {

#ifndef PFX_DM_LANES
#define PFX_DM_LANES 32
#endif
constexpr int DM_LANES = PFX_DM_LANES;

template <int N>
__global__ void __launch_bounds__(128, 3)
k_dmfit(const void* __restrict__ dmcurve, int dm_dtype, long long n, int n_dm, int family,
        const double* __restrict__ params, double* feat, long long ld, int8_t* ier, int32_t* status,
        u64* counter) {
    extern __shared__ double smem[];
    LaneWs ws;
    ws.carve(smem + threadIdx.x, blockDim.x, N);
    typedef LaneGroup<DM_LANES> G;
    DmLane<N, DM_LANES> g;
    g.phase = GP_IDLE;
    g.cand = -1;
    bool more = true;
    for (;;) {
        if (g.phase == GP_IDLE) {
            if (g.cand >= 0) {
                if (G::rank() == 0) {
                    g.write(feat + g.cand * ld);
                    if (ier) ier[g.cand * PFX_NIER + PFX_FIT_DM] = g.ier;
                    if (g.status) atomicOr(status + g.cand, (int)g.status);
                }
                g.cand = -1;
            }
            if (more) {
                long long c = 0;
                if (G::rank() == 0) c = (long long)atomicAdd(counter, 1ULL);
                c = G::bcast(c);
                if (c < n) g.start(dmcurve, dm_dtype, c, n_dm, family, params + c * PFX_NPARAM);
                else more = false;
            }
        }
        if (!__any_sync(0xffffffffu, g.phase != GP_IDLE || more)) break;
        g.step(ws);
    }
}

template <int N>
int launch_n(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* dmcurve, const double* params,
             double* feat, int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, pfx_scratch_set* set, int ck) {
    // one warp per block: a block leaves when its 32 lanes are done, so the tail of slow candidates
    // holds little shared memory and the Gaussian kernels' blocks fit next to it
    const int threads = 32;
    const size_t smem = (size_t)lane_ws_doubles(N) * 8 * threads;
    int grid;
    int rc = persistent_grid(ctx, k_dmfit<N>, threads, smem, &grid);
    if (rc) return rc;
    const long long groups = (long long)grid * threads / DM_LANES;
    if (groups > n) grid = (int)((n * DM_LANES + threads - 1) / threads);
    k_dmfit<N><<<grid, threads, smem, st>>>(dmcurve, s->dm_dtype, n, s->n_dm, s->family, params, feat, ld, ier,
                                             status, take_counter(set, ck, st));
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}

__global__ void k_pfd_params(const void* __restrict__ profile, int dtype, long long n, int nbin, double* params) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double snr, width;
    pfd_params(profile, dtype, c, nbin, snr, width);
    params[c * PFX_NPARAM + 0] = snr;
    params[c * PFX_NPARAM + 3] = width;
}

}  // namespace

int pfx::launch_pfd_params(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* params,
                           cudaStream_t st) {
    if (s->nbin < 2) return set_err(ctx, PFX_ERR_UNSUPPORTED, "nbin < 2");
    k_pfd_params<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(profile, s->dtype, n, s->nbin, params);
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}

int pfx::launch_dmfit(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* dmcurve, const double* params,
                      double* feat, int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, pfx_scratch_set* set, int ck) {
    if (s->n_dm < 4) return set_err(ctx, PFX_ERR_UNSUPPORTED, "n_dm < 4");
    // default: the bit-exact form (pfx_k_dmfit_exact.cu); pfx_set_dm_exact(ctx, 0) selects the lane form
    if (ctx->dm_exact) return launch_dmfit_exact(ctx, s, n, dmcurve, params, feat, ld, ier, status, st, set, ck);
    if (s->family == PFX_FAMILY_PFD)
        return launch_n<4>(ctx, s, n, dmcurve, params, feat, ld, ier, status, st, set, ck);
    return launch_n<3>(ctx, s, n, dmcurve, params, feat, ld, ier, status, st, set, ck);
}
