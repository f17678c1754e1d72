// pfx_k_dmfit.cu -- scores 12-19 (PHCXOperations.py:81-233, PFDOperations.py:274-393,
// CandidateFileInterface.py:84-112): pass-through parameters and the DM-curve fit, one lane per
// candidate (state machine of pfx_lane_misc.cuh).  N = 3 parameters for the PHCX family, 4 for PFD.
#include "pfx_internal.cuh"
#include "pfx_lane_misc.cuh"

using namespace pfx;

namespace {

template <int N>
__global__ void __launch_bounds__(128, 3)
k_dmfit(const void* __restrict__ dmcurve, int dm_dtype, long long n, int n_dm, int family,
        const double* __restrict__ params, double* feat, long long ld, int8_t* ier, int32_t* status,
        u64* counter) {
    extern __shared__ double smem[];
    LaneWs ws;
    ws.carve(smem + threadIdx.x, blockDim.x, N);
    DmLane<N> g;
    g.phase = GP_IDLE;
    g.cand = -1;
    bool more = true;
    for (;;) {
        if (g.phase == GP_IDLE) {
            if (g.cand >= 0) {
                g.write(feat + g.cand * ld);
                if (ier) ier[g.cand * PFX_NIER + PFX_FIT_DM] = g.ier;
                if (g.status) atomicOr(status + g.cand, (int)g.status);
                g.cand = -1;
            }
            if (more) {
                const long long c = (long long)atomicAdd(counter, 1ULL);
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
             double* feat, int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, int ck) {
    const int threads = 128;
    const size_t smem = (size_t)lane_ws_doubles(N) * 8 * threads;
    int grid;
    int rc = persistent_grid(ctx, k_dmfit<N>, threads, smem, &grid);
    if (rc) return rc;
    if ((long long)grid * threads > n) grid = (int)((n + threads - 1) / threads);
    k_dmfit<N><<<grid, threads, smem, st>>>(dmcurve, s->dm_dtype, n, s->n_dm, s->family, params, feat, ld, ier,
                                             status, take_counter(ctx, ck, st));
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
                      double* feat, int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, int ck) {
    if (s->n_dm < 4) return set_err(ctx, PFX_ERR_UNSUPPORTED, "n_dm < 4");
    if (s->family == PFX_FAMILY_PFD)
        return launch_n<4>(ctx, s, n, dmcurve, params, feat, ld, ier, status, st, ck);
    return launch_n<3>(ctx, s, n, dmcurve, params, feat, ld, ier, status, st, ck);
}
