// pfx_k_sinusoid.cu -- scores 1-4 (ProfileOperations.py:190-587): profile statistics, peak count and
// the sine / sine-squared fits, a group of four lanes per candidate (LaneGroup; state machine of
// pfx_lane_misc.cuh).
#include "pfx_internal.cuh"
#include "pfx_lane_misc.cuh"

using namespace pfx;

namespace {

#ifndef PFX_SIN_LANES
#define PFX_SIN_LANES 4
#endif
constexpr int SIN_LANES = PFX_SIN_LANES;      // lanes per candidate (LaneGroup)

__global__ void __launch_bounds__(128, 3)
k_sinusoid(const void* __restrict__ profile, int dtype, long long n, int m, double* feat, long long ld,
           int8_t* ier, u64* counter, int chunk) {
    extern __shared__ double smem[];
    // chunk > 0: the block owns candidates [blockIdx.x * chunk, + chunk) (short-lived blocks that the block
    // scheduler can place in the gaps other kernels leave, see score_dev_impl); else one global work queue
    __shared__ int s_ctr;
    if (threadIdx.x == 0) s_ctr = 0;
    __syncthreads();
    const long long base = chunk > 0 ? (long long)blockIdx.x * chunk : 0;
    const long long limit = chunk > 0 ? ((base + chunk < n) ? base + chunk : n) : n;
    LaneWs ws;
    ws.carve(smem + threadIdx.x, blockDim.x, 2);
    typedef LaneGroup<SIN_LANES> G;
    SinLaneT<SIN_LANES> g;
    g.phase = GP_IDLE;
    g.cand = -1;
    bool more = true;
    for (;;) {
        if (g.phase == GP_IDLE) {
            if (g.cand >= 0) {
                if (G::rank() == 0) {
                    g.write(feat + g.cand * ld);
                    if (ier) {
                        ier[g.cand * PFX_NIER + PFX_FIT_SINE] = g.ier[0];
                        ier[g.cand * PFX_NIER + PFX_FIT_SINESQ] = g.ier[1];
                    }
                }
                g.cand = -1;
            }
            if (more) {
                long long c = 0;
                if (G::rank() == 0) c = chunk > 0 ? base + atomicAdd(&s_ctr, 1) : (long long)atomicAdd(counter, 1ULL);
                c = G::bcast(c);
                if (c < limit) g.start(profile, dtype, c, m);
                else more = false;
            }
        }
        if (!__any_sync(0xffffffffu, g.phase != GP_IDLE || more)) break;
        g.step(ws);
    }
}

}  // namespace

int pfx::launch_sinusoid(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* feat,
                         int64_t ld, int8_t* ier, cudaStream_t st, pfx_scratch_set* set, int ck, int chunk) {
    if (s->nbin < 2) return set_err(ctx, PFX_ERR_UNSUPPORTED, "nbin < 2");
    // one warp per block: a block leaves when its 32 lanes are done, so the tail of slow candidates
    // holds little shared memory and the Gaussian kernels' blocks fit next to it
    const int threads = 32;
    const size_t smem = (size_t)lane_ws_doubles(2) * 8 * threads;
    int grid;
    int rc = persistent_grid(ctx, k_sinusoid, threads, smem, &grid);
    if (rc) return rc;
    const int per_block = threads / SIN_LANES;
    if ((long long)grid * per_block > n) grid = (int)((n + per_block - 1) / per_block);
    if (chunk > 0) grid = (int)((n + chunk - 1) / chunk);
    // (development knob: at most sin_per_sm blocks per SM, so that the kernel runs in the background of the
    // Gaussian kernels with what they leave of an SM)
    if (chunk <= 0 && ctx->sin_per_sm > 0 && grid > ctx->sin_per_sm * ctx->sm_count) grid = ctx->sin_per_sm * ctx->sm_count;
    k_sinusoid<<<grid, threads, smem, st>>>(profile, s->dtype, n, s->nbin, feat, ld, ier,
                                             take_counter(set, ck, st), chunk);
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}
