// pfx_k_gauss.cu -- scores 8-11: Gaussian + background fit and the double-Gaussian procedure on the
// rotated profile (ProfileOperations.py:724-768,1061-1483), ONE LANE PER CANDIDATE.
//
// Two persistent kernels (see pfx_lane_gauss.cuh):
//   k_gauss_rounds   the nine 4-parameter fits (T1 + eight rounds) as a per-lane state machine;
//                    one trip of the loop = one Jacobian pass + one trial point for 32 candidates.
//   k_gauss_final    the 8-parameter fit, the combi/final choice and scores 10, 11.
// Lanes pull candidates from a global counter one at a time, so a lane whose fits converge
// quickly simply takes more candidates (the spread in function evaluations is ~100x).
// Per-lane data (float64 data vector + uint16 bin numbers) sit in shared memory with the lane
// index fastest, i.e. bank-conflict free; the n x n trust-region state lives in local memory.
#include "pfx_internal.cuh"
#include "pfx_lane_gauss.cuh"

using namespace pfx;

namespace {

// shared memory of a block of T lanes: ws[nws][T] float64 | dat[m][T] float64 | xs[m][T] uint16
__device__ __forceinline__ LaneBuf lane_buf(double* smem, int m, int nws) {
    LaneBuf b;
    b.stride = blockDim.x;
    b.dat = smem + (size_t)nws * blockDim.x + threadIdx.x;
    b.xs = (unsigned short*)(smem + (size_t)(nws + m) * blockDim.x) + threadIdx.x;
    return b;
}
__device__ __forceinline__ LaneWs lane_ws(double* smem, int n) {
    LaneWs w;
    w.carve(smem + threadIdx.x, blockDim.x, n);
    return w;
}

__global__ void __launch_bounds__(128, 1)
k_gauss_rounds(const void* __restrict__ profile, int dtype, long long n, int m, double* feat, long long ld,
               int8_t* ier, int32_t* status, u64* counter, GaussCarry* carry) {
    extern __shared__ double smem[];
    const LaneBuf b = lane_buf(smem, m, lane_ws_doubles(4));
    const LaneWs ws = lane_ws(smem, 4);
    GaussLane g;
    g.phase = GP_IDLE;
    g.finished = false;
    g.cand = -1;
    bool more = true;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (;;) {
        if (g.phase == GP_IDLE) {
            if (g.cand >= 0) {
                // flush the finished candidate
                double* row = feat + g.cand * ld;
                row[7] = g.t1_fwhm;
                row[8] = g.t1_chi;
                GaussCarry& cy = carry[g.cand];
                cy.skip = (g.sentinel || (g.status & PFX_ST_FAIL_MASK)) ? 1 : 0;
                if (g.sentinel) {
                    row[9] = 1000000.0;
                    row[10] = 1000000.0;
                }
                if (!cy.skip) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        cy.p_t1[k] = g.p_t1[k];
                        cy.p1[k] = g.p1[k];
                        cy.p2[k] = g.p2[k];
                    }
                    cy.minbg = g.minbg;
                    cy.sd = g.sd;
                    cy.t1_fwhm = g.t1_fwhm;
                }
                if (ier) {
                    int8_t* ir = ier + g.cand * PFX_NIER + PFX_FIT_T1;
                    for (int k = 0; k < 9; ++k) ir[k] = g.ier[k];
                }
                if (g.status) atomicOr(status + g.cand, (int)g.status);
                g.cand = -1;
            }
            if (more) {
                const long long c = (long long)atomicAdd(counter, 1ULL);
                if (c < n) {
                    double* row = feat + c * ld;
                    const double ph = row[7];      // parked by k_hist
                    if (status[c] & (int)PFX_ST_FAIL_MASK) {
                        row[7] = qnan;             // the reference raised before reaching score 8
                        carry[c].skip = 1;
                    } else {
                        g.start(profile, dtype, c, m, ph);
                    }
                } else {
                    more = false;
                }
            }
        }
        if (!__any_sync(0xffffffffu, g.phase != GP_IDLE || more)) break;
        g.step(b, ws);
    }
}

__global__ void __launch_bounds__(128, 1)
k_gauss_final(const void* __restrict__ profile, int dtype, long long n, int m, double* feat, long long ld,
              int8_t* ier, u64* counter, const GaussCarry* __restrict__ carry) {
    extern __shared__ double smem[];
    const LaneBuf b = lane_buf(smem, m, lane_ws_doubles(8));
    const LaneWs ws = lane_ws(smem, 8);
    Gauss8Lane g;
    g.phase = GP_IDLE;
    g.cand = -1;
    bool more = true;
    for (;;) {
        if (g.phase == GP_IDLE) {
            if (g.cand >= 0) {
                double* row = feat + g.cand * ld;
                row[9] = g.s10;
                row[10] = g.s11;
                if (ier) ier[g.cand * PFX_NIER + PFX_FIT_DG_FINAL] = g.ier;
                g.cand = -1;
            }
            if (more) {
                const long long c = (long long)atomicAdd(counter, 1ULL);
                if (c < n) {
                    if (!carry[c].skip) g.start(profile, dtype, c, m, carry[c], b);
                } else {
                    more = false;
                }
            }
        }
        if (!__any_sync(0xffffffffu, g.phase != GP_IDLE || more)) break;
        g.step(b, ws);
    }
}

}  // namespace

static int lane_block(pfx_ctx* ctx, int m, int np, int* threads, size_t* smem) {
    // per lane: LM workspace + m float64 + m uint16
    const size_t per_lane = (size_t)m * 10 + (size_t)lane_ws_doubles(np) * 8;
    int t = 128;
    const size_t budget = (size_t)ctx->smem_optin - 1024;
    while (t > 32 && per_lane * t > budget) t -= 32;
    if (per_lane * t > budget)
        return set_err(ctx, PFX_ERR_UNSUPPORTED, "nbin=%d needs %zu bytes of shared memory per lane", m, per_lane);
    *threads = t;
    *smem = (per_lane * t + 15) / 16 * 16;
    return PFX_OK;
}

int pfx::launch_gauss(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* feat,
                      int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, int ck) {
    const int m = s->nbin;
    if (m < 8 || m > 65535) return set_err(ctx, PFX_ERR_UNSUPPORTED, "nbin out of range");
    int threads;
    size_t smem;
    int rc = lane_block(ctx, m, 4, &threads, &smem);
    if (rc) return rc;
    // carry records
    const size_t need = (size_t)n * sizeof(GaussCarry);
    if (ctx->carry_bytes < need) {
        if (ctx->carry) {
            cudaDeviceSynchronize();
            cudaFree(ctx->carry);
            ctx->scratch_bytes -= ctx->carry_bytes;
            ctx->carry = nullptr;
            ctx->carry_bytes = 0;
        }
        PFX_CUDA(ctx, cudaMalloc(&ctx->carry, need * 2));      // two halves: calls in flight alternate
        ctx->carry_bytes = need;
        ctx->scratch_bytes += need * 2;
    }
    GaussCarry* carry = (GaussCarry*)((char*)ctx->carry + (size_t)(ctx->counter_slot & 1) * ctx->carry_bytes);
    int grid;
    rc = persistent_grid(ctx, k_gauss_rounds, threads, smem, &grid);
    if (rc) return rc;
    const long long lanes = (long long)grid * threads;
    if (lanes > n) grid = (int)((n + threads - 1) / threads);
    k_gauss_rounds<<<grid, threads, smem, st>>>(profile, s->dtype, n, m, feat, ld, ier, status,
                                                 take_counter(ctx, ck, st), carry);
    PFX_CUDA(ctx, cudaGetLastError());
    rc = lane_block(ctx, m, 8, &threads, &smem);
    if (rc) return rc;
    rc = persistent_grid(ctx, k_gauss_final, threads, smem, &grid);
    if (rc) return rc;
    if ((long long)grid * threads > n) grid = (int)((n + threads - 1) / threads);
    k_gauss_final<<<grid, threads, smem, st>>>(profile, s->dtype, n, m, feat, ld, ier,
                                                take_counter(ctx, ck + 4, st), carry);
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}
