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

#if defined(PFX_PROFILE)
// development build (make prof): cycles per section of the lane state machine
static __device__ unsigned long long pfx_prof[16];
extern "C" int pfx_prof_dump(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, pfx_prof, sizeof(unsigned long long) * 16);
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(pfx_prof, z, sizeof(z));
    }
    return 0;
}
#endif

namespace {

// Per-lane data vector and bin numbers live in a global scratch area laid out [row][lane] over all
// lanes of the grid, so a warp reads 32 consecutive float64 per row (coalesced, L1/L2 resident);
// the small n x n workspace is in shared memory, [element][lane of the block].
__device__ __forceinline__ LaneBuf lane_buf(double* gdat, unsigned short* gxs) {
    LaneBuf b;
    const size_t lanes = (size_t)gridDim.x * blockDim.x;
    const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    b.stride = (int)lanes;
    b.dat = gdat + lane;
    b.xs = gxs + lane;
    return b;
}
// shared-memory variant: ws[nws][T] float64 | dat[m][T] float64 | xs[m][T] uint16.  Measured
// faster than the global scratch whenever one block of >= 96 lanes fits: the kernels are bound by
// the latency of a candidate's dependent chain, so fewer, faster lanes win over more resident warps.
__device__ __forceinline__ LaneBuf lane_buf_smem(double* smem, int m, int nws) {
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

__global__ void __launch_bounds__(128, 2)
k_gauss_rounds(const void* __restrict__ profile, int dtype, long long n, int m, double* feat, long long ld,
               int8_t* ier, int32_t* status, u64* counter, GaussCarry* carry, double* gdat,
               unsigned short* gxs) {
    extern __shared__ double smem[];
    // lane data in shared memory when it fits (gdat == nullptr), else in the global scratch
    const LaneBuf b = gdat ? lane_buf(gdat, gxs) : lane_buf_smem(smem, m, lane_ws_doubles(4));
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
#if defined(PFX_PROFILE)
        {
            // development build: cycles per section, summed over the lanes that executed it
            long long t0 = clock64();
            if (g.phase == GP_SETUP) { g.setup(b); long long t1 = clock64(); atomicAdd(&pfx_prof[0], (u64)(t1 - t0)); atomicAdd(&pfx_prof[8], 1ULL); t0 = t1; }
            if (g.phase == GP_JAC) { g.jac_pass(b, ws); long long t1 = clock64(); atomicAdd(&pfx_prof[1], (u64)(t1 - t0)); atomicAdd(&pfx_prof[9], 1ULL); t0 = t1; }
            if (g.phase == GP_TRIAL) { g.trial_pass(b, ws); long long t1 = clock64(); atomicAdd(&pfx_prof[2], (u64)(t1 - t0)); atomicAdd(&pfx_prof[10], 1ULL); t0 = t1; }
            if (g.phase == GP_FIN) { g.finish(b); long long t1 = clock64(); atomicAdd(&pfx_prof[3], (u64)(t1 - t0)); atomicAdd(&pfx_prof[11], 1ULL); t0 = t1; }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) { atomicAdd(&pfx_prof[12], 1ULL); }
        }
#else
        g.step(b, ws);
#endif
    }
}

__global__ void __launch_bounds__(128, 2)
k_gauss_final(const void* __restrict__ profile, int dtype, long long n, int m, double* feat, long long ld,
              int8_t* ier, u64* counter, const GaussCarry* __restrict__ carry, double* gdat,
              unsigned short* gxs) {
    extern __shared__ double smem[];
    const LaneBuf b = gdat ? lane_buf(gdat, gxs) : lane_buf_smem(smem, m, lane_ws_doubles(8));
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

#if defined(PFX_PROFILE)
// micro-benchmark: every lane repeats the T1 Jacobian pass / trial pass of its candidate `reps`
// times (no LM logic), i.e. the row loops under ideal conditions: no divergence, no tail.
__global__ void __launch_bounds__(128, 2)
k_micro_rows(const void* __restrict__ profile, int dtype, long long n, int m, int reps, int what,
             double* gdat, unsigned short* gxs, double* sink) {
    extern __shared__ double smem[];
    const LaneBuf b = lane_buf(gdat, gxs);
    const LaneWs ws = lane_ws(smem, 4);
    GaussLane g;
    const long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) % n;
    g.start(profile, dtype, c, m, 100.0);
    g.setup(b);
    double acc = 0.0;
    for (int r = 0; r < reps; ++r) {
        if (what == 0) {
            g.jac_pass(b, ws);
            acc += g.lm.gnorm;
            g.phase = GP_JAC;
        } else if (what == 1) {
            g.lm.trial_begin(ws);
            G4Eval ev;
            ev.set(g.lm.xnew, g.fl);
            double s = 0.0;
#pragma unroll 4
            for (int i = 0; i < g.nrows; ++i) {
                const double f = PFX_SUB(b.d(i), ev.value((double)b.x(i)));
                s = PFX_FMA(f, f, s);
            }
            acc += s;
        } else {
            g.lm.trial_begin(ws);
            acc += g.lm.pnorm;
        }
    }
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
extern "C" float pfx_micro_rows(const void* profile, int dtype, long long n, int m, int reps, int what, int blocks_per_sm) {
    int dev, sms;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = 128, grid = sms * blocks_per_sm;
    const size_t lanes = (size_t)grid * threads;
    double *gdat, *sink;
    cudaMalloc(&gdat, lanes * m * 10 + 256);
    cudaMalloc(&sink, lanes * 8);
    unsigned short* gxs = (unsigned short*)((char*)gdat + lanes * m * 8);
    const size_t smem = (size_t)lane_ws_doubles(4) * 8 * threads;
    cudaFuncSetAttribute(k_micro_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_micro_rows<<<grid, threads, smem>>>(profile, dtype, n, m, 2, what, gdat, gxs, sink);
    float ms = 1e30f;
    for (int t = 0; t < 3; ++t) {
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k_micro_rows<<<grid, threads, smem>>>(profile, dtype, n, m, reps, what, gdat, gxs, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float t_ms = -1;
        cudaEventElapsedTime(&t_ms, e0, e1);
        if (t_ms < ms) ms = t_ms;
    }
    if (cudaGetLastError() != cudaSuccess) ms = -2;
    cudaFree(gdat);
    cudaFree(sink);
    return ms;
}
#endif

static const int LANE_THREADS = 128;

// Block size and dynamic shared memory of a lane kernel.  Returns true when the per-lane data
// vectors go to shared memory (one block of 128 or 96 lanes fits), false for the global scratch.
static bool lane_layout(pfx_ctx* ctx, int m, int np, int* threads, size_t* smem) {
    const size_t ws = (size_t)lane_ws_doubles(np) * 8;
    const size_t per_lane = ws + (size_t)m * 10;
    const size_t budget = (size_t)ctx->smem_optin;
    for (int t = 128; t >= 96; t -= 32) {
        if (per_lane * t + 64 <= budget) {
            *threads = t;
            *smem = (per_lane * t + 15) / 16 * 16;
            return true;
        }
    }
    *threads = LANE_THREADS;
    *smem = ws * LANE_THREADS;
    return false;
}

// scratch for the per-lane data vectors of one launch: lanes x m x (8 + 2) bytes
static int lane_scratch(pfx_ctx* ctx, size_t lanes, int m, int slot, double** gdat, unsigned short** gxs) {
    const size_t need = ((lanes * (size_t)m * 10) + 255) / 256 * 256;
    if (ctx->lane_bytes[slot] < need) {
        if (ctx->lane_buf[slot]) {
            cudaDeviceSynchronize();
            cudaFree(ctx->lane_buf[slot]);
            ctx->scratch_bytes -= ctx->lane_bytes[slot];
            ctx->lane_buf[slot] = nullptr;
            ctx->lane_bytes[slot] = 0;
        }
        PFX_CUDA(ctx, cudaMalloc(&ctx->lane_buf[slot], need));
        ctx->lane_bytes[slot] = need;
        ctx->scratch_bytes += need;
    }
    *gdat = (double*)ctx->lane_buf[slot];
    *gxs = (unsigned short*)((char*)ctx->lane_buf[slot] + lanes * (size_t)m * 8);
    return PFX_OK;
}

int pfx::launch_gauss(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* feat,
                      int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, int ck) {
    const int m = s->nbin;
    if (m < 8 || m > 65535) return set_err(ctx, PFX_ERR_UNSUPPORTED, "nbin out of range");
    const int par = ctx->counter_slot & 1;      // calls in flight alternate between two scratch sets
    // carry records
    const size_t need = (size_t)n * sizeof(GaussCarry);
    if (ctx->carry_bytes < need) {
        if (ctx->carry) {
            cudaDeviceSynchronize();
            cudaFree(ctx->carry);
            ctx->scratch_bytes -= 2 * ctx->carry_bytes;
            ctx->carry = nullptr;
            ctx->carry_bytes = 0;
        }
        PFX_CUDA(ctx, cudaMalloc(&ctx->carry, need * 2));
        ctx->carry_bytes = need;
        ctx->scratch_bytes += need * 2;
    }
    GaussCarry* carry = (GaussCarry*)((char*)ctx->carry + (size_t)par * ctx->carry_bytes);
    double* gdat;
    unsigned short* gxs;
    int grid, rc;
    {
        int threads;
        size_t smem;
        const bool in_smem = lane_layout(ctx, m, 4, &threads, &smem);
        rc = persistent_grid(ctx, k_gauss_rounds, threads, smem, &grid);
        if (rc) return rc;
        if ((long long)grid * threads > n) grid = (int)((n + threads - 1) / threads);
        gdat = nullptr;
        gxs = nullptr;
        if (!in_smem) {
            rc = lane_scratch(ctx, (size_t)grid * threads, m, par * 2 + 0, &gdat, &gxs);
            if (rc) return rc;
        }
        k_gauss_rounds<<<grid, threads, smem, st>>>(profile, s->dtype, n, m, feat, ld, ier, status,
                                                     take_counter(ctx, ck, st), carry, gdat, gxs);
        PFX_CUDA(ctx, cudaGetLastError());
    }
    {
        int threads;
        size_t smem;
        const bool in_smem = lane_layout(ctx, m, 8, &threads, &smem);
        rc = persistent_grid(ctx, k_gauss_final, threads, smem, &grid);
        if (rc) return rc;
        if ((long long)grid * threads > n) grid = (int)((n + threads - 1) / threads);
        gdat = nullptr;
        gxs = nullptr;
        if (!in_smem) {
            rc = lane_scratch(ctx, (size_t)grid * threads, m, par * 2 + 1, &gdat, &gxs);
            if (rc) return rc;
        }
        k_gauss_final<<<grid, threads, smem, st>>>(profile, s->dtype, n, m, feat, ld, ier,
                                                    take_counter(ctx, ck + 4, st), carry, gdat, gxs);
        PFX_CUDA(ctx, cudaGetLastError());
    }
    return PFX_OK;
}
