// pfx_k_hist.cu -- scores 5-7: Freedman-Diaconis histograms of the profile and of its derivative and
// the three Gaussian fits on them (ProfileOperations.py:654-722,774-1057,
// ProfileOperationsInterface.py:138-186).  One warp per candidate; the histogram lives in shared
// memory, or in the warp's global scratch slot when it has more bins than fit there.
// The expectation of the profile-histogram fit, needed by scores 8-11, is parked in the score-8
// column of the feature row; the next kernel (pfx_k_gauss.cu) consumes it.
// the n x n trust-region solve runs from the rolled shared-memory routines here (smaller code;
// measured slightly faster than the register-resident SmallLM for this kernel)
#define PFX_SMALL_LM 0
#include "pfx_internal.cuh"
#include "pfx_lane_hist.cuh"

using namespace pfx;

namespace {

__global__ void __launch_bounds__(BLOCK_THREADS, 4)
k_hist(const void* __restrict__ profile, int dtype, long long n, int m, double pow_m, double pow_m1,
       double* feat, long long ld, int8_t* ier, int32_t* status, u64* counter, int work_doubles,
       int cap_smem, double* gscratch, int cap_global, const int* __restrict__ list,
       const u64* __restrict__ list_count) {
    extern __shared__ double smem[];
    Group<WARP> g;
    const int gid = threadIdx.x / WARP;
    double* ws = smem + gid * work_doubles;
    double* gs = gscratch
                     ? gscratch + ((size_t)blockIdx.x * GROUPS_PER_BLOCK + gid) * hist_scratch_doubles(cap_global)
                     : nullptr;
    // with an index list the kernel scores only those candidates (histograms that overflowed the
    // per-lane capacity of k_hist_lane)
    const long long total = list ? (long long)*list_count : n;
    for (;;) {
        const long long q = next_candidate(g, counter);
        if (q >= total) break;
        const long long c = list ? (long long)list[q] : q;
        double ph_expect;
        unsigned st = hist_candidate(g, ws, cap_smem, gs, cap_global, profile, dtype, c, m, pow_m, pow_m1,
                                     feat + c * ld, ier ? ier + c * PFX_NIER : nullptr, &ph_expect);
        if (g.leader()) {
            feat[c * ld + 7] = ph_expect;
            if (st) atomicOr(status + c, (int)st);
        }
    }
}

__global__ void __launch_bounds__(128, 3)
k_hist_lane(const void* __restrict__ profile, int dtype, long long n, int m, double pow_m, double pow_m1,
            double* feat, long long ld, int8_t* ier, int32_t* status, u64* counter, double* gbuf,
            int* ovf_list, u64* ovf_count) {
    extern __shared__ double smem[];
    LaneWs ws;
    ws.carve(smem + threadIdx.x, blockDim.x, 3);
    HistLaneBuf b;
    b.carve(gbuf + (size_t)blockIdx.x * blockDim.x + threadIdx.x, (int)(gridDim.x * blockDim.x), m);
    HistLane g;
    g.phase = GP_IDLE;
    g.cand = -1;
    bool more = true;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (;;) {
        if (g.phase == GP_IDLE) {
            if (g.cand >= 0) {
                double* row = feat + g.cand * ld;
                if (g.overflow) {
                    ovf_list[atomicAdd(ovf_count, 1ULL)] = (int)g.cand;
                } else {
                    double ph = qnan;
                    if (g.complete()) g.write(row, &ph);
                    row[7] = ph;                   // parked for the Gaussian kernels
                    if (ier) {
                        int8_t* ir = ier + g.cand * PFX_NIER;
                        ir[PFX_FIT_DYHIST] = g.ier[0];
                        ir[PFX_FIT_PROFHIST] = g.ier[1];
                        ir[PFX_FIT_FIXED] = g.ier[2];
                        ir[PFX_FIT_DYHIST_COUNT] = g.counts[0];
                        ir[PFX_FIT_PROFHIST_COUNT] = g.counts[1];
                    }
                    if (g.status) atomicOr(status + g.cand, (int)g.status);
                }
                g.cand = -1;
            }
            if (more) {
                const long long c = (long long)atomicAdd(counter, 1ULL);
                if (c < n) g.start(c, m, dtype == PFX_U8, pow_m, pow_m1);
                else more = false;
            }
        }
        if (!__any_sync(0xffffffffu, g.phase != GP_IDLE || more)) break;
        g.step(b, ws, profile, dtype);
    }
}


// ---- uint8 profiles of at most 128 bins (HTRU, SUPERB): no sort buffer, no global scratch -------------
// Round 1 gave every lane private y / dy / sort / edge / count arrays in a global scratch (5.4 KB per
// lane, ncu: 11.3 KB of DRAM traffic per candidate against 165 B of input and output) and let the lane
// sort and bin alone while the other 31 waited (4.7 of 32 lanes active per instruction).  Here the set-up
// of a histogram is done by the whole warp for one candidate at a time: four bins per lane, the two
// order-statistics sorts as bitonic networks over the warp's registers (shuffles), the counts added
// into a 256-byte table of the owner in shared memory; the fits then read counts from that table
// and compute the edges (HistBufU8).  Same arithmetic as the lane kernel, identical rows.
constexpr int HU8_THREADS = 128;

__device__ __forceinline__ void bitonic128(int (&v)[4], int lane) {
#pragma unroll 1
    for (int k = 2; k <= 128; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 4) {
                const int lj = j >> 2;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int idx = 4 * lane + e;
                    const int o = __shfl_xor_sync(0xffffffffu, v[e], lj);
                    const bool up = (idx & k) == 0;
                    const bool lower = (idx & j) == 0;
                    const int mn = min(v[e], o), mx = max(v[e], o);
                    v[e] = (lower == up) ? mn : mx;
                }
            } else {
                // partners inside the lane: (0,1)(2,3) for j = 1, (0,2)(1,3) for j = 2
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int pe = e ^ j;
                    if (pe > e) {
                        const bool up = ((4 * lane + e) & k) == 0;
                        const int a = v[e], b2 = v[pe];
                        const int mn = min(a, b2), mx = max(a, b2);
                        v[e] = up ? mn : mx;
                        v[pe] = up ? mx : mn;
                    }
                }
            }
        }
    }
}
// element of rank r of the sorted sequence (rank 4 * lane + e is register e of lane `lane`)
__device__ __forceinline__ double sorted_at(const int (&v)[4], int r) {
    const int sel = (r & 3) == 0 ? v[0] : ((r & 3) == 1 ? v[1] : ((r & 3) == 2 ? v[2] : v[3]));
    return (double)__shfl_sync(0xffffffffu, sel, r >> 2);
}
__device__ __forceinline__ double percentile128(const int (&v)[4], int n, double per) {
    const double idx = per / 100.0 * (double)(n - 1);
    const int i = (int)idx;
    if ((double)i == idx) return sorted_at(v, i) * 1.0 / 1.0;
    const int j = i + 1;
    const double w0 = (double)j - idx, w1 = idx - (double)i;
    const double a = sorted_at(v, i), b2 = sorted_at(v, i + 1 < n ? i + 1 : i);
    return (a * w0 + b2 * w1) / (w0 + w1);
}
// POI:138-166 on the integer data d[0..n) held four per lane (d beyond n must be INT_MAX)
__device__ __forceinline__ long long fd_bins128(const int (&d)[4], int lane, int n, double pow_n, double& lo, double& hi) {
    int v[4] = {d[0], d[1], d[2], d[3]};
    bitonic128(v, lane);
    const double iqr = percentile128(v, n, 75.0) - percentile128(v, n, 25.0);
    lo = sorted_at(v, 0);
    hi = sorted_at(v, n - 1);
    const double binwidth = 2.0 * iqr * pow_n;
    const double rnge = hi - lo;
    double q;
    if (binwidth <= 0.0) q = floor(rnge / 60.0);
    else q = rnge / binwidth;
    const double c = ceil(q);
    if (!(c < 9.0e15)) return 9000000000000000LL;
    return (long long)c;
}
// numpy.histogram (uniform-bin path) of the warp's data into the owner's byte table
__device__ __forceinline__ void histogram128(const int (&d)[4], int lane, int n, int nbins, double first, double last,
                                             unsigned* tbl, int rot, double& o_first, double& o_step) {
    if (first == last) {
        first -= 0.5;
        last += 0.5;
    }
    const double step = (last - first) / (double)nbins;
    tbl[lane] = 0u;
    tbl[lane + 32] = 0u;
    __syncwarp();
    const double denom = last - first;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        if (4 * lane + e >= n) continue;
        const double a = (double)d[e];
        if (!(a >= first && a <= last)) continue;
        const double fi = PFX_MUL(PFX_DIV(PFX_SUB(a, first), denom), (double)nbins);
        int idx = (int)fi;
        if (idx == nbins) idx -= 1;
        const double e_lo = PFX_ADD(PFX_MUL((double)idx, step), first);
        if (a < e_lo) idx -= 1;
        else {
            const double e_hi = (idx + 1 == nbins) ? last : PFX_ADD(PFX_MUL((double)(idx + 1), step), first);
            if (a >= e_hi && idx != nbins - 1) idx += 1;
        }
        const int pos = (idx + rot) & 255;
        atomicAdd(tbl + (pos >> 2), 1u << (8 * (pos & 3)));
    }
    __syncwarp();
    o_first = first;
    o_step = step;
}

typedef HistLaneT<HistBufU8> HistLaneU8;

__global__ void __launch_bounds__(HU8_THREADS, 3)
k_hist_lane_u8(const unsigned char* __restrict__ profile, long long n, int m, double pow_m, double pow_m1,
               double* feat, long long ld, int8_t* ier, int32_t* status, u64* counter, int* ovf_list,
               u64* ovf_count) {
    extern __shared__ double smem[];
    LaneWs ws;
    ws.carve(smem + threadIdx.x, blockDim.x, 3);
    unsigned* tables = (unsigned*)(smem + (size_t)lane_ws_doubles(3) * blockDim.x);     // 64 words per lane
    const int lane = (int)(threadIdx.x & 31);
    unsigned* my_tbl = tables + (size_t)threadIdx.x * 64;
    const int rot = 4 * lane;
    HistLaneU8 g;
    g.phase = GP_IDLE;
    g.cand = -1;
    g.stage = 0;
    g.h_first = 0.0;
    g.h_step = 0.0;
    g.h_nb = 0;
    bool more = true;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (;;) {
        if (g.phase == GP_IDLE) {
            if (g.cand >= 0) {
                double* row = feat + g.cand * ld;
                if (g.overflow) {
                    ovf_list[atomicAdd(ovf_count, 1ULL)] = (int)g.cand;
                } else {
                    double ph = qnan;
                    if (g.complete()) g.write(row, &ph);
                    row[7] = ph;                   // parked for the Gaussian kernels
                    if (ier) {
                        int8_t* ir = ier + g.cand * PFX_NIER;
                        ir[PFX_FIT_DYHIST] = g.ier[0];
                        ir[PFX_FIT_PROFHIST] = g.ier[1];
                        ir[PFX_FIT_FIXED] = g.ier[2];
                        ir[PFX_FIT_DYHIST_COUNT] = g.counts[0];
                        ir[PFX_FIT_PROFHIST_COUNT] = g.counts[1];
                    }
                    if (g.status) atomicOr(status + g.cand, (int)g.status);
                }
                g.cand = -1;
            }
            if (more) {
                const long long c = (long long)atomicAdd(counter, 1ULL);
                if (c < n) g.start(c, m, true, pow_m, pow_m1);
                else more = false;
            }
        }
        if (!__any_sync(0xffffffffu, g.phase != GP_IDLE || more)) break;
        // histogram set-ups (stages 0 and 1) with the whole warp, one candidate at a time
        unsigned su = __ballot_sync(0xffffffffu, g.phase == GP_SETUP && g.stage <= 1);
        while (su) {
            const int src = __ffs(su) - 1;
            su &= su - 1;
            const long long cd = __shfl_sync(0xffffffffu, g.cand, src);
            const int st = __shfl_sync(0xffffffffu, g.stage, src);
            unsigned* tbl = tables + (size_t)(threadIdx.x - lane + src) * 64;
            const unsigned char* p = profile + cd * (long long)m;
            int y[4], dy[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) y[e] = (4 * lane + e < m) ? (int)__ldg(p + 4 * lane + e) : 0x7fffffff;
            double first = 0.0, step = 0.0;
            int nb = 0;
            if (st == 0) {
                const int ynext = __shfl_down_sync(0xffffffffu, y[0], 1);
                dy[0] = y[0] - y[1];
                dy[1] = y[1] - y[2];
                dy[2] = y[2] - y[3];
                dy[3] = y[3] - ynext;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (4 * lane + e >= m - 1) dy[e] = 0x7fffffff;
                double plo, phi, dlo, dhi;
                const long long np_ = fd_bins128(y, lane, m, pow_m, plo, phi);
                const long long nd_ = fd_bins128(dy, lane, m - 1, pow_m1, dlo, dhi);
                int outcome = 0;                                   // 1: zero bins, 2: overflow to the warp kernel
                if (nd_ < 1 || np_ < 1) outcome = 1;
                else if (nd_ > HIST_LANE_CAP || np_ > HIST_LANE_CAP) outcome = 2;
                if (outcome == 0) {
                    nb = (int)nd_;
                    histogram128(dy, lane, m - 1, nb, dlo, dhi, tbl, 4 * src, first, step);
                }
                if (lane == src) {
                    g.pmin = plo;
                    g.pmax = phi;
                    g.dmin = dlo;
                    g.dmax = dhi;
                    if (outcome == 1) {
                        g.status |= PFX_ST_HIST_ZERO_BINS;
                        g.phase = GP_IDLE;
                    } else if (outcome == 2) {
                        g.overflow = true;
                        g.phase = GP_IDLE;
                    } else {
                        g.nbp = (int)np_;
                        g.nbd = (int)nd_;
                    }
                }
            } else {
                nb = __shfl_sync(0xffffffffu, g.nbp, src);
                const double plo = __shfl_sync(0xffffffffu, g.pmin, src), phi = __shfl_sync(0xffffffffu, g.pmax, src);
                histogram128(y, lane, m, nb, plo, phi, tbl, 4 * src, first, step);
            }
            if (lane == src) {
                g.h_first = first;
                g.h_step = step;
                g.h_nb = nb;
            }
        }
        HistBufU8 b;
        b.hx.first = g.h_first;
        b.hx.step = g.h_step;
        b.hx.nb = g.h_nb;
        b.hy.p = (const unsigned char*)my_tbl;
        b.hy.rot = rot;
        b.hy.nb = g.h_nb;
        if (g.phase == GP_SETUP) g.setup_fit(b);
        if (g.phase == GP_JAC) g.jac_pass(b, ws);
        if (g.phase == GP_TRIAL) g.trial_pass(b, ws);
        if (g.phase == GP_FIN) g.finish(b);
    }
}

}  // namespace

namespace {
struct HistPlan {
    int grid, lgrid, wd, cap_smem;
    size_t smem, lsmem;
    bool u8fast;           // k_hist_lane_u8: uint8 profile of at most 128 bins
};
}  // namespace

static int hist_plan(pfx_ctx* ctx, pfx_scratch_set* set, int m, int dtype, int64_t n, HistPlan* p) {
    p->cap_smem = 192;                              // typical histograms have 6..110 bins
    while (3 * p->cap_smem < m) p->cap_smem += 64;  // the derivative is staged in the Jacobian area
    p->wd = align_up(hist_work_doubles(m, p->cap_smem), 2);
    p->smem = (size_t)p->wd * 8 * GROUPS_PER_BLOCK;
    int rc = persistent_grid(ctx, k_hist, BLOCK_THREADS, p->smem, &p->grid);
    if (rc) return rc;
    // global histogram slots of the warp-form kernel: one per warp of its grid, per scratch set
    rc = ensure_buf(ctx, &set->hist_warp, (size_t)p->grid * GROUPS_PER_BLOCK * hist_scratch_doubles(HIST_CAP_GLOBAL) * 8);
    if (rc) return rc;
    const int threads = 128;
    p->u8fast = ctx->hist_u8 && dtype == PFX_U8 && m <= 128;
    if (p->u8fast) {
        p->lsmem = ((size_t)lane_ws_doubles(3) * 8 + 256) * HU8_THREADS;
        rc = persistent_grid(ctx, k_hist_lane_u8, HU8_THREADS, p->lsmem, &p->lgrid);
        if (rc) return rc;
        if ((long long)p->lgrid * HU8_THREADS > n) p->lgrid = (int)((n + HU8_THREADS - 1) / HU8_THREADS);
        return ensure_buf(ctx, &set->ovf, (size_t)n * sizeof(int));
    }
    p->lsmem = (size_t)lane_ws_doubles(3) * 8 * threads;
    rc = persistent_grid(ctx, k_hist_lane, threads, p->lsmem, &p->lgrid);
    if (rc) return rc;
#if defined(PFX_HIST_BLOCKS_PER_SM)
    if (p->lgrid > PFX_HIST_BLOCKS_PER_SM * ctx->sm_count) p->lgrid = PFX_HIST_BLOCKS_PER_SM * ctx->sm_count;
#endif
    if ((long long)p->lgrid * threads > n) p->lgrid = (int)((n + threads - 1) / threads);
    rc = ensure_buf(ctx, &set->hist_lane, (size_t)p->lgrid * threads * (size_t)hist_lane_doubles(m) * 8);
    if (rc) return rc;
    return ensure_buf(ctx, &set->ovf, (size_t)n * sizeof(int));
}

int pfx::reserve_hist(pfx_ctx* ctx, pfx_scratch_set* set, const pfx_shape* s, int64_t n) {
    HistPlan p;
    return hist_plan(ctx, set, s->nbin, s->dtype, n, &p);
}

int pfx::launch_hist(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* feat,
                     int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, pfx_scratch_set* set, int ck) {
    const int m = s->nbin;
    if (m < 8) return set_err(ctx, PFX_ERR_UNSUPPORTED, "nbin < 8");
    HistPlan p;
    int rc = hist_plan(ctx, set, m, s->dtype, n, &p);
    if (rc) return rc;
    const double pow_m = pow((double)m, -0.3333333);        // POI:151, evaluated by the C library
    const double pow_m1 = pow((double)(m - 1), -0.3333333);
    // pass 1: one lane per candidate, histograms up to HIST_LANE_CAP bins
    u64* ovf_count = take_counter(set, 7, st);
    if (p.u8fast)
        k_hist_lane_u8<<<p.lgrid, HU8_THREADS, p.lsmem, st>>>((const unsigned char*)profile, n, m, pow_m, pow_m1, feat, ld,
                                                               ier, status, take_counter(set, ck, st), (int*)set->ovf.p,
                                                               ovf_count);
    else
        k_hist_lane<<<p.lgrid, 128, p.lsmem, st>>>(profile, s->dtype, n, m, pow_m, pow_m1, feat, ld, ier, status,
                                                    take_counter(set, ck, st), (double*)set->hist_lane.p,
                                                    (int*)set->ovf.p, ovf_count);
    PFX_CUDA(ctx, cudaGetLastError());
    // pass 2: the few candidates whose histogram has more bins go through the warp-form kernel
    // (4096-bin scratch slot per warp); a few per cent of a synthetic batch end up here (1-2 k
    // candidates of 100 k): one warp each, so the whole persistent grid is launched; blocks without
    // work leave at once
    k_hist<<<p.grid, BLOCK_THREADS, p.smem, st>>>(profile, s->dtype, n, m, pow_m, pow_m1, feat, ld, ier, status,
                                                   take_counter(set, 6, st), p.wd, p.cap_smem,
                                                   (double*)set->hist_warp.p, HIST_CAP_GLOBAL, (int*)set->ovf.p, ovf_count);
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}
