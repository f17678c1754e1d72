// pfx_k_gauss.cu -- scores 8-11: Gaussian + background fit and the double-Gaussian procedure on the
// rotated profile (ProfileOperations.py:724-768,1061-1483).
//
// Two persistent kernels per call (fit logic in pfx_lane_gauss.cuh, Levenberg-Marquardt in pfx_lane_lm.cuh):
//   rounds   the nine 4-parameter fits (T1 + eight rounds of the double-Gaussian procedure)
//   final    the 8-parameter fit, the combi/final choice and scores 10, 11
// THE KERNELS IN USE are the block-synchronous pair k_gauss_rounds_s / k_gauss_final_s (further down): the
// state of a fit in the registers of an owner lane, data vectors and a small workspace per candidate slot in
// shared memory, the phases of a trip block-wide, row passes and set-up / close tasks pulled from block lists by
// groups of eight lanes, and a hand-over list between the two kernels so that their tails overlap.
// The forms they grew out of stay selectable for comparison:
//   k_gauss_rounds     one lane (or a group of C lanes) per candidate from its first fit to its last  (PFX_SLOTS=0)
//   k_gauss_rounds_h   hybrid: owner lanes + eight-lane row groups inside a warp                       (PFX_HYBRID=1)
//   k_gauss_wave       fit state in global memory, block-wide phases over all candidates of a block  (PFX_WAVE=1)
//   k_gauss_final      two lanes per candidate, one n x n workspace per pair                           (PFX_FINAL_SLOTS=0)
// All pull candidates from a global counter, because work per candidate is heavy-tailed (~100x).
#include "pfx_internal.cuh"
#include "pfx_lane_gauss.cuh"

using namespace pfx;

namespace {

// Per-lane data vector and bin numbers live in a global scratch area laid out [row][lane] over all
// lanes of the grid, so a warp reads 32 consecutive float64 per row (coalesced, L1/L2 resident);
// the small n x n workspace is in shared memory, [element][lane of the block].
template <int C>
__device__ __forceinline__ LaneBuf lane_buf(double* gdat, unsigned short* gxs) {
    LaneBuf b;
    const size_t groups = (size_t)gridDim.x * blockDim.x / C;
    const size_t group = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / C;
    b.stride = (int)groups;
    b.dat = gdat + group;
    b.xs = gxs + group;
    return b;
}
// shared-memory variant: ws[nws][T] float64 | dat[m][T] float64 | xs[m][T] uint16.  Measured
// faster than the global scratch whenever one block of >= 96 lanes fits: the kernels are bound by
// the latency of a candidate's dependent chain, so fewer, faster lanes win over more resident warps.
template <int C>
__device__ __forceinline__ LaneBuf lane_buf_smem(double* smem, int m, int nws) {
    LaneBuf b;
    const int groups = blockDim.x / C, group = threadIdx.x / C;
    b.stride = groups;
    double* dat = smem + (size_t)nws * blockDim.x;
    b.dat = dat + group;
    b.xs = (unsigned short*)(dat + (size_t)m * groups) + group;
    return b;
}
__device__ __forceinline__ LaneWs lane_ws(double* smem, int n) {
    LaneWs w;
    w.carve(smem + threadIdx.x, blockDim.x, n);
    return w;
}

#ifndef PFX_ROUNDS_LANES
#define PFX_ROUNDS_LANES 1
#endif
#ifndef PFX_FINAL_LANES
#define PFX_FINAL_LANES 2
#endif
constexpr int ROUNDS_LANES = PFX_ROUNDS_LANES;      // lanes per candidate in k_gauss_rounds
constexpr int FINAL_LANES = PFX_FINAL_LANES;        // and in k_gauss_final
// largest block considered (bounded by 64K registers per SM)
constexpr int ROUNDS_MAXT = ROUNDS_LANES == 1 ? 128 : (ROUNDS_LANES == 2 ? 224 : 320);
constexpr int FINAL_MAXT = FINAL_LANES == 1 ? 128 : 256;
constexpr int FINAL_WS_PER_GROUP = 1;          // the lanes of a group share the n x n workspace

// results of a candidate that finished its nine 4-parameter fits: scores 8, 9, the carry record of the
// final-fit kernel, ier slots, status bits (called by the owner lane)
template <class GL>
__device__ __forceinline__ void flush_rounds(GL& g, double* feat, long long ld, int8_t* ier, int32_t* status,
                                             GaussCarry* carry) {
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

__global__ void __launch_bounds__(ROUNDS_MAXT, 1)
k_gauss_rounds(const void* __restrict__ profile, int dtype, long long n, int m, double* feat, long long ld,
               int8_t* ier, int32_t* status, u64* counter, GaussCarry* carry, double* gdat,
               unsigned short* gxs) {
    extern __shared__ double smem[];
    // lane data in shared memory when it fits (gdat == nullptr), else in the global scratch
    typedef LaneGroup<ROUNDS_LANES> G;
    const LaneBuf b = gdat ? lane_buf<ROUNDS_LANES>(gdat, gxs) : lane_buf_smem<ROUNDS_LANES>(smem, m, lane_ws_doubles(4));
    const LaneWs ws = lane_ws(smem, 4);
    GaussLaneT<ROUNDS_LANES> g;
    g.phase = GP_IDLE;
    g.finished = false;
    g.cand = -1;
    bool more = true;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    PFX_PROF_T0(pk0);
#if defined(PFX_PROFILE)
    long long ptail = 0;
#endif
    for (;;) {
        PFX_PROF_T0(pq0);
#if defined(PFX_PROFILE)
        // the tail of a warp starts when its first lane finds the queue empty
        if (!ptail && __any_sync(0xffffffffu, !more)) ptail = clock64();
#endif
        if (g.phase == GP_IDLE) {
            if (g.cand >= 0 && G::rank() != 0) g.cand = -1;
            if (g.cand >= 0) flush_rounds(g, feat, ld, ier, status, carry);
            if (more) {
                long long c = 0;
                if (G::rank() == 0) c = (long long)atomicAdd(counter, 1ULL);
                c = G::bcast(c);
                if (c < n) {
                    double* row = feat + c * ld;
                    const double ph = row[7];      // parked by k_hist
                    G::sync();
                    // only failures of the histogram stage count here: the DM and sub-band kernels
                    // run concurrently on side streams and own their bits (the reference reaches
                    // them after the Gaussian fits)
                    const int before = (int)(PFX_ST_FAIL_MASK & ~(PFX_ST_DM_NDEG_ZERO | PFX_ST_SUBBAND_NO_PAIRS |
                                                                  PFX_ST_SUBBAND_BAD_WIDTH));
                    if (status[c] & before) {
                        if (G::rank() == 0) {
                            row[7] = qnan;         // the reference raised before reaching score 8
                            carry[c].skip = 1;
                        }
                    } else {
                        g.start(profile, dtype, c, m, ph);
                    }
                } else {
                    more = false;
                }
            }
        }
        __syncwarp();
        PFX_PROF_ADD(PROF_QUEUE, pq0);
        if (!__any_sync(0xffffffffu, g.phase != GP_IDLE || more)) break;
        if (ROUNDS_LANES == 1 && PFX_COOP) {
            // close and set up fits with the whole warp, one candidate at a time (pfx_lane_gauss.cuh)
            PFX_PROF_T0(pf0);
            unsigned fin = __ballot_sync(0xffffffffu, g.phase == GP_FIN);
            while (fin) {
                const int src = __ffs(fin) - 1;
                fin &= fin - 1;
                g.coop_finish(src, b);
            }
            PFX_PROF_ADD(PROF_FINISH, pf0);
            PFX_PROF_T0(ps0);
            unsigned set = __ballot_sync(0xffffffffu, g.phase == GP_SETUP);
            while (set) {
                const int src = __ffs(set) - 1;
                set &= set - 1;
                g.coop_setup(src, b, ws);
            }
            PFX_PROF_ADD(PROF_SETUP, ps0);
        }
        g.step(b, ws);
    }
    __syncwarp();
    PFX_PROF_ADD(PROF_TOTAL, pk0);
#if defined(PFX_PROFILE)
    if (ptail) PFX_PROF_ADD(PROF_TAIL, ptail);
#endif
}

// ---- hybrid form of the rounds kernel (pfx_lane_gauss.cuh, RowFactor4) --------------------------
// A warp owns K candidates (lanes 0 .. K-1 hold their fit state and run the n x n work); all 32 lanes
// serve the row passes, eight per candidate, four candidates per turn.  K must be 32 (the staging block of
// a worker lane lives in its own n x n workspace).  EXPERIMENTAL, off by default (PFX_HYBRID=1 selects it):
// measured on B200 at 100 000 candidates, histogram + rounds + final 621 ms against 633 ms for the lane
// kernel -- the tail of slow candidates shrinks (20 % of a warp's life instead of 28 %) and 15 % fewer
// instructions are issued, but four-row leaves plus three merge levels are 56 dependent Householder
// chains per Jacobian pass and candidate where the lane kernel runs 16, and those chains, not the
// exponentials, bound a pass (profiles/r02_phase_profile_hybrid_k32.txt, r02_rounds_hybrid_summary.txt).  Shared memory: ws[44][slots] | dat[slot][pitch] float64
// | xs[slot][pitch] uint16, one contiguous data vector per candidate (pitch = bins + 8: neighbouring
// candidates start 16 banks apart); long profiles keep the vectors in a global scratch of the same shape.
// Layout "packed" (K < 32, the one in use): a warp owns K < 32 candidates and a block has EIGHT warps, i.e. two
// per scheduler, which is what hides the dependent Householder chains of the row passes (with K = 32 and four
// warps the kernel only matched the lane kernel).  Every thread has a private 4-row staging block; a slot keeps
// only what lives between the phases (r, diag, qtf, the trial constants: HY_WSN doubles).
//     ws[HY_WSN][slots] | stg[20][threads] | dat[slot][pitch] float64 | xs[slot][pitch] uint16
constexpr int HY_MAX_SLOTS = 128;     // K = 32 layout
constexpr int HY_MAXT = 256;

__global__ void __launch_bounds__(HY_MAXT, 1)
k_gauss_rounds_h(const void* __restrict__ profile, int dtype, long long n, int m, double* feat, long long ld,
                 int8_t* ier, int32_t* status, u64* counter, GaussCarry* carry, double* gdat,
                 unsigned short* gxs, int pitch, int HY_K) {
    extern __shared__ double smem[];
    const int lane = (int)(threadIdx.x & 31), warp = (int)(threadIdx.x >> 5);
    const int nslots = (int)(blockDim.x >> 5) * HY_K;
    const int slot = warp * HY_K + lane;          // lanes >= K: never dereferenced, only used as a base for offsets
    const bool packed = HY_K < 32;
    LaneWs ws;
    ws.carve(smem + slot, nslots, 4);
    double* rest = smem + (size_t)(packed ? HY_WSN : lane_ws_doubles(4)) * nslots;
    if (packed) {
        ws.stg.p = rest + threadIdx.x;
        ws.stg.stride = (int)blockDim.x;
        ws.ipvt.p = (int*)(smem + slot + (size_t)HY_TC * nslots);      // dead before trial_prep writes the constants
        rest += (size_t)20 * blockDim.x;
    }
    LaneBuf b;
    b.stride = 1;
    b.pitch = pitch;
    if (gdat) {
        const size_t gslot = (size_t)blockIdx.x * nslots + slot;
        b.dat = gdat + gslot * pitch;
        b.xs = gxs + gslot * pitch;
    } else {
        b.dat = rest + (size_t)slot * pitch;
        b.xs = (unsigned short*)(rest + (size_t)nslots * pitch) + (size_t)slot * pitch;
    }
    GaussLaneT<1> g;
    g.phase = GP_IDLE;
    g.finished = false;
    g.cand = -1;
    g.nrows = 0;
    bool more = lane < HY_K;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    PFX_PROF_T0(pk0);
#if defined(PFX_PROFILE)
    long long ptail = 0;
#endif
    for (;;) {
        PFX_PROF_T0(pq0);
#if defined(PFX_PROFILE)
        if (!ptail && __any_sync(0xffffffffu, lane < HY_K && !more)) ptail = clock64();
#endif
        if (g.phase == GP_IDLE && lane < HY_K) {
            if (g.cand >= 0) flush_rounds(g, feat, ld, ier, status, carry);
            if (more) {
                const long long c = (long long)atomicAdd(counter, 1ULL);
                if (c < n) {
                    double* row = feat + c * ld;
                    const double ph = row[7];      // parked by k_hist
                    const int before = (int)(PFX_ST_FAIL_MASK & ~(PFX_ST_DM_NDEG_ZERO | PFX_ST_SUBBAND_NO_PAIRS |
                                                                  PFX_ST_SUBBAND_BAD_WIDTH));
                    if (status[c] & before) {
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
        __syncwarp();
        PFX_PROF_ADD(PROF_QUEUE, pq0);
        if (!__any_sync(0xffffffffu, g.phase != GP_IDLE || more)) break;
        {
            PFX_PROF_T0(pf0);
            unsigned fin = __ballot_sync(0xffffffffu, g.phase == GP_FIN);
            while (fin) {
                const int src = __ffs(fin) - 1;
                fin &= fin - 1;
                g.coop_finish(src, b);
            }
            PFX_PROF_ADD(PROF_FINISH, pf0);
            PFX_PROF_T0(ps0);
            unsigned set = __ballot_sync(0xffffffffu, g.phase == GP_SETUP);
            while (set) {
                const int src = __ffs(set) - 1;
                set &= set - 1;
                g.coop_setup(src, b, ws);
            }
            PFX_PROF_ADD(PROF_SETUP, ps0);
        }
        // Jacobian pass: constants by the owners, rows by the warp, n x n by the owners
        {
            const bool isjac = g.phase == GP_JAC;
            if (isjac) g.jac_prep(ws);
            __syncwarp();
            const unsigned jm = __ballot_sync(0xffffffffu, isjac);
            PFX_PROF_T0(pj0);
            if (jm) GaussLaneT<1>::hybrid_jac_rows(jm, g.nrows, b, ws);
            PFX_PROF_ADD(PROF_JAC_ROWS, pj0);
            if (isjac) g.jac_close(ws);
        }
        // trial point
        {
            const bool istr = g.phase == GP_TRIAL;
            if (istr) g.trial_prep(ws);
            __syncwarp();
            const unsigned tm = __ballot_sync(0xffffffffu, istr);
            PFX_PROF_T0(pt0);
            if (tm) GaussLaneT<1>::hybrid_trial_rows(tm, g.nrows, b, ws);
            PFX_PROF_ADD(PROF_TRIAL_ROWS, pt0);
            if (istr) g.trial_close(ws);
        }
    }
    __syncwarp();
    PFX_PROF_ADD(PROF_TOTAL, pk0);
#if defined(PFX_PROFILE)
    if (ptail) PFX_PROF_ADD(PROF_TAIL, ptail);
#endif
}

// ---- block-synchronous form of the hybrid kernel ("slots") ------------------------------------------
// What bounds the lane and hybrid kernels is instruction FETCH: a trip walks ~200 KB of mostly straight-line
// code (the unrolled n x n solve, the merge levels, set-up and close of a fit) and every warp of an SM is at its
// own place in it, so every warp instruction is fetched for one use (ncu: no_instruction 30 % of the stalls,
// more resident warps make the kernel SLOWER: packed hybrid layout with eight warps 643 ms against 525 ms with
// four).  The wave kernel showed the cure -- with all warps of a block in the same phase no_instruction falls
// to 2 % -- but paid for its global-memory state and its draining blocks with barrier and memory stalls.
// Here the fit state stays in the REGISTERS of an owner lane as in the hybrid kernel (K owners per warp, eight
// warps, slot = warp * K + lane), the data vectors and the small per-slot workspace stay in shared memory, the
// slots are refilled from the work queue, and only the PHASES of a trip are block-wide:
//     owners   trial_close of the previous trip, flush / refill, close and set up fits (warp-cooperative),
//              jac_prep; owners whose fit needs a Jacobian pass enter the block's list
//     rows     groups of eight lanes pull slots from the list: Jacobian rows + TSQR merge (group_jac_pass)
//     owners   jac_close, trial_prep (lmpar); owners with a trial point enter the second list
//     rows     groups pull: residual norm at the trial point
// Same arithmetic as the hybrid kernel (rows of a fit are dealt to eight lanes by row index only).
//     ws[HY_WSN][slots] | stg[20][threads] | nrows[slots], list[2][slots] int | dat[slot][pitch] | xs[slot][pitch]
#ifndef PFX_SL_MAXT
#define PFX_SL_MAXT 256
#endif
constexpr int SL_MAXT = PFX_SL_MAXT;      // (development: more warps per block at fewer registers per thread)

// ---- hand-over of finished candidates from the rounds kernel to the final-fit kernel --------------------
// Both kernels end in a tail set by their slowest candidates (34 and 42 ms at 100 000 candidates), and a kernel
// boundary between them puts the two tails in series.  With the hand-over the final-fit kernel is launched next
// to the rounds kernel (lower stream priority, so the rounds kernel's blocks are placed first): its blocks
// start on the SMs the rounds kernel's blocks leave and take candidates from a list the rounds kernel appends
// to as candidates finish.  Producer: carry record, fence, slot of the list (initialised to -1).  Consumer:
// an owner holds a ticket (its position in the list) and polls it once per trip, never blocking its block;
// when the last block of the rounds kernel has left, `done` holds the final count + 1.
struct PipeOut {
    int* list;              // nullptr: no hand-over (the final-fit kernel runs after the rounds kernel)
    u64 *produced, *exited, *done;
};
struct PipeIn {
    const int* list;
    u64* taken;
    const u64* done;
};

// Set-up and close of a fit as block-wide TASKS: run by the owner's warp (as in the hybrid kernel) they
// serialise inside that warp, and a trip waits for the warp with the most fits ending (ncu: 11.5 % of all
// samples at the barrier after the owners' phase).  Owners instead post a record in the idle workspace of
// their slot (rows SR .. SR+14: what to do, the fit parameters; the results come back in it), any warp
// takes the tasks before it pulls row passes, and the owner commits the results in the next trip.
constexpr int SR = 14;       // first workspace row of the task record; the bit set of the peak stripping keeps rows 0 .. SR-1
enum { SR_OP = SR, SR_ST_FIN, SR_ST_SET, SR_CAND, SR_P, SR_MEAN = SR_P, SR_SD, SR_PH = SR_P + 4, SR_CHI, SR_X0, SR_NR = SR_X0 + 4 };
static_assert(SR_NR < HY_WSN, "task record exceeds the slot workspace");

__global__ void __launch_bounds__(SL_MAXT, 1)
k_gauss_rounds_s(const void* __restrict__ profile, int dtype, long long n, int m, double* feat, long long ld,
                 int8_t* ier, int32_t* status, u64* counter, GaussCarry* carry, double* gdat,
                 unsigned short* gxs, int pitch, int K, int lmpar_budget, PipeOut po) {
    extern __shared__ double smem[];
    __shared__ int s_cnt[3], s_next[3];
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = (int)(blockDim.x >> 5) * K;
    const int slot = warp * K + lane;             // lanes >= K: only a base for offsets, never dereferenced
    const bool owner = lane < K;
    const int gs = lane & (HY_GROUP - 1);
    const unsigned gmask = 0xffu << (8 * (lane >> 3));
    LaneWs ws;
    ws.carve(smem + slot, S, 4);
    double* rest = smem + (size_t)HY_WSN * S;
    ws.stg.p = rest + tid;
    ws.stg.stride = (int)blockDim.x;
    ws.ipvt.p = (int*)(smem + slot + (size_t)HY_TC * S);     // dead before trial_prep writes the constants
    rest += (size_t)20 * blockDim.x;
    int* snrows = (int*)rest;
    int* list0 = snrows + S;
    int* list1 = list0 + S;
    int* list2 = list1 + S;
    rest += 2 * S;
    double* datbase;
    unsigned short* xsbase;
    if (gdat) {
        datbase = gdat + (size_t)blockIdx.x * S * pitch;
        xsbase = gxs + (size_t)blockIdx.x * S * pitch;
    } else {
        datbase = rest;
        xsbase = (unsigned short*)(rest + (size_t)S * pitch);
    }
    GaussLaneT<1> g;
    g.phase = GP_IDLE;
    g.finished = false;
    g.cand = -1;
    g.nrows = 0;
    bool more = owner;
    bool istr = false, lpend = false;
    LmparState lst;
    lst.parl = lst.paru = lst.fp = 0.0;
    lst.iter = 0;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    if (tid == 0) {
        s_cnt[0] = s_cnt[1] = s_cnt[2] = 0;
        s_next[0] = s_next[1] = s_next[2] = 0;
    }
    __syncthreads();
    for (;;) {
        // ---- owners: result of the trial pass, results of the tasks, new tasks, constants of the Jacobian pass ----
        if (istr) g.trial_close(ws);
        if (owner) {
            if (g.phase == GP_WAIT) {
                if (((int)ws.blk[SR_OP] & 1) && (int)ws.blk[SR_ST_FIN] == 0) g.t1_chi = ws.blk[SR_CHI];
                GaussLaneT<1>::SetupOut o;
#pragma unroll
                for (int k = 0; k < 4; ++k) o.x0[k] = ws.blk[SR_X0 + k];
                o.nr = (int)ws.blk[SR_NR];
                o.mean = ws.blk[SR_MEAN];
                o.sd = ws.blk[SR_SD];
                o.minbg = (o.mean < g.ph_expect) ? o.mean : g.ph_expect;      // builtin min(ph_expect, mean)
                g.commit_setup(g.stage, o);
            }
            int op = 0, st_fin = 0;
            double p[4] = {0.0, 0.0, 0.0, 0.0};
            if (g.phase == GP_FIN) {
                st_fin = g.stage;
#pragma unroll
                for (int k = 0; k < 4; ++k) p[k] = g.lm.x[k];
                g.commit_finish();
                if (st_fin < 8) op = 1;          // (the close of the last round has no pass over the data)
            }
            if (g.phase == GP_IDLE) {
                if (g.cand >= 0) {
                    const long long cf = g.cand;
                    flush_rounds(g, feat, ld, ier, status, carry);
                    if (po.list && !carry[cf].skip) {      // hand the candidate to the final-fit kernel
                        __threadfence();
                        const u64 pos = atomicAdd(po.produced, 1ULL);
                        ((volatile int*)po.list)[pos] = (int)cf;
                    }
                }
                if (more) {
                    const long long c = (long long)atomicAdd(counter, 1ULL);
                    if (c < n) {
                        double* row = feat + c * ld;
                        const double ph = row[7];      // parked by k_hist
                        const int before = (int)(PFX_ST_FAIL_MASK & ~(PFX_ST_DM_NDEG_ZERO | PFX_ST_SUBBAND_NO_PAIRS |
                                                                      PFX_ST_SUBBAND_BAD_WIDTH));
                        if (status[c] & before) {
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
            if (g.phase == GP_SETUP) op |= 2;
            if (op) {
                ws.blk[SR_OP] = (double)op;
                ws.blk[SR_ST_FIN] = (double)st_fin;
                ws.blk[SR_ST_SET] = (double)g.stage;
                ws.blk[SR_CAND] = (double)g.cand;
#pragma unroll
                for (int k = 0; k < 4; ++k) ws.blk[SR_P + k] = p[k];
                ws.blk[SR_PH] = g.ph_expect;
                list2[atomicAdd(&s_cnt[2], 1)] = slot;
                g.phase = GP_WAIT;
            }
        }
        const bool isjac = g.phase == GP_JAC;
        if (isjac) {
            g.jac_prep(ws);
            snrows[slot] = g.nrows;
            list0[atomicAdd(&s_cnt[0], 1)] = slot;
        }
        if (!__syncthreads_or((g.phase != GP_IDLE || more) ? 1 : 0)) break;
        // ---- tasks (a warp each), then the rows of the Jacobian passes (eight lanes each) ----
        {
            const int F = s_cnt[2];
            for (;;) {
                int k = 0;
                if (lane == 0) k = atomicAdd(&s_next[2], 1);
                k = __shfl_sync(0xffffffffu, k, 0);
                if (k >= F) break;
                const int sl = list2[k];
                double* w = smem + sl;
                const int op = (int)w[(size_t)SR_OP * S], st_fin = (int)w[(size_t)SR_ST_FIN * S], st_set = (int)w[(size_t)SR_ST_SET * S];
                const long long cd = (long long)w[(size_t)SR_CAND * S];
                double p[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) p[q] = w[(size_t)(SR_P + q) * S];
                const double ph = w[(size_t)SR_PH * S];
                LaneBuf bs;
                bs.dat = datbase + (size_t)sl * pitch;
                bs.xs = xsbase + (size_t)sl * pitch;
                bs.stride = 1;
                __syncwarp();      // inputs read by every lane before lane 0 writes results over them
                double chi = 0.0;
                if (op & 1) chi = GaussLaneT<1>::work_finish(bs, st_fin, p, profile, dtype, cd, m, m / 2);
                GaussLaneT<1>::SetupOut o;
                SPtr<double> bits;
                bits.p = w;
                bits.stride = S;
                GaussLaneT<1>::work_setup(0, bs, bits, SR, st_set, ph, profile, dtype, cd, m, m / 2, o);
                if (lane == 0) {
                    w[(size_t)SR_CHI * S] = chi;
                    w[(size_t)SR_MEAN * S] = o.mean;
                    w[(size_t)SR_SD * S] = o.sd;
#pragma unroll
                    for (int q = 0; q < 4; ++q) w[(size_t)(SR_X0 + q) * S] = o.x0[q];
                    w[(size_t)SR_NR * S] = (double)o.nr;
                }
                __syncwarp();
            }
        }
        {
            const int A = s_cnt[0];
            for (;;) {
                int k = 0;
                if (gs == 0) k = atomicAdd(&s_next[0], 1);
                k = __shfl_sync(gmask, k, 0, HY_GROUP);
                if (k >= A) break;
                const int sl = list0[k];
                LaneBuf bs;
                bs.dat = datbase + (size_t)sl * pitch;
                bs.xs = xsbase + (size_t)sl * pitch;
                bs.stride = 1;
                GaussLaneT<1>::group_jac_pass(bs, snrows[sl], gs, gmask, smem + sl, S, ws.stg);
            }
        }
        __syncthreads();
        if (tid == 0) {
            s_cnt[0] = s_cnt[2] = 0;
            s_next[0] = s_next[2] = 0;
        }
        // ---- owners: pivoting / Q^T f, lmpar ----
        if (isjac) g.jac_close(ws);
        // (at most lmpar_budget iterations of lmpar per trip: a phase lasts as long as its slowest owner, and one
        // sub-problem in twenty-five needs more than two -- such a fit stays in this phase for another trip)
        istr = false;
        if (g.phase == GP_TRIAL) {
            istr = g.trial_prep_budget(ws, lst, lpend, lmpar_budget);
            lpend = !istr;
            if (istr) list1[atomicAdd(&s_cnt[1], 1)] = slot;
        }
        __syncthreads();
        // ---- rows of the trial points ----
        {
            const int A = s_cnt[1];
            for (;;) {
                int k = 0;
                if (gs == 0) k = atomicAdd(&s_next[1], 1);
                k = __shfl_sync(gmask, k, 0, HY_GROUP);
                if (k >= A) break;
                const int sl = list1[k];
                LaneBuf bs;
                bs.dat = datbase + (size_t)sl * pitch;
                bs.xs = xsbase + (size_t)sl * pitch;
                bs.stride = 1;
                GaussLaneT<1>::group_trial_pass(bs, snrows[sl], gs, gmask, smem + sl, S);
            }
        }
        __syncthreads();
        if (tid == 0) {
            s_cnt[1] = 0;
            s_next[1] = 0;
        }
    }
    if (po.list) {
        // the last block to leave publishes the final count (every list entry is visible before it)
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            if (atomicAdd(po.exited, 1ULL) == (u64)gridDim.x - 1) {
                const u64 total = atomicAdd(po.produced, 0ULL);
                __threadfence();
                *(volatile u64*)po.done = total + 1;
            }
        }
    }
}

// ---- wave form of the rounds kernel ---------------------------------------------------------------
// The lane kernel keeps a candidate on one lane from its first fit to its last and lets the lanes of a
// warp wander through 240 KB of code independently: ncu shows instruction fetch (29 %) and fixed-latency
// dependencies (35 %) as its top stalls with one warp per scheduler, 12.5 of 32 lanes active per
// instruction, and a tail of slow candidates of 28 % of a warp's life.  Here the fit state lives in
// GLOBAL memory (fit state, n x n workspace and data vector: 2.3 KB per candidate), every candidate of
// the call is in flight from the start, and a block (one per SM, eight warps) walks ITS share of the
// candidates in phases separated by block barriers:
//     rows-jac    eight lanes per candidate: Jacobian rows, TSQR (RowFactor4), all lanes busy
//     lane A      one lane per candidate: pivoting / Q^T f (jac_end), then lmpar for every fit at a trial point
//     rows-trial  eight lanes per candidate: the residual norm at the trial point
//     lane B      one lane per candidate: the step decision; fits that ended are closed and the next one
//                 set up with the whole warp (coop_finish / coop_setup); finished candidates leave the list
// All warps of an SM run the same phase (one code region in the instruction cache), the cost of an
// iteration is proportional to the fits that still run, and the slowest candidate -- which bounds the
// number of iterations -- has been running since the first one: no queue, no tail of idle lanes.
// Arithmetic = the hybrid kernel's (which rows meet in which factor depends on the row count only).
constexpr int WV_THREADS = 256;
constexpr int WV_GROUPS = WV_THREADS / HY_GROUP;

struct WaveBufs {
    GaussLaneT<1>* state;      // [n]
    double* ws;                // [n][44]
    double* dat;               // [n][pitch]
    unsigned short* xs;        // [n][pitch]
    int* list;                 // [grid][2][cap]
    int pitch, cap;
};

__device__ __forceinline__ LaneWs wave_ws(const WaveBufs& wb, long long c) {
    LaneWs ws;
    ws.carve(wb.ws + c * lane_ws_doubles(4), 1, 4);
    return ws;
}
__device__ __forceinline__ LaneBuf wave_buf(const WaveBufs& wb, long long c) {
    LaneBuf b;
    b.dat = wb.dat + c * wb.pitch;
    b.xs = wb.xs + c * wb.pitch;
    b.stride = 1;
    b.pitch = -1;
    return b;
}

// shared memory of the wave kernel: staging blocks of the row passes [20][threads] | n x n workspaces of
// the lane phases [44][threads] | data vector of the candidate a group is working on [groups][pitch] float64
// + uint16 (only when it fits: profiles up to 256 bins; longer ones are read from global memory)
__host__ __device__ inline size_t wave_smem_bytes(int pitch, bool stage_data) {
    size_t b = (size_t)(20 + 44) * 8 * WV_THREADS;
    if (stage_data) b += (size_t)WV_GROUPS * pitch * 10;
    return b;
}

__global__ void __launch_bounds__(WV_THREADS, 1)
k_gauss_wave(const void* __restrict__ profile, int dtype, long long n, int m, double* feat, long long ld,
             int8_t* ier, int32_t* status, GaussCarry* carry, WaveBufs wb, int stage_data) {
    extern __shared__ double smem[];
    __shared__ int s_count, s_next[3];
    const int tid = (int)threadIdx.x, lane = tid & 31;
    SPtr<double> stg;
    stg.p = smem + tid;
    stg.stride = WV_THREADS;
    LaneWs wss;                                       // this thread's n x n workspace in shared memory
    wss.carve(smem + 20 * WV_THREADS + tid, WV_THREADS, 4);
    const int gi0 = tid / HY_GROUP, gs = tid & (HY_GROUP - 1);
    const unsigned gmask = 0xffu << (8 * ((tid & 31) >> 3));
    double* sdat = smem + (20 + 44) * WV_THREADS + (size_t)gi0 * wb.pitch;
    unsigned short* sxs = (unsigned short*)(smem + (20 + 44) * WV_THREADS + (size_t)WV_GROUPS * wb.pitch) + (size_t)gi0 * wb.pitch;
    int* list[2] = {wb.list + (size_t)blockIdx.x * 2 * wb.cap, wb.list + (size_t)blockIdx.x * 2 * wb.cap + wb.cap};
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const int nws = lane_ws_doubles(4);
    // candidates of this block: blockIdx.x, blockIdx.x + gridDim.x, ...
    const long long mine = (n > (long long)blockIdx.x) ? (n - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    for (long long k = tid; k < mine; k += WV_THREADS) {
        const long long c = (long long)blockIdx.x + k * gridDim.x;
        GaussLaneT<1> g;
        g.phase = GP_IDLE;
        g.finished = false;
        g.cand = -1;
        g.nrows = 0;
        double* row = feat + c * ld;
        const double ph = row[7];      // parked by k_hist
        const int before = (int)(PFX_ST_FAIL_MASK & ~(PFX_ST_DM_NDEG_ZERO | PFX_ST_SUBBAND_NO_PAIRS | PFX_ST_SUBBAND_BAD_WIDTH));
        if (status[c] & before) {
            row[7] = qnan;             // the reference raised before reaching score 8
            carry[c].skip = 1;
        } else {
            g.start(profile, dtype, c, m, ph);
        }
        wb.state[c] = g;
        list[0][k] = (int)c;
    }
    if (tid == 0) {
        s_count = 0;
        s_next[0] = s_next[1] = s_next[2] = 0;
    }
    __syncthreads();
    int cur = 0;
    int A = (int)mine;
    for (;;) {
        // ---- lane B: step decision, close / set up fits, constants of the next Jacobian pass, new list ----
        for (int base = (tid & ~31); base < A; base += WV_THREADS) {
            const int k = base + lane;
            const bool valid = k < A;
            const long long c = valid ? (long long)list[cur][k] : 0;
            GaussLaneT<1> g;
            if (valid) g = wb.state[c];
            else {
                g.phase = GP_IDLE;
                g.cand = -1;
                g.stage = 0;
                g.nrows = 0;
            }
            const LaneWs ws = wave_ws(wb, c);
            const LaneBuf b = wave_buf(wb, c);
            if (valid && g.phase == GP_TRIAL) g.trial_close(ws);
            unsigned fin = __ballot_sync(0xffffffffu, valid && g.phase == GP_FIN);
            while (fin) {
                const int src = __ffs(fin) - 1;
                fin &= fin - 1;
                g.coop_finish(src, b);
            }
            unsigned set = __ballot_sync(0xffffffffu, valid && g.phase == GP_SETUP);
            while (set) {
                const int src = __ffs(set) - 1;
                set &= set - 1;
                g.coop_setup(src, b, wss);        // (the workspace is only the owner's bit set of the peak stripping)
            }
            bool keep = false;
            if (valid) {
                if (g.phase == GP_JAC) g.jac_prep(ws);
                if (g.phase == GP_IDLE) {
                    if (g.cand >= 0) flush_rounds(g, feat, ld, ier, status, carry);
                } else {
                    wb.state[c] = g;
                    keep = true;
                }
            }
            const unsigned km = __ballot_sync(0xffffffffu, keep);
            int pos = 0;
            if (lane == 0 && km) pos = atomicAdd(&s_count, __popc(km));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (keep) list[cur ^ 1][pos + __popc(km & ((1u << lane) - 1u))] = (int)c;
        }
        __syncthreads();
        A = s_count;
        cur ^= 1;
        __syncthreads();
        if (tid == 0) {
            s_count = 0;
            s_next[2] = 0;
        }
        if (A == 0) break;
        // ---- rows-jac: groups of eight lanes pull candidates ------------------------------------------
        for (;;) {
            int k = 0;
            if (gs == 0) k = atomicAdd(&s_next[0], 1);
            k = __shfl_sync(gmask, k, 0, HY_GROUP);
            if (k >= A) break;
            const long long c = (long long)list[cur][k];
            const GaussLaneT<1>& st = wb.state[c];
            if (st.phase != GP_JAC) continue;
            const int nr = st.nrows;
            LaneBuf bs = wave_buf(wb, c);
            if (stage_data) {
                __syncwarp(gmask);
#pragma unroll 4
                for (int i = gs; i < nr; i += HY_GROUP) {
                    sdat[i] = bs.dat[i];
                    sxs[i] = bs.xs[i];
                }
                __syncwarp(gmask);
                bs.dat = sdat;
                bs.xs = sxs;
            }
            GaussLaneT<1>::group_jac_pass(bs, nr, gs, gmask, wb.ws + c * nws, 1, stg);
        }
        __syncthreads();
        if (tid == 0) s_next[0] = 0;
        // ---- lane A: close the Jacobian pass, lmpar for every fit at a trial point (warps pull 32 at a time) ----
        for (;;) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_next[1], 32);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= A) break;
            const int k = base + lane;
            if (k < A) {
                const long long c = (long long)list[cur][k];
                GaussLaneT<1> g = wb.state[c];
                double* gw = wb.ws + c * nws;
                if (g.phase == GP_JAC) {
#pragma unroll
                    for (int q = 0; q < 19; ++q) wss.blk[HY_JR + q] = gw[HY_JR + q];
                    g.jac_close(wss);
                } else {
#pragma unroll
                    for (int q = 0; q < 24; ++q) wss.blk[q] = gw[q];       // r and qtf of the factorisation in use
                }
                if (g.phase == GP_TRIAL) g.trial_prep(wss);
#pragma unroll
                for (int q = 0; q < 28; ++q) gw[q] = wss.blk[q];           // r, diag, qtf, trial constants
                wb.state[c] = g;
            }
        }
        __syncthreads();
        if (tid == 0) s_next[1] = 0;
        // ---- rows-trial ------------------------------------------------------------------------------
        for (;;) {
            int k = 0;
            if (gs == 0) k = atomicAdd(&s_next[2], 1);
            k = __shfl_sync(gmask, k, 0, HY_GROUP);
            if (k >= A) break;
            const long long c = (long long)list[cur][k];
            const GaussLaneT<1>& st = wb.state[c];
            if (st.phase != GP_TRIAL) continue;
            const int nr = st.nrows;
            LaneBuf bs = wave_buf(wb, c);
            if (stage_data) {
                __syncwarp(gmask);
#pragma unroll 4
                for (int i = gs; i < nr; i += HY_GROUP) {
                    sdat[i] = bs.dat[i];
                    sxs[i] = bs.xs[i];
                }
                __syncwarp(gmask);
                bs.dat = sdat;
                bs.xs = sxs;
            }
            GaussLaneT<1>::group_trial_pass(bs, nr, gs, gmask, wb.ws + c * nws, 1);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(FINAL_MAXT, 1)
k_gauss_final(const void* __restrict__ profile, int dtype, long long n, int m, double* feat, long long ld,
              int8_t* ier, u64* counter, const GaussCarry* __restrict__ carry, double* gdat,
              unsigned short* gxs) {
    extern __shared__ double smem[];
    typedef LaneGroup<FINAL_LANES> G;
    // one n x n workspace per group: [element][group of the block]
    const int groups = blockDim.x / FINAL_LANES;
    LaneBuf b;
    if (gdat) {
        b = lane_buf<FINAL_LANES>(gdat, gxs);
    } else {
        double* dat = smem + (size_t)lane_ws_doubles(8) * groups;
        b.stride = groups;
        b.dat = dat + threadIdx.x / FINAL_LANES;
        b.xs = nullptr;         // the 8-parameter fit runs on all bins in order: no bin numbers needed
    }
    LaneWs ws;
    ws.carve(smem + threadIdx.x / FINAL_LANES, groups, 8);
    Gauss8LaneT<FINAL_LANES> g;
    g.phase = GP_IDLE;
    g.cand = -1;
    bool more = true;
    PFX_PROF_T0(pk0);
    for (;;) {
        if (g.phase == GP_IDLE) {
            if (g.cand >= 0) {
                if (G::rank() == 0) {
                    double* row = feat + g.cand * ld;
                    row[9] = g.s10;
                    row[10] = g.s11;
                    if (ier) ier[g.cand * PFX_NIER + PFX_FIT_DG_FINAL] = g.ier;
                }
                g.cand = -1;
            }
            if (more) {
                long long c = 0;
                if (G::rank() == 0) c = (long long)atomicAdd(counter, 1ULL);
                c = G::bcast(c);
                if (c < n) {
                    PFX_PROF_T0(ps);
                    if (!carry[c].skip) g.start(profile, dtype, c, m, carry[c], b);
                    PFX_PROF_ADD(PROF8_START, ps);
                } else {
                    more = false;
                }
            }
        }
        if (!__any_sync(0xffffffffu, g.phase != GP_IDLE || more)) break;
        g.step(b, ws);
    }
    __syncwarp();
    PFX_PROF_ADD(PROF8_TOTAL, pk0);
}

// ---- block-synchronous form of the final-fit kernel ----------------------------------------------
// Same organisation as k_gauss_rounds_s for the 8-parameter fit: the scalars of a fit (par, delta, norms,
// counters) stay in the registers of an owner lane, its vectors (x, diag, the trial point) and the n x n
// workspace of the rolled MINPACK routines live in the slot's shared-memory workspace, and the phases of a
// trip are block-wide: owners (trial_end, start / close tasks posted, x published) | tasks by whole warps,
// then Jacobian rows by groups of FS_GROUP lanes (Gauss8LaneT<FS_GROUP>::group_jac: blocks of two rows
// eliminated against the lane's factor, butterfly merge, the leader leaves the factor unpacked in the
// workspace) | owners (jac_end on the unpacked factor, lmpar) | trial rows by groups.
// The lane-pair kernel it replaces ran 8.4 of 32 lanes per instruction with 255 registers and 21 KB of
// spill traffic per candidate, and at 100 000 candidates a fifth of its time was the tail of fits that run
// to maxfev (200 trips): here a trip of a lone fit is a third as long.
//     ws[G8_WSN][slots] | nrows-free lists: list[3][slots] int (+pad) | dat[slot][pitch] float64
#ifndef PFX_FS_GROUP
#define PFX_FS_GROUP 8
#endif
constexpr int FS_GROUP = PFX_FS_GROUP;
constexpr int FS_MAXT = 256;
typedef Gauss8LaneT<FS_GROUP> G8;

__global__ void __launch_bounds__(FS_MAXT, 1)
k_gauss_final_s(const void* __restrict__ profile, int dtype, long long n, int m, double* feat, long long ld,
                int8_t* ier, u64* counter, const GaussCarry* carry, double* gdat, int pitch, int K,
                int lmpar_budget, PipeIn pin) {
    extern __shared__ double smem[];
    __shared__ int s_cnt[3], s_next[3];
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = (int)(blockDim.x >> 5) * K;
    const int slot = warp * K + lane;             // lanes >= K: only a base for offsets, never dereferenced
    const bool owner = lane < K;
    const int gs = lane & (FS_GROUP - 1);
    const unsigned gmask = (FS_GROUP == 32) ? 0xffffffffu : (((1u << (FS_GROUP & 31)) - 1u) << (lane & ~(FS_GROUP - 1)));
    LaneWs ws;
    ws.carve(smem + slot, S, 8);
    const SPtr<double> wsl = ws.blk;              // row k of this slot's workspace: wsl[k]
    double* rest = smem + (size_t)G8::G8_WSN * S;
    int* list0 = (int*)rest;
    int* list1 = list0 + S;
    int* list2 = list1 + S;
    rest += 2 * S;
    double* datbase = gdat ? gdat + (size_t)blockIdx.x * S * pitch : rest;
    LaneLM<8> lm;
    long long cand = -1;
    int phase = GP_IDLE;
    bool more = owner, istr = false, lpend = false;
    long long ticket = -1;       // hand-over mode: this owner's position in the list of finished candidates
    LmparState lst;
    lst.parl = lst.paru = lst.fp = 0.0;
    lst.iter = 0;
    if (tid == 0) {
        s_cnt[0] = s_cnt[1] = s_cnt[2] = 0;
        s_next[0] = s_next[1] = s_next[2] = 0;
    }
    __syncthreads();
    for (;;) {
        // ---- owners ----
        int op = 0;
        if (owner) {
            if (istr) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    lm.x[k] = wsl[G8::G8_X + k];
                    lm.diag[k] = ws.diag[k];
                    lm.xnew[k] = wsl[G8::G8_XN + k];
                }
                const int r = lm.trial_end(sqrt(wsl[G8::G8_TS]));
#pragma unroll
                for (int k = 0; k < 8; ++k) wsl[G8::G8_X + k] = lm.x[k];
                phase = (r == LANE_JAC) ? GP_JAC : (r == LANE_TRIAL ? GP_TRIAL : GP_FIN);
            }
            if (phase == GP_WAIT) phase = GP_JAC;          // the profile of the new candidate is in place
            long long cfin = -1;
            if (phase == GP_FIN) {
                // close: parameters to the trial-point rows (dead now), the task writes scores 10 and 11
                if (ier) ier[cand * PFX_NIER + PFX_FIT_DG_FINAL] = (int8_t)lm.info;
#pragma unroll
                for (int k = 0; k < 8; ++k) wsl[G8::G8_XN + k] = wsl[G8::G8_X + k];
                cfin = cand;
                cand = -1;
                op = 1;
                phase = GP_IDLE;
            }
            if (phase == GP_IDLE && more) {
                if (pin.list) {
                    // hand-over: poll this owner's ticket once per trip (never block the block)
                    if (ticket < 0) ticket = (long long)atomicAdd(pin.taken, 1ULL);
                    const int c = (ticket < n) ? ((const volatile int*)pin.list)[ticket] : -1;
                    if (c >= 0) {
                        __threadfence();
                        cand = c;
                        ticket = -1;
                    } else {
                        const u64 d = *(const volatile u64*)pin.done;
                        if (d != 0 && (u64)ticket >= d - 1) more = false;      // the rounds kernel is done and the list is shorter
                    }
                } else {
                    // candidates the rounds kernel finished itself (failure, sentinel) are skipped
                    for (;;) {
                        const long long c = (long long)atomicAdd(counter, 1ULL);
                        if (c >= n) {
                            more = false;
                            break;
                        }
                        if (!carry[c].skip) {
                            cand = c;
                            break;
                        }
                    }
                }
                if (cand >= 0) {
                    const double* cy = (const double*)(carry + cand);      // p_t1[4] | p1[4] | p2[4] | ...
                    double x0[8];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        x0[k] = __ldcg(cy + 4 + k);
                        x0[4 + k] = __ldcg(cy + 8 + k);
                    }
                    lm.init(x0);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        wsl[G8::G8_X + k] = x0[k];
                        ws.diag[k] = 0.0;
                    }
                    op |= 2;
                    phase = GP_WAIT;
                }
            }
            if (op) {
                wsl[G8::G8_OP] = (double)op;
                wsl[G8::G8_CFIN] = (double)cfin;
                wsl[G8::G8_CNEW] = (double)cand;
                list2[atomicAdd(&s_cnt[2], 1)] = slot;
            }
        }
        const bool isjac = phase == GP_JAC;
        if (isjac) list0[atomicAdd(&s_cnt[0], 1)] = slot;
        if (!__syncthreads_or((phase != GP_IDLE || more || op) ? 1 : 0)) break;
        if (pin.list && s_cnt[0] == 0 && s_cnt[2] == 0 && !__syncthreads_or(phase == GP_TRIAL ? 1 : 0)) {
            // nothing to do yet (the rounds kernel has not finished a candidate for this block): wait a little
            __nanosleep(2000);
            continue;
        }
        // ---- tasks (a warp each), then the rows of the Jacobian passes ----
        {
            const int F = s_cnt[2];
            for (;;) {
                int k = 0;
                if (lane == 0) k = atomicAdd(&s_next[2], 1);
                k = __shfl_sync(0xffffffffu, k, 0);
                if (k >= F) break;
                const int sl = list2[k];
                double* w = smem + sl;
                const int top = (int)w[(size_t)G8::G8_OP * S];
                const long long cfin = (long long)w[(size_t)G8::G8_CFIN * S], cnew = (long long)w[(size_t)G8::G8_CNEW * S];
                LaneBuf bs;
                bs.dat = datbase + (size_t)sl * pitch;
                bs.xs = nullptr;
                bs.stride = 1;
                if (top & 1) {
                    double q[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) q[j] = w[(size_t)(G8::G8_XN + j) * S];
                    double s10, s11;
                    G8::task_finish(bs, q, carry[cfin], m, s10, s11);
                    if (lane == 0) {
                        feat[cfin * ld + 9] = s10;
                        feat[cfin * ld + 10] = s11;
                    }
                }
                if (top & 2) G8::task_start(bs, profile, dtype, cnew, m);
                __syncwarp();
            }
        }
        {
            const int A = s_cnt[0];
            for (;;) {
                int k = 0;
                if (gs == 0) k = atomicAdd(&s_next[0], 1);
                k = __shfl_sync(gmask, k, 0, FS_GROUP);
                if (k >= A) break;
                const int sl = list0[k];
                LaneBuf bs;
                bs.dat = datbase + (size_t)sl * pitch;
                bs.xs = nullptr;
                bs.stride = 1;
                G8::group_jac(bs, m, smem + sl, S);
            }
        }
        __syncthreads();
        if (tid == 0) {
            s_cnt[0] = s_cnt[2] = 0;
            s_next[0] = s_next[2] = 0;
        }
        // ---- owners: pivoting / Q^T f, lmpar ----
        if (isjac) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                lm.x[k] = wsl[G8::G8_X + k];
                lm.diag[k] = ws.diag[k];
                lm.q[k] = wsl[G8::G8_Q + k];
                lm.acn2[k] = wsl[G8::G8_ACN2 + k];
            }
            lm.f2 = wsl[G8::G8_F2];
            const int r = lm.jac_end(ws, m, true);
            phase = (r == LANE_TRIAL) ? GP_TRIAL : GP_FIN;
        }
        // lmpar: at most lmpar_budget iterations per trip.  One sub-problem in six needs more than two, one in
        // fifty more than four, and a phase lasts as long as its slowest owner: a fit whose sub-problem is hard
        // stays in this phase for another trip (its iteration state is four scalars) instead of holding up the block.
        istr = false;
        if (phase == GP_TRIAL) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                lm.x[k] = wsl[G8::G8_X + k];
                lm.diag[k] = ws.diag[k];
            }
            bool done = lpend ? false : lm.lmpar_start(ws, lst);
            for (int it = 0; !done && it < lmpar_budget; ++it) done = lm.lmpar_step(ws, lst);
            lpend = !done;
            if (done) {
                lm.trial_point_seq(ws);
#pragma unroll
                for (int k = 0; k < 8; ++k) wsl[G8::G8_XN + k] = lm.xnew[k];
                list1[atomicAdd(&s_cnt[1], 1)] = slot;
                istr = true;
            }
        }
        __syncthreads();
        // ---- rows of the trial points ----
        {
            const int A = s_cnt[1];
            for (;;) {
                int k = 0;
                if (gs == 0) k = atomicAdd(&s_next[1], 1);
                k = __shfl_sync(gmask, k, 0, FS_GROUP);
                if (k >= A) break;
                const int sl = list1[k];
                LaneBuf bs;
                bs.dat = datbase + (size_t)sl * pitch;
                bs.xs = nullptr;
                bs.stride = 1;
                G8::group_trial(bs, m, smem + sl, S);
            }
        }
        __syncthreads();
        if (tid == 0) {
            s_cnt[1] = 0;
            s_next[1] = 0;
        }
    }
}

}  // namespace


// Block size and dynamic shared memory of a lane kernel with C lanes per candidate: the largest
// block (multiple of 32 lanes, at most max_threads) whose groups fit with their data vectors in
// shared memory.  Returns false when not even two warps fit (long profiles): the data vectors then
// go to the global scratch and only the n x n workspaces stay in shared memory.
static bool lane_layout(pfx_ctx* ctx, int m, int np, int lanes_per_cand, int max_threads, int ws_per_group,
                        int bytes_per_bin, int* threads, size_t* smem) {
    const size_t ws = (size_t)lane_ws_doubles(np) * 8;
    const size_t per_group = ws * ws_per_group + (size_t)m * bytes_per_bin;
    const size_t budget = (size_t)ctx->smem_optin;
    for (int t = max_threads; t >= 64; t -= 32) {
        if (per_group * (t / lanes_per_cand) + 64 <= budget) {
            *threads = t;
            *smem = (per_group * (t / lanes_per_cand) + 64 + 15) / 16 * 16;
            return true;
        }
    }
    *threads = 128;
    *smem = ws * 128;
    return false;
}

// scratch for the data vectors of one launch: groups x m x (8 + 2) bytes
static int lane_scratch(pfx_ctx* ctx, pfx_buf* buf, size_t lanes, int m, double** gdat, unsigned short** gxs) {
    const size_t need = ((lanes * (size_t)m * 10) + 255) / 256 * 256;
    const int rc = ensure_buf(ctx, buf, need);
    if (rc) return rc;
    *gdat = (double*)buf->p;
    *gxs = (unsigned short*)((char*)buf->p + lanes * (size_t)m * 8);
    return PFX_OK;
}

namespace {
struct GaussPlan {
    int threads[2], grid[2];
    size_t smem[2];
    bool in_smem[2];
    int per_block[2];      // candidates resident per block
    int pitch;             // hybrid rounds kernel: elements between the data vectors of two candidates
    int hy_k;              // hybrid rounds kernel: candidates per warp
    int final_pitch, final_k;   // block-synchronous final-fit kernel: elements between data vectors, candidates per warp
};
}  // namespace

// hybrid rounds kernel: candidates per warp (K), warps, shared memory.  k_req = 32 selects the first layout
// (every lane an owner, staging inside the lane's workspace, up to four warps); otherwise the packed layout:
// eight warps and as many candidates per warp as fit (at most k_req; 0 = as many as fit, at most 24).
static bool hybrid_layout(pfx_ctx* ctx, int m, int k_req, int* threads, size_t* smem, int* slots_out, int* pitch, int* k_out) {
    *pitch = align_up(m, 8) + 8;
    const size_t budget = (size_t)ctx->smem_optin;
    if (k_req == 32) {
        const size_t ws = (size_t)lane_ws_doubles(4) * 8;
        const size_t per_slot = ws + (size_t)*pitch * 10;
        *k_out = 32;
        for (int slots = HY_MAX_SLOTS; slots >= 64; slots -= 32) {
            if (per_slot * slots + 64 <= budget) {
                *threads = slots;
                *slots_out = slots;
                *smem = (per_slot * slots + 64 + 15) / 16 * 16;
                return true;
            }
        }
        *threads = HY_MAX_SLOTS;
        *slots_out = HY_MAX_SLOTS;
        *smem = ws * HY_MAX_SLOTS;
        return false;
    }
    const int warps = ctx->hy_warps > 0 ? ctx->hy_warps : 8;
    const int kmax = k_req > 0 ? (k_req < 31 ? k_req : 31) : 24;
    const size_t fixed = (size_t)20 * 8 * warps * 32 + 64;
    const size_t per_slot = (size_t)HY_WSN * 8 + (size_t)*pitch * 10;
    *threads = warps * 32;
    int k = (budget > fixed) ? (int)((budget - fixed) / per_slot / warps) : 0;
    if (k > kmax) k = kmax;
    if (k >= 8) {
        *k_out = k;
        *slots_out = k * warps;
        *smem = (fixed + per_slot * k * warps + 15) / 16 * 16;
        return true;
    }
    // long profiles: data vectors in the global scratch, workspaces and staging blocks in shared memory
    *k_out = kmax < 16 ? kmax : 16;
    *slots_out = *k_out * warps;
    *smem = (fixed + (size_t)HY_WSN * 8 * *slots_out + 15) / 16 * 16;
    return false;
}

// block-synchronous kernel: eight warps, as many candidates per warp as fit (at most k_req; 0: at most 24)
static bool slots_layout(pfx_ctx* ctx, int m, int k_req, int* threads, size_t* smem, int* slots_out, int* pitch, int* k_out) {
    *pitch = align_up(m, 8) + ctx->hy_pad;
    const size_t budget = (size_t)ctx->smem_optin;
    const int warps = ctx->hy_warps > 0 ? ctx->hy_warps : 8;
    const int kmax = k_req > 0 ? (k_req < 32 ? k_req : 32) : 24;
    const size_t fixed = (size_t)20 * 8 * warps * 32 + 64;
    const size_t per_slot = (size_t)HY_WSN * 8 + (size_t)*pitch * 10 + 16;
    *threads = warps * 32;
    int k = (budget > fixed) ? (int)((budget - fixed) / per_slot / warps) : 0;
    if (k > kmax) k = kmax;
    if (k >= 6) {
        *k_out = k;
        *slots_out = k * warps;
        *smem = (fixed + per_slot * k * warps + 15) / 16 * 16;
        return true;
    }
    *k_out = kmax < 16 ? kmax : 16;
    *slots_out = *k_out * warps;
    *smem = (fixed + ((size_t)HY_WSN * 8 + 16) * *slots_out + 15) / 16 * 16;
    return false;
}

// block-synchronous final-fit kernel: eight warps, as many candidates per warp as fit (at most 24)
static bool final_slots_layout(pfx_ctx* ctx, int m, int* threads, size_t* smem, int* slots_out, int* pitch, int* k_out) {
    *pitch = align_up(m, 8) + 8;
    const size_t budget = (size_t)ctx->smem_optin;
    const int warps = 8;
    const size_t per_slot = (size_t)G8::G8_WSN * 8 + (size_t)*pitch * 8 + 16;
    *threads = warps * 32;
    int k = (budget > 64) ? (int)((budget - 64) / per_slot / warps) : 0;
    if (k > 24) k = 24;
    if (ctx->fs_k > 0 && k > ctx->fs_k) k = ctx->fs_k;
    if (k >= 4) {
        *k_out = k;
        *slots_out = k * warps;
        *smem = (64 + per_slot * k * warps + 15) / 16 * 16;
        return true;
    }
    *k_out = 16;
    *slots_out = 16 * warps;
    *smem = (64 + ((size_t)G8::G8_WSN * 8 + 16) * *slots_out + 15) / 16 * 16;
    return false;
}

// launch geometry of the two kernels for n candidates of m bins, and the scratch they need
static int gauss_plan(pfx_ctx* ctx, pfx_scratch_set* set, int m, int64_t n, GaussPlan* p) {
    int rc = ensure_buf(ctx, &set->carry, (size_t)n * sizeof(GaussCarry));
    if (rc) return rc;
    p->pitch = m;
    if (ctx->gauss_wave) {
        // wave kernel: one block per SM, fit state / workspaces / data vectors / lists of the whole call in global memory
        p->pitch = align_up(m, 8);
        p->threads[0] = WV_THREADS;
        p->smem[0] = wave_smem_bytes(p->pitch, true);
        if (p->smem[0] > (size_t)ctx->smem_optin) p->smem[0] = wave_smem_bytes(p->pitch, false);
        p->in_smem[0] = true;
        p->per_block[0] = 1 << 30;
        int g0;
        rc = persistent_grid(ctx, k_gauss_wave, WV_THREADS, p->smem[0], &g0);      // (sets the shared-memory attribute)
        if (rc) return rc;
        p->grid[0] = ctx->sm_count;
        if (p->grid[0] > n) p->grid[0] = (int)(n > 0 ? n : 1);
        const size_t cap = (size_t)((n + p->grid[0] - 1) / p->grid[0]);
        if ((rc = ensure_buf(ctx, &set->wave_state, (size_t)n * sizeof(GaussLaneT<1>)))) return rc;
        if ((rc = ensure_buf(ctx, &set->wave_ws, (size_t)n * lane_ws_doubles(4) * 8))) return rc;
        if ((rc = ensure_buf(ctx, &set->wave_list, (size_t)p->grid[0] * 2 * cap * sizeof(int)))) return rc;
        double* gd;
        unsigned short* gx;
        if ((rc = lane_scratch(ctx, &set->lane[0], (size_t)n, p->pitch, &gd, &gx))) return rc;
    } else if (ctx->gauss_slots) {
        p->in_smem[0] = slots_layout(ctx, m, ctx->hy_k, &p->threads[0], &p->smem[0], &p->per_block[0], &p->pitch, &p->hy_k);
        rc = persistent_grid(ctx, k_gauss_rounds_s, p->threads[0], p->smem[0], &p->grid[0]);
    } else if (ctx->gauss_hybrid) {
        p->in_smem[0] = hybrid_layout(ctx, m, ctx->hy_k, &p->threads[0], &p->smem[0], &p->per_block[0], &p->pitch, &p->hy_k);
        rc = persistent_grid(ctx, k_gauss_rounds_h, p->threads[0], p->smem[0], &p->grid[0]);
    } else {
        p->in_smem[0] = lane_layout(ctx, m, 4, ROUNDS_LANES, ROUNDS_MAXT, ROUNDS_LANES, 10, &p->threads[0], &p->smem[0]);
        p->per_block[0] = p->threads[0] / ROUNDS_LANES;
        rc = persistent_grid(ctx, k_gauss_rounds, p->threads[0], p->smem[0], &p->grid[0]);
    }
    if (rc) return rc;
    p->final_pitch = m;
    if (ctx->pipe && ctx->gauss_slots && ctx->final_slots && !ctx->gauss_wave) {
        if ((rc = ensure_buf(ctx, &set->ready, (size_t)(n > 0 ? n : 1) * sizeof(int)))) return rc;
    }
    if (ctx->final_slots) {
        p->in_smem[1] = final_slots_layout(ctx, m, &p->threads[1], &p->smem[1], &p->per_block[1], &p->final_pitch, &p->final_k);
        rc = persistent_grid(ctx, k_gauss_final_s, p->threads[1], p->smem[1], &p->grid[1]);
    } else {
        p->in_smem[1] = lane_layout(ctx, m, 8, FINAL_LANES, FINAL_MAXT, FINAL_WS_PER_GROUP, 8, &p->threads[1], &p->smem[1]);
        p->per_block[1] = p->threads[1] / FINAL_LANES;
        rc = persistent_grid(ctx, k_gauss_final, p->threads[1], p->smem[1], &p->grid[1]);
    }
    if (rc) return rc;
    for (int k = 0; k < 2; ++k) {
        if (k == 0 && ctx->gauss_wave) continue;
        const int per_block = p->per_block[k];
        if ((long long)p->grid[k] * per_block > n) p->grid[k] = (int)((n + per_block - 1) / per_block);
        if (!p->in_smem[k]) {
            double* gdat;
            unsigned short* gxs;
            rc = lane_scratch(ctx, &set->lane[k], (size_t)p->grid[k] * per_block, k == 0 ? p->pitch : p->final_pitch, &gdat, &gxs);
            if (rc) return rc;
        }
    }
    return PFX_OK;
}

int pfx::reserve_gauss(pfx_ctx* ctx, pfx_scratch_set* set, const pfx_shape* s, int64_t n) {
    GaussPlan p;
    return gauss_plan(ctx, set, s->nbin, n, &p);
}

int pfx::launch_gauss(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* feat,
                      int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, pfx_scratch_set* set, int ck) {
    const int m = s->nbin;
    if (m < 8 || m > 65535) return set_err(ctx, PFX_ERR_UNSUPPORTED, "nbin out of range");
    GaussPlan p;
    int rc = gauss_plan(ctx, set, m, n, &p);
    if (rc) return rc;
    ctx->gauss_launches = 2;
    GaussCarry* carry = (GaussCarry*)set->carry.p;
    double* gdat = nullptr;
    unsigned short* gxs = nullptr;
    // hand-over of finished candidates to the final-fit kernel (see PipeOut): only when the chain runs on the
    // context's high-priority stream, so that the rounds kernel's blocks are placed before the final-fit kernel's
    const bool piped = ctx->pipe && ctx->chain_on_hi && ctx->gauss_slots && ctx->final_slots && !ctx->gauss_wave && n > 0;
    PipeOut po = {nullptr, nullptr, nullptr, nullptr};
    PipeIn pin = {nullptr, nullptr, nullptr};
    if (!p.in_smem[0] && !ctx->gauss_wave) {
        rc = lane_scratch(ctx, &set->lane[0], (size_t)p.grid[0] * p.per_block[0], p.pitch, &gdat, &gxs);
        if (rc) return rc;
    }
    if (ctx->gauss_wave) {
        WaveBufs wb;
        wb.state = (GaussLaneT<1>*)set->wave_state.p;
        wb.ws = (double*)set->wave_ws.p;
        rc = lane_scratch(ctx, &set->lane[0], (size_t)n, p.pitch, &wb.dat, &wb.xs);
        if (rc) return rc;
        wb.list = (int*)set->wave_list.p;
        wb.pitch = p.pitch;
        wb.cap = (int)((n + p.grid[0] - 1) / p.grid[0]);
        k_gauss_wave<<<p.grid[0], p.threads[0], p.smem[0], st>>>(profile, s->dtype, n, m, feat, ld, ier, status, carry, wb,
                                                                  p.smem[0] == wave_smem_bytes(p.pitch, true) ? 1 : 0);
    } else if (ctx->gauss_slots) {
        if (piped) {
            // list of finished candidates (-1 = not yet), counters; the final-fit kernel may start once they are set
            cudaMemsetAsync(set->ready.p, 0xff, (size_t)n * sizeof(int), st);
            po.list = (int*)set->ready.p;
            po.produced = take_counter(set, 8, st);
            pin.taken = take_counter(set, 9, st);
            po.exited = take_counter(set, 10, st);
            po.done = take_counter(set, 11, st);
            pin.list = po.list;
            pin.done = po.done;
            cudaEventRecord(ctx->pipe_ev[ctx->chain_par], st);
        }
        k_gauss_rounds_s<<<p.grid[0], p.threads[0], p.smem[0], st>>>(profile, s->dtype, n, m, feat, ld, ier, status,
                                                                      take_counter(set, ck, st), carry, gdat, gxs, p.pitch, p.hy_k,
                                                                      ctx->lmpar_budget > 0 ? ctx->lmpar_budget : 10, po);
    }
    else if (ctx->gauss_hybrid)
        k_gauss_rounds_h<<<p.grid[0], p.threads[0], p.smem[0], st>>>(profile, s->dtype, n, m, feat, ld, ier, status,
                                                                      take_counter(set, ck, st), carry, gdat, gxs, p.pitch, p.hy_k);
    else
        k_gauss_rounds<<<p.grid[0], p.threads[0], p.smem[0], st>>>(profile, s->dtype, n, m, feat, ld, ier, status,
                                                                    take_counter(set, ck, st), carry, gdat, gxs);
    PFX_CUDA(ctx, cudaGetLastError());
    gdat = nullptr;
    gxs = nullptr;
    if (!p.in_smem[1]) {
        rc = lane_scratch(ctx, &set->lane[1], (size_t)p.grid[1] * p.per_block[1], p.final_pitch, &gdat, &gxs);
        if (rc) return rc;
    }
    if (ctx->final_slots && piped) {
        // next to the rounds kernel, on a stream of default (lower) priority; a few SMs are never taken by it, so
        // the rounds kernel makes progress whatever order the block scheduler chooses
        cudaStream_t sf = ctx->pf[ctx->chain_par];
        cudaStreamWaitEvent(sf, ctx->pipe_ev[ctx->chain_par], 0);
        int gridf = p.grid[1];
        if (gridf > ctx->sm_count - ctx->pipe_hold) gridf = ctx->sm_count - ctx->pipe_hold;
        if (gridf < 1) gridf = 1;
        k_gauss_final_s<<<gridf, p.threads[1], p.smem[1], sf>>>(profile, s->dtype, n, m, feat, ld, ier, nullptr, carry, gdat,
                                                                p.final_pitch, p.final_k,
                                                                ctx->lmpar_budget8 > 0 ? ctx->lmpar_budget8 : 10, pin);
        cudaEventRecord(ctx->pipe_done_ev[ctx->chain_par], sf);
        // the SMs held back above join in once the rounds kernel has ended (same list, same tickets)
        if (p.grid[1] > gridf) ctx->gauss_launches = 3;
        if (p.grid[1] > gridf)
            k_gauss_final_s<<<p.grid[1] - gridf, p.threads[1], p.smem[1], st>>>(profile, s->dtype, n, m, feat, ld, ier, nullptr, carry,
                                                                                gdat ? gdat + (size_t)gridf * p.per_block[1] * p.final_pitch : nullptr,
                                                                                p.final_pitch, p.final_k,
                                                                                ctx->lmpar_budget8 > 0 ? ctx->lmpar_budget8 : 10, pin);
        cudaStreamWaitEvent(st, ctx->pipe_done_ev[ctx->chain_par], 0);
    } else if (ctx->final_slots)
        k_gauss_final_s<<<p.grid[1], p.threads[1], p.smem[1], st>>>(profile, s->dtype, n, m, feat, ld, ier,
                                                                     take_counter(set, ck + 4, st), carry, gdat, p.final_pitch, p.final_k,
                                                                     ctx->lmpar_budget8 > 0 ? ctx->lmpar_budget8 : 10, pin);
    else
        k_gauss_final<<<p.grid[1], p.threads[1], p.smem[1], st>>>(profile, s->dtype, n, m, feat, ld, ier,
                                                                   take_counter(set, ck + 4, st), carry, gdat, gxs);
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}

#if defined(PFX_PROFILE)
// development builds only (not part of include/pfx.h): per-phase counters of k_gauss_rounds
extern "C" int pfx_debug_profile(unsigned long long* out32, int reset) {
    if (out32 && cudaMemcpyFromSymbol(out32, g_pfx_prof, sizeof(unsigned long long) * 32) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long z[32] = {0};
        if (cudaMemcpyToSymbol(g_pfx_prof, z, sizeof(z)) != cudaSuccess) return -1;
    }
    return 0;
}
#endif
