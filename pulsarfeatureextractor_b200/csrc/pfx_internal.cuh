// pfx_internal.cuh -- shared by the translation units of libpfx (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "pfx_driver.cuh"

namespace pfx {

constexpr int WARP = 32;
constexpr int BLOCK_THREADS = 128;            // 4 warps = 4 candidates in flight per block
constexpr int GROUPS_PER_BLOCK = BLOCK_THREADS / WARP;
constexpr int LYON_W = 8;                     // lanes per candidate in the Lyon kernel
constexpr int LYON_THREADS = 256;
constexpr int HIST_CAP_GLOBAL = 4096;         // bins a global scratch slot can hold
constexpr int N_SETS = 10;                    // scratch sets: launches in flight per context (see ScratchSet)
constexpr int COUNTERS_PER_SET = 16;
constexpr int N_EVENTS = 8;
constexpr int MAX_SUB = 4;                    // sub-batches of one pfx_score_dev call (own stream + set each)

typedef unsigned long long u64;

}  // namespace pfx

// Scratch of one launch group (one *_dev call, or one sub-batch of pfx_score_dev): work-queue
// counters, the carry records between the two Gaussian kernels, per-lane data vectors, histogram
// buffers.  A context owns N_SETS of them and hands them out round robin; `done` is recorded on
// the stream of the call after its last kernel, and the next user of the set makes its stream wait
// for that event first.  So any number of calls may be in flight on any streams of one context:
// more than N_SETS of them serialise on the device instead of sharing scratch (no host
// synchronisation either way).
struct pfx_buf {
    void* p;
    size_t bytes;
};
struct pfx_scratch_set {
    pfx::u64* counters;       // COUNTERS_PER_SET
    pfx_buf carry;            // per-candidate records passed between the two Gaussian kernels
    pfx_buf lane[2];          // per-lane data vectors of the Gaussian lane kernels (long profiles)
    pfx_buf wave_state, wave_ws, wave_list;   // wave form of the rounds kernel: fit state, n x n workspaces, active lists
    pfx_buf hist_lane;        // per-lane arrays of k_hist_lane
    pfx_buf sub_scratch;      // box-car sums of the sub-band kernel for long profiles (per warp)
    pfx_buf ready;            // candidates the rounds kernel has finished, in order of completion (hand-over to the final fit)
    pfx_buf ovf;              // candidates whose histogram overflowed the per-lane capacity
    pfx_buf hist_warp;        // global histogram slots of the warp-form kernel
    cudaEvent_t done;
    bool used;
};
struct pfx_grid_cache_entry {
    const void* fn;
    int threads;
    size_t smem;
    int grid;
};

struct pfx_ctx {
    int device;
    int sm_count;
    int smem_optin;
    char err[512];
    pfx::u64* counters;       // N_SETS * COUNTERS_PER_SET
    pfx_scratch_set sets[pfx::N_SETS];
    int set_rr;
    pfx_grid_cache_entry grid_cache[32];
    int n_grid_cache;
    int hist_u8;              // 1: warp-cooperative histogram set-up for uint8 profiles of <= 128 bins (default)
    int gauss_wave;           // 1: wave form of the rounds kernel (state in global memory, phases per block)
    int hy_pad;               // elements between the data vectors of neighbouring slots beyond the bins (bank spread)
    int hy_k, hy_warps;       // hybrid rounds kernel: candidates per warp (32: first layout, 0: as many as fit), warps per block
    int pipe;                 // 1: the final-fit kernel runs next to the rounds kernel and takes candidates as they finish
    int sub_global, sub_blocks;   // sub-band kernel: 1 = global scratch also for mid-size shapes; blocks per SM there (development)
    int sin_per_sm;           // > 0: cap of the sinusoid kernel's resident blocks per SM (development)
    int gauss_launches;       // kernels the last launch_gauss started (2, or 3 with the hand-over's held-back blocks)
    int pipe_suspend;         // set by pfx_score_host while a batch of several chunks is in flight
    int chain_on_hi, chain_par;   // set by score_dev_impl around the chain: it runs on hi[chain_par]
    cudaStream_t pf[2];
    cudaEvent_t pipe_ev[2], pipe_done_ev[2];
    int fill;                 // 1: main chain on a high-priority stream, sinusoid kernel in short blocks behind the histogram stage
    cudaStream_t hi[2];
    cudaEvent_t fill_ev[2], main_ev[2];
    int lmpar_budget8, fs_k, pipe_hold;   // final-fit kernel: lmpar iterations per trip, candidates per warp (0: as many as fit), SMs held back
    int lmpar_budget;         // block-synchronous kernels: lmpar iterations per trip (<= 0: no limit)
    int final_slots;          // 1: block-synchronous final-fit kernel, 0: lane-pair kernel
    int gauss_slots;          // 1: block-synchronous hybrid kernel (phases of a trip are block-wide)
    int gauss_hybrid;         // 1: hybrid rounds kernel (eight lanes per candidate in the row passes), 0: lane kernel
    int dm_exact;             // 1: DM-curve fit in MINPACK's own summation order (bit-exact form), 0: lane form
    int subbatch;             // sub-batches of the histogram/Gaussian chain per pfx_score_dev call
    bool reserved;            // pfx_reserve was called: growing scratch inside a call is reported
    cudaEvent_t ev[pfx::N_EVENTS];
    // fork/join: the independent score groups of one call run on side streams
    cudaStream_t side[2], side2[2], sub[pfx::MAX_SUB];
    cudaEvent_t fork_ev[2], join_ev[2], join2_ev[2], hist_ev[2], sub_ev[pfx::MAX_SUB];
    int fork_rr;
    bool ev_valid, ev_trace;
    int stage_timing;         // 1: serialise the stages of pfx_score_dev and time each (pfx_last_timing)
    int last_launches;
    // host-path staging
    cudaStream_t streams[2];
    void* stage[2];
    size_t stage_bytes[2];
    size_t scratch_bytes;
};

namespace pfx {

int set_err(pfx_ctx* ctx, int code, const char* fmt, ...);

#define PFX_CUDA(ctx, call)                                                                      \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return pfx::set_err(ctx, PFX_ERR_CUDA, "%s failed: %s (%s:%d)", #call,               \
                                cudaGetErrorString(e_), __FILE__, __LINE__);                     \
    } while (0)

// Next scratch set of the context, ordered behind its previous user on the device (see
// pfx_scratch_set).  release_set records the completion event on the stream that covers every
// kernel of the launch group.
pfx_scratch_set* take_set(pfx_ctx* ctx, cudaStream_t st);
void release_set(pfx_ctx* ctx, pfx_scratch_set* set, cudaStream_t st);
// grow-only scratch buffer.  Growing inside a call frees the old block, which synchronises the
// device: pfx_reserve() sizes every set up front so that this never happens on the hot path.
int ensure_buf(pfx_ctx* ctx, pfx_buf* b, size_t need);

// k-th work-queue counter of a set, zeroed on the launching stream
inline u64* take_counter(pfx_scratch_set* set, int k, cudaStream_t st) {
    u64* c = set->counters + k;
    cudaMemsetAsync(c, 0, sizeof(u64), st);
    return c;
}

// persistent launch: grid = resident blocks per SM x SM count.  The attribute call and the
// occupancy query run once per (kernel, block size, shared memory) and are cached in the context.
int persistent_grid_impl(pfx_ctx* ctx, const void* fn, int threads, size_t smem, int* grid);
template <class K>
inline int persistent_grid(pfx_ctx* ctx, K kernel, int threads, size_t smem, int* grid) {
    return persistent_grid_impl(ctx, (const void*)kernel, threads, smem, grid);
}

#if defined(__CUDACC__)
template <class G>
__device__ __forceinline__ long long next_candidate(const G& g, u64* counter) {
    u64 c = 0;
    if (g.leader()) c = atomicAdd(counter, 1ULL);
    return (long long)g.bcast(c, 0);
}
#endif

// launchers, one per translation unit
int launch_sinusoid(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* feat,
                    int64_t ld, int8_t* ier, cudaStream_t st, pfx_scratch_set* set, int ck, int chunk = 0);
int launch_hist(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* feat,
                int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, pfx_scratch_set* set, int ck);
int launch_gauss(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* feat,
                 int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, pfx_scratch_set* set, int ck);
int launch_dmfit(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* dmcurve, const double* params,
                 double* feat, int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, pfx_scratch_set* set, int ck);
int launch_dmfit_exact(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* dmcurve, const double* params,
                       double* feat, int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, pfx_scratch_set* set, int ck);
int launch_subband(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, const void* subbands,
                   const double* params, double* feat, int64_t ld, int32_t* status, cudaStream_t st, pfx_scratch_set* set, int ck);
int launch_lyon(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, const void* dmstat,
                double* feat, int64_t ld, cudaStream_t st);
int launch_pfd_params(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* params,
                      cudaStream_t st);
// size the scratch of a set for n candidates of this shape (what the launchers would grow lazily)
int reserve_hist(pfx_ctx* ctx, pfx_scratch_set* set, const pfx_shape* s, int64_t n);
int reserve_subband(pfx_ctx* ctx, pfx_scratch_set* set, const pfx_shape* s);
int reserve_gauss(pfx_ctx* ctx, pfx_scratch_set* set, const pfx_shape* s, int64_t n);
int launch_fill_nan(pfx_ctx* ctx, double* feat, int64_t n, int64_t ld, int ncol, cudaStream_t st);

}  // namespace pfx
