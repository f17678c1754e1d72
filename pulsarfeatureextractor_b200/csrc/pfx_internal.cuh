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
constexpr int COUNTER_SLOTS = 64;             // ring of work-queue counter sets
constexpr int COUNTERS_PER_SLOT = 8;
constexpr int N_EVENTS = 8;

typedef unsigned long long u64;

}  // namespace pfx

struct pfx_ctx {
    int device;
    int sm_count;
    int smem_optin;
    char err[512];
    pfx::u64* counters;       // COUNTER_SLOTS * COUNTERS_PER_SLOT
    int counter_slot;
    double* hist_scratch;
    size_t hist_scratch_bytes;
    void* carry;              // per-candidate records passed between the two Gaussian kernels
    size_t carry_bytes;
    void* lane_buf[4];        // per-lane data vectors of the lane-per-candidate kernels
    size_t lane_bytes[4];
    void* hist_lane_buf[2];   // per-lane arrays of k_hist_lane
    size_t hist_lane_bytes[2];
    int* ovf_list[2];         // candidates whose histogram overflowed the per-lane capacity
    size_t ovf_bytes[2];
    cudaEvent_t ev[pfx::N_EVENTS];
    // fork/join: the independent score groups of one call run on a side stream
    cudaStream_t side[2], side2[2];
    cudaEvent_t fork_ev[2], join_ev[2], join2_ev[2];
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

// k-th counter of the current slot; slots rotate per call so that calls in flight on different
// streams never share a counter
inline u64* take_counter(pfx_ctx* ctx, int k, cudaStream_t st) {
    u64* c = ctx->counters + (size_t)ctx->counter_slot * COUNTERS_PER_SLOT + k;
    cudaMemsetAsync(c, 0, sizeof(u64), st);
    return c;
}

// persistent launch: grid = resident blocks per SM x SM count
template <class K>
inline int persistent_grid(pfx_ctx* ctx, K kernel, int threads, size_t smem, int* grid) {
    PFX_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PFX_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1)
        return set_err(ctx, PFX_ERR_UNSUPPORTED,
                       "kernel needs %zu bytes of shared memory per block, limit %d", smem, ctx->smem_optin);
    *grid = per_sm * ctx->sm_count;
    return PFX_OK;
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
                    int64_t ld, int8_t* ier, cudaStream_t st, int ck);
int launch_hist(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* feat,
                int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, int ck);
int launch_gauss(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* feat,
                 int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, int ck);
int launch_dmfit(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* dmcurve, const double* params,
                 double* feat, int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, int ck);
int launch_subband(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, const void* subbands,
                   const double* params, double* feat, int64_t ld, int32_t* status, cudaStream_t st, int ck);
int launch_lyon(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, const void* dmstat,
                double* feat, int64_t ld, cudaStream_t st);
int launch_pfd_params(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* params,
                      cudaStream_t st);
int launch_fill_nan(pfx_ctx* ctx, double* feat, int64_t n, int64_t ld, int ncol, cudaStream_t st);

}  // namespace pfx
