// pfx_k_gauss.cu -- scores 8-11: Gaussian + background fit and the double-Gaussian procedure (peak
// stripping, eight 4-parameter rounds, one 8-parameter fit) on the rotated profile
// (ProfileOperations.py:724-768,1061-1483).  One warp per candidate, everything in shared memory.
// This is the most expensive kernel of the Bates path: 10 Levenberg-Marquardt solves per candidate.
#include "pfx_internal.cuh"

using namespace pfx;

namespace {

__global__ void __launch_bounds__(BLOCK_THREADS, 3)
k_gauss(const void* __restrict__ profile, int dtype, long long n, int m, double* feat, long long ld,
        int8_t* ier, int32_t* status, u64* counter, int work_doubles) {
    extern __shared__ double smem[];
    Group<WARP> g;
    double* ws = smem + (threadIdx.x / WARP) * work_doubles;
    for (;;) {
        const long long c = next_candidate(g, counter);
        if (c >= n) break;
        double* row = feat + c * ld;
        // the histogram kernel parked mu of the profile-histogram fit in the score-8 column
        const double ph_expect = row[7];
        const int prior = status[c];
        g.sync();
        if (g.leader()) row[7] = __longlong_as_double(0x7ff8000000000000LL);
        if (prior & (int)PFX_ST_FAIL_MASK) continue;   // the reference raised before reaching score 8
        unsigned st = gauss_candidate(g, ws, profile, dtype, c, m, ph_expect, row, ier ? ier + c * PFX_NIER : nullptr);
        if (st && g.leader()) atomicOr(status + c, (int)st);
    }
}

}  // namespace

int pfx::launch_gauss(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile, double* feat,
                      int64_t ld, int8_t* ier, int32_t* status, cudaStream_t st, int ck) {
    const int m = s->nbin;
    if (m < 8) return set_err(ctx, PFX_ERR_UNSUPPORTED, "nbin < 8");
    const int wd = align_up(gauss_work_doubles(m), 2);
    const size_t smem = (size_t)wd * 8 * GROUPS_PER_BLOCK;
    int grid;
    int rc = persistent_grid(ctx, k_gauss, BLOCK_THREADS, smem, &grid);
    if (rc) return rc;
    k_gauss<<<grid, BLOCK_THREADS, smem, st>>>(profile, s->dtype, n, m, feat, ld, ier, status,
                                                take_counter(ctx, ck, st), wd);
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}
