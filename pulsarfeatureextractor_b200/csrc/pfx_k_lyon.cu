// pfx_k_lyon.cu -- Lyon statistics (PHCXFile.py:320-379), 8 lanes per candidate.
#include "pfx_internal.cuh"

using namespace pfx;

namespace {
// ---- Lyon statistics of uint8 rows: exact integer power sums --------------------------------
// For bytes the four raw power sums S1 = sum x, S2 = sum x^2, S3 = sum x^3, S4 = sum x^4 are exact
// integers (S4 <= 65535 * 255^4 < 2^48), and so are the central sums scaled by powers of n:
//     N2 = n S2 - S1^2                                   = n^2 m2
//     N3 = n^2 S3 - 3 n S1 S2 + 2 S1^3                   = n^3 m3
//     N4 = n^3 S4 - 4 n^2 S1 S3 + 6 n S1^2 S2 - 3 S1^4   = n^4 m4
// (m_k = biased central moments, scipy.stats semantics).  They are formed once per row in 128-bit
// integer arithmetic, so nothing cancels in floating point and the FP64 pipe -- B200 issues one
// FP64 warp instruction every other cycle -- is out of the per-byte loop: round 1 spent four FP64
// instructions per input byte on the third and fourth central sums and ran at 40 % of HBM
// bandwidth, issue and FP64 bound.  Per byte now: one byte extract, x^2, S3 += x^2 x (IMAD) and
// S4 += x^2 x^2 (IMAD.WIDE); S1 and S2 come from dp4a on whole words.
typedef unsigned long long u64_t;
typedef unsigned __int128 u128_t;
typedef __int128 i128_t;

__device__ __forceinline__ void pow_word(unsigned w, unsigned& s3, u64_t& s4) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const unsigned x = __byte_perm(w, 0, 0x4440 + b);
        const unsigned x2 = x * x;
        s3 += x2 * x;                       // <= 128 bytes per lane: 128 * 255^3 < 2^32
        s4 += (u64_t)x2 * (u64_t)x2;
    }
}

__device__ __forceinline__ double i128_to_double(i128_t v) {
    const bool neg = v < 0;
    const u128_t a = neg ? (u128_t)(-v) : (u128_t)v;
    const double d = (double)(u64_t)(a >> 64) * 18446744073709551616.0 + (double)(u64_t)a;
    return neg ? -d : d;
}

// mean, std, skew, kurtosis of one row from its exact power sums.  The sums are first re-centred on
// the integer c nearest to the mean, in 64-bit integer arithmetic (exact):
//     T_k = sum (x - c)^k,   e = S1 - n c  (|e| <= n/2, so delta = mean - c = e/n, |delta| <= 1/2)
// and the central sums follow from the shifted-data identities
//     n m2 = T2 - e^2/n,   n m3 = T3 - 3 e T2/n + 2 e^3/n^2,   n m4 = T4 - 4 e T3/n + 6 e^2 T2/n^2 - 3 e^4/n^3
// whose terms are exact doubles (integer numerators below 2^53) and whose corrections are small
// against the leading term because |delta| <= 1/2: no cancellation, no 128-bit arithmetic.
// in1 = 1/n (exact when n is a power of two).
__device__ __forceinline__ void lyon_from_sums(long long n, long long S1, long long S2, long long S3, long long S4,
                                               double in1, double& mean, double& sd, double& skew, double& kurt) {
    const long long c = (2 * S1 + n) / (2 * n);
    const long long e = S1 - n * c;
    const long long c2 = c * c;
    const long long T2 = S2 - c * (2 * S1 - n * c);
    const long long T3 = S3 - 3 * c * S2 + 3 * c2 * S1 - n * c2 * c;
    const long long T4 = S4 - 4 * c * S3 + 6 * c2 * S2 - 4 * c2 * c * S1 + n * c2 * c2;
    const long long e2 = e * e;
    const double in2 = in1 * in1;
    mean = (double)S1 * in1;
    const double m2 = (double)(n * S2 - S1 * S1) * in2;                 // exact integer numerator
    const double M3 = ((double)T3 - (double)(3 * e * T2) * in1) + (double)(2 * e2 * e) * in2;
    const double M4 = (((double)T4 - (double)(4 * e * T3) * in1) + (double)(6 * e2 * T2) * in2) -
                      (double)(3 * e2 * e2) * (in2 * in1);
    const double m3 = M3 * in1, m4 = M4 * in1;
    sd = sqrt(m2);
    const double thr = (LM_EPSMCH * mean) * (LM_EPSMCH * mean);
    const bool degenerate = m2 <= thr;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    skew = degenerate ? qnan : m3 / (m2 * sd);
    kurt = degenerate ? qnan : m4 / (m2 * m2) - 3.0;
}

// The same for n = 2^L <= 1024 with the operand widths spelled out (S1 < 2^18, S2 < 2^26, c <= 255,
// |e| <= 512): every product is one or two integer multiply-adds and every multiplication or
// division by n is a shift.
__device__ __forceinline__ void lyon_from_sums_pow2(int L, int S1, int S2, long long S3, long long S4,
                                                    double& mean, double& sd, double& skew, double& kurt) {
    const int c = (2 * S1 + (1 << L)) >> (L + 1);
    const int e = S1 - (c << L);
    const int c2 = c * c, c3 = c2 * c;
    const unsigned c4 = (unsigned)c2 * (unsigned)c2;
    const long long T2 = (long long)S2 - (long long)c * (long long)(2 * S1 - (c << L));
    const long long T3 = S3 - (long long)(3 * c) * (long long)S2 + (long long)(3 * c2) * (long long)S1 -
                         ((long long)c3 << L);
    const long long T4 = S4 - (long long)(4 * c) * S3 + (long long)(6 * c2) * (long long)S2 -
                         (long long)(4 * c3) * (long long)S1 + (long long)((unsigned long long)c4 << L);
    const int e2 = e * e;
    const double in1 = __hiloint2double((1023 - L) << 20, 0);           // 2^-L
    const double in2 = in1 * in1;
    mean = (double)S1 * in1;
    const double m2 = (double)(((long long)S2 << L) - (long long)S1 * (long long)S1) * in2;
    const double M3 = ((double)T3 - (double)((long long)(3 * e) * T2) * in1) + (double)(2 * e2 * e) * in2;
    const double M4 = (((double)T4 - (double)((long long)(4 * e) * T3) * in1) + (double)((long long)(6 * e2) * T2) * in2) -
                      (double)(3 * (long long)e2 * (long long)e2) * (in2 * in1);
    const double m3 = M3 * in1, m4 = M4 * in1;
    sd = sqrt(m2);
    const double thr = (LM_EPSMCH * mean) * (LM_EPSMCH * mean);
    const bool degenerate = m2 <= thr;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    skew = degenerate ? qnan : m3 / (m2 * sd);
    kurt = degenerate ? qnan : m4 / (m2 * m2) - 3.0;
}

// generic uint8 row whose length is a multiple of 16 and at most 65535: 8 lanes, 16-byte loads
__device__ __forceinline__ void lyon_u8_row(const Group<LYON_W>& g, const uint4* __restrict__ row,
                                            int nvec, int len, double* out) {
    unsigned us1 = 0, us2 = 0;
    u64_t s3 = 0, s4 = 0;
    for (int v = g.lane(); v < nvec; v += LYON_W) {
        const uint4 q = __ldg(row + v);
        const unsigned wds[4] = {q.x, q.y, q.z, q.w};
        unsigned t3 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            us1 = __dp4a(wds[k], 0x01010101u, us1);
            us2 = __dp4a(wds[k], wds[k], us2);
            pow_word(wds[k], t3, s4);
        }
        s3 += t3;
    }
    const u64_t S1 = (u64_t)g.sum((long long)us1), S2 = (u64_t)g.sum((long long)us2);
    s3 = (u64_t)g.sum((long long)s3);
    s4 = (u64_t)g.sum((long long)s4);
    if (g.leader()) {
        double mean, sd, skew, kurt;
        lyon_from_sums((long long)len, (long long)S1, (long long)S2, (long long)s3, (long long)s4,
                       1.0 / (double)len, mean, sd, skew, kurt);
        out[0] = mean;
        out[1] = sd;
        out[2] = skew;
        out[3] = kurt;
    }
}

// butterflies over the 8 lanes of a candidate with the full-warp mask: all 32 lanes run the same
// code (the loop below is warp uniform), and a compile-time mask spares the collective
// synchronisation the compiler wraps around shuffles with a run-time mask
__device__ __forceinline__ unsigned bfly8(unsigned v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}
__device__ __forceinline__ u64_t bfly8(u64_t v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// ---- fast path: uint8 rows held in registers ---------------------------------------------
// PV / DV = 16-byte vectors per lane for the profile / DM row (row length = 128 * PV resp.
// 128 * DV bytes with 8 lanes per candidate).  A warp takes 16 consecutive candidates per pass, in
// four quarters of four candidates: per quarter all loads are issued up front, the power sums are
// integer (see above) and closed by butterflies over the 8 lanes of a candidate, and the totals of
// the eight rows of the quarter are handed to the lanes 8t .. 8t+7.  After the four quarters every
// lane owns one of the 32 rows (lane l: candidate l/2, row l&1) and finishes it alone: the closing
// arithmetic (re-centring, two divisions, a square root -- about as many instructions as the byte
// loop of a quarter) runs once per 16 candidates instead of once per 4, and a lane stores its four
// statistics as 32 contiguous bytes.  The kernel is bound by instruction issue, not by a pipe:
// measured at one instruction per cycle and scheduler, so instructions per candidate is the figure
// of merit (round 1: 335, now about 130).  The pass loop is uniform per block, which spares the
// collective synchronisation the compiler wraps around shuffles in a loop with a per-thread bound.
template <int PV, int DV>
__global__ void __launch_bounds__(LYON_THREADS)
k_lyon8_u8(const uint4* __restrict__ profile, const uint4* __restrict__ dmstat, long long n, double* feat,
           long long ld) {
    const int l32 = threadIdx.x & 31;
    const int lane = l32 & (LYON_W - 1);
    constexpr int WPB = LYON_THREADS / 32;                 // warps per block
    constexpr long long CPW = 16;                          // candidates per warp pass
    const long long per_pass = (long long)gridDim.x * WPB * CPW;
    const long long passes = (n + per_pass - 1) / per_pass;     // the same for every thread of the grid
    const bool third32 = (PV <= 2) && (DV <= 2);   // 256 * 255^3 < 2^32: third power sums of rows <= 256 bytes fit 32 bits
    for (long long it = 0; it < passes; ++it) {
        const long long cbase = it * per_pass + ((long long)blockIdx.x * WPB + (threadIdx.x >> 5)) * CPW;
        unsigned r1 = 0, r2 = 0;
        u64_t r3 = 0, r4 = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const long long cw = cbase + 4 * t + (l32 >> 3);
            const long long c = cw < n ? cw : n - 1;           // clamped for the loads, masked at the store
            uint4 vp[PV], vd[DV];
#pragma unroll
            for (int k = 0; k < PV; ++k) vp[k] = __ldcs(profile + c * (8 * PV) + k * 8 + lane);
#pragma unroll
            for (int k = 0; k < DV; ++k) vd[k] = __ldcs(dmstat + c * (8 * DV) + k * 8 + lane);
            unsigned a1 = 0, a2 = 0, a3 = 0, b1 = 0, b2 = 0, b3 = 0;
            u64_t a4 = 0, b4 = 0;
#pragma unroll
            for (int k = 0; k < PV; ++k) {
                const unsigned w[4] = {vp[k].x, vp[k].y, vp[k].z, vp[k].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    a1 = __dp4a(w[j], 0x01010101u, a1);
                    a2 = __dp4a(w[j], w[j], a2);
                    pow_word(w[j], a3, a4);
                }
            }
#pragma unroll
            for (int k = 0; k < DV; ++k) {
                const unsigned w[4] = {vd[k].x, vd[k].y, vd[k].z, vd[k].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    b1 = __dp4a(w[j], 0x01010101u, b1);
                    b2 = __dp4a(w[j], w[j], b2);
                    pow_word(w[j], b3, b4);
                }
            }
            const unsigned ps1 = bfly8(a1), ps2 = bfly8(a2), ds1 = bfly8(b1), ds2 = bfly8(b2);
            const u64_t ps3 = third32 ? (u64_t)bfly8(a3) : bfly8((u64_t)a3);
            const u64_t ds3 = third32 ? (u64_t)bfly8(b3) : bfly8((u64_t)b3);
            const u64_t ps4 = bfly8(a4), ds4 = bfly8(b4);
            // hand the totals of candidate g of this quarter (held by lanes 8g .. 8g+7) to the owners of
            // its two rows, lanes 8t + 2g (profile) and 8t + 2g + 1 (DM row)
            // (the source lane picks what it sends, so the DM rows are fetched from an odd lane of
            // the group, which offers the DM totals, and the profile rows from an even one)
            const int src = ((l32 & 7) >> 1) * 8 + (l32 & 1);
            const bool isd = (l32 & 1) != 0;
            const unsigned g1 = __shfl_sync(0xffffffffu, isd ? ds1 : ps1, src);
            const unsigned g2 = __shfl_sync(0xffffffffu, isd ? ds2 : ps2, src);
            const u64_t g3 = third32 ? (u64_t)__shfl_sync(0xffffffffu, (unsigned)(isd ? ds3 : ps3), src)
                                     : __shfl_sync(0xffffffffu, isd ? ds3 : ps3, src);
            const u64_t g4 = __shfl_sync(0xffffffffu, isd ? ds4 : ps4, src);
            if ((l32 >> 3) == t) {
                r1 = g1;
                r2 = g2;
                r3 = g3;
                r4 = g4;
            }
        }
        // lane l: candidate cbase + l/2, row l&1
        constexpr int LP = (PV == 1) ? 7 : (PV == 2) ? 8 : (PV == 4) ? 9 : 10;
        constexpr int LD = (DV == 1) ? 7 : (DV == 2) ? 8 : (DV == 4) ? 9 : 10;
        static_assert((128 * PV) == (1 << LP) && (128 * DV) == (1 << LD), "row lengths must be powers of two");
        double mean, sd, skew, kurt;
        lyon_from_sums_pow2((l32 & 1) ? LD : LP, (int)r1, (int)r2, (long long)r3, (long long)r4, mean, sd, skew, kurt);
        const long long cmine = cbase + (l32 >> 1);
        if (cmine < n) {
            double* o = feat + cmine * ld + 4 * (l32 & 1);
            if ((ld & 1) == 0) {          // 16-byte aligned rows: two vector stores
                __stcs((double2*)o, make_double2(mean, sd));
                __stcs((double2*)o + 1, make_double2(skew, kurt));
            } else {
                __stcs(o, mean);
                __stcs(o + 1, sd);
                __stcs(o + 2, skew);
                __stcs(o + 3, kurt);
            }
        }
    }
}

__global__ void __launch_bounds__(LYON_THREADS)
k_lyon8(const void* __restrict__ profile, int dtype, int nbin, const void* __restrict__ dmstat,
        int ds_dtype, int n_ds, long long n, double* feat, long long ld) {
    Group<LYON_W> g;
    const long long gpb = LYON_THREADS / LYON_W;
    const long long stride = (long long)gridDim.x * gpb;
    const bool fast_p = (dtype == PFX_U8) && (nbin % 16 == 0);
    const bool fast_d = (ds_dtype == PFX_U8) && (n_ds % 16 == 0);
    for (long long c = (long long)blockIdx.x * gpb + threadIdx.x / LYON_W; c < n; c += stride) {
        double* out = feat + c * ld;
        if (fast_p)
            lyon_u8_row(g, (const uint4*)((const unsigned char*)profile + c * nbin), nbin / 16, nbin, out);
        else
            lyon4_row(g, profile, dtype, c, nbin, out);
        if (fast_d)
            lyon_u8_row(g, (const uint4*)((const unsigned char*)dmstat + c * n_ds), n_ds / 16, n_ds, out + 4);
        else
            lyon4_row(g, dmstat, ds_dtype, c, n_ds, out + 4);
    }
}

__global__ void k_fill_nan(double* feat, long long n, long long ld, int ncol) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n * ncol) feat[(t / ncol) * ld + (t % ncol)] = __longlong_as_double(0x7ff8000000000000LL);
}

}  // namespace

int pfx::launch_lyon(pfx_ctx* ctx, const pfx_shape* s, int64_t n, const void* profile,
                       const void* dmstat, double* feat, int64_t ld, cudaStream_t st) {
    if (s->nbin < 1 || s->n_dmstat < 1) return set_err(ctx, PFX_ERR_UNSUPPORTED, "empty Lyon rows");
    const long long cap = (long long)ctx->sm_count * 32;
    const bool u8 = s->dtype == PFX_U8 && s->dmstat_dtype == PFX_U8 &&
                    ((uintptr_t)profile % 16 == 0) && ((uintptr_t)dmstat % 16 == 0) && ((uintptr_t)feat % 16 == 0);
    // fast kernels: a warp takes 16 candidates per pass
    const long long cpb_fast = (LYON_THREADS / 32) * 16;
    long long fblocks = (n + cpb_fast - 1) / cpb_fast;
    if (fblocks > cap) fblocks = cap;
    if (fblocks < 1) fblocks = 1;
    if (u8 && s->nbin == 128 && s->n_dmstat == 256)
        k_lyon8_u8<1, 2><<<(int)fblocks, LYON_THREADS, 0, st>>>((const uint4*)profile, (const uint4*)dmstat, n, feat, ld);
    else if (u8 && s->nbin == 128 && s->n_dmstat == 128)
        k_lyon8_u8<1, 1><<<(int)fblocks, LYON_THREADS, 0, st>>>((const uint4*)profile, (const uint4*)dmstat, n, feat, ld);
    else if (u8 && s->nbin == 256 && s->n_dmstat == 256)
        k_lyon8_u8<2, 2><<<(int)fblocks, LYON_THREADS, 0, st>>>((const uint4*)profile, (const uint4*)dmstat, n, feat, ld);
    else if (u8 && s->nbin == 1024 && s->n_dmstat == 256)
        k_lyon8_u8<8, 2><<<(int)fblocks, LYON_THREADS, 0, st>>>((const uint4*)profile, (const uint4*)dmstat, n, feat, ld);
    else {
        const long long gpb = LYON_THREADS / LYON_W;
        long long blocks = (n + gpb - 1) / gpb;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        k_lyon8<<<(int)blocks, LYON_THREADS, 0, st>>>(profile, s->dtype, s->nbin, dmstat, s->dmstat_dtype,
                                                       s->n_dmstat, n, feat, ld);
    }
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}


int pfx::launch_fill_nan(pfx_ctx* ctx, double* feat, int64_t n, int64_t ld, int ncol, cudaStream_t st) {
    const long long tot = (long long)n * ncol;
    k_fill_nan<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(feat, n, ld, ncol);
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}
