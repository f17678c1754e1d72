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

// mean, std, skew, kurtosis of one row from the scaled central sums N2, N3 (exact int64) and N4
// (exact 128-bit, sign + magnitude as two 64-bit halves); in1 = 1/n
__device__ __forceinline__ void lyon_finish_scaled(u64_t S1, u64_t N2, long long N3, u64_t n4_hi, u64_t n4_lo,
                                                   double in1, double& mean, double& sd, double& skew, double& kurt) {
    const double in2 = in1 * in1;
    mean = (double)S1 * in1;
    const double m2 = (double)N2 * in2;
    const double m3 = (double)N3 * (in2 * in1);
    const double m4 = ((double)n4_hi * 18446744073709551616.0 + (double)n4_lo) * (in2 * in2);   // N4 >= 0
    sd = sqrt(m2);
    const double thr = (LM_EPSMCH * mean) * (LM_EPSMCH * mean);
    const bool degenerate = m2 <= thr;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    skew = degenerate ? qnan : m3 / (m2 * sd);
    kurt = degenerate ? qnan : m4 / (m2 * m2) - 3.0;
}

// n = 2^L <= 1024: every multiplication by a power of n is a shift and every product but S1^4 fits
// 64 bits (S1 < 2^18, S2 < 2^26, S3 < 2^34, S4 < 2^42)
__device__ __forceinline__ void lyon_from_sums_pow2(int L, unsigned S1, unsigned S2, u64_t S3, u64_t S4,
                                                    double& mean, double& sd, double& skew, double& kurt) {
    const u64_t s1sq = (u64_t)S1 * S1;                                  // < 2^36
    const u64_t N2 = ((u64_t)S2 << L) - s1sq;
    const u64_t s1s2 = (u64_t)S1 * S2;                                  // < 2^44
    const long long N3 = (long long)(S3 << (2 * L)) - (long long)((3 * s1s2) << L) + (long long)(2 * s1sq * S1);
    // N4 = (S4 << 3L) - ((4 S1 S3) << 2L) + ((6 S1^2 S2) << L) - 3 S1^4 in 128 bits (it is a sum of
    // fourth powers, so the result is non-negative)
    const u64_t t1 = 4 * (u64_t)S1 * S3;                                // < 2^54
    const u64_t t2 = 6 * s1sq * S2;                                     // < 2^65 ?  6 * 2^36 * 2^26: kept in 128 bits
    (void)t2;
    u128_t acc = ((u128_t)S4 << (3 * L)) + (((u128_t)(3 * s1sq) * (u64_t)(2 * S2)) << L);
    acc -= ((u128_t)t1 << (2 * L));
    acc -= (u128_t)(3 * s1sq) * s1sq;
    lyon_finish_scaled(S1, N2, N3, (u64_t)(acc >> 64), (u64_t)acc, 1.0 / (double)(1u << L), mean, sd, skew, kurt);
}

// any n <= 65535: 128-bit products throughout
__device__ __forceinline__ void lyon_from_sums(u64_t n, u64_t S1, u64_t S2, u64_t S3, u64_t S4,
                                               double& mean, double& sd, double& skew, double& kurt) {
    const u64_t n2 = n * n;
    const u64_t N2 = n * S2 - S1 * S1;                                   // < 2^53
    const u128_t s1sq = (u128_t)S1 * S1;
    const i128_t N3 = (i128_t)((u128_t)n2 * S3) - (i128_t)((u128_t)(3 * n * S1) * S2) + (i128_t)(2 * s1sq * S1);
    const u128_t N4 = (u128_t)(n2 * n) * S4 + 6 * n * s1sq * S2 - (u128_t)(4 * n2 * S1) * S3 - 3 * s1sq * s1sq;
    // |N3| < 2^63 for n <= 4096; beyond that its low bits are rounded like any double
    const bool neg = N3 < 0;
    const u128_t a3 = neg ? (u128_t)(-N3) : (u128_t)N3;
    const double d3 = (double)(u64_t)(a3 >> 64) * 18446744073709551616.0 + (double)(u64_t)a3;
    const double in1 = 1.0 / (double)n, in2 = in1 * in1;
    mean = (double)S1 * in1;
    const double m2 = (double)N2 * in2;
    const double m3 = (neg ? -d3 : d3) * (in2 * in1);
    const double m4 = ((double)(u64_t)(N4 >> 64) * 18446744073709551616.0 + (double)(u64_t)N4) * (in2 * in2);
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
        lyon_from_sums((u64_t)len, S1, S2, s3, s4, mean, sd, skew, kurt);
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
// 128 * DV bytes with 8 lanes per candidate).  All loads of a candidate are issued up front, the
// power sums are integer (see above), and the eight statistics are finished by all eight lanes at
// once (lanes 0-3 own the profile, 4-7 the DM row; lane k writes statistic k & 3), so a warp writes
// 4 x 64 contiguous bytes.
template <int PV, int DV>
__global__ void __launch_bounds__(LYON_THREADS)
k_lyon8_u8(const uint4* __restrict__ profile, const uint4* __restrict__ dmstat, long long n, double* feat,
           long long ld) {
    const int lane = threadIdx.x & (LYON_W - 1);
    const long long gpb = LYON_THREADS / LYON_W;
    const long long stride = (long long)gridDim.x * gpb;
    // warp-uniform loop: the four candidates of a warp are consecutive; candidates past the end
    // are clamped for the loads and masked for the store
    for (long long c0 = (long long)blockIdx.x * gpb + (threadIdx.x / 32) * 4; c0 < n; c0 += stride) {
        const long long cw = c0 + ((threadIdx.x & 31) / LYON_W);
        const bool live = cw < n;
        const long long c = live ? cw : n - 1;
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
        // lanes 0-3 finish the profile, lanes 4-7 the DM row: exchange so that after the butterfly
        // every lane holds the sums of the row it finishes.  (The butterfly adds all eight lanes, so
        // both rows are reduced in full; the select happens afterwards.)
        const unsigned ps1 = bfly8(a1), ps2 = bfly8(a2), ds1 = bfly8(b1), ds2 = bfly8(b2);
        const u64_t ps3 = bfly8((u64_t)a3), ds3 = bfly8((u64_t)b3);
        const u64_t ps4 = bfly8(a4), ds4 = bfly8(b4);
        const bool isd = (lane & 4) != 0;
        double mean, sd, skew, kurt;
        constexpr int LP = (PV == 1) ? 7 : (PV == 2) ? 8 : (PV == 4) ? 9 : 10;
        constexpr int LD = (DV == 1) ? 7 : (DV == 2) ? 8 : (DV == 4) ? 9 : 10;
        static_assert((128 * PV) == (1 << LP) && (128 * DV) == (1 << LD), "row lengths must be powers of two");
        lyon_from_sums_pow2(isd ? LD : LP, isd ? ds1 : ps1, isd ? ds2 : ps2, isd ? ds3 : ps3, isd ? ds4 : ps4,
                            mean, sd, skew, kurt);
        const int k = lane & 3;
        const double v = (k == 0) ? mean : (k == 1) ? sd : (k == 2) ? skew : kurt;
        if (live) __stcs(feat + c * ld + lane, v);
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
    const long long gpb = LYON_THREADS / LYON_W;
    long long blocks = (n + gpb - 1) / gpb;
    const long long cap = (long long)ctx->sm_count * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const bool u8 = s->dtype == PFX_U8 && s->dmstat_dtype == PFX_U8 &&
                    ((uintptr_t)profile % 16 == 0) && ((uintptr_t)dmstat % 16 == 0);
    if (u8 && s->nbin == 128 && s->n_dmstat == 256)
        k_lyon8_u8<1, 2><<<(int)blocks, LYON_THREADS, 0, st>>>((const uint4*)profile, (const uint4*)dmstat, n, feat, ld);
    else if (u8 && s->nbin == 128 && s->n_dmstat == 128)
        k_lyon8_u8<1, 1><<<(int)blocks, LYON_THREADS, 0, st>>>((const uint4*)profile, (const uint4*)dmstat, n, feat, ld);
    else if (u8 && s->nbin == 256 && s->n_dmstat == 256)
        k_lyon8_u8<2, 2><<<(int)blocks, LYON_THREADS, 0, st>>>((const uint4*)profile, (const uint4*)dmstat, n, feat, ld);
    else if (u8 && s->nbin == 64 * 16 && s->n_dmstat == 256)
        k_lyon8_u8<8, 2><<<(int)blocks, LYON_THREADS, 0, st>>>((const uint4*)profile, (const uint4*)dmstat, n, feat, ld);
    else
        k_lyon8<<<(int)blocks, LYON_THREADS, 0, st>>>(profile, s->dtype, s->nbin, dmstat, s->dmstat_dtype,
                                                       s->n_dmstat, n, feat, ld);
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}


int pfx::launch_fill_nan(pfx_ctx* ctx, double* feat, int64_t n, int64_t ld, int ncol, cudaStream_t st) {
    const long long tot = (long long)n * ncol;
    k_fill_nan<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(feat, n, ld, ncol);
    PFX_CUDA(ctx, cudaGetLastError());
    return PFX_OK;
}
