// pfx_k_lyon.cu -- Lyon statistics (PHCXFile.py:320-379), 8 lanes per candidate.
#include "pfx_internal.cuh"

using namespace pfx;

namespace {
// ---- Lyon statistics -----------------------------------------------------------------------
// uint8 rows whose length is a multiple of 16: 8 lanes per candidate, 16-byte loads, the first
// moment in exact integer arithmetic (dp4a), the central sums on D = n*x - sum(x) (an exact
// integer, turned into a double with one add) so that no int->double conversion instruction
// and no rounding of the mean enter: m_k = sum(D^k) / n^(k+1).
__device__ __forceinline__ void lyon_u8_row(const Group<LYON_W>& g, const uint4* __restrict__ row,
                                            int nvec, int len, double* out) {
    unsigned us1 = 0;
    for (int v = g.lane(); v < nvec; v += LYON_W) {
        const uint4 q = __ldg(row + v);
        us1 = __dp4a(q.x, 0x01010101u, us1);
        us1 = __dp4a(q.y, 0x01010101u, us1);
        us1 = __dp4a(q.z, 0x01010101u, us1);
        us1 = __dp4a(q.w, 0x01010101u, us1);
    }
    const int s1 = g.sum((int)us1);
    long long s2 = 0;
    double s3 = 0.0, s4 = 0.0;
    const double magic = 4503601774854144.0;   // 2^52 + 2^31
    for (int v = g.lane(); v < nvec; v += LYON_W) {
        const uint4 q = __ldg(row + v);
        const unsigned wds[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int x = (int)((wds[k] >> (8 * b)) & 0xffu);
                const int D = len * x - s1;
                s2 += (long long)D * (long long)D;
                const double d = __hiloint2double(0x43300000, D ^ 0x80000000) - magic;
                const double d2 = d * d;
                s3 = fma(d2, d, s3);
                s4 = fma(d2, d2, s4);
            }
        }
    }
    s2 = g.sum(s2);
    s3 = g.sum(s3);
    s4 = g.sum(s4);
    if (g.leader()) {
        const double n = (double)len;
        const double n2 = n * n;
        lyon_finish(n, (double)s1 / n, (double)s2 / n2, s3 / (n2 * n), s4 / (n2 * n2), out);
    }
}

// ---- fast path: uint8 rows held in registers ---------------------------------------------
// PV / DV = 16-byte vectors per lane for the profile / DM row (row length = 128 * PV resp.
// 128 * DV bytes with 8 lanes per candidate).  All loads of a candidate are issued up front,
// the raw first and second moments are exact integer dot products (dp4a), the third and fourth
// central sums run on D = n*x - S1 held exactly in a double, and the eight statistics are
// finished by all eight lanes at once (lanes 0-3 own the profile, 4-7 the DM row; lane k writes
// statistic k & 3), so a warp writes 4 x 64 contiguous bytes.
//
// kbias = 2^31 - S1 (mod 2^32): n*x + kbias is D = n*x - S1 with the sign bit flipped, which is
// the low word of the double 2^52 + 2^31 + D; subtracting the magic constant leaves D exactly
// (no int->double conversion instruction, no rounding of the mean).
__device__ __forceinline__ void lyon_accum(unsigned w, unsigned n, unsigned kbias, double& s3, double& s4) {
    const double magic = 4503601774854144.0;   // 2^52 + 2^31
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const unsigned x = __byte_perm(w, 0, 0x4440 + b);
        const double d = __hiloint2double(0x43300000, (int)(n * x + kbias)) - magic;
        const double d2 = d * d;
        s3 = fma(d2, d, s3);
        s4 = fma(d2, d2, s4);
    }
}

// butterflies over the 8 lanes of a candidate with the full-warp mask: all 32 lanes run the same
// code (the loop below is warp uniform), and a compile-time mask spares the collective
// synchronisation the compiler wraps around shuffles with a run-time mask
__device__ __forceinline__ int bfly8(int v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}
__device__ __forceinline__ double bfly8(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

template <int PV, int DV>
__global__ void __launch_bounds__(LYON_THREADS)
k_lyon8_u8(const uint4* __restrict__ profile, const uint4* __restrict__ dmstat, long long n, double* feat,
           long long ld) {
    const int lane = threadIdx.x & (LYON_W - 1);
    const long long gpb = LYON_THREADS / LYON_W;
    const long long stride = (long long)gridDim.x * gpb;
    constexpr unsigned NP = 128 * PV, ND = 128 * DV;
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
        unsigned a1 = 0, a2 = 0, b1 = 0, b2 = 0;
#pragma unroll
        for (int k = 0; k < PV; ++k) {
            const unsigned w[4] = {vp[k].x, vp[k].y, vp[k].z, vp[k].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a1 = __dp4a(w[j], 0x01010101u, a1);
                a2 = __dp4a(w[j], w[j], a2);
            }
        }
#pragma unroll
        for (int k = 0; k < DV; ++k) {
            const unsigned w[4] = {vd[k].x, vd[k].y, vd[k].z, vd[k].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                b1 = __dp4a(w[j], 0x01010101u, b1);
                b2 = __dp4a(w[j], w[j], b2);
            }
        }
        const int ps1 = bfly8((int)a1), ps2 = bfly8((int)a2);
        const int ds1 = bfly8((int)b1), ds2 = bfly8((int)b2);
        double p3 = 0.0, p4 = 0.0, d3 = 0.0, d4 = 0.0;
        const unsigned pk = 0x80000000u - (unsigned)ps1, dk = 0x80000000u - (unsigned)ds1;
#if defined(PFX_LYON_SPLIT)
        // two independent accumulator pairs per row (shorter dependency chains)
        double p3b = 0.0, p4b = 0.0, d3b = 0.0, d4b = 0.0;
#pragma unroll
        for (int k = 0; k < PV; ++k) {
            lyon_accum(vp[k].x, NP, pk, p3, p4);
            lyon_accum(vp[k].y, NP, pk, p3b, p4b);
            lyon_accum(vp[k].z, NP, pk, p3, p4);
            lyon_accum(vp[k].w, NP, pk, p3b, p4b);
        }
#pragma unroll
        for (int k = 0; k < DV; ++k) {
            lyon_accum(vd[k].x, ND, dk, d3, d4);
            lyon_accum(vd[k].y, ND, dk, d3b, d4b);
            lyon_accum(vd[k].z, ND, dk, d3, d4);
            lyon_accum(vd[k].w, ND, dk, d3b, d4b);
        }
        p3 += p3b; p4 += p4b; d3 += d3b; d4 += d4b;
#else
#pragma unroll
        for (int k = 0; k < PV; ++k) {
            lyon_accum(vp[k].x, NP, pk, p3, p4);
            lyon_accum(vp[k].y, NP, pk, p3, p4);
            lyon_accum(vp[k].z, NP, pk, p3, p4);
            lyon_accum(vp[k].w, NP, pk, p3, p4);
        }
#pragma unroll
        for (int k = 0; k < DV; ++k) {
            lyon_accum(vd[k].x, ND, dk, d3, d4);
            lyon_accum(vd[k].y, ND, dk, d3, d4);
            lyon_accum(vd[k].z, ND, dk, d3, d4);
            lyon_accum(vd[k].w, ND, dk, d3, d4);
        }
#endif
        p3 = bfly8(p3);
        p4 = bfly8(p4);
        d3 = bfly8(d3);
        d4 = bfly8(d4);
        // lanes 0-3: profile, lanes 4-7: DM row; row lengths are powers of two here, so scaling by
        // 1/n^k is exact
        const bool isd = (lane & 4) != 0;
        const double in = isd ? 1.0 / (double)ND : 1.0 / (double)NP;
        const long long S1 = isd ? ds1 : ps1, S2 = isd ? ds2 : ps2;
        const double c3 = isd ? d3 : p3, c4 = isd ? d4 : p4;
        const long long nint = isd ? ND : NP;
        const double in2 = in * in;
        const double mean = (double)S1 * in;
        const double m2 = (double)(nint * S2 - S1 * S1) * in2;      // exact integer numerator
        const double m3 = c3 * (in2 * in2);                          // sum(D^3) / n^3 / n
        const double m4 = c4 * (in2 * in2 * in);                     // sum(D^4) / n^4 / n
        const double sd = sqrt(m2);
        const double thr = (LM_EPSMCH * mean) * (LM_EPSMCH * mean);
        const bool degenerate = m2 <= thr;
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        const double skew = degenerate ? qnan : m3 / (m2 * sd);
        const double kurt = degenerate ? qnan : m4 / (m2 * m2) - 3.0;
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
