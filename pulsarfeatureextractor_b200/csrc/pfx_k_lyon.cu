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
