// pfx_group.cuh -- lane-group abstraction shared by every scoring routine.
//
// One candidate is scored by a *group* of W lanes.  Rows of every m-long vector (profile bins,
// histogram bins, DM trials) are striped over the lanes: lane l owns rows l, l+W, l+2W, ...
// Sums over rows are lane-partial sums combined by an xor-butterfly, so every lane of the
// group ends up with bit-identical values and all control flow that depends on them is
// group-uniform.
//
//   Group<32>, Group<16>, Group<8>   device: W lanes of one warp (shuffles, __syncwarp)
//   Group<1>                         one lane does everything; compiles for the host as well
//                                    and is what tests/hostsim builds to check the kernel logic
//                                    against the oracle on machines without a GPU.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PFX_HD __host__ __device__
#define PFX_D __device__
#define PFX_NOINLINE __noinline__
#else
#define PFX_HD
#define PFX_D
#define PFX_NOINLINE __attribute__((noinline))
#endif

#if defined(__CUDA_ARCH__)
// Un-fused arithmetic for the model expressions: numpy never contracts a*b+c, nvcc does.
#define PFX_MUL(a, b) __dmul_rn((a), (b))
#define PFX_ADD(a, b) __dadd_rn((a), (b))
#define PFX_SUB(a, b) __dsub_rn((a), (b))
#define PFX_DIV(a, b) __ddiv_rn((a), (b))
#else
// host build is compiled with -ffp-contract=off
#define PFX_MUL(a, b) ((a) * (b))
#define PFX_ADD(a, b) ((a) + (b))
#define PFX_SUB(a, b) ((a) - (b))
#define PFX_DIV(a, b) ((a) / (b))
#endif

// Inner products use fused multiply-add on the device.  The host logic check can be built with
// -DPFX_NO_FMA so that, with a one-lane group, sums run in MINPACK's own order and rounding.
#if defined(PFX_NO_FMA)
#define PFX_FMA(a, b, c) ((a) * (b) + (c))
#else
#define PFX_FMA(a, b, c) fma((a), (b), (c))
#endif

namespace pfx {

struct ValIdx {
    double v;
    int i;
};

// numpy.argmax ordering: a NaN counts as the maximum, ties and multiple NaNs go to the
// lowest index.
PFX_HD inline bool argmax_better(double av, int ai, double bv, int bi) {
    bool an = av != av, bn = bv != bv;
    if (an || bn) {
        if (an && bn) return ai < bi;
        return an;
    }
    if (av > bv) return true;
    if (av < bv) return false;
    return ai < bi;
}

template <int W_>
struct Group;

template <>
struct Group<1> {
    static constexpr int W = 1;
    static constexpr bool ordered = false;     // (one lane: sums run in row order anyway)
    PFX_HD Group() {}
    PFX_HD int lane() const { return 0; }
    PFX_HD bool leader() const { return true; }
    PFX_HD void sync() const {}
    PFX_HD double sum(double v) const { return v; }
    PFX_HD double max(double v) const { return v; }
    PFX_HD double min(double v) const { return v; }
    PFX_HD int sum(int v) const { return v; }
    PFX_HD int max(int v) const { return v; }
    PFX_HD int min(int v) const { return v; }
    PFX_HD long long sum(long long v) const { return v; }
    PFX_HD int any(int p) const { return p; }
    PFX_HD ValIdx argmax(ValIdx a) const { return a; }
    template <class T>
    PFX_HD T bcast(T v, int /*src*/) const { return v; }
};

#if defined(__CUDACC__)
template <int W_>
struct Group {
    static constexpr int W = W_;
    static constexpr bool ordered = false;
    unsigned mask;
    int ln;
    __device__ Group() {
        const int wl = threadIdx.x & 31;
        ln = wl & (W - 1);
        mask = (W == 32) ? 0xffffffffu : (((1u << W) - 1u) << (wl & ~(W - 1)));
    }
    __device__ int lane() const { return ln; }
    __device__ bool leader() const { return ln == 0; }
    __device__ void sync() const { __syncwarp(mask); }
    __device__ double sum(double v) const {
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
        return v;
    }
    __device__ double max(double v) const {   // fmax semantics (NaN ignored)
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o));
        return v;
    }
    __device__ double min(double v) const {
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(mask, v, o));
        return v;
    }
    __device__ int sum(int v) const {
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
        return v;
    }
    __device__ int max(int v) const {
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) v = ::max(v, __shfl_xor_sync(mask, v, o));
        return v;
    }
    __device__ int min(int v) const {
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) v = ::min(v, __shfl_xor_sync(mask, v, o));
        return v;
    }
    __device__ long long sum(long long v) const {
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
        return v;
    }
    __device__ int any(int p) const { return (__ballot_sync(mask, p) & mask) != 0; }
    __device__ ValIdx argmax(ValIdx a) const {
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) {
            ValIdx b;
            b.v = __shfl_xor_sync(mask, a.v, o);
            b.i = __shfl_xor_sync(mask, a.i, o);
            if (argmax_better(b.v, b.i, a.v, a.i)) a = b;
        }
        return a;
    }
    template <class T>
    __device__ T bcast(T v, int src) const { return __shfl_sync(mask, v, src, W); }
};
#endif

#if defined(__CUDACC__)
// A group whose sums over rows run IN ROW ORDER, un-fused, on the leader (the other lanes wait):
// element-wise work is still striped over the lanes, but every inner product rounds exactly like the
// sequential loops of MINPACK and numpy do.  With it the DM-curve fit, whose model needs only
// correctly rounded operations (+ - * / sqrt), reproduces scipy.optimize.leastsq BIT FOR BIT
// (pfx_k_dmfit_exact.cu; checked on the host with the one-lane build in tests/test_hostsim_logic.py).
template <int W_>
struct OrderedGroup : Group<W_> {
    static constexpr bool ordered = true;
};
#endif

// sum over rows lo..m-1 of term(i), in row order, by the leader; every lane gets the result
template <class G, class F>
PFX_HD inline double ordered_rows(const G& g, int lo, int m, F term) {
    g.sync();
    double s = 0.0;
    if (g.leader())
        for (int i = lo; i < m; ++i) s = PFX_ADD(s, term(i));
    return g.bcast(s, 0);
}

// ---- row-striped reductions --------------------------------------------------------------
template <class G>
PFX_HD inline double rows_sum(const G& g, const double* a, int lo, int m) {
    if constexpr (G::ordered) return ordered_rows(g, lo < 0 ? 0 : lo, m, [&](int i) { return a[i]; });
    double s = 0.0;
    for (int i = g.lane(); i < m; i += G::W)
        if (i >= lo) s += a[i];
    return g.sum(s);
}

template <class G>
PFX_HD inline double rows_sumsq(const G& g, const double* a, int lo, int m) {
    if constexpr (G::ordered) return ordered_rows(g, lo < 0 ? 0 : lo, m, [&](int i) { return PFX_MUL(a[i], a[i]); });
    double s = 0.0;
    for (int i = g.lane(); i < m; i += G::W)
        if (i >= lo) s = PFX_FMA(a[i], a[i], s);
    return g.sum(s);
}

template <class G>
PFX_HD inline double rows_dot(const G& g, const double* a, const double* b, int lo, int m) {
    if constexpr (G::ordered) return ordered_rows(g, lo < 0 ? 0 : lo, m, [&](int i) { return PFX_MUL(a[i], b[i]); });
    double s = 0.0;
    for (int i = g.lane(); i < m; i += G::W)
        if (i >= lo) s = PFX_FMA(a[i], b[i], s);
    return g.sum(s);
}

// numpy.argmax over rows [0, m)
template <class G>
PFX_HD inline ValIdx rows_argmax(const G& g, const double* a, int m) {
    ValIdx best;
    best.v = -INFINITY;
    best.i = 0x7fffffff;
    bool have = false;
    for (int i = g.lane(); i < m; i += G::W) {
        if (!have || argmax_better(a[i], i, best.v, best.i)) {
            best.v = a[i];
            best.i = i;
            have = true;
        }
    }
    return g.argmax(best);
}

// Python's builtin max(list): sequential '>' comparisons, so a NaN survives only when it is
// the first element; other NaNs are skipped.
template <class G>
PFX_HD inline double rows_builtin_max(const G& g, const double* a, int m) {
    double v = -INFINITY;
    for (int i = g.lane(); i < m; i += G::W)
        if (a[i] == a[i]) v = fmax(v, a[i]);
    v = g.max(v);
    double a0 = a[0];
    return (a0 != a0) ? a0 : v;
}

template <class G>
PFX_HD inline double rows_builtin_min(const G& g, const double* a, int m) {
    double v = INFINITY;
    for (int i = g.lane(); i < m; i += G::W)
        if (a[i] == a[i]) v = fmin(v, a[i]);
    v = g.min(v);
    double a0 = a[0];
    return (a0 != a0) ? a0 : v;
}

// mean and population standard deviation the way numpy computes them (two passes).
template <class G>
PFX_HD inline void rows_mean_std(const G& g, const double* a, int m, double& mean, double& sd) {
    mean = rows_sum(g, a, 0, m) / (double)m;
    double s = 0.0;
    for (int i = g.lane(); i < m; i += G::W) {
        double d = a[i] - mean;
        s = PFX_FMA(d, d, s);
    }
    sd = sqrt(g.sum(s) / (double)m);
}

}  // namespace pfx
