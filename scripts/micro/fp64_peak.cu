// FP64 micro-benchmarks on the GPU box: DFMA throughput vs resident warps, dependent-DFMA latency,
// exp() and division cost.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double x[ILP];
    for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 1e-9 + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0;
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_exp(double* out, int iters, double a) {
    double x = threadIdx.x * 1e-3, s = 0;
    for (int i = 0; i < iters; ++i) { s += exp(-x); x += a; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_div(double* out, int iters, double a) {
    double x = 1.0 + threadIdx.x * 1e-3, s = 0;
    for (int i = 0; i < iters; ++i) { s += a / x; x += 1e-3; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F>
float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 32 * 1024 * 8);
    int sm = 148; const int iters = 20000;
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        int threads = 32 * (warps > 32 ? 32 : warps);
        float ms1 = timeit([&] { k_dfma<1><<<sm, threads>>>(out, iters, 1.0000001, 1e-9); });
        float ms8 = timeit([&] { k_dfma<8><<<sm, threads>>>(out, iters, 1.0000001, 1e-9); });
        double fl1 = 2.0 * sm * threads * (double)iters / (ms1 * 1e-3) / 1e12;
        double fl8 = 2.0 * 8 * sm * threads * (double)iters / (ms8 * 1e-3) / 1e12;
        printf("warps/SM %2d: ILP1 %.3f ms %.2f TFLOP/s (%.1f cyc/dep-DFMA @1.965GHz) | ILP8 %.3f ms %.2f TFLOP/s\n",
               warps, ms1, fl1, ms1 * 1e-3 * 1.965e9 / iters, ms8, fl8);
    }
    float me = timeit([&] { k_exp<<<sm, 512>>>(out, 2000, 1e-3); });
    printf("exp: %.1f ns per call per thread-stream; %.2f G exp/s chip (16 warps/SM)\n", me * 1e6 / 2000, sm * 512 * 2000.0 / (me * 1e-3) / 1e9);
    float md = timeit([&] { k_div<<<sm, 512>>>(out, 2000, 3.0); });
    printf("div: %.2f G div/s chip (16 warps/SM)\n", sm * 512 * 2000.0 / (md * 1e-3) / 1e9);
    return 0;
}
