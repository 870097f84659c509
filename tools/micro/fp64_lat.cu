// microbenchmark: FP64 dependent-issue latency and throughput vs warps per SM sub-partition (B200)
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void dfma_lat(double* out, long long* cyc, int iters, double a, double b)
{
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void dadd_lat(double* out, long long* cyc, int iters, double a)
{
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = x + a;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void shfl_lat(double* out, long long* cyc, int iters)
{
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 8 << 20); cudaMallocManaged(&cyc, 8);
    const int iters = 2000;
    for (int warps : {1, 2, 4, 8, 12, 16, 24, 32}) {
        dfma_lat<1><<<1, 32 * warps>>>(out, cyc, iters, 0.999, 1e-9); cudaDeviceSynchronize();
        double c1 = (double)*cyc / (iters * 16);
        dfma_lat<2><<<1, 32 * warps>>>(out, cyc, iters, 0.999, 1e-9); cudaDeviceSynchronize();
        double c2 = (double)*cyc / (iters * 16);
        dfma_lat<4><<<1, 32 * warps>>>(out, cyc, iters, 0.999, 1e-9); cudaDeviceSynchronize();
        double c4 = (double)*cyc / (iters * 16);
        dfma_lat<8><<<1, 32 * warps>>>(out, cyc, iters, 0.999, 1e-9); cudaDeviceSynchronize();
        double c8 = (double)*cyc / (iters * 16);
        printf("warps/SM %2d: cycles per loop step (1 DFMA per chain): chains1 %.2f chains2 %.2f chains4 %.2f chains8 %.2f | DFMA/clk/SM at 8 chains: %.1f\n",
               warps, c1, c2, c4, c8, warps * 32 * 8 / c8);
    }
    dadd_lat<<<1, 32>>>(out, cyc, iters, 1e-9); cudaDeviceSynchronize();
    printf("DADD dependent latency %.2f cycles\n", (double)*cyc / (iters * 16));
    shfl_lat<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("double SHFL dependent latency %.2f cycles (2 SHFL.32)\n", (double)*cyc / (iters * 16));
    return 0;
}
