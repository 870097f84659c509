// peak.cu -- measures the device's FP64 FMA throughput (roofline denominator of the integration kernel; the
// driver-written MEASURED_PEAKS.json holds HBM and bf16 peaks only).
#include "../../include/mio_cuda.h"
#include <cuda_runtime.h>

namespace
{
__global__ void __launch_bounds__(256) dfma_chain_kernel(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b);
            x1 = fma(x1, a, b);
            x2 = fma(x2, a, b);
            x3 = fma(x3, a, b);
            x4 = fma(x4, a, b);
            x5 = fma(x5, a, b);
            x6 = fma(x6, a, b);
            x7 = fma(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
} // namespace

extern "C" int mio_cuda_measure_fp64_peak(int device, double seconds, double* tflops)
{
    if (!tflops || cudaSetDevice(device) != cudaSuccess) {
        return MIO_UNKNOWN_ERROR;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    double* out      = nullptr;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) {
        return MIO_UNKNOWN_ERROR;
    }
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    int iters = 2000;
    dfma_chain_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9); // warm-up
    cudaDeviceSynchronize();
    double best = 0.0, elapsed_total = 0.0;
    while (elapsed_total < seconds * 1e3) {
        cudaEventRecord(a);
        dfma_chain_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        elapsed_total += ms;
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        best               = flops / (ms * 1e-3) / 1e12 > best ? flops / (ms * 1e-3) / 1e12 : best;
        if (ms < 20.f) {
            iters *= 2;
        }
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    *tflops = best;
    return cudaGetLastError() == cudaSuccess ? MIO_OK : MIO_UNKNOWN_ERROR;
}
