// peak_fp64.cu — measures the FP64 vector-pipe FMA peak of the device, live, for the roofline denominator.
// MEASURED_PEAKS.json carries HBM and bf16 only; K1 is bound by the FP64 pipe, so bench.py needs
// this number from the same box, the same clocks and the same run.  Pure register DFMA chains: 8 independent
// accumulators per thread, 1024 threads per SM x 2 CTAs, no memory traffic.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vpmb200.h"

namespace {

__global__ void __launch_bounds__(512) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[0] = s;  // keep the chains alive without a store in the common case
}

}  // namespace

extern "C" int32_t vpmb200_measure_fp64_peak(int32_t device, int32_t iters, int32_t repeats, double* tflops, double* ms) {
    if (!tflops || iters <= 0 || repeats <= 0) return VPMB200_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return VPMB200_ENODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return VPMB200_ECUDA;
    double* d = nullptr;
    if (cudaMalloc(&d, 64) != cudaSuccess) return VPMB200_ECUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = prop.multiProcessorCount * 4, block = 512;
    dfma_peak_kernel<<<grid, block>>>(d, iters, 0.999999, 1e-7);  // warm-up
    double best = 1e30;
    for (int r = 0; r < repeats; ++r) {
        cudaEventRecord(e0);
        dfma_peak_kernel<<<grid, block>>>(d, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return VPMB200_ECUDA; }
        float t = 0;
        cudaEventElapsedTime(&t, e0, e1);
        if (t < best) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    double flops = 2.0 * 8 * 16 * (double)iters * (double)grid * block;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms) *ms = best;
    return VPMB200_OK;
}
