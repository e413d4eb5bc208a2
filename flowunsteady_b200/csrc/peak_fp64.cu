// peak_fp64.cu — the FP64 vector-pipe FMA peak of the device, live, for the roofline denominator.
// MEASURED_PEAKS.json carries HBM and bf16 only; K1 is bound by the FP64 pipe, so bench.py needs this number
// from the same box, the same clocks and the same run.
//
// Two figures are produced:
//   * `measured` = the best of several register-only DFMA chain kernels (different chain counts / CTA shapes) — the roofline
//     denominator bench.py uses.  ncu reads 99.97 % FP64 pipe active on it (profiles/r02a_dfma_peak.txt) and it reaches
//     99.8 % of the nominal pipe rate (round 1's single 4 ms launch read 92 %: too short, one shape);
//   * `pipe` = SMs x 64 FP64 lanes x 2 flop x the driver's nominal max SM clock: the issue-rate ceiling ncu's
//     sm__inst_executed_pipe_fp64 counts against (one warp-wide DFMA per 2 cycles per SM sub-partition), for context.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vpmb200.h"

namespace {

// NCH independent chains per thread, two alternating multipliers; every shape below keeps ALL its CTAs resident at once
// (registers: 2 NCH + ~10), so block 0's clock64 span is the span of the launch.
template <int NCH>
__global__ void dfma_peak_kernel(double* out, long long* clk, int iters, double a, double b) {
    double x[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) x[c] = threadIdx.x * 1e-3 + c;
    const double a2 = a - 1e-9;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 128 / NCH; ++u) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) x[c] = fma(x[c], (c & 1) ? a2 : a, b);
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < NCH; ++c) s += x[c];
    if (s == 123.456) out[0] = s;  // keep the chains alive without a store in the common case
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

struct Shape {
    int nch, block, ctas_per_sm;
};

template <int NCH>
cudaError_t run_shape(int grid, int block, double* d, long long* clk, int iters) {
    dfma_peak_kernel<NCH><<<grid, block>>>(d, clk, iters, 0.999999, 1e-7);
    return cudaGetLastError();
}

cudaError_t launch(const Shape& s, int sms, double* d, long long* clk, int iters) {
    const int grid = sms * s.ctas_per_sm;
    switch (s.nch) {
    case 4: return run_shape<4>(grid, s.block, d, clk, iters);
    case 8: return run_shape<8>(grid, s.block, d, clk, iters);
    default: return run_shape<16>(grid, s.block, d, clk, iters);
    }
}

}  // namespace

extern "C" int32_t vpmb200_measure_fp64_peak2(int32_t device, int32_t iters, int32_t repeats, double* out8) {
    if (!out8 || iters <= 0 || repeats <= 0) return VPMB200_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return VPMB200_ENODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return VPMB200_ECUDA;
    double* d = nullptr;
    long long* clk = nullptr;
    if (cudaMalloc(&d, 64) != cudaSuccess) return VPMB200_ECUDA;
    if (cudaMalloc(&clk, 64) != cudaSuccess) { cudaFree(d); return VPMB200_ECUDA; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    // 8 chains x 16 warps/scheduler (round 1's shape), 16 chains x 8 warps, 4 chains x 16 warps, 8 chains x 8 warps
    const Shape shapes[4] = {{8, 512, 4}, {16, 256, 4}, {4, 512, 4}, {8, 256, 4}};
    double best_tf = 0.0, best_ms = 0.0, best_mhz = 0.0;
    int best_shape = -1;
    for (int k = 0; k < 4; ++k) {
        const Shape& s = shapes[k];
        if (launch(s, prop.multiProcessorCount, d, clk, iters) != cudaSuccess) continue;   // warm-up
        cudaDeviceSynchronize();
        for (int r = 0; r < repeats; ++r) {
            cudaEventRecord(e0);
            launch(s, prop.multiProcessorCount, d, clk, iters);
            cudaEventRecord(e1);
            if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); cudaFree(clk); return VPMB200_ECUDA; }
            float t = 0;
            cudaEventElapsedTime(&t, e0, e1);
            long long cyc = 0;
            cudaMemcpy(&cyc, clk, sizeof(cyc), cudaMemcpyDeviceToHost);
            const double flops = 2.0 * 128.0 * (double)iters * (double)prop.multiProcessorCount * s.ctas_per_sm * s.block;
            const double tf = flops / (t * 1e-3) / 1e12;
            if (tf > best_tf) {
                best_tf = tf;
                best_ms = t;
                best_shape = k;
                // block 0's own span in SM cycles (diagnostic only: CTAs of one launch do not all start at once)
                best_mhz = (double)cyc / (t * 1e-3) / 1e6;
            }
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    cudaFree(clk);
    if (best_shape < 0) return VPMB200_ECUDA;
    const double pipe_tf = (double)prop.multiProcessorCount * 64.0 * 2.0 * (double)prop.clockRate * 1e3 / 1e12;
    out8[0] = best_tf;                 // best measured DFMA throughput, TFLOP/s
    out8[1] = best_ms;                 // its launch time
    out8[2] = best_mhz;                // block 0's clock64 span / event span, MHz (diagnostic)
    out8[3] = pipe_tf;                 // SMs x 64 lanes x 2 x nominal max SM clock
    out8[4] = pipe_tf > 0 ? best_tf / pipe_tf : 0.0;
    out8[5] = (double)best_shape;      // index into {8ch x 512 x 4, 16ch x 256 x 4, 4ch x 512 x 4, 8ch x 256 x 4}
    out8[6] = (double)prop.multiProcessorCount;
    out8[7] = (double)prop.clockRate * 1e-3;   // the driver's nominal max SM clock, MHz
    return VPMB200_OK;
}

extern "C" int32_t vpmb200_measure_fp64_peak(int32_t device, int32_t iters, int32_t repeats, double* tflops, double* ms) {
    if (!tflops) return VPMB200_EINVAL;
    double o[8];
    int32_t rc = vpmb200_measure_fp64_peak2(device, iters, repeats, o);
    if (rc) return rc;
    *tflops = o[0];
    if (ms) *ms = o[1];
    return VPMB200_OK;
}
