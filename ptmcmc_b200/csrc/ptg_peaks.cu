// ptg_peaks.cu -- FP64 peak microbenchmarks (the FP64 roofline denominators are not in MEASURED_PEAKS.json and
// must be measured on the box): dependent-free DFMA chains, the DMUL+DADD pairs the engine's -fmad=false arithmetic issues,
// and mma.sync.m8n8k4.f64 (DMMA).  Exposed through the C ABI as ptg_measure_fp64_peaks; bench.py --peaks records them.
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/ptmcmc_b200.h"

#define PEAK_CHAINS 8
__global__ void __launch_bounds__(256) peak_dfma_kernel(double *out, int iters, double b, double c) {
  double a[PEAK_CHAINS];
#pragma unroll
  for (int k = 0; k < PEAK_CHAINS; k++) a[k] = threadIdx.x * 1e-3 + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; k++) a[k] = fma(a[k], b, c);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < PEAK_CHAINS; k++) s += a[k];
  if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) peak_dmul_dadd_kernel(double *out, int iters, double b, double c) {
  double a[PEAK_CHAINS];
#pragma unroll
  for (int k = 0; k < PEAK_CHAINS; k++) a[k] = threadIdx.x * 1e-3 + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; k++) a[k] = __dadd_rn(__dmul_rn(a[k], b), c);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < PEAK_CHAINS; k++) s += a[k];
  if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) peak_dmma_kernel(double *out, int iters, double av, double bv) {
  double c[4][2];
#pragma unroll
  for (int k = 0; k < 4; k++) { c[k][0] = threadIdx.x; c[k][1] = k; }
  const double a = av + (threadIdx.x & 3) * 1e-9, b = bv;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 4; k++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) s += c[k][0] + c[k][1];
  if (s == 12345.678) out[0] = s;
}

// out[0] = DFMA TFLOP/s (2 flops per fma), out[1] = DMUL+DADD TFLOP/s (2 flops per pair), out[2] = DMMA TFLOP/s, out[3] = SM count
extern "C" int ptg_measure_fp64_peaks(int32_t device, double *out) {
  if (!out) return PTG_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return PTG_ECUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PTG_ECUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
  double *d = nullptr;
  if (cudaMalloc(&d, 64) != cudaSuccess) return PTG_ENOMEM;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  double best[3] = {0, 0, 0};
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0); peak_dfma_kernel<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double v = 2.0 * PEAK_CHAINS * (double)iters * blocks * threads / (ms * 1e-3) / 1e12; if (rep && v > best[0]) best[0] = v;
    cudaEventRecord(e0); peak_dmul_dadd_kernel<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    v = 2.0 * PEAK_CHAINS * (double)iters * blocks * threads / (ms * 1e-3) / 1e12; if (rep && v > best[1]) best[1] = v;
    cudaEventRecord(e0); peak_dmma_kernel<<<blocks, threads>>>(d, iters, 1.0000001, 0.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    v = 512.0 * 4 * (double)iters * blocks * (threads / 32) / (ms * 1e-3) / 1e12; if (rep && v > best[2]) best[2] = v;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  if (cudaGetLastError() != cudaSuccess) return PTG_ECUDA;
  out[0] = best[0]; out[1] = best[1]; out[2] = best[2]; out[3] = prop.multiProcessorCount;
  return 0;
}
