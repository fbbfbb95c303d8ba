// ptg_inst.cuh -- per-dimension instantiation of the thread-per-chain kernels (one translation unit per D so the
// build parallelises).  PTG_INSTANTIATE(D) defines the two launchers declared in ptg_launch.h.
#pragma once
#include <cstdlib>
#include "ptg_kernels.cuh"
#include "ptg_warp.cuh"
#include "ptg_fast.cuh"
#include "ptg_launch.h"

template <int D, int MODE>
static cudaError_t launch_step_t(const PtgModel &m, const PtgState &s, long long step0, int n_steps, int lpb, size_t smem, cudaStream_t st, int phase = 0) {
  auto k = ptg_step_kernel<D, MODE>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  int blocks = (int)((m.n_ladders + lpb - 1) / lpb);
  k<<<blocks, lpb * m.n_rungs, smem, st>>>(m, s, step0, n_steps, lpb, phase);
  return cudaGetLastError();
}
// second-generation kernel: ladder-in-a-warp (n_rungs <= W <= 32), 4 warps per CTA
template <int D, int MODE>
static cudaError_t launch_wstep_t(const PtgModel &m, const PtgState &s, long long step0, int n_steps, int W, cudaStream_t st) {
  const long long warps = (m.n_ladders + (32 / W) - 1) / (32 / W);
  const int blocks = (int)((warps + 3) / 4);
  const size_t smem = (size_t)m.n_rungs * m.n_bins * sizeof(double);
  ptg_wstep_kernel<D, MODE><<<blocks, 128, smem, st>>>(m, s, step0, n_steps, W);
  return cudaGetLastError();
}
// production kernel (Philox draws): ladder-in-a-warp; proposal table, bins, per-thread counters / MAP and the per-warp pool slots in
// dynamic shared memory (FShared).  lk = -1: every feature at run time; lk >= 0: the streamlined instantiation for likelihood kind lk
// (the host checks the configuration, ptg_api.cu:fstep_streamlined_kind)
template <int D, int XCHG, int LK, int MAXT>
static cudaError_t launch_fstep_k(const PtgModel &m, const PtgState &s, long long step0, int n_steps, int W, const PtgXchg &xc, int blocks, int threads, size_t smem, cudaStream_t st) {
  auto k = ptg_fstep_kernel<D, XCHG, LK, MAXT>;
  static size_t smem_set[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > 48 * 1024 && smem > smem_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    smem_set[dev & 63] = smem;
  }
  if (XCHG == 2) {
    // the in-launch exchange pairs warps across GPUs: every CTA of the grid must be resident at once (measured, not assumed)
    int per_sm = 0, sms = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, smem);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if ((long long)blocks > (long long)per_sm * sms) return cudaErrorCooperativeLaunchTooLarge;
  }
  k<<<blocks, threads, smem, st>>>(m, s, step0, n_steps, W, xc);
  return cudaGetLastError();
}
template <int D>
static cudaError_t launch_fstep_t(const PtgModel &m, const PtgState &s, long long step0, int n_steps, int W, const PtgXchg &xc, int lk, cudaStream_t st) {
  // data chi^2 likelihoods: one ladder per warp whatever its size; the data sums of the chains that passed the prior gate are spread over all
  // 32 lanes (ptg_fast.cuh: flike_data_compact).  A property of the workload, not of the batch size.
  if ((lk == PTG_LIKE_POLY_CHI2 || lk == PTG_LIKE_SINUSOID_CHI2) && !xc.on) W = 32;
  const long long warps = (m.n_ladders + (32 / W) - 1) / (32 / W);
  // CTA size: the largest of 28 / 14 / 4 warps that still puts a CTA on (nearly) every SM.  The kernel is compiled for 72
  // registers (launch bound 896 x 1), so 896 resident threads per SM in every geometry; one 28-warp CTA per SM measured
  // 5 % faster than seven 4-warp CTAs on the 4096 x 32 workload.
  int threads = ptg_fstep_threads(warps);
  if ((lk == PTG_LIKE_POLY_CHI2 || lk == PTG_LIKE_SINUSOID_CHI2) && threads > 448) threads = 448; // the data-likelihood instantiations are compiled for <= 448 threads
  const int wpb = threads / 32;
  const int blocks = (int)((warps + wpb - 1) / wpb);
  const size_t smem = FShared<D>::bytes(threads);
#define GO(X, L, T) return launch_fstep_k<D, X, L, T>(m, s, step0, n_steps, W, xc, blocks, threads, smem, st)
  const bool small = threads <= 448; // at most 14 warps per SM: the 144-register instantiations
  if (xc.on && xc.every > 0) GO(2, -1, 896);
  if (xc.on) { if (lk == PTG_LIKE_SINES) { if (small) GO(1, PTG_LIKE_SINES, 448); GO(1, PTG_LIKE_SINES, 896); } GO(1, -1, 896); }
  if (lk == PTG_LIKE_SINES) { if (small) GO(0, PTG_LIKE_SINES, 448); GO(0, PTG_LIKE_SINES, 896); }
  if (lk == PTG_LIKE_GAUSS_ISO) { if (small) GO(0, PTG_LIKE_GAUSS_ISO, 448); GO(0, PTG_LIKE_GAUSS_ISO, 896); }
  if constexpr (D >= 2) { if (lk == PTG_LIKE_POLY_CHI2) GO(0, PTG_LIKE_POLY_CHI2, 448); }       // fp64-bound: at most 14 warps per SM, the full register file
  if constexpr (D % 3 == 0) { if (lk == PTG_LIKE_SINUSOID_CHI2) GO(0, PTG_LIKE_SINUSOID_CHI2, 448); }
  GO(0, -1, 896);
#undef GO
}
template <int D, int MODE>
static cudaError_t launch_init_t(const PtgModel &m, const PtgState &s, const double *init_x, cudaStream_t st) {
  int blocks = (int)((m.n_chains + 127) / 128);
  ptg_init_kernel<D, MODE><<<blocks, 128, 0, st>>>(m, s, init_x);
  return cudaGetLastError();
}

#define PTG_INSTANTIATE(D)                                                                                                   \
  cudaError_t ptg_launch_step_d##D(int mode, const PtgModel &m, const PtgState &s, long long step0, int n_steps, int lpb,    \
                                   size_t smem, cudaStream_t st) {                                                           \
    return mode == PTG_RNG_TAPE ? launch_step_t<D, PTG_RNG_TAPE>(m, s, step0, n_steps, lpb, smem, st)                        \
                                : launch_step_t<D, PTG_RNG_PHILOX>(m, s, step0, n_steps, lpb, smem, st);                     \
  }                                                                                                                          \
  cudaError_t ptg_launch_cb_d##D(int what, const PtgModel &m, const PtgState &s, long long step, int lpb, size_t smem, int k, int32_t *attempt,   \
                                 int32_t *n_open, cudaStream_t st) {                                                        \
    const unsigned nb = (unsigned)((m.n_chains + 127) / 128);                                                                \
    if (what == 0) return launch_step_t<D, PTG_RNG_PHILOX>(m, s, step, 1, lpb, smem, st, 1);                                 \
    if (what == 1) ptg_cb_finish_kernel<D><<<nb, 128, 0, st>>>(m, s, step);                                                  \
    else if (what == 2) ptg_cb_init_draw_kernel<D><<<nb, 128, 0, st>>>(m, s, k, attempt);                                    \
    else ptg_cb_init_accept_kernel<D><<<nb, 128, 0, st>>>(m, s, attempt, n_open);                                            \
    return cudaGetLastError();                                                                                               \
  }                                                                                                                          \
  cudaError_t ptg_launch_wstep_d##D(int mode, const PtgModel &m, const PtgState &s, long long step0, int n_steps, int W,     \
                                    cudaStream_t st) {                                                                       \
    return mode == PTG_RNG_TAPE ? launch_wstep_t<D, PTG_RNG_TAPE>(m, s, step0, n_steps, W, st)                               \
                                : launch_wstep_t<D, PTG_RNG_PHILOX>(m, s, step0, n_steps, W, st);                            \
  }                                                                                                                          \
  cudaError_t ptg_launch_fstep_d##D(const PtgModel &m, const PtgState &s, long long step0, int n_steps, int W, const PtgXchg &xc, int lk, cudaStream_t st) { \
    return launch_fstep_t<D>(m, s, step0, n_steps, W, xc, lk, st);                                                                   \
  }                                                                                                                          \
  cudaError_t ptg_launch_init_d##D(int mode, const PtgModel &m, const PtgState &s, const double *init_x, cudaStream_t st) {  \
    return mode == PTG_RNG_TAPE ? launch_init_t<D, PTG_RNG_TAPE>(m, s, init_x, st)                                           \
                                : launch_init_t<D, PTG_RNG_PHILOX>(m, s, init_x, st);                                        \
  }                                                                                                                          \
  cudaError_t ptg_launch_eval_d##D(const PtgModel &m, const double *x, long long n, double *ll, double *lp, cudaStream_t st) { \
    ptg_eval_kernel<D><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(m, x, n, ll, lp);                                        \
    return cudaGetLastError();                                                                                               \
  }
