// ptg_launch.h -- launchers exported by the per-dimension translation units (ptg_inst_dN.cu)
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptg_types.h"
// dynamic shared memory of one ladder in the shared-memory step kernel (must match LadderShared::carve in ptg_kernels.cuh)
static inline size_t ptg_ladder_shared_bytes(int D, int R) {
  size_t b = sizeof(double) * ((size_t)R * D * 2 + (size_t)R * 10 + PTG_SWAP_SLOTS * 3) + sizeof(int) * ((size_t)R * 5 + PTG_SWAP_SLOTS) +
             sizeof(long long) * (size_t)R * 2;
  return (b + 15) & ~(size_t)15;
}
#ifdef PTG_DEV_DIM3  // developer build: `make DEV=1` compiles only dim = 3, 5, 9 (seconds instead of minutes)
#define PTG_DIM_LIST(X) X(3) X(5) X(9)
#else
#define PTG_DIM_LIST(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
#endif
// CTA size of the production kernel (ptg_fstep_kernel) for a batch of `warps` ladder-warps
static inline int ptg_fstep_threads(long long warps) {
  static const int forced = [] { const char *e = getenv("PTG_FSTEP_THREADS"); return e ? atoi(e) : 0; }(); // experiments only
  if (forced >= 32 && forced <= 896 && forced % 32 == 0) return forced;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long need = (long long)sms * 95 / 100;
  if ((warps + 27) / 28 >= need) return 896;
  if ((warps + 13) / 14 >= need) return 448;
  return 128;
}
// 1 if every CTA of that grid is resident at once (72 registers: 896 threads per SM), else 0
static inline int ptg_fstep_grid_is_resident(long long warps) {
  const int threads = ptg_fstep_threads(warps);
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long blocks = (warps + threads / 32 - 1) / (threads / 32);
  return blocks <= (long long)(896 / threads) * sms ? 1 : 0;
}

#define PTG_DECLARE(D)                                                                                                      \
  cudaError_t ptg_launch_step_d##D(int mode, const PtgModel &m, const PtgState &s, long long step0, int n_steps, int lpb,   \
                                   size_t smem, cudaStream_t st);                                                           \
  cudaError_t ptg_launch_wstep_d##D(int mode, const PtgModel &m, const PtgState &s, long long step0, int n_steps, int W,    \
                                    cudaStream_t st);                                                                       \
  /* host-callback likelihood mode: what = 0 propose (shared-memory kernel, phase 1), 1 finish, 2 init draw, 3 init accept */                 \
  cudaError_t ptg_launch_cb_d##D(int what, const PtgModel &m, const PtgState &s, long long step, int lpb, size_t smem, int k, int32_t *attempt,   \
                                 int32_t *n_open, cudaStream_t st);                                                          \
  cudaError_t ptg_launch_fstep_d##D(const PtgModel &m, const PtgState &s, long long step0, int n_steps, int W, const PtgXchg &xc, int lk, cudaStream_t st); \
  cudaError_t ptg_launch_init_d##D(int mode, const PtgModel &m, const PtgState &s, const double *init_x, cudaStream_t st);       \
  cudaError_t ptg_launch_eval_d##D(const PtgModel &m, const double *x, long long n, double *ll, double *lp, cudaStream_t st);
PTG_DIM_LIST(PTG_DECLARE)
#undef PTG_DECLARE

// warp-per-chain kernels for 17 <= dim <= 128 (ptg_wide.cu)
cudaError_t ptg_launch_xstep(int mode, const PtgModel &m, const PtgState &s, long long step0, int n_steps, cudaStream_t st);
cudaError_t ptg_launch_xmstep(const PtgModel &m, const PtgState &s, long long step0, int n_steps, int trans_off, cudaStream_t st);
cudaError_t ptg_launch_xinit(int mode, const PtgModel &m, const PtgState &s, const double *init_x, cudaStream_t st);
cudaError_t ptg_launch_xeval(const PtgModel &m, const double *x, long long n, double *ll, double *lp, cudaStream_t st);
