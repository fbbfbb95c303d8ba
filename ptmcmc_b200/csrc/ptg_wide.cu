// ptg_wide.cu -- instantiation and launchers of the warp-per-chain kernels (17 <= dim <= 128), see ptg_wide.cuh
#include "ptg_wide_mma.cuh"
#include "ptg_launch.h"

static int cpl_for(int dim) { return dim <= 32 ? 1 : (dim <= 64 ? 2 : 4); }

template <int CPL, int MODE, int MAXT>
static cudaError_t xstep_tt(const PtgModel &m, const PtgState &s, long long step0, int n_steps, cudaStream_t st) {
  auto k = ptg_xstep_kernel<CPL, MODE, MAXT>;
  const size_t smem = ptg_xshared_bytes(m.n_rungs, 32 * CPL, m.n_props);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<m.n_ladders, 32 * m.n_rungs, smem, st>>>(m, s, step0, n_steps);
  return cudaGetLastError();
}
template <int CPL, int MODE>
static cudaError_t xstep_t(const PtgModel &m, const PtgState &s, long long step0, int n_steps, cudaStream_t st) {
  if (m.n_rungs <= 16) return xstep_tt<CPL, MODE, 512>(m, s, step0, n_steps, st);
  if (m.n_rungs <= 24) return xstep_tt<CPL, MODE, 768>(m, s, step0, n_steps, st);
  return xstep_tt<CPL, MODE, 1024>(m, s, step0, n_steps, st);
}
template <int CPL, int MODE>
static cudaError_t xinit_t(const PtgModel &m, const PtgState &s, const double *init_x, cudaStream_t st) {
  const size_t smem = (size_t)8 * 32 * CPL * sizeof(double);
  ptg_xinit_kernel<CPL, MODE><<<(unsigned)((m.n_chains + 3) / 4), 128, smem, st>>>(m, s, init_x);
  return cudaGetLastError();
}
template <int CPL>
static cudaError_t xeval_t(const PtgModel &m, const double *x, long long n, double *ll, double *lp, cudaStream_t st) {
  const size_t smem = (size_t)8 * 32 * CPL * sizeof(double);
  ptg_xeval_kernel<CPL><<<(unsigned)((n + 3) / 4), 128, smem, st>>>(m, x, n, ll, lp);
  return cudaGetLastError();
}

template <int CPL, int MAXT>
static cudaError_t xmstep_tt(const PtgModel &m, const PtgState &s, long long step0, int n_steps, int trans_off, cudaStream_t st) {
  auto k = ptg_xmstep_kernel<CPL, MAXT>;
  const size_t smem = ptg_xshared_bytes(m.n_rungs, 32 * CPL, m.n_props);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<m.n_ladders, 32 * m.n_rungs, smem, st>>>(m, s, step0, n_steps, trans_off);
  return cudaGetLastError();
}
template <int CPL>
static cudaError_t xmstep_t(const PtgModel &m, const PtgState &s, long long step0, int n_steps, int trans_off, cudaStream_t st) {
  if (m.n_rungs <= 16) return xmstep_tt<CPL, 512>(m, s, step0, n_steps, trans_off, st);
  if (m.n_rungs <= 24) return xmstep_tt<CPL, 768>(m, s, step0, n_steps, trans_off, st);
  return xmstep_tt<CPL, 1024>(m, s, step0, n_steps, trans_off, st);
}
// Philox production kernel with the DMMA-batched contractions
cudaError_t ptg_launch_xmstep(const PtgModel &m, const PtgState &s, long long step0, int n_steps, int trans_off, cudaStream_t st) {
  switch (cpl_for(m.dim)) {
  case 1: return xmstep_t<1>(m, s, step0, n_steps, trans_off, st);
  case 2: return xmstep_t<2>(m, s, step0, n_steps, trans_off, st);
  default: return xmstep_t<4>(m, s, step0, n_steps, trans_off, st);
  }
}

#define XDISPATCH(CALL_TAPE, CALL_PHILOX)                       \
  switch (cpl_for(m.dim)) {                                      \
  case 1: { constexpr int C = 1; return mode == PTG_RNG_TAPE ? CALL_TAPE : CALL_PHILOX; } \
  case 2: { constexpr int C = 2; return mode == PTG_RNG_TAPE ? CALL_TAPE : CALL_PHILOX; } \
  default: { constexpr int C = 4; return mode == PTG_RNG_TAPE ? CALL_TAPE : CALL_PHILOX; } \
  }

cudaError_t ptg_launch_xstep(int mode, const PtgModel &m, const PtgState &s, long long step0, int n_steps, cudaStream_t st) {
  XDISPATCH((xstep_t<C, PTG_RNG_TAPE>(m, s, step0, n_steps, st)), (xstep_t<C, PTG_RNG_PHILOX>(m, s, step0, n_steps, st)))
}
cudaError_t ptg_launch_xinit(int mode, const PtgModel &m, const PtgState &s, const double *init_x, cudaStream_t st) {
  XDISPATCH((xinit_t<C, PTG_RNG_TAPE>(m, s, init_x, st)), (xinit_t<C, PTG_RNG_PHILOX>(m, s, init_x, st)))
}
cudaError_t ptg_launch_xeval(const PtgModel &m, const double *x, long long n, double *ll, double *lp, cudaStream_t st) {
  switch (cpl_for(m.dim)) {
  case 1: return xeval_t<1>(m, x, n, ll, lp, st);
  case 2: return xeval_t<2>(m, x, n, ll, lp, st);
  default: return xeval_t<4>(m, x, n, ll, lp, st);
  }
}
