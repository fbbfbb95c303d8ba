// ptg_wide_pipe.cu -- instantiation and launcher of the pipelined full-covariance kernel (ptg_wide_pipe.cuh); its own translation unit so
// that the other warp-per-chain kernels (ptg_wide.cu) do not recompile with it
#include "ptg_wide_pipe.cuh"
#include "ptg_launch.h"

static int cpl_for(int dim) { return dim <= 32 ? 1 : (dim <= 64 ? 2 : 4); }

// Pipelined production kernel for the full-covariance Gaussian workload (ptg_wide_pipe.cuh): both matrices resident in shared memory,
// one warp per chain, the swap phase overlapped with the rotation.  ptg_xpstep_fits tells the caller whether the configuration qualifies.
template <int CPL, int MAXT>
static cudaError_t xpstep_tt(const PtgModel &m, const PtgState &s, long long step0, int n_steps, int trans_off, double *scratch, cudaStream_t st) {
  static const bool pool = [] { const char *e = getenv("PTG_XP_POOL"); return e && atoi(e) != 0; }(); // experiment switch: pooled Gaussian offsets
  auto k = pool ? ptg_xpstep_kernel<CPL, MAXT, true> : ptg_xpstep_kernel<CPL, MAXT, false>;
  const size_t smem = ptg_xp_shared_bytes(m.n_rungs, m.dim, m.n_props, CPL);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<m.n_ladders, 32 * m.n_rungs, smem, st>>>(m, s, step0, n_steps, trans_off, scratch, ptg_xp_layout(m.n_rungs, m.dim, m.n_props, CPL));
  return cudaGetLastError();
}
template <int CPL>
static cudaError_t xpstep_t(const PtgModel &m, const PtgState &s, long long step0, int n_steps, int trans_off, double *scratch, cudaStream_t st) {
  if (m.n_rungs <= 16) return xpstep_tt<CPL, 512>(m, s, step0, n_steps, trans_off, scratch, st);
  if (m.n_rungs <= 24) return xpstep_tt<CPL, 768>(m, s, step0, n_steps, trans_off, scratch, st);
  return xpstep_tt<CPL, 1024>(m, s, step0, n_steps, trans_off, scratch, st);
}
int ptg_xpstep_fits(const PtgModel &m, int trans_off) {
  if (m.like_kind != PTG_LIKE_GAUSS_FULLCOV || trans_off < 0 || m.n_rungs > 32 || m.maxswaps > 32 || (m.dim & 3) != 0) return 0;
  return ptg_xp_shared_bytes(m.n_rungs, m.dim, m.n_props, cpl_for(m.dim)) <= (size_t)227 * 1024 ? 1 : 0;
}
size_t ptg_xpstep_scratch_doubles(const PtgModel &m) { return (size_t)m.n_chains * 32 * cpl_for(m.dim); }
cudaError_t ptg_launch_xpstep(const PtgModel &m, const PtgState &s, long long step0, int n_steps, int trans_off, double *scratch, cudaStream_t st) {
  switch (cpl_for(m.dim)) {
  case 1: return xpstep_t<1>(m, s, step0, n_steps, trans_off, scratch, st);
  case 2: return xpstep_t<2>(m, s, step0, n_steps, trans_off, scratch, st);
  default: return xpstep_t<4>(m, s, step0, n_steps, trans_off, scratch, st);
  }
}

