// ptg_api.cu -- the C ABI of include/ptmcmc_b200.h: handle management, model set-up, kernel dispatch, read-back.
// Host logic only; all chain arithmetic happens in the kernels of ptg_kernels.cuh / ptg_wide.cuh.
// There is no CPU fallback: every entry point that computes needs a CUDA device.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <vector>
#include "ptg_launch.h"
#include "ptg_device.cuh"

static thread_local char g_err[512] = "";
extern "C" const char *ptg_last_error(void) { return g_err; }
extern "C" int ptg_abi_version(void) { return PTG_ABI_VERSION; }
static int fail(int code, const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                                     \
  do {                                                                                                     \
    cudaError_t e__ = (expr);                                                                              \
    if (e__ != cudaSuccess) return fail(PTG_ECUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

struct HostProp {
  ptg_proposal p;
  std::vector<double> sigmas, transform;
};

struct ptg_handle {
  ptg_config cfg;
  PtgModel m;
  PtgState s;
  cudaStream_t stream;
  bool own_stream;
  bool have_space, have_prior, have_like, have_props, inited, model_uploaded;
  bool model_dirty;                // a set_* call changed the model since the last upload
  std::vector<HostProp> props;
  double Tpow;
  double adapt_rate; int de_mixing; double de_Tmix;   // ptg_set_proposal_options
  int nest_first, nest_count; double nest_share, nest_hot, nest_adapt;   // ptg_set_nested_set
  std::vector<double> lparams, ldata, betas;
  std::vector<int32_t> lower, upper; std::vector<double> xmin, xmax; std::vector<PtgPrior1D> prior; // every dimension (dim may exceed 16)
  int wide_trans_off;              // offset of the (first) eigen-rotation matrix in prop_data, -1 = none
  bool wide;                       // dim > 16: warp-per-chain kernels (ptg_wide.cuh)
  std::vector<void *> allocs;       // every device allocation (freed in destroy)
  double *d_lparams, *d_ldata, *d_prop_data, *d_bins;
  double *d_xscratch;              // pipelined wide kernel: one published-state row per chain
  double *d_tape_u, *d_tape_z; long long *d_u_end, *d_z_end;
  long long istep;
  double *d_scratch; size_t scratch_bytes;
  double *h_pinned; size_t pinned_bytes;
  int kernel_choice; long long launches;
  ptg_batch_loglike_fn cb_fn; void *cb_user;   // host-callback likelihood
  int32_t *d_attempt, *d_nopen;
  std::vector<double> cb_x, cb_like, cb_xc, cb_lc; std::vector<int32_t> cb_flags; std::vector<int64_t> cb_idx;
  // rung-sharded ladders with the exchange fused into the step kernel (ptg_xchg_*)
  void *d_xchg; size_t xchg_bytes; void *peer_lo, *peer_hi; bool peer_lo_ipc, peer_hi_ipc;
  PtgXchg xchg, xchg_cur; bool xchg_launch;  // xchg_cur: parameters of the exchange launch in flight through ptg_step (xchg_launch = false: plain step)
  cudaStream_t copy_stream; cudaEvent_t ev_gathered[2], ev_copied[2]; double *d_stage[2]; size_t stage_bytes[2]; bool stage_used[2]; int stage_next; // ptg_step_host_begin
  int *h_abort; int *d_abort;               // watchdog word of the fused exchange (mapped pinned host memory) and its device alias
  bool dead;                                // a boundary exchange was aborted: the ladders of this rank are out of step with their neighbours
};

template <typename T>
static int dev_alloc(ptg_handle *h, T **p, size_t n, bool zero = true) {
  void *q = nullptr;
  size_t bytes = (n ? n : 1) * sizeof(T);
  cudaError_t e = cudaMalloc(&q, bytes);
  if (e != cudaSuccess) return fail(PTG_ENOMEM, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
  if (zero) cudaMemsetAsync(q, 0, bytes, h->stream);
  h->allocs.push_back(q);
  *p = (T *)q;
  return 0;
}

__global__ void ptg_uniform_lprior_kernel(const PtgPrior1D *prior, int dim, double *out) {
  // log(prod_i 1/(b_i-a_i)): same operation order as uniform_dist_product::evaluate (probability_function.cc:156-166)
  double result = 1;
  for (int i = 0; i < dim; i++) result *= 1 / (prior[i].b - prior[i].a);
  *out = log(result);
}
__global__ void ptg_fill_kernel(double *p, long long n, double v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void ptg_fill_ll_kernel(long long *p, long long n, long long v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void ptg_fill_i_kernel(int32_t *p, long long n, int32_t v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void ptg_ladder_init_kernel(PtgState s, int R, long long n_chains) {
  // chain.cc:1345-1358
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chains) return;
  int r = (int)(c % R);
  s.instances[c] = r;
  s.directions[c] = (r == 0) ? -1 : (r == R - 1 ? 1 : 0);
  s.ups[c] = 0; s.downs[c] = 0; s.swap_count[c] = 0; s.swap_accept[c] = 0;
}
__global__ void ptg_sum_nhist_kernel(const long long *nhist, long long n, unsigned long long *out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long v = (i < n) ? (unsigned long long)nhist[i] : 0ull;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}
// newest n_out stored samples of every ladder's cold chain -> out_x[L][n_out][D], out_lp[L][n_out], out_ll[L][n_out]
__global__ void ptg_gather_cold_kernel(PtgModel m, PtgState s, int n_out, double *out_x, double *out_lp, double *out_ll) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)m.n_ladders * n_out;
  if (t >= total) return;
  long long l = t / n_out; int k = (int)(t - l * n_out);
  long long c = l * m.n_rungs;
  long long nsize = s.nsize[c];
  long long logical = nsize - n_out + k; // k-th of the newest n_out
  const int D = m.dim;
  if (logical < 0) { for (int i = 0; i < D; i++) out_x[t * D + i] = CUDART_NAN; out_lp[t] = CUDART_NAN; out_ll[t] = CUDART_NAN; return; }
  const long long rec = c * m.hist_cap + (logical % m.hist_cap);
  const double *h = s.hist + rec * m.hx;
  for (int i = 0; i < D; i++) out_x[t * D + i] = h[i];
  out_lp[t] = s.hist_lp[2 * rec]; out_ll[t] = s.hist_lp[2 * rec + 1];
}

// [chain][dim] -> [dim][chain]
__global__ void ptg_transpose_kernel(const double *src, double *dst, long long n, int d) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * d) return;
  long long c = t / d; int k = (int)(t - c * d);
  dst[(long long)k * n + c] = src[t];
}

static inline int grid_for(long long n, int b = 256) { return (int)((n + b - 1) / b); }

// ------------------------------------------------------------------------------------------------- lifecycle
extern "C" int ptg_create(const ptg_config *cfg, ptg_handle **out) {
  if (!cfg || !out) return fail(PTG_EINVAL, "null argument");
  if (cfg->abi_version != PTG_ABI_VERSION) return fail(PTG_EINVAL, "abi version mismatch (%d != %d)", cfg->abi_version, PTG_ABI_VERSION);
  if (cfg->dim < 1 || cfg->dim > PTG_MAX_DIM) return fail(PTG_EINVAL, "dim out of range");
  if (cfg->n_rungs < 1 || cfg->n_rungs > PTG_MAX_RUNGS || cfg->n_ladders < 1) return fail(PTG_EINVAL, "bad ladder shape");
  if (cfg->save_every < 1 || cfg->n_init < 1) return fail(PTG_EINVAL, "save_every and n_init must be >= 1");
  if (cfg->swap_mode != PTG_SWAP_REFERENCE && cfg->swap_mode != PTG_SWAP_EVEN_ODD) return fail(PTG_EINVAL, "bad swap_mode");
  if (cfg->swap_mode == PTG_SWAP_EVEN_ODD && cfg->evolve_rate > 0)
    return fail(PTG_EINVAL, "temperature evolution is defined for the reference swap schedule only");
  if (cfg->rng_mode != PTG_RNG_PHILOX && cfg->rng_mode != PTG_RNG_TAPE) return fail(PTG_EINVAL, "bad rng_mode");
  if ((int64_t)cfg->n_ladders * cfg->n_rungs >= (1ll << 31) || cfg->ladder_offset < 0 || cfg->ladder_offset + cfg->n_ladders >= (1ll << 31))
    return fail(PTG_EINVAL, "chain and global ladder ids must stay below 2^31");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(PTG_ECUDA, "no CUDA device: %s (the ptg engine has no CPU fallback)", e != cudaSuccess ? cudaGetErrorString(e) : "count=0");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(PTG_EINVAL, "device %d out of range (%d devices)", cfg->device, ndev);
  CUDA_TRY(cudaSetDevice(cfg->device));
  {
    // DE gathers read single 32-byte sectors at random ring slots: ask L2 not to fetch a second sector with every miss (the default
    // fetch granularity is 64 bytes).  A hint; PTG_L2_FETCH_GRANULARITY = 32 | 64 | 128 overrides it for experiments.
    const char *eg = getenv("PTG_L2_FETCH_GRANULARITY");
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, eg ? (size_t)atoi(eg) : (size_t)32);
    cudaGetLastError();
  }
  bool supported_dim = false;
#define X(D) if (cfg->dim == D) supported_dim = true;
  PTG_DIM_LIST(X)
#undef X
  if (!supported_dim && cfg->dim <= PTG_TPC_MAX_DIM)
    return fail(PTG_EINVAL, "dim=%d: thread-per-chain kernels are instantiated for 1-16; dims 17-128 run the warp-per-chain kernels", cfg->dim);
  if (cfg->dim > PTG_TPC_MAX_DIM && cfg->n_rungs > 32) return fail(PTG_EINVAL, "dim > 16 needs n_rungs <= 32 (one CTA per ladder, one warp per rung)");

  ptg_handle *h = new ptg_handle();
  memset(&h->m, 0, sizeof(h->m)); memset(&h->s, 0, sizeof(h->s));
  h->cfg = *cfg;
  h->wide = cfg->dim > PTG_TPC_MAX_DIM; h->wide_trans_off = -1;
  h->lower.assign(cfg->dim, PTG_BOUND_OPEN); h->upper.assign(cfg->dim, PTG_BOUND_OPEN);
  h->xmin.assign(cfg->dim, -INFINITY); h->xmax.assign(cfg->dim, INFINITY); h->prior.resize(cfg->dim);
  h->have_space = h->have_prior = h->have_like = h->have_props = h->inited = h->model_uploaded = false; h->model_dirty = true;
  h->d_lparams = h->d_ldata = h->d_prop_data = h->d_bins = nullptr;
  h->d_tape_u = h->d_tape_z = nullptr; h->d_u_end = h->d_z_end = nullptr;
  h->copy_stream = nullptr; h->d_stage[0] = h->d_stage[1] = nullptr; h->stage_bytes[0] = h->stage_bytes[1] = 0; h->stage_used[0] = h->stage_used[1] = false; h->stage_next = 0;
  h->xchg_launch = false; h->h_abort = h->d_abort = nullptr; h->dead = false; h->d_xchg = nullptr; h->peer_lo = h->peer_hi = nullptr; h->peer_lo_ipc = h->peer_hi_ipc = false;
  h->kernel_choice = PTG_KERNEL_AUTO; h->launches = 0; h->cb_fn = nullptr; h->cb_user = nullptr; h->d_attempt = h->d_nopen = nullptr; h->istep = 0; h->Tpow = 0; h->d_scratch = nullptr; h->scratch_bytes = 0; h->h_pinned = nullptr; h->pinned_bytes = 0;
  cudaError_t es = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (es != cudaSuccess) { delete h; return fail(PTG_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(es)); }
  h->own_stream = true;

  PtgModel &m = h->m;
  m.dim = cfg->dim; m.n_rungs = cfg->n_rungs; m.n_ladders = cfg->n_ladders;
  m.n_chains = (int64_t)cfg->n_rungs * cfg->n_ladders;
  m.save_every = cfg->save_every; m.n_init = cfg->n_init;
  int cap = cfg->hist_capacity > 0 ? cfg->hist_capacity : cfg->n_init + 1024;
  if (cap < cfg->n_init) cap = cfg->n_init;
  m.hist_cap = cap;
  int maxswaps = (int)(1 + 2 * cfg->swap_rate * cfg->n_rungs); // chain.cc:1192
  if (maxswaps > PTG_SWAP_SLOTS) maxswaps = PTG_SWAP_SLOTS;
  if (maxswaps < 1) maxswaps = 1;
  m.maxswaps = maxswaps;
  m.swap_mode = cfg->swap_mode; m.record_full = cfg->record_level == PTG_RECORD_FULL; m.wrap_in_set = 1; m.zero_valid = 1;
  m.trace_steps = cfg->trace_steps;
  m.swap_rate = cfg->swap_rate; m.dprior_min = cfg->dprior_min; m.evolve_rate = cfg->evolve_rate; m.evolve_lpost_cut = cfg->evolve_lpost_cut;
  m.seed = cfg->seed; m.ladder_offset = cfg->ladder_offset;
  for (int i = 0; i < PTG_TPC_MAX_DIM; i++) { m.lower[i] = m.upper[i] = PTG_BOUND_OPEN; m.xmin[i] = -INFINITY; m.xmax[i] = INFINITY; }

  // device state
  PtgState &s = h->s;
  const size_t n = (size_t)m.n_chains, d = (size_t)m.dim;
  int rc = 0;
  rc |= dev_alloc(h, &s.cur_x, n * d); rc |= dev_alloc(h, &s.lpost, n); rc |= dev_alloc(h, &s.llike, n);
  rc |= dev_alloc(h, &s.lprior, n); rc |= dev_alloc(h, &s.beta, n); rc |= dev_alloc(h, &s.map_lpost, n); rc |= dev_alloc(h, &s.map_x, n * d);
  rc |= dev_alloc(h, &s.nhist, n); rc |= dev_alloc(h, &s.nsize, n); rc |= dev_alloc(h, &s.ntries, n); rc |= dev_alloc(h, &s.naccept, n);
  rc |= dev_alloc(h, &s.last_type, n);
  m.hx = PTG_HX(cfg->dim);
  rc |= dev_alloc(h, &s.hist, n * (size_t)cap * (size_t)m.hx, false); rc |= dev_alloc(h, &s.hist_lp, n * (size_t)cap * 2, false);
  if (m.record_full) {
    rc |= dev_alloc(h, &s.hist_acc, n * (size_t)cap, false); rc |= dev_alloc(h, &s.hist_beta, n * (size_t)cap, false);
    rc |= dev_alloc(h, &s.hist_type, n * (size_t)cap, false);
  }
  rc |= dev_alloc(h, &s.swap_count, n); rc |= dev_alloc(h, &s.swap_accept, n);
  rc |= dev_alloc(h, &s.directions, n); rc |= dev_alloc(h, &s.ups, n); rc |= dev_alloc(h, &s.downs, n); rc |= dev_alloc(h, &s.instances, n);
  rc |= dev_alloc(h, &s.u_pos, n + (size_t)m.n_ladders); rc |= dev_alloc(h, &s.z_pos, n + (size_t)m.n_ladders);
  rc |= dev_alloc(h, &s.err, 1);
  if (cfg->trace_steps > 0) { rc |= dev_alloc(h, &s.trace_lhr, n * (size_t)cfg->trace_steps); rc |= dev_alloc(h, &s.trace_code, n * (size_t)cfg->trace_steps); }
  if (rc) { ptg_destroy(h); return PTG_ENOMEM; }

  // geometric ladder (chain.cc:1181-1183,1339): temps[i]=temps[i-1]*tratio ; invtemp = 1/temps[i]
  h->betas.resize(n);
  double tratio = (m.n_rungs > 1) ? exp(log(cfg->Tmax) / (m.n_rungs - 1)) : 1.0;
  for (int l = 0; l < m.n_ladders; l++) {
    double temp = 1;
    for (int r = 0; r < m.n_rungs; r++) { if (r > 0) temp = temp * tratio; h->betas[(size_t)l * m.n_rungs + r] = 1 / temp; }
  }
  *out = h;
  return 0;
}

extern "C" int ptg_destroy(ptg_handle *h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  if (h->peer_lo && h->peer_lo_ipc) cudaIpcCloseMemHandle(h->peer_lo);
  if (h->peer_hi && h->peer_hi_ipc) cudaIpcCloseMemHandle(h->peer_hi);
  for (void *p : h->allocs) cudaFree(p);
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  if (h->h_abort) cudaFreeHost(h->h_abort);
  if (h->copy_stream) {
    cudaStreamSynchronize(h->copy_stream);
    for (int b = 0; b < 2; b++) { cudaEventDestroy(h->ev_gathered[b]); cudaEventDestroy(h->ev_copied[b]); if (h->d_stage[b]) cudaFree(h->d_stage[b]); }
    cudaStreamDestroy(h->copy_stream);
  }
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

// ------------------------------------------------------------------------------------------------- set-up
extern "C" int ptg_set_space(ptg_handle *h, const int32_t *lt, const int32_t *ut, const double *xmin, const double *xmax) {
  if (!h || !lt || !ut || !xmin || !xmax) return fail(PTG_EINVAL, "null argument");
  PtgModel &m = h->m;
  bool zero_valid = true;
  for (int i = 0; i < m.dim; i++) {
    h->lower[i] = lt[i]; h->upper[i] = ut[i]; h->xmin[i] = xmin[i]; h->xmax[i] = xmax[i];
    if (i < PTG_TPC_MAX_DIM) { m.lower[i] = lt[i]; m.upper[i] = ut[i]; m.xmin[i] = xmin[i]; m.xmax[i] = xmax[i]; }
    // validity of the zero vector decides the validity of every state built by state::add / scalar_mult
    // (states.cc:168-176,194-204).  Pure comparisons: no arithmetic result is kept.
    if (zero_valid) {
      bool wl = lt[i] == PTG_BOUND_WRAP, wu = ut[i] == PTG_BOUND_WRAP;
      if (wl != wu) zero_valid = false;
      else if (wl) { if (xmax[i] - xmin[i] <= 0) zero_valid = false; }
      else if (lt[i] == PTG_BOUND_REFLECT && ut[i] == PTG_BOUND_REFLECT) { if (xmax[i] - xmin[i] <= 0) zero_valid = false; }
      else {
        double x = 0;
        if (lt[i] == PTG_BOUND_REFLECT && x < xmin[i]) x = xmin[i] + (xmin[i] - x);
        else if (ut[i] == PTG_BOUND_REFLECT && x > xmax[i]) x = xmax[i] - (x - xmax[i]);
        if (lt[i] == PTG_BOUND_LIMIT && x < xmin[i]) zero_valid = false;
        if (ut[i] == PTG_BOUND_LIMIT && x > xmax[i]) zero_valid = false;
      }
    }
  }
  m.zero_valid = zero_valid ? 1 : 0;
  m.any_bound = 0;
  for (int i = 0; i < m.dim; i++) if (lt[i] != PTG_BOUND_OPEN || ut[i] != PTG_BOUND_OPEN) m.any_bound = 1;
  h->have_space = true;
  h->model_dirty = true;
  return 0;
}

extern "C" int ptg_set_prior(ptg_handle *h, const int32_t *type, const double *a, const double *b) {
  if (!h || !type || !a || !b) return fail(PTG_EINVAL, "null argument");
  PtgModel &m = h->m;
  bool all_uniform = true;
  for (int i = 0; i < m.dim; i++) {
    PtgPrior1D &p = h->prior[i];
    p.kind = type[i]; p.pad = 0; p.a = a[i]; p.b = b[i]; p.norm = p.cdfoff = p.la = p.lb = 0;
    if (p.kind != PTG_PRIOR_UNIFORM) all_uniform = false;
    if (p.kind == PTG_PRIOR_POLAR) { // ProbabilityDist.h:187-192
      double lo = a[i], hi = b[i];
      if (lo < 0) lo = 0;
      if (hi > M_PI) hi = M_PI;
      p.norm = -cos(hi) + cos(lo); p.cdfoff = -cos(lo) / p.norm;
    } else if (p.kind == PTG_PRIOR_COPOLAR) { // ProbabilityDist.h:228-233
      double lo = a[i], hi = b[i];
      if (lo < -M_PI / 2) lo = -M_PI / 2;
      if (hi > M_PI / 2) hi = M_PI / 2;
      p.norm = sin(hi) - sin(lo); p.cdfoff = sin(lo) / p.norm;
    } else if (p.kind == PTG_PRIOR_LOG) {
      if (a[i] <= 0 || b[i] <= a[i]) return fail(PTG_EINVAL, "log prior needs 0 < xmin < xmax");
      p.la = log(a[i]); p.lb = log(b[i]);
    } else if (p.kind != PTG_PRIOR_UNIFORM && p.kind != PTG_PRIOR_GAUSSIAN && p.kind != PTG_PRIOR_GAUSSIAN_WRAPPED) return fail(PTG_EINVAL, "bad prior type %d", p.kind);
  }
  for (int i = 0; i < m.dim && i < PTG_TPC_MAX_DIM; i++) m.prior[i] = h->prior[i];
  m.all_uniform_prior = all_uniform ? 1 : 0;
  h->have_prior = true;
  h->model_dirty = true;
  return 0;
}

extern "C" int ptg_set_likelihood(ptg_handle *h, int32_t kind, const double *params, int32_t n_params, const double *data, int64_t n_data) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  PtgModel &m = h->m;
  const int d = m.dim;
  int need = 0;
  switch (kind) {
  case PTG_LIKE_FLAT: need = 0; break;
  case PTG_LIKE_GAUSS_ISO: need = 2 + d; break;
  case PTG_LIKE_SINES: need = 2 + 3 * d; break;
  case PTG_LIKE_SHELL2D: need = 5; break;
  case PTG_LIKE_SHELLS: need = 7; break;
  case PTG_LIKE_POLY_CHI2: case PTG_LIKE_SINUSOID_CHI2: case PTG_LIKE_GAUSS_FULLCOV: need = 1; break;
  case PTG_LIKE_HOST_CALLBACK: return fail(PTG_EINVAL, "use ptg_register_evaluate_log for a host-callback likelihood");
  default: return fail(PTG_EINVAL, "unknown likelihood kind %d", kind);
  }
  if (n_params < need || (need && !params)) return fail(PTG_EINVAL, "likelihood kind %d needs %d params", kind, need);
  if ((kind == PTG_LIKE_POLY_CHI2 || kind == PTG_LIKE_SINUSOID_CHI2) && (n_data <= 0 || n_data % 3 || !data))
    return fail(PTG_EINVAL, "chi2 likelihoods need data = [x[N], y[N], var[N]] (n_data = 3N)");
  if (kind == PTG_LIKE_SINUSOID_CHI2 && d % 3) return fail(PTG_EINVAL, "sinusoid model needs dim = 3k");
  if (kind == PTG_LIKE_SHELL2D && d != 2) return fail(PTG_EINVAL, "the 2-D shell likelihood needs dim = 2 (example.cc:147-222)");
  if ((kind == PTG_LIKE_SHELL2D || kind == PTG_LIKE_SHELLS) && h->wide) return fail(PTG_EINVAL, "shell likelihoods run in the thread-per-chain kernels (dim <= 16)");
  if (kind == PTG_LIKE_GAUSS_FULLCOV && (n_data != (int64_t)d * d || !data)) return fail(PTG_EINVAL, "fullcov needs Cinv[dim*dim]");
  h->lparams.assign(params, params + (n_params > 0 ? n_params : 0));
  if (h->lparams.empty()) h->lparams.push_back(0.0);
  h->ldata.assign(data ? data : nullptr, data ? data + n_data : nullptr);
  m.like_kind = kind; m.n_lparams = n_params; m.n_ldata = n_data;
  m.like_nsum = 0;
  if (kind == PTG_LIKE_POLY_CHI2 || kind == PTG_LIKE_SINUSOID_CHI2) {
    // nsum = sum_i log(S_i) (bayesian.hh:613) is independent of the state: evaluated once at set-up
    int64_t N = n_data / 3; double nsum = 0;
    for (int64_t i = 0; i < N; i++) nsum += log(data[2 * N + i]);
    m.like_nsum = nsum;
    // uniform abscissae (a sampled time series): the production sinusoid functor advances sin / cos by a rotation instead of calling
    // libm at every sample (ptg_wide_mma.cuh)
    m.like_uniform_t = 0; m.like_t0 = data[0]; m.like_dt = 0;
    if (N >= 2) {
      const double dt = (data[N - 1] - data[0]) / (double)(N - 1);
      double worst = 0, scale = fabs(data[0]) > fabs(data[N - 1]) ? fabs(data[0]) : fabs(data[N - 1]);
      for (int64_t i = 0; i < N; i++) { const double e = fabs(data[i] - (data[0] + (double)i * dt)); if (e > worst) worst = e; }
      if (dt != 0 && worst <= 1e-13 * (scale > fabs(dt) ? scale : fabs(dt))) { m.like_uniform_t = 1; m.like_dt = dt; }
    }
  }
  h->have_like = true;
  h->model_dirty = true;
  return 0;
}

extern "C" int ptg_register_evaluate_log(ptg_handle *h, ptg_batch_loglike_fn fn, void *user) {
  if (!h || !fn) return fail(PTG_EINVAL, "null argument");
  if (h->inited) return fail(PTG_EINVAL, "register the likelihood before initialising");
  if (h->wide) return fail(PTG_EINVAL, "host-callback likelihoods need dim <= 16");
  if (h->cfg.rng_mode != PTG_RNG_PHILOX) return fail(PTG_EINVAL, "host-callback likelihoods run with Philox draws");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  PtgModel &m = h->m; PtgState &s = h->s;
  const size_t n = (size_t)m.n_chains, d = (size_t)m.dim;
  if (!s.pend_x) {
    int rc = 0;
    rc |= dev_alloc(h, &s.pend_x, n * d); rc |= dev_alloc(h, &s.pend_lprior, n); rc |= dev_alloc(h, &s.pend_lh, n); rc |= dev_alloc(h, &s.pend_like, n);
    rc |= dev_alloc(h, &s.pend_type, n); rc |= dev_alloc(h, &s.pend_flags, n); rc |= dev_alloc(h, &s.pend_w, 2 * n);
    rc |= dev_alloc(h, &h->d_attempt, n); rc |= dev_alloc(h, &h->d_nopen, 1);
    if (rc) return PTG_ENOMEM;
  }
  h->cb_fn = fn; h->cb_user = user;
  h->cb_x.resize(n * d); h->cb_like.resize(n); h->cb_xc.resize(n * d); h->cb_lc.resize(n); h->cb_flags.resize(n); h->cb_idx.resize(n);
  h->lparams.assign(1, 0.0); h->ldata.clear();
  m.like_kind = PTG_LIKE_HOST_CALLBACK; m.n_lparams = 0; m.n_ldata = 0; m.like_nsum = 0;
  h->have_like = true; h->model_dirty = true;
  return 0;
}

// host-callback mode: fetch the parked proposals, evaluate the caller's likelihood for the gated ones in one call, send the values back
static int callback_round(ptg_handle *h) {
  PtgModel &m = h->m; PtgState &s = h->s;
  const size_t n = (size_t)m.n_chains, d = (size_t)m.dim;
  CUDA_TRY(cudaMemcpyAsync(h->cb_x.data(), s.pend_x, n * d * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->cb_flags.data(), s.pend_flags, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  int64_t cnt = 0;
  for (size_t c = 0; c < n; c++)
    if ((h->cb_flags[c] & 2) && h->cb_flags[c] != 8) { memcpy(&h->cb_xc[(size_t)cnt * d], &h->cb_x[c * d], d * sizeof(double)); h->cb_idx[cnt++] = (int64_t)c; }
  if (cnt) h->cb_fn(h->cb_user, h->cb_xc.data(), cnt, h->cb_lc.data());
  for (size_t c = 0; c < n; c++) h->cb_like[c] = -INFINITY;
  for (int64_t k = 0; k < cnt; k++) h->cb_like[(size_t)h->cb_idx[k]] = h->cb_lc[k];
  CUDA_TRY(cudaMemcpyAsync(s.pend_like, h->cb_like.data(), n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return 0;
}

extern "C" int ptg_set_proposals(ptg_handle *h, int32_t n, const ptg_proposal *props, double Tpow, int32_t wrap_in_set) {
  if (!h || !props) return fail(PTG_EINVAL, "null argument");
  if (n < 1 || n > PTG_MAX_PROPOSALS) return fail(PTG_EINVAL, "bad proposal count %d", n);
  if (!wrap_in_set && n != 1) return fail(PTG_EINVAL, "a bare proposal must be single");
  h->props.clear();
  for (int i = 0; i < n; i++) {
    HostProp hp; hp.p = props[i];
    if (props[i].kind == PTG_PROP_GAUSS) {
      if (!props[i].sigmas) return fail(PTG_EINVAL, "gaussian proposal %d without sigmas", i);
      hp.sigmas.assign(props[i].sigmas, props[i].sigmas + h->m.dim);
      if (props[i].transform) hp.transform.assign(props[i].transform, props[i].transform + (size_t)h->m.dim * h->m.dim);
    } else if (props[i].kind == PTG_PROP_DE) {
      if (!(props[i].reduce_gamma > 0)) return fail(PTG_EINVAL, "DE proposal %d needs reduce_gamma > 0", i);
    } else if (props[i].kind != PTG_PROP_PRIOR_DRAW) return fail(PTG_EINVAL, "unknown proposal kind %d", props[i].kind);
    h->props.push_back(hp);
  }
  h->Tpow = Tpow; h->m.wrap_in_set = wrap_in_set ? 1 : 0; h->m.n_props = n;
  h->adapt_rate = 0; h->de_mixing = 0; h->de_Tmix = 1;
  h->nest_first = h->nest_count = 0; h->nest_share = h->nest_hot = h->nest_adapt = 0;
  h->have_props = true;
  h->model_dirty = true;
  return 0;
}

static int restrict_to_warp_kernel(const ptg_handle *h, const char *what);
// proposal_distribution_set's adapt_rate (proposal_distribution.cc:61,132-166) and differential_evolution::support_mixing /
// mix_temperatures_more (proposal_distribution.hh:403,412)
extern "C" int ptg_set_proposal_options(ptg_handle *h, double adapt_rate, int32_t de_mixing, double de_Tmix) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  if (!h->have_props) return fail(PTG_EINVAL, "set the proposals first");
  if (h->inited) return fail(PTG_EINVAL, "proposal options must be set before initialising");
  if (adapt_rate != 0 || de_mixing) { int rc = restrict_to_warp_kernel(h, "adaptive shares / temperature mixing"); if (rc) return rc; }
  if (adapt_rate != 0 && !h->m.wrap_in_set) return fail(PTG_EINVAL, "adaptive shares need a proposal set (wrap_in_set = 1)");
  if (de_mixing) {
    if (h->m.wrap_in_set || h->props[0].p.kind != PTG_PROP_DE)
      return fail(PTG_EINVAL, "temperature mixing needs a bare differential-evolution proposal: inside a set the reference never mixes (chain.cc:1375)");
    if (h->props[0].p.unlikely_alpha != 0) return fail(PTG_EINVAL, "temperature mixing needs unlikely_alpha = 0");
    if (!(de_Tmix > 0)) return fail(PTG_EINVAL, "de_Tmix must be positive");
  }
  h->adapt_rate = adapt_rate; h->de_mixing = de_mixing ? 1 : 0; h->de_Tmix = de_Tmix;
  h->model_dirty = true;
  return 0;
}

static int restrict_to_warp_kernel(const ptg_handle *h, const char *what) {
  if (h->wide) return fail(PTG_EINVAL, "%s run in the thread-per-chain warp kernel: dim <= 16", what);
  if (h->m.n_rungs > 32 || h->m.maxswaps > 32) return fail(PTG_EINVAL, "%s need n_rungs <= 32 and at most 32 swap trials per step", what);
  if (h->m.like_kind == PTG_LIKE_HOST_CALLBACK) return fail(PTG_EINVAL, "%s are not available with a host-callback likelihood", what);
  return 0;
}
// a nested proposal_distribution_set inside the top-level one (ptmcmc.cc:70-72,123-143)
extern "C" int ptg_set_nested_set(ptg_handle *h, int32_t first, int32_t count, double share, double hot_share, double adapt_rate) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  if (!h->have_props || !h->m.wrap_in_set) return fail(PTG_EINVAL, "a nested set needs a proposal set (ptg_set_proposals with wrap_in_set = 1 first)");
  if (h->inited) return fail(PTG_EINVAL, "the nested set must be declared before initialising");
  if (first < 0 || count < 1 || first + count > h->m.n_props) return fail(PTG_EINVAL, "nested set [%d, %d) is outside the %d members", first, first + count, h->m.n_props);
  if (!(share >= 0)) return fail(PTG_EINVAL, "bad share");
  int rc = restrict_to_warp_kernel(h, "nested proposal sets"); if (rc) return rc;
  h->nest_first = first; h->nest_count = count; h->nest_share = share; h->nest_hot = hot_share; h->nest_adapt = adapt_rate;
  h->model_dirty = true;
  return 0;
}

static void compute_set(int n, const double *shares_in, const double *hot_in, double Tpow, double beta, double *bin_max, double *shares_out, double *hot_out);
static void top_level_shares(const ptg_handle *h, double *sh, double *hot);
extern "C" int ptg_get_proposal_shares(ptg_handle *h, double *shares) {
  if (!h || !shares) return fail(PTG_EINVAL, "null argument");
  if (!h->model_uploaded) return fail(PTG_EINVAL, "not initialised");
  const PtgModel &m = h->m;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  if (h->s.ad_shares) {
    CUDA_TRY(cudaMemcpyAsync(shares, h->s.ad_shares, (size_t)m.n_chains * m.n_bins * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
  } else {
    double sh[PTG_MAX_PROPOSALS], hot[PTG_MAX_PROPOSALS], bn[PTG_MAX_PROPOSALS], out[2 * PTG_MAX_PROPOSALS];
    top_level_shares(h, sh, hot);
    compute_set(m.n_slots, sh, hot, h->Tpow, 1.0, bn, out, nullptr);
    for (int j = 0; j < m.nest_count; j++) { sh[j] = h->props[m.nest_first + j].p.share; hot[j] = 0; }
    if (m.nest_count) compute_set(m.nest_count, sh, hot, 0.0, 1.0, bn, out + m.n_slots, nullptr);
    for (int64_t c = 0; c < m.n_chains; c++) for (int k = 0; k < m.n_bins; k++) shares[c * m.n_bins + k] = out[k];
  }
  return 0;
}

extern "C" int ptg_set_betas(ptg_handle *h, const double *betas) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  if (betas) h->betas.assign(betas, betas + h->m.n_chains);
  h->model_dirty = true;
  return 0;
}

extern "C" int ptg_seed(ptg_handle *h, uint64_t seed) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  h->cfg.seed = seed; h->m.seed = seed;
  return 0;
}

extern "C" int ptg_inject_tapes(ptg_handle *h, const double *u, const int64_t *u_off, const double *z, const int64_t *z_off) {
  if (!h || !u || !u_off || !z || !z_off) return fail(PTG_EINVAL, "null argument");
  if (h->cfg.rng_mode != PTG_RNG_TAPE) return fail(PTG_EINVAL, "engine was not created with PTG_RNG_TAPE");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t ns = (size_t)h->m.n_chains + h->m.n_ladders;
  int rc = 0;
  rc |= dev_alloc(h, &h->d_tape_u, (size_t)u_off[ns], false); rc |= dev_alloc(h, &h->d_tape_z, (size_t)z_off[ns], false);
  rc |= dev_alloc(h, &h->d_u_end, ns, false); rc |= dev_alloc(h, &h->d_z_end, ns, false);
  if (rc) return PTG_ENOMEM;
  CUDA_TRY(cudaMemcpyAsync(h->d_tape_u, u, (size_t)u_off[ns] * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_tape_z, z, (size_t)z_off[ns] * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->s.u_pos, u_off, ns * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->s.z_pos, z_off, ns * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_u_end, u_off + 1, ns * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_z_end, z_off + 1, ns * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->s.tape_u = h->d_tape_u; h->s.tape_z = h->d_tape_z; h->s.u_end = h->d_u_end; h->s.z_end = h->d_z_end;
  return 0;
}

extern "C" int ptg_inject_tape_marks(ptg_handle *h, int64_t n_steps, const int64_t *u_mark, const int64_t *z_mark) {
  if (!h || !u_mark || !z_mark || n_steps < 1) return fail(PTG_EINVAL, "bad argument");
  if (h->cfg.rng_mode != PTG_RNG_TAPE || !h->s.tape_u) return fail(PTG_EINVAL, "inject the tapes first");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t cnt = (size_t)n_steps * ((size_t)h->m.n_chains + h->m.n_ladders);
  long long *du = nullptr, *dz = nullptr;
  if (dev_alloc(h, &du, cnt, false) || dev_alloc(h, &dz, cnt, false)) return PTG_ENOMEM;
  CUDA_TRY(cudaMemcpyAsync(du, u_mark, cnt * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(dz, z_mark, cnt * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->s.u_mark = du; h->s.z_mark = dz; h->s.n_mark_steps = n_steps;
  return 0;
}

// proposal_distribution_set::reset_bins (proposal_distribution.cc:37-59): run once by the constructor (no chain yet,
// Tfac = 0) and once more by set_chain on each rung's clone (proposal_distribution.hh:336); for one set of n members
static void compute_set(int n, const double *shares_in, const double *hot_in, double Tpow, double beta, double *bin_max, double *shares_out, double *hot_out) {
  double shares[PTG_MAX_PROPOSALS], hot[PTG_MAX_PROPOSALS];
  for (int i = 0; i < n; i++) { shares[i] = shares_in[i]; hot[i] = hot_in[i]; }
  if (Tpow > 0) {
    double sum = 0;
    for (int i = 0; i < n; i++) sum += hot[i];
    if (sum <= 0) for (int i = 0; i < n; i++) hot[i] = shares[i];
    else for (int i = 0; i < n; i++) hot[i] /= sum;
  }
  for (int pass = 0; pass < 2; pass++) {
    double Tfac = 0;
    if (Tpow > 0 && pass == 1) Tfac = 1 - pow(beta, Tpow);
    double sum = 0;
    for (int i = 0; i < n; i++) sum += shares[i];
    double last = 0;
    for (int i = 0; i < n; i++) {
      shares[i] /= sum;
      bin_max[i] = last + shares[i];
      if (Tpow > 0) bin_max[i] += (hot[i] - shares[i]) * Tfac;
      last = bin_max[i];
    }
    double back = bin_max[n - 1];
    for (int i = 0; i < n; i++) bin_max[i] /= back;
  }
  for (int i = 0; i < n; i++) { if (shares_out) shares_out[i] = shares[i]; if (hot_out) hot_out[i] = hot[i]; }
}
// shares / hot shares of the top-level slots: the members outside the nested set, and the nested set itself at slot nest_first
static void top_level_shares(const ptg_handle *h, double *sh, double *hot) {
  const int n = (int)h->props.size(), nc = h->nest_count, nf = h->nest_first;
  const int nslots = nc ? n - nc + 1 : n;
  for (int sl = 0; sl < nslots; sl++) {
    const int mem = (nc == 0 || sl < nf) ? sl : (sl == nf ? -1 : sl + nc - 1);
    sh[sl] = mem < 0 ? h->nest_share : h->props[mem].p.share;
    hot[sl] = mem < 0 ? h->nest_hot : h->props[mem].p.hot_share;
  }
}
// bins (and optionally normalised shares) of every table entry for a chain at inverse temperature beta: [n_slots] top level, then [nest_count]
static void compute_bins(const ptg_handle *h, double beta, double *bin_max, double *shares_out, double *hot_out) {
  const int n = (int)h->props.size(), nc = h->nest_count;
  const int nslots = nc ? n - nc + 1 : n;
  double sh[PTG_MAX_PROPOSALS], hot[PTG_MAX_PROPOSALS];
  top_level_shares(h, sh, hot);
  compute_set(nslots, sh, hot, h->Tpow, beta, bin_max, shares_out, hot_out);
  if (nc) {
    for (int j = 0; j < nc; j++) { sh[j] = h->props[h->nest_first + j].p.share; hot[j] = 0; }
    compute_set(nc, sh, hot, 0.0, beta, bin_max + nslots, shares_out ? shares_out + nslots : nullptr, nullptr); // no thermal weighting inside (ptmcmc.cc:130)
  }
}

static int upload_model(ptg_handle *h) {
  PtgModel &m = h->m;
  if (!h->have_prior || !h->have_like || !h->have_props) return fail(PTG_EINVAL, "set prior, likelihood and proposals before initialising");
  const int d = m.dim;
  if (h->wide) {
    if (m.like_kind != PTG_LIKE_FLAT && m.like_kind != PTG_LIKE_GAUSS_ISO && m.like_kind != PTG_LIKE_GAUSS_FULLCOV && m.like_kind != PTG_LIKE_POLY_CHI2 &&
        m.like_kind != PTG_LIKE_SINUSOID_CHI2)
      return fail(PTG_EINVAL, "dim > 16: likelihood functors flat, gaussian, full-covariance gaussian and the data chi^2 models are available");
  }
  if ((h->adapt_rate != 0 || h->de_mixing) && m.like_kind == PTG_LIKE_HOST_CALLBACK)
    return fail(PTG_EINVAL, "adaptive shares / temperature mixing are not available with a host-callback likelihood");
  // proposals
  h->wide_trans_off = -1;
  std::vector<double> pdata;
  for (int i = 0; i < m.n_props; i++) {
    const HostProp &hp = h->props[i];
    PtgProp &p = m.props[i];
    memset(&p, 0, sizeof(p));
    p.kind = hp.p.kind; p.snooker = hp.p.snooker; p.g1frac = hp.p.gamma_one_frac; p.ignore_frac = hp.p.ignore_frac;
    p.unlikely_alpha = hp.p.unlikely_alpha; p.reduce_gamma = hp.p.reduce_gamma; p.one_d_frac = hp.p.one_d_frac;
    if (p.kind == PTG_PROP_DE) p.gamma_std = 1.68 / sqrt((double)d) / hp.p.reduce_gamma; // proposal_distribution.cc:495
    if (p.kind == PTG_PROP_GAUSS) {
      p.sigma_off = (int)pdata.size(); pdata.insert(pdata.end(), hp.sigmas.begin(), hp.sigmas.end());
      p.has_transform = hp.transform.empty() ? 0 : 1;
      if (p.has_transform) {
        p.trans_off = (int)pdata.size(); pdata.insert(pdata.end(), hp.transform.begin(), hp.transform.end());
        if (h->wide_trans_off < 0) h->wide_trans_off = p.trans_off;
      }
    }
  }
  if (pdata.empty()) pdata.push_back(0.0);
  m.adapt_rate = h->adapt_rate; m.de_mixing = h->de_mixing; m.de_Tmix = h->de_Tmix; m.Tpow = h->Tpow;
  m.nest_first = h->nest_first; m.nest_count = h->nest_count; m.nest_adapt = h->nest_adapt;
  m.n_slots = h->nest_count ? m.n_props - h->nest_count + 1 : m.n_props; m.n_bins = m.n_slots + h->nest_count;
  if (h->nest_count && (h->wide || m.like_kind == PTG_LIKE_HOST_CALLBACK)) return fail(PTG_EINVAL, "nested proposal sets need dim <= 16 and a device likelihood");
  if (m.adapt_rate != 0 || (m.nest_count && m.nest_adapt != 0)) {
    // adaptive shares: every chain's clone of the set starts from the normalised shares and the bins of its own rung's temperature
    // (constructor + set_chain, proposal_distribution.cc:61-93, .hh:336), all members "last accepted" (:88), adapt_count 0
    const size_t nc = (size_t)m.n_chains, np = (size_t)m.n_bins;
    std::vector<double> sh(nc * np), bn(nc * np);
    for (size_t c = 0; c < nc; c++) compute_bins(h, h->betas[c], &bn[c * np], &sh[c * np], m.hot_norm);
    int rc2 = 0;
    if (!h->s.ad_shares) {
      rc2 |= dev_alloc(h, &h->s.ad_shares, nc * np, false); rc2 |= dev_alloc(h, &h->s.ad_bins, nc * np, false);
      rc2 |= dev_alloc(h, &h->s.ad_last, nc, false); rc2 |= dev_alloc(h, &h->s.ad_count, nc); rc2 |= dev_alloc(h, &h->s.ad_count2, nc);
    }
    if (rc2) return PTG_ENOMEM;
    CUDA_TRY(cudaMemcpyAsync(h->s.ad_shares, sh.data(), sh.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->s.ad_bins, bn.data(), bn.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->s.ad_last, 0xff, nc * sizeof(int32_t), h->stream));
    CUDA_TRY(cudaMemsetAsync(h->s.ad_count, 0, nc * sizeof(int32_t), h->stream));
    CUDA_TRY(cudaMemsetAsync(h->s.ad_count2, 0, nc * sizeof(int32_t), h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  // bins per rung, from ladder 0's initial inverse temperatures
  std::vector<double> bins((size_t)m.n_rungs * m.n_bins);
  for (int r = 0; r < m.n_rungs; r++) compute_bins(h, h->betas[r], &bins[(size_t)r * m.n_bins], nullptr, m.hot_norm);
  if (h->Tpow > 0)
    for (int l = 1; l < m.n_ladders; l++)
      for (int r = 0; r < m.n_rungs; r++)
        if (h->betas[(size_t)l * m.n_rungs + r] != h->betas[r]) return fail(PTG_EINVAL, "Tpow > 0 needs identical ladders");
  int rc = 0;
  rc |= dev_alloc(h, &h->d_lparams, h->lparams.size(), false); // data chi-squared likelihoods: the reciprocals of the N variances ride behind the data block (x | y | S | 1/S); the warp-per-chain
  // production path multiplies by them, the tape-replay kernels keep the reference's division
  std::vector<double> ldata_up = h->ldata;
  if ((m.like_kind == PTG_LIKE_POLY_CHI2 || m.like_kind == PTG_LIKE_SINUSOID_CHI2) && !h->ldata.empty()) {
    const size_t N = h->ldata.size() / 3;
    for (size_t i = 0; i < N; i++) ldata_up.push_back(1.0 / h->ldata[2 * N + i]);
  }
  rc |= dev_alloc(h, &h->d_ldata, ldata_up.size(), false);
  rc |= dev_alloc(h, &h->d_prop_data, pdata.size(), false); rc |= dev_alloc(h, &h->d_bins, bins.size(), false);
  if (rc) return PTG_ENOMEM;
  CUDA_TRY(cudaMemcpyAsync(h->d_lparams, h->lparams.data(), h->lparams.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (!ldata_up.empty()) { CUDA_TRY(cudaMemcpyAsync(h->d_ldata, ldata_up.data(), ldata_up.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream)); CUDA_TRY(cudaStreamSynchronize(h->stream)); }
  CUDA_TRY(cudaMemcpyAsync(h->d_prop_data, pdata.data(), pdata.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->d_bins, bins.data(), bins.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  m.lparams = h->d_lparams; m.ldata = h->d_ldata; m.prop_data = h->d_prop_data; m.bins = h->d_bins;
  {
    int32_t *dl = nullptr, *du = nullptr; double *dmin = nullptr, *dmax = nullptr; PtgPrior1D *dp = nullptr;
    if (dev_alloc(h, &dl, (size_t)d, false) || dev_alloc(h, &du, (size_t)d, false) || dev_alloc(h, &dmin, (size_t)d, false) ||
        dev_alloc(h, &dmax, (size_t)d, false) || dev_alloc(h, &dp, (size_t)d, false)) return PTG_ENOMEM;
    CUDA_TRY(cudaMemcpyAsync(dl, h->lower.data(), d * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(du, h->upper.data(), d * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(dmin, h->xmin.data(), d * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(dmax, h->xmax.data(), d * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(dp, h->prior.data(), d * sizeof(PtgPrior1D), cudaMemcpyHostToDevice, h->stream));
    m.lower_w = dl; m.upper_w = du; m.xmin_w = dmin; m.xmax_w = dmax; m.prior_w = dp;
  }
  m.uniform_lprior = 0;
  if (m.all_uniform_prior) {
    double *d_out; rc = dev_alloc(h, &d_out, 1); if (rc) return rc;
    ptg_uniform_lprior_kernel<<<1, 1, 0, h->stream>>>(m.prior_w, m.dim, d_out);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(&m.uniform_lprior, d_out, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->model_uploaded = true; h->model_dirty = false;
  return 0;
}

static int ensure_scratch(ptg_handle *h, size_t bytes);
static int check_device_error(ptg_handle *h) {
  int32_t e = 0;
  CUDA_TRY(cudaMemcpyAsync(&e, h->s.err, sizeof(e), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (e) CUDA_TRY(cudaMemsetAsync(h->s.err, 0, sizeof(int32_t), h->stream)); // reported once, then cleared
  if (e == 1) return fail(PTG_ETAPE, "injected tape exhausted");
  if (e == 2) return fail(PTG_EINVAL, "proposal set: no member ready");
  if (e == 4) return fail(PTG_ESTUCK, "init: cannot draw a valid state");
  if (e == PTG_ERR_XCHG_TIMEOUT) {
    h->dead = true; // the trial was skipped on this side only: the boundary pair is no longer consistent, stop stepping
    return fail(PTG_EXCHANGE, "cross-GPU boundary exchange aborted (neighbour did not publish before the watchdog fired): this handle accepts no further steps");
  }
  if (e) return fail(PTG_ECUDA, "device error flag %d", e);
  return 0;
}

static int do_init(ptg_handle *h, const double *x_host) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  if (h->inited) return fail(PTG_EINVAL, "already initialised (chain.cc:847-850)");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  if (h->cfg.rng_mode == PTG_RNG_TAPE && !h->s.tape_u && !x_host) return fail(PTG_EINVAL, "PTG_RNG_TAPE: inject tapes before init");
  int rc = (h->model_uploaded && !h->model_dirty) ? 0 : upload_model(h); if (rc) return rc; // uploaded already by ptg_eval and unchanged since
  PtgModel &m = h->m; PtgState &s = h->s;
  const long long n = m.n_chains;
  CUDA_TRY(cudaMemcpyAsync(s.beta, h->betas.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  ptg_fill_kernel<<<grid_for(n), 256, 0, h->stream>>>(s.map_lpost, n, -1e200);        // chain.hh:69
  ptg_fill_ll_kernel<<<grid_for(n), 256, 0, h->stream>>>(s.ntries, n, 1);              // chain.cc:649
  ptg_fill_ll_kernel<<<grid_for(n), 256, 0, h->stream>>>(s.naccept, n, 1);
  ptg_fill_i_kernel<<<grid_for(n), 256, 0, h->stream>>>(s.last_type, n, -1);
  ptg_ladder_init_kernel<<<grid_for(n), 256, 0, h->stream>>>(s, m.n_rungs, n);
  double *d_x = nullptr;
  if (x_host) {
    size_t cnt = (size_t)n * m.n_init * m.dim;
    rc = dev_alloc(h, &d_x, cnt, false); if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(d_x, x_host, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  if (m.like_kind == PTG_LIKE_HOST_CALLBACK) {
    // MH_chain::initialize (chain.cc:846-876) with the caller's likelihood: one prior draw per chain and round, redrawn while invalid
    // or log L < -1e100, the same Philox addresses as the fused init kernel
    if (x_host) return fail(PTG_EINVAL, "ptg_init_states is not available with a host-callback likelihood");
    const int lpb0 = 1; (void)lpb0;
    for (int k = 0; k < m.n_init; k++) {
      CUDA_TRY(cudaMemsetAsync(s.pend_flags, 0, (size_t)n * sizeof(int32_t), h->stream));
      CUDA_TRY(cudaMemsetAsync(h->d_attempt, 0, (size_t)n * sizeof(int32_t), h->stream));
      for (int round = 0;; round++) {
        cudaError_t ec = cudaErrorInvalidValue;
        switch (m.dim) {
#define X(D) case D: ec = ptg_launch_cb_d##D(2, m, s, 0, 1, 0, k, h->d_attempt, h->d_nopen, h->stream); break;
          PTG_DIM_LIST(X)
#undef X
        }
        CUDA_TRY(ec);
        rc = callback_round(h); if (rc) return rc;
        CUDA_TRY(cudaMemsetAsync(h->d_nopen, 0, sizeof(int32_t), h->stream));
        switch (m.dim) {
#define X(D) case D: ec = ptg_launch_cb_d##D(3, m, s, 0, 1, 0, k, h->d_attempt, h->d_nopen, h->stream); break;
          PTG_DIM_LIST(X)
#undef X
        }
        CUDA_TRY(ec);
        int32_t open = 0;
        CUDA_TRY(cudaMemcpyAsync(&open, h->d_nopen, sizeof(open), cudaMemcpyDeviceToHost, h->stream));
        rc = check_device_error(h); if (rc) return rc;
        if (open == 0) break;
      }
    }
    CUDA_TRY(cudaMemsetAsync(s.pend_flags, 0, (size_t)n * sizeof(int32_t), h->stream));
    h->inited = true; h->istep = 0;
    return 0;
  }
  cudaError_t e = cudaErrorInvalidValue;
  if (h->wide) e = ptg_launch_xinit(h->cfg.rng_mode, m, s, d_x, h->stream);
  else switch (m.dim) {
#define X(D) case D: e = ptg_launch_init_d##D(h->cfg.rng_mode, m, s, d_x, h->stream); break;
    PTG_DIM_LIST(X)
#undef X
  }
  CUDA_TRY(e);
  rc = check_device_error(h); if (rc) return rc;
  h->inited = true; h->istep = 0;
  return 0;
}
extern "C" int ptg_init_from_prior(ptg_handle *h) { return do_init(h, nullptr); }
extern "C" int ptg_init_states(ptg_handle *h, const double *x) {
  if (!x) return fail(PTG_EINVAL, "null states");
  return do_init(h, x);
}

// ------------------------------------------------------------------------------------------------- stepping
// ladders per CTA of the shared-memory kernel: ~128 threads, bounded by the shared memory a ladder needs
static int ladders_per_block(const PtgModel &m) {
  int lpb = 128 / m.n_rungs;
  const size_t per_ladder = ptg_ladder_shared_bytes(m.dim, m.n_rungs), budget = 96 * 1024;
  if ((size_t)lpb * per_ladder > budget) lpb = (int)(budget / per_ladder);
  if (lpb < 1) lpb = 1;
  if (lpb > m.n_ladders) lpb = m.n_ladders;
  return lpb;
}

// lanes per ladder of the warp kernels: the smallest power of two that holds n_rungs; 0 = the ladder does not fit one
// warp (or has more swap trials per step than lanes) and the shared-memory kernel runs instead
static int warp_kernel_width(const ptg_handle *h) {
  const PtgModel &m = h->m;
  if (m.n_rungs > 32) return 0;
  int W = 1;
  while (W < m.n_rungs) W <<= 1;
  if (m.swap_mode == PTG_SWAP_REFERENCE && m.maxswaps > W) return 0;
  return W;
}
// which step kernel runs: PTG_KERNEL_FAST (Philox draws, <= 32 rungs), PTG_KERNEL_WARP (tape replay, <= 32 rungs),
// PTG_KERNEL_SHARED otherwise; ptg_select_kernel can pin WARP or SHARED where they apply
static const PtgXchg xchg_off = {};
// Data chi-squared likelihoods (BASELINE configs B, C2) run one WARP per chain with the data loop spread over its lanes
// (ptg_wide_mma.cuh) whenever the run is a Philox run with automatic kernel selection.  The rule depends on the WORKLOAD only, never on
// the batch size: the warp-per-chain reduction sums in a different order than the thread-per-chain loop, so a size-dependent choice
// would make a ladder's samples depend on how many ladders share its GPU (shard vs full batch).
static bool routes_warp_per_chain(const ptg_handle *h) {
  const PtgModel &m = h->m;
  const bool data_like = m.like_kind == PTG_LIKE_POLY_CHI2 || m.like_kind == PTG_LIKE_SINUSOID_CHI2;
  bool has_prior_draw = false;
  for (const HostProp &hp : h->props) if (hp.p.kind == PTG_PROP_PRIOR_DRAW) has_prior_draw = true;
  // Since round 2 the thread-per-chain production kernel carries fused data functors (ptg_fast.cuh: ~220 fp64 instructions per chain-step
  // for config B against the ~1000 instructions of per-chain machinery a whole warp spends here), so the warp-per-chain route is taken only
  // on request (PTG_DATA_WARP_PER_CHAIN=1, experiments).
  static const bool want = [] { const char *e = getenv("PTG_DATA_WARP_PER_CHAIN"); return e && atoi(e) != 0; }();
  return want && !h->wide && data_like && !has_prior_draw && h->cfg.rng_mode == PTG_RNG_PHILOX && h->kernel_choice == PTG_KERNEL_AUTO && m.n_rungs <= 32 &&
         m.dim <= (m.like_kind == PTG_LIKE_POLY_CHI2 ? 16 : 18);
}
// The production kernel has streamlined instantiations (ptg_fast.cuh, LK >= 0) for the common production configuration; returns the
// likelihood kind to instantiate for, or -1 when the configuration needs the general instantiation.  Chains are bit-identical either way.
static int fstep_streamlined_kind(const ptg_handle *h) {
  const PtgModel &m = h->m;
  if (h->kernel_choice == PTG_KERNEL_FAST_GENERAL) return -1;
  // data chi-squared likelihoods: the production functors (fused multiply-adds, rotation recurrences) under automatic selection only;
  // pinning PTG_KERNEL_FAST keeps the reference's unfused arithmetic (bit-identical with the tape-capable kernels)
  if (m.like_kind == PTG_LIKE_POLY_CHI2 || m.like_kind == PTG_LIKE_SINUSOID_CHI2) return h->kernel_choice == PTG_KERNEL_AUTO ? m.like_kind : -1;
  if (m.like_kind != PTG_LIKE_SINES && m.like_kind != PTG_LIKE_GAUSS_ISO) return -1;
  if (m.swap_mode != PTG_SWAP_REFERENCE || m.evolve_rate > 0 || !m.wrap_in_set || m.record_full || m.trace_steps > 0) return -1;
  if (m.any_bound || !m.all_uniform_prior || !m.zero_valid) return -1;
  for (const HostProp &hp : h->props) {
    if (hp.p.kind != PTG_PROP_DE && hp.p.kind != PTG_PROP_GAUSS) return -1;
    if (hp.p.kind == PTG_PROP_DE && hp.p.unlikely_alpha > 0) return -1;
    if (hp.p.kind == PTG_PROP_GAUSS && !hp.transform.empty()) return -1;
  }
  return m.like_kind;
}
// ptg_wide.cu: the pipelined full-covariance kernel (ptg_wide_pipe.cuh)
int ptg_xpstep_fits(const PtgModel &m, int trans_off);
size_t ptg_xpstep_scratch_doubles(const PtgModel &m);
cudaError_t ptg_launch_xpstep(const PtgModel &m, const PtgState &s, long long step0, int n_steps, int trans_off, double *scratch, cudaStream_t st);
static int pick_kernel(const ptg_handle *h, int *W) {
  *W = warp_kernel_width(h);
  int k = PTG_KERNEL_SHARED;
  if (*W) k = (h->cfg.rng_mode == PTG_RNG_PHILOX) ? PTG_KERNEL_FAST : PTG_KERNEL_WARP;
  if (h->kernel_choice == PTG_KERNEL_SHARED) k = PTG_KERNEL_SHARED;
  if (h->kernel_choice == PTG_KERNEL_WARP && *W) k = PTG_KERNEL_WARP;
  // adaptive shares and temperature mixing exist in the tape-capable warp kernel (either RNG mode); ptg_set_proposal_options made sure it applies
  if ((h->m.adapt_rate != 0 || h->m.de_mixing || h->m.nest_count) && *W) k = PTG_KERNEL_WARP;
  return k;
}

extern "C" int ptg_step(ptg_handle *h, int64_t n_steps) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  if (!h->inited) return fail(PTG_EINVAL, "MH_chain:step: Can't step before initializing chain (chain.cc:967-971)");
  if (n_steps < 0) return fail(PTG_EINVAL, "negative step count");
  if (h->dead) return fail(PTG_EXCHANGE, "handle stopped after an aborted boundary exchange");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  PtgModel &m = h->m;
  if (m.like_kind == PTG_LIKE_HOST_CALLBACK) {
    // per PT iteration: swap phase + proposals on the device, the caller's likelihood on the host, Metropolis test + append on the device
    const int lpb = ladders_per_block(m);
    const size_t smem = (size_t)lpb * ptg_ladder_shared_bytes(m.dim, m.n_rungs) + (size_t)m.n_rungs * m.n_props * sizeof(double);
    for (int64_t it = 0; it < n_steps; it++) {
      cudaError_t ec = cudaErrorInvalidValue;
      switch (m.dim) {
#define X(D) case D: ec = ptg_launch_cb_d##D(0, m, h->s, h->istep, lpb, smem, 0, nullptr, nullptr, h->stream); break;
        PTG_DIM_LIST(X)
#undef X
      }
      CUDA_TRY(ec);
      int rc = callback_round(h); if (rc) return rc;
      switch (m.dim) {
#define X(D) case D: ec = ptg_launch_cb_d##D(1, m, h->s, h->istep, lpb, smem, 0, nullptr, nullptr, h->stream); break;
        PTG_DIM_LIST(X)
#undef X
      }
      CUDA_TRY(ec);
      h->istep++; h->launches += 2;
    }
    return 0;
  }
  int W = 0;
  const int kern = pick_kernel(h, &W);
  const int lk = fstep_streamlined_kind(h);
  const int lpb = ladders_per_block(m);
  const size_t smem = (size_t)lpb * ptg_ladder_shared_bytes(m.dim, m.n_rungs) + (size_t)m.n_rungs * m.n_props * sizeof(double);
  const int max_chunk = (kern == PTG_KERNEL_FAST) ? 16384 : (1 << 20); // the fast kernel keeps 16-bit launch-local statistics
  int64_t left = n_steps;
  while (left > 0) {
    int chunk = (int)(left > max_chunk ? max_chunk : left);
    cudaError_t e = cudaErrorInvalidValue;
    const bool warp_per_chain = routes_warp_per_chain(h);
    if (warp_per_chain) e = ptg_launch_xmstep(m, h->s, h->istep, chunk, -1, h->stream);
    else if (h->wide) {
      // Philox runs take the DMMA-batched kernel; tape replay (and PTG_KERNEL_WARP) the exact-summation-order kernel
      // (data chi^2 likelihoods and the prior-draw member exist in the exact-order kernel only)
      bool exact_only = m.like_kind == PTG_LIKE_POLY_CHI2 || m.like_kind == PTG_LIKE_SINUSOID_CHI2;
      for (const HostProp &hp : h->props) if (hp.p.kind == PTG_PROP_PRIOR_DRAW) exact_only = true;
      if (h->cfg.rng_mode == PTG_RNG_PHILOX && h->kernel_choice != PTG_KERNEL_WARP && !exact_only) {
        // full-covariance Gaussian with one rotation matrix (BASELINE config D): the pipelined kernel with both matrices resident in shared
        // memory; PTG_KERNEL_SHARED pins the round-1 L1-streamed DMMA kernel (bit-identical chains, tests)
        if (h->kernel_choice != PTG_KERNEL_SHARED && ptg_xpstep_fits(m, h->wide_trans_off)) {
          if (!h->d_xscratch) { int rc = dev_alloc(h, &h->d_xscratch, ptg_xpstep_scratch_doubles(m)); if (rc) return rc; }
          e = ptg_launch_xpstep(m, h->s, h->istep, chunk, h->wide_trans_off, h->d_xscratch, h->stream);
        } else e = ptg_launch_xmstep(m, h->s, h->istep, chunk, h->wide_trans_off, h->stream);
      }
      else e = ptg_launch_xstep(h->cfg.rng_mode, m, h->s, h->istep, chunk, h->stream);
    }
    else switch (m.dim) {
#define X(D) case D:                                                                                                       \
      if (kern == PTG_KERNEL_FAST) e = ptg_launch_fstep_d##D(m, h->s, h->istep, chunk, W, h->xchg_launch ? h->xchg_cur : xchg_off, lk, h->stream); \
      else if (kern == PTG_KERNEL_WARP) e = ptg_launch_wstep_d##D(h->cfg.rng_mode, m, h->s, h->istep, chunk, W, h->stream); \
      else e = ptg_launch_step_d##D(h->cfg.rng_mode, m, h->s, h->istep, chunk, lpb, smem, h->stream);                       \
      break;
      PTG_DIM_LIST(X)
#undef X
    }
    CUDA_TRY(e);
    h->istep += chunk; left -= chunk; h->launches++;
  }
  return 0;
}

extern "C" int ptg_select_kernel(ptg_handle *h, int32_t kernel) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  if (kernel < PTG_KERNEL_AUTO || kernel > PTG_KERNEL_FAST_GENERAL) return fail(PTG_EINVAL, "bad kernel id %d", kernel);
  h->kernel_choice = kernel;
  return 0;
}
/* number of step-kernel launches issued since creation */
extern "C" int ptg_get_launch_count(ptg_handle *h, int64_t *n) {
  if (!h || !n) return fail(PTG_EINVAL, "null argument");
  *n = h->launches;
  return 0;
}

extern "C" int ptg_synchronize(ptg_handle *h) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return check_device_error(h);
}

static int ensure_scratch(ptg_handle *h, size_t bytes) {
  if (bytes <= h->scratch_bytes) return 0;
  void *p = nullptr;
  CUDA_TRY(cudaMalloc(&p, bytes));
  h->allocs.push_back(p); h->d_scratch = (double *)p; h->scratch_bytes = bytes;
  return 0;
}

// End-to-end block for a host facade (the run loop's dump of every cold sample, ptmcmc.cc:601-616): n_steps iterations, then the cold
// chains' newest n_out stored samples of every ladder travel to the caller's host buffers.  _begin only ENQUEUES: the gather into one of two
// device staging buffers runs on the engine's stream, the three device-to-host copies on a separate copy stream behind an event, so they
// overlap whatever the caller enqueues next (the next block's kernel); _wait blocks until every begun block's samples have landed.
// One cudaMemcpyAsync per output array; pass pinned host memory for the overlap to be real.
static int stage_setup(ptg_handle *h, size_t bytes) {
  if (!h->copy_stream) {
    CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_gathered[b], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_copied[b], cudaEventDisableTiming));
    }
  }
  for (int b = 0; b < 2; b++) if (h->stage_bytes[b] < bytes) {
    if (h->d_stage[b]) { CUDA_TRY(cudaStreamSynchronize(h->copy_stream)); CUDA_TRY(cudaStreamSynchronize(h->stream)); CUDA_TRY(cudaFree(h->d_stage[b])); h->d_stage[b] = nullptr; }
    CUDA_TRY(cudaMalloc((void **)&h->d_stage[b], bytes));
    h->stage_bytes[b] = bytes;
  }
  return 0;
}
extern "C" int ptg_step_host_begin(ptg_handle *h, int64_t n_steps, int32_t n_out, double *x_out, double *lpost_out, double *llike_out) {
  int rc = ptg_step(h, n_steps); if (rc) return rc;
  if (n_out <= 0) return 0;
  if (!x_out || !lpost_out || !llike_out) return fail(PTG_EINVAL, "null output");
  if (n_out > h->m.hist_cap) return fail(PTG_EINVAL, "n_out = %d exceeds the history ring (%d slots)", n_out, h->m.hist_cap);
  PtgModel &m = h->m;
  const long long total = (long long)m.n_ladders * n_out;
  rc = stage_setup(h, (size_t)total * (m.dim + 2) * sizeof(double)); if (rc) return rc;
  const int b = h->stage_next; h->stage_next ^= 1;
  double *dx = h->d_stage[b], *dlp = dx + total * m.dim, *dll = dlp + total;
  if (h->stage_used[b]) CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_copied[b], 0)); // the copy that last read this buffer
  ptg_gather_cold_kernel<<<grid_for(total), 256, 0, h->stream>>>(m, h->s, n_out, dx, dlp, dll);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(h->ev_gathered[b], h->stream));
  CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_gathered[b], 0));
  CUDA_TRY(cudaMemcpyAsync(x_out, dx, (size_t)total * m.dim * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
  CUDA_TRY(cudaMemcpyAsync(lpost_out, dlp, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
  CUDA_TRY(cudaMemcpyAsync(llike_out, dll, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
  CUDA_TRY(cudaEventRecord(h->ev_copied[b], h->copy_stream));
  h->stage_used[b] = true;
  return 0;
}
extern "C" int ptg_step_host_wait(ptg_handle *h) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  if (h->copy_stream) CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
  return 0;
}
extern "C" int ptg_step_host(ptg_handle *h, int64_t n_steps, int32_t n_out, double *x_out, double *lpost_out, double *llike_out) {
  int rc = ptg_step_host_begin(h, n_steps, n_out, x_out, lpost_out, llike_out); if (rc) return rc;
  rc = ptg_step_host_wait(h); if (rc) return rc;
  return ptg_synchronize(h);
}
/* effective ring capacity (ptg_create maps hist_capacity = 0 to n_init + 1024) */
extern "C" int ptg_get_hist_capacity(ptg_handle *h, int32_t *capacity) {
  if (!h || !capacity) return fail(PTG_EINVAL, "null argument");
  *capacity = h->m.hist_cap;
  return 0;
}

// batched device evaluation of the likelihood functor / the prior at caller-provided states x[n][dim]
extern "C" int ptg_eval(ptg_handle *h, const double *x, int64_t n, double *loglike, double *logprior) {
  if (!h || !x || n < 0) return fail(PTG_EINVAL, "bad argument");
  if (n == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  if (!h->model_uploaded || h->model_dirty) { int rc = upload_model(h); if (rc) return rc; }
  PtgModel &m = h->m;
  if (m.like_kind == PTG_LIKE_HOST_CALLBACK && loglike) { // the caller's own function: evaluated where it lives
    h->cb_fn(h->cb_user, x, n, loglike);
    for (int64_t i = 0; i < n; i++) if (!std::isfinite(loglike[i])) loglike[i] = -INFINITY;
    if (!logprior) return 0;
    loglike = nullptr;
  }
  const size_t d = (size_t)m.dim;
  int rc = ensure_scratch(h, (size_t)n * (d + 2) * sizeof(double)); if (rc) return rc;
  double *dx = h->d_scratch, *dll = dx + (size_t)n * d, *dlp = dll + n;
  CUDA_TRY(cudaMemcpyAsync(dx, x, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  cudaError_t e = cudaErrorInvalidValue;
  if (h->wide) e = ptg_launch_xeval(m, dx, n, loglike ? dll : nullptr, logprior ? dlp : nullptr, h->stream);
  else switch (m.dim) {
#define X(D) case D: e = ptg_launch_eval_d##D(m, dx, n, loglike ? dll : nullptr, logprior ? dlp : nullptr, h->stream); break;
    PTG_DIM_LIST(X)
#undef X
  }
  CUDA_TRY(e);
  if (loglike) CUDA_TRY(cudaMemcpyAsync(loglike, dll, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (logprior) CUDA_TRY(cudaMemcpyAsync(logprior, dlp, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------- read-back
template <typename T>
static int d2h(ptg_handle *h, T *dst, const T *src, size_t n) {
  if (!dst) return 0;
  CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
  return 0;
}

extern "C" int ptg_get_current(ptg_handle *h, double *x, double *lpost, double *llike, double *beta) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t n = (size_t)h->m.n_chains, d = (size_t)h->m.dim;
  std::vector<double> tmp;
  if (x) { tmp.resize(n * d); int rc = d2h(h, tmp.data(), h->s.cur_x, n * d); if (rc) return rc; }
  int rc = d2h(h, lpost, h->s.lpost, n); if (rc) return rc;
  rc = d2h(h, llike, h->s.llike, n); if (rc) return rc;
  rc = d2h(h, beta, h->s.beta, n); if (rc) return rc;
  rc = ptg_synchronize(h); if (rc) return rc;
  if (x) for (size_t c = 0; c < n; c++) for (size_t k = 0; k < d; k++) x[c * d + k] = tmp[k * n + c]; // [dim][chain] -> [chain][dim]
  return 0;
}

extern "C" int ptg_get_lprior(ptg_handle *h, double *lprior) {
  if (!h || !lprior) return fail(PTG_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  int rc = d2h(h, lprior, h->s.lprior, (size_t)h->m.n_chains); if (rc) return rc;
  return ptg_synchronize(h);
}

extern "C" int ptg_set_current(ptg_handle *h, const double *x, const double *lpost, const double *llike, const double *lprior) {
  if (!h || !x || !lpost || !llike || !lprior) return fail(PTG_EINVAL, "null argument");
  if (!h->inited) return fail(PTG_EINVAL, "not initialised");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t n = (size_t)h->m.n_chains, d = (size_t)h->m.dim;
  // x arrives [chain][dim]; the device layout is [dim][chain]: transpose on the device from a staged copy
  int rc = ensure_scratch(h, n * d * sizeof(double)); if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(h->d_scratch, x, n * d * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  ptg_transpose_kernel<<<grid_for((long long)(n * d)), 256, 0, h->stream>>>(h->d_scratch, h->s.cur_x, (long long)n, (int)d);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(h->s.lpost, lpost, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->s.llike, llike, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->s.lprior, lprior, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return 0;
}

extern "C" int ptg_get_counters(ptg_handle *h, int64_t *nhist, int64_t *nsize, int64_t *ntries, int64_t *naccept, int32_t *last_type, double *map_lpost) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t n = (size_t)h->m.n_chains;
  int rc = 0;
  rc |= d2h(h, (long long *)nhist, h->s.nhist, n); rc |= d2h(h, (long long *)nsize, h->s.nsize, n);
  rc |= d2h(h, (long long *)ntries, h->s.ntries, n); rc |= d2h(h, (long long *)naccept, h->s.naccept, n);
  rc |= d2h(h, last_type, h->s.last_type, n); rc |= d2h(h, map_lpost, h->s.map_lpost, n);
  if (rc) return rc;
  return ptg_synchronize(h);
}

extern "C" int ptg_get_history(ptg_handle *h, int32_t ladder, int32_t rung, int64_t first, int64_t count,
                               double *x, double *lpost, double *llike, double *acc, double *beta, int32_t *type) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  PtgModel &m = h->m;
  if (ladder < 0 || ladder >= m.n_ladders || rung < 0 || rung >= m.n_rungs) return fail(PTG_EINVAL, "bad chain (%d,%d)", ladder, rung);
  if ((acc || beta || type) && !m.record_full) return fail(PTG_EINVAL, "acc/beta/type need record_level = PTG_RECORD_FULL");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const long long c = (long long)ladder * m.n_rungs + rung;
  long long nsize = 0;
  CUDA_TRY(cudaMemcpyAsync(&nsize, h->s.nsize + c, sizeof(nsize), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (first < 0 || count < 0 || first + count > nsize) return fail(PTG_EINVAL, "history range [%lld,%lld) outside [0,%lld)", (long long)first, (long long)(first + count), nsize);
  if (nsize - first > m.hist_cap) return fail(PTG_EINVAL, "history element %lld already overwritten (ring capacity %d)", (long long)first, m.hist_cap);
  if (count == 0) return 0;
  const int D = m.dim;
  const int HX = m.hx;
  std::vector<double> rec((size_t)count * HX), lp((size_t)count * 2);
  // the range may wrap around the ring: at most two contiguous pieces
  long long p0 = first % m.hist_cap, n0 = (p0 + count <= m.hist_cap) ? count : m.hist_cap - p0;
  const double *base = h->s.hist + c * m.hist_cap * HX, *lbase = h->s.hist_lp + c * m.hist_cap * 2;
  if (x) {
    CUDA_TRY(cudaMemcpyAsync(rec.data(), base + p0 * HX, (size_t)n0 * HX * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (n0 < count) CUDA_TRY(cudaMemcpyAsync(rec.data() + n0 * HX, base, (size_t)(count - n0) * HX * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  if (lpost || llike) {
    CUDA_TRY(cudaMemcpyAsync(lp.data(), lbase + p0 * 2, (size_t)n0 * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (n0 < count) CUDA_TRY(cudaMemcpyAsync(lp.data() + n0 * 2, lbase, (size_t)(count - n0) * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  auto piece = [&](auto *dst, const auto *src) -> int {
    if (!dst) return 0;
    CUDA_TRY(cudaMemcpyAsync(dst, src + c * m.hist_cap + p0, (size_t)n0 * sizeof(*dst), cudaMemcpyDeviceToHost, h->stream));
    if (n0 < count) CUDA_TRY(cudaMemcpyAsync(dst + n0, src + c * m.hist_cap, (size_t)(count - n0) * sizeof(*dst), cudaMemcpyDeviceToHost, h->stream));
    return 0;
  };
  int rc = 0;
  if (m.record_full) { rc |= piece(acc, h->s.hist_acc); rc |= piece(beta, h->s.hist_beta); rc |= piece(type, h->s.hist_type); }
  if (rc) return rc;
  rc = ptg_synchronize(h); if (rc) return rc;
  for (long long k = 0; k < count; k++) {
    if (x) for (int i = 0; i < D; i++) x[k * D + i] = rec[k * HX + i];
    if (lpost) lpost[k] = lp[2 * k];
    if (llike) llike[k] = lp[2 * k + 1];
  }
  return 0;
}

extern "C" int ptg_get_swap_stats(ptg_handle *h, int64_t *swap_count, int64_t *swap_accept, int32_t *directions, int32_t *ups, int32_t *downs, int32_t *instances) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t n = (size_t)h->m.n_chains;
  const int R = h->m.n_rungs, L = h->m.n_ladders;
  std::vector<long long> sc(n), sa(n);
  int rc = 0;
  rc |= d2h(h, sc.data(), h->s.swap_count, n); rc |= d2h(h, sa.data(), h->s.swap_accept, n);
  rc |= d2h(h, directions, h->s.directions, n); rc |= d2h(h, ups, h->s.ups, n); rc |= d2h(h, downs, h->s.downs, n); rc |= d2h(h, instances, h->s.instances, n);
  if (rc) return rc;
  rc = ptg_synchronize(h); if (rc) return rc;
  for (int l = 0; l < L; l++)
    for (int r = 0; r < R - 1; r++) {
      if (swap_count) swap_count[(size_t)l * (R - 1) + r] = sc[(size_t)l * R + r];
      if (swap_accept) swap_accept[(size_t)l * (R - 1) + r] = sa[(size_t)l * R + r];
    }
  return 0;
}

extern "C" int ptg_get_trace(ptg_handle *h, int64_t first, int64_t count, double *lhr, int32_t *code) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  if (first < 0 || count < 0 || first + count > h->cfg.trace_steps || first + count > h->istep) return fail(PTG_EINVAL, "trace range");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t n = (size_t)h->m.n_chains;
  int rc = 0;
  rc |= d2h(h, lhr, h->s.trace_lhr + (size_t)first * n, (size_t)count * n);
  rc |= d2h(h, code, h->s.trace_code + (size_t)first * n, (size_t)count * n);
  if (rc) return rc;
  return ptg_synchronize(h);
}

extern "C" int ptg_get_total_steps(ptg_handle *h, int64_t *total) {
  if (!h || !total) return fail(PTG_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  int rc = ensure_scratch(h, sizeof(unsigned long long)); if (rc) return rc;
  CUDA_TRY(cudaMemsetAsync(h->d_scratch, 0, sizeof(unsigned long long), h->stream));
  ptg_sum_nhist_kernel<<<grid_for(h->m.n_chains), 256, 0, h->stream>>>(h->s.nhist, h->m.n_chains, (unsigned long long *)h->d_scratch);
  CUDA_TRY(cudaGetLastError());
  unsigned long long v = 0;
  CUDA_TRY(cudaMemcpyAsync(&v, h->d_scratch, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
  rc = ptg_synchronize(h); if (rc) return rc;
  *total = (int64_t)v;
  return 0;
}

extern "C" int ptg_get_device_views(ptg_handle *h, void **hist_dev, void **cur_x_dev, void **stream, int64_t *hist_stride) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  if (hist_dev) *hist_dev = h->s.hist;
  if (cur_x_dev) *cur_x_dev = h->s.cur_x;
  if (stream) *stream = (void *)h->stream;
  if (hist_stride) *hist_stride = (int64_t)h->m.hist_cap * h->m.hx;
  return 0;
}

extern "C" int ptg_set_stream(ptg_handle *h, void *cuda_stream) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (h->own_stream) cudaStreamDestroy(h->stream);
  h->stream = (cudaStream_t)cuda_stream; h->own_stream = false;
  return 0;
}

// ------------------------------------------------------------------------------------------------- evidence
// mean log-likelihood of every chain over its newest `n_last` stored samples (one warp per chain, lanes stride the ring)
__global__ void ptg_mean_llike_kernel(PtgModel m, PtgState s, int n_last, double *out) {
  const long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= m.n_chains) return;
  const long long nsize = s.nsize[c];
  long long n = n_last;
  if (n > nsize) n = nsize;
  if (n > m.hist_cap) n = m.hist_cap;
  const double *base = s.hist_lp + c * m.hist_cap * 2;
  double sum = 0;
  for (long long k = lane; k < n; k += 32) sum += base[((nsize - n + k) % m.hist_cap) * 2 + 1];
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) out[c] = n > 0 ? sum / (double)n : CUDART_NAN;
}
// thermodynamic-integration evidence of every ladder from the per-rung means, the reference's trapezoid
// (parallel_tempering_chains::log_evidence_ratio and its use, chain.cc:1582-1600,1984-2012)
__global__ void ptg_log_evidence_kernel(PtgModel m, PtgState s, const double *mean_ll, double *out) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= m.n_ladders) return;
  const int R = m.n_rungs;
  const double *ml = mean_ll + l * R, *b = s.beta + l * R;
  double evidence = 0, up = 0, down = 0;
  for (int i = 0; i < R - 1; i++) {
    up = ml[i + 1] * (b[i] - b[i + 1]);        // log_evidence_ratio(i, i+1): mean over rung i+1's samples of llike * (beta_i - beta_{i+1})
    down = -(ml[i] * (b[i + 1] - b[i]));       // -log_evidence_ratio(i+1, i)
    evidence += (up + down) / 2.0;
  }
  if (R > 1) evidence += (up + down) / 2.0 / (b[R - 2] / b[R - 1] - 1); // the tail below the hottest rung (chain.cc:1597)
  out[l] = evidence;
}

// integrated autocorrelation time of every parameter of rung `rung` of every ladder over its newest n_last stored samples:
// one CTA per ladder; the series is staged in shared memory, one thread per lag computes the (biased) autocovariance, thread 0
// applies Sokal's self-consistent window M >= c*tau(M) to tau(M) = 1 + 2 sum_{t<=M} rho_t  (the estimator of
// ptmcmc_b200/analysis.py:integrated_act, per chain).  Not the reference's own recipe (chain.cc:126-643), which is a different
// finite-sample estimator of the same quantity.
__global__ void __launch_bounds__(256) ptg_act_kernel(PtgModel m, PtgState s, int rung, int n_last, int max_lag, double cwin, double *tau_out) {
  extern __shared__ double sh[];
  double *y = sh, *rho = sh + n_last, *red = rho + max_lag;
  const long long l = blockIdx.x, c = l * m.n_rungs + rung;
  const long long nsize = s.nsize[c];
  int n = n_last;
  if (n > nsize) n = (int)nsize;
  if (n > m.hist_cap) n = m.hist_cap;
  const double *base = s.hist + c * m.hist_cap * m.hx;
  for (int j = 0; j < m.dim; j++) {
    double part = 0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) { const double v = base[((nsize - n + k) % m.hist_cap) * m.hx + j]; y[k] = v; part += v; }
    red[threadIdx.x] = part;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    const double mean = red[0] / n;
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) y[k] -= mean;
    __syncthreads();
    for (int lag = threadIdx.x; lag < max_lag; lag += blockDim.x) {
      double acc = 0;
      for (int k = 0; k + lag < n; k++) acc += y[k] * y[k + lag];
      rho[lag] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const double c0 = rho[0];
      double tau = 1.0, cum = 0;
      const int lim = max_lag < n ? max_lag : n;
      for (int M = 0; M < lim; M++) {
        cum += (c0 > 0 ? rho[M] / c0 : 1.0);
        tau = 2.0 * cum - 1.0;
        if (M >= cwin * tau) break;
      }
      tau_out[l * m.dim + j] = tau > 1.0 ? tau : 1.0;
    }
    __syncthreads();
  }
}

extern "C" int ptg_get_act(ptg_handle *h, int32_t rung, int32_t n_last, int32_t max_lag, double *tau) {
  if (!h || !tau || n_last < 8 || max_lag < 2) return fail(PTG_EINVAL, "bad argument");
  if (!h->inited) return fail(PTG_EINVAL, "not initialised");
  if (rung < 0 || rung >= h->m.n_rungs) return fail(PTG_EINVAL, "rung %d out of range", rung);
  if (max_lag > n_last) max_lag = n_last;
  const size_t smem = ((size_t)n_last + max_lag + 256) * sizeof(double);
  if (smem > 200 * 1024) return fail(PTG_EINVAL, "n_last + max_lag too large for shared memory (%zu bytes)", smem);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t cnt = (size_t)h->m.n_ladders * h->m.dim;
  int rc = ensure_scratch(h, cnt * sizeof(double)); if (rc) return rc;
  if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(ptg_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ptg_act_kernel<<<h->m.n_ladders, 256, smem, h->stream>>>(h->m, h->s, rung, n_last, max_lag, 5.0, h->d_scratch);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(tau, h->d_scratch, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  return ptg_synchronize(h);
}

// The reference's own ESS recipe, device part: the windowed lag statistics of chain::compute_autocovar_windows
// (chain.cc:126-289) for rung `rung` of every ladder.  One CTA per (window, ladder); thread j owns lag j and walks the window's
// samples in index order -- the reference's own summation order (`xsum += fi + f[i]; xxsum += fi * f[i]`, chain.cc:268-279), so the
// numbers equal the C++ reference's to the last bit, and
//     means[l][f][k][j] = sum(f_i + f_{i-lag_j}) / n / 2 ,  covar[l][f][k][j] = sum(f_i f_{i-lag_j}) / n - means^2        (lag 0: mean, variance)
// Windows are counted in stored records: window k of ladder l covers records [end_l - (n_win - k) swidth, +swidth), end_l =
// end_rec[l] (default: all records stored so far).  The combination over windows (chain.cc:292-449) is cheap and stays on the host.
__global__ void __launch_bounds__(128) ptg_autocovar_kernel(PtgModel m, PtgState s, int rung, int swidth, int n_win, int n_lag, const int *__restrict__ lag_rec,
                                                            const long long *__restrict__ end_rec, int n_feat, double *means, double *covar, int *bad) {
  const int k = blockIdx.x;
  const long long l = blockIdx.y, c = l * m.n_rungs + rung;
  const long long nsize = s.nsize[c], end = end_rec ? end_rec[l] : nsize;
  const long long start = end - (long long)(n_win - k) * swidth;
  const int D = m.dim, cap = m.hist_cap;
  const int HX = m.hx;
  const double *base = s.hist + c * (long long)cap * HX;
  for (int j = threadIdx.x; j < n_lag; j += blockDim.x) {
    const long long first = start - lag_rec[j];
    if (first < 0 || first < nsize - cap || end > nsize) { if (bad) atomicExch(bad, 1); continue; } // not resident in the ring
    int p0 = (int)(start % cap), p1 = (int)(first % cap);
    for (int f0 = 0; f0 < n_feat; f0 += 4) {
      double sx[4] = {0, 0, 0, 0}, sxx[4] = {0, 0, 0, 0};
      int q0 = p0, q1 = p1;
      for (int i = 0; i < swidth; i++) {
        const double *a = base + (long long)q0 * HX + f0, *b = base + (long long)q1 * HX + f0;
#pragma unroll
        for (int f = 0; f < 4; f++) if (f0 + f < n_feat) { const double fi = a[f], fl = b[f]; sx[f] += (fl + fi); sxx[f] += fl * fi; }
        if (++q0 == cap) q0 = 0;
        if (++q1 == cap) q1 = 0;
      }
#pragma unroll
      for (int f = 0; f < 4; f++) if (f0 + f < n_feat) {
        const size_t o = (((size_t)l * n_feat + f0 + f) * n_win + k) * n_lag + j;
        const double mean = sx[f] / swidth / 2;
        means[o] = mean;
        covar[o] = sxx[f] / swidth - mean * mean;
      }
    }
  }
}
extern "C" int ptg_get_autocovar_windows(ptg_handle *h, int32_t rung, int32_t swidth, int32_t n_win, int32_t n_lag, const int32_t *lag_rec,
                                         const int64_t *end_rec, int32_t n_feat, double *means, double *covar) {
  if (!h || !lag_rec || !means || !covar || swidth < 1 || n_win < 1 || n_lag < 1 || n_feat < 1) return fail(PTG_EINVAL, "bad argument");
  if (!h->inited) return fail(PTG_EINVAL, "not initialised");
  if (rung < 0 || rung >= h->m.n_rungs) return fail(PTG_EINVAL, "rung %d out of range", rung);
  if (n_feat > h->m.dim) return fail(PTG_EINVAL, "n_feat exceeds dim");
  if (n_win > 65535) return fail(PTG_EINVAL, "too many windows");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t L = (size_t)h->m.n_ladders, cnt = L * n_feat * n_win * n_lag;
  // scratch: means | covar | lags (int) | end_rec (int64) | bad flag
  const size_t bytes = 2 * cnt * sizeof(double) + ((size_t)n_lag + 2) * sizeof(int) + (L + 1) * sizeof(long long) + 64;
  int rc = ensure_scratch(h, bytes); if (rc) return rc;
  double *dm = h->d_scratch, *dc = dm + cnt;
  long long *dend = (long long *)(dc + cnt);
  int *dlag = (int *)(dend + L + 1), *dbad = dlag + n_lag;
  CUDA_TRY(cudaMemcpyAsync(dlag, lag_rec, (size_t)n_lag * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if (end_rec) CUDA_TRY(cudaMemcpyAsync(dend, end_rec, L * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemsetAsync(dbad, 0, sizeof(int), h->stream));
  CUDA_TRY(cudaMemsetAsync(dm, 0, 2 * cnt * sizeof(double), h->stream));
  ptg_autocovar_kernel<<<dim3((unsigned)n_win, (unsigned)L), 128, 0, h->stream>>>(h->m, h->s, rung, swidth, n_win, n_lag, dlag, end_rec ? dend : nullptr, n_feat, dm, dc, dbad);
  CUDA_TRY(cudaGetLastError());
  int bad = 0;
  CUDA_TRY(cudaMemcpyAsync(means, dm, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaMemcpyAsync(covar, dc, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaMemcpyAsync(&bad, dbad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  rc = ptg_synchronize(h); if (rc) return rc;
  if (bad) return fail(PTG_EINVAL, "a requested window (or its lagged copy) is no longer in the history ring, or lies beyond the stored records");
  return 0;
}

extern "C" int ptg_get_mean_loglike(ptg_handle *h, int32_t n_last, double *mean_ll) {
  if (!h || !mean_ll || n_last < 1) return fail(PTG_EINVAL, "bad argument");
  if (!h->inited) return fail(PTG_EINVAL, "not initialised");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t n = (size_t)h->m.n_chains;
  int rc = ensure_scratch(h, (n + h->m.n_ladders) * sizeof(double)); if (rc) return rc;
  ptg_mean_llike_kernel<<<grid_for((long long)n * 32), 256, 0, h->stream>>>(h->m, h->s, n_last, h->d_scratch);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(mean_ll, h->d_scratch, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  return ptg_synchronize(h);
}
extern "C" int ptg_get_log_evidence(ptg_handle *h, int32_t n_last, double *log_evidence) {
  if (!h || !log_evidence || n_last < 1) return fail(PTG_EINVAL, "bad argument");
  if (!h->inited) return fail(PTG_EINVAL, "not initialised");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const size_t n = (size_t)h->m.n_chains, L = (size_t)h->m.n_ladders;
  int rc = ensure_scratch(h, (n + L) * sizeof(double)); if (rc) return rc;
  ptg_mean_llike_kernel<<<grid_for((long long)n * 32), 256, 0, h->stream>>>(h->m, h->s, n_last, h->d_scratch);
  ptg_log_evidence_kernel<<<grid_for((long long)L), 256, 0, h->stream>>>(h->m, h->s, h->d_scratch, h->d_scratch + n);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(log_evidence, h->d_scratch + n, L * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  return ptg_synchronize(h);
}

// ------------------------------------------------------------------------------------------------- rung-sharded ladders
__global__ void ptg_boundary_pack_kernel(PtgModel m, PtgState s, int rung, double *out) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= m.n_ladders) return;
  const long long c = l * m.n_rungs + rung;
  const int D = m.dim;
  double *rec = out + l * (D + 3);
  for (int k = 0; k < D; k++) rec[k] = s.cur_x[(long long)k * m.n_chains + c];
  rec[D] = s.llike[c]; rec[D + 1] = s.lprior[c]; rec[D + 2] = s.beta[c];
}
// one thread per ladder: swap trial between my rung and the neighbour's packed rung (chain.cc:1459-1490), then add_state
__global__ void ptg_boundary_swap_kernel(PtgModel m, PtgState s, int rung, const double *__restrict__ nb, int i_am_lower, uint64_t shared_seed,
                                         long long boundary_id, long long exchange_index) {
  const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= m.n_ladders) return;
  const long long c = l * m.n_rungs + rung;
  const int D = m.dim;
  const double *rec = nb + l * (D + 3);
  const double my_ll = s.llike[c], my_beta = s.beta[c], nb_ll = rec[D], nb_lprior = rec[D + 1], nb_beta = rec[D + 2];
  // pair (i, i+1): i = the colder chain
  double lla = i_am_lower ? my_ll : nb_ll; if (!(lla > -1e200)) lla = -1e200;
  double llb = i_am_lower ? nb_ll : my_ll; if (!(llb > -1e200)) llb = -1e200;
  const double ba = i_am_lower ? my_beta : nb_beta, bb = i_am_lower ? nb_beta : my_beta;
  const double lhr = -(bb - ba) * (llb - lla);
  bool accept = true;
  if (lhr < 0) {
    uint32_t w[4];
    ptg_philox_draw(shared_seed, (uint64_t)(m.ladder_offset + l) * PTG_STREAM_STRIDE + PTG_STREAM_LADDER, PTG_DOMAIN_BOUNDARY, (uint64_t)exchange_index,
                    (uint32_t)boundary_id, w);
    accept = (log(ptg_u52_to_unit(w[0], w[1])) < lhr);
  }
  double lpost = s.lpost[c], llike = my_ll;
  if (accept) {
    for (int k = 0; k < D; k++) s.cur_x[(long long)k * m.n_chains + c] = rec[k];
    llike = nb_ll;
    lpost = nb_lprior + my_beta * nb_ll; // lpost recomputed = lprior(x) + beta*llike (chain.cc:925-928)
    s.llike[c] = llike; s.lprior[c] = nb_lprior; s.lpost[c] = lpost;
  }
  // MH_chain::add_state (chain.cc:916-949)
  if (lpost > s.map_lpost[c]) {
    s.map_lpost[c] = lpost;
    for (int k = 0; k < D; k++) s.map_x[(long long)k * m.n_chains + c] = s.cur_x[(long long)k * m.n_chains + c];
  }
  const long long nhist = s.nhist[c];
  if (nhist % m.save_every == 0) {
    const long long nsize = s.nsize[c], slot = nsize % m.hist_cap, r = c * m.hist_cap + slot;
    double *hrec = s.hist + r * m.hx;
    for (int k = 0; k < D; k++) hrec[k] = s.cur_x[(long long)k * m.n_chains + c];
    s.hist_lp[2 * r] = lpost; s.hist_lp[2 * r + 1] = llike;
    if (m.record_full) { s.hist_acc[r] = s.naccept[c] / (double)s.ntries[c]; s.hist_beta[r] = my_beta; s.hist_type[r] = s.last_type[c]; }
    s.nsize[c] = nsize + 1;
  }
  s.nhist[c] = nhist + 1;
  if (i_am_lower) { s.swap_count[c] += 1; if (accept) s.swap_accept[c] += 1; } // counted at the pair's lower rung (this rank's last rung)
}

extern "C" int ptg_boundary_pack(ptg_handle *h, int32_t rung, void *out_dev) {
  if (!h || !out_dev) return fail(PTG_EINVAL, "null argument");
  if (!h->inited) return fail(PTG_EINVAL, "not initialised");
  if (rung < 0 || rung >= h->m.n_rungs) return fail(PTG_EINVAL, "rung %d out of range", rung);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  ptg_boundary_pack_kernel<<<grid_for(h->m.n_ladders), 256, 0, h->stream>>>(h->m, h->s, rung, (double *)out_dev);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int ptg_boundary_swap(ptg_handle *h, int32_t my_rung, const void *neighbour_pack_dev, int32_t i_am_lower, uint64_t shared_seed,
                                 int64_t boundary_id, int64_t exchange_index) {
  if (!h || !neighbour_pack_dev) return fail(PTG_EINVAL, "null argument");
  if (!h->inited) return fail(PTG_EINVAL, "not initialised");
  if (my_rung < 0 || my_rung >= h->m.n_rungs) return fail(PTG_EINVAL, "rung %d out of range", my_rung);
  if (h->cfg.rng_mode != PTG_RNG_PHILOX) return fail(PTG_EINVAL, "boundary swaps draw from the Philox layout");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  ptg_boundary_swap_kernel<<<grid_for(h->m.n_ladders), 256, 0, h->stream>>>(h->m, h->s, my_rung, (const double *)neighbour_pack_dev, i_am_lower ? 1 : 0,
                                                                            shared_seed, (long long)boundary_id, (long long)exchange_index);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---- exchange fused into the production step kernel over peer memory (NVLink): see PtgXchg in ptg_types.h
static size_t xchg_edge_doubles(const ptg_handle *h) { return (size_t)4 * h->m.n_ladders * (h->m.dim + 3); }
static int xchg_alloc(ptg_handle *h) {
  if (h->d_xchg) return 0;
  h->xchg_bytes = xchg_edge_doubles(h) * sizeof(double) + (size_t)2 * h->m.n_ladders * sizeof(int);
  // a dedicated cudaMalloc (IPC handles cover whole allocations); freed in ptg_destroy with every other allocation
  CUDA_TRY(cudaMalloc(&h->d_xchg, h->xchg_bytes));
  h->allocs.push_back(h->d_xchg);
  CUDA_TRY(cudaMemsetAsync(h->d_xchg, 0, h->xchg_bytes, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  // the watchdog word lives in mapped host memory: the host can raise it while a launch is spinning (ptg_xchg_abort)
  CUDA_TRY(cudaHostAlloc((void **)&h->h_abort, sizeof(int), cudaHostAllocMapped));
  *h->h_abort = 0;
  CUDA_TRY(cudaHostGetDevicePointer((void **)&h->d_abort, h->h_abort, 0));
  return 0;
}
// Watchdog of the fused exchange: every boundary wait in flight (and every later one) gives up, the launch ends with the sticky
// error PTG_EXCHANGE.  Callable from any host thread while ptg_synchronize blocks in another (it only writes one mapped word).
extern "C" int ptg_xchg_abort(ptg_handle *h) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  if (h->h_abort) { *(volatile int *)h->h_abort = 1; __sync_synchronize(); }
  return 0;
}
static int xchg_check(ptg_handle *h) {
  if (!h) return fail(PTG_EINVAL, "null handle");
  if (!h->inited) return fail(PTG_EINVAL, "not initialised");
  int W = 0;
  if (h->wide || h->m.like_kind == PTG_LIKE_HOST_CALLBACK || pick_kernel(h, &W) != PTG_KERNEL_FAST || routes_warp_per_chain(h))
    return fail(PTG_EINVAL, "the fused exchange runs in the production kernel only (Philox draws, n_rungs <= 32, dim <= 16, no data-chi^2 likelihood under automatic kernel selection)");
  return 0;
}
extern "C" int ptg_xchg_export(ptg_handle *h, void *ipc_handle_64_bytes, void **local_ptr) {
  int rc = xchg_check(h); if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  rc = xchg_alloc(h); if (rc) return rc;
  if (ipc_handle_64_bytes) {
    cudaIpcMemHandle_t hd;
    CUDA_TRY(cudaIpcGetMemHandle(&hd, h->d_xchg));
    static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(ipc_handle_64_bytes, &hd, sizeof(hd));
  }
  if (local_ptr) *local_ptr = h->d_xchg;
  return 0;
}
static int xchg_open(ptg_handle *h, const void *handle_or_ptr, int is_ipc, void **out, bool *out_ipc) {
  *out = nullptr; *out_ipc = false;
  if (!handle_or_ptr) return 0;
  if (!is_ipc) { *out = const_cast<void *>(handle_or_ptr); return 0; }
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle_or_ptr, sizeof(hd));
  CUDA_TRY(cudaIpcOpenMemHandle(out, hd, cudaIpcMemLazyEnablePeerAccess));
  *out_ipc = true;
  return 0;
}
extern "C" int ptg_xchg_connect(ptg_handle *h, const void *colder, const void *hotter, int32_t handles_are_ipc, uint64_t shared_seed,
                                int64_t colder_boundary_id, int64_t hotter_boundary_id) {
  int rc = xchg_check(h); if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  rc = xchg_alloc(h); if (rc) return rc;
  rc = xchg_open(h, colder, handles_are_ipc, &h->peer_lo, &h->peer_lo_ipc); if (rc) return rc;
  rc = xchg_open(h, hotter, handles_are_ipc, &h->peer_hi, &h->peer_hi_ipc); if (rc) return rc;
  const size_t ed = xchg_edge_doubles(h);
  PtgXchg &x = h->xchg;
  x = PtgXchg{};
  x.on = 1; x.has_lo = h->peer_lo != nullptr; x.has_hi = h->peer_hi != nullptr; x.index = 0;
  x.my_edges = (double *)h->d_xchg; x.my_flags = (int *)((double *)h->d_xchg + ed);
  x.lo_edges = (const double *)h->peer_lo; x.lo_flags = h->peer_lo ? (const int *)((const double *)h->peer_lo + ed) : nullptr;
  x.hi_edges = (const double *)h->peer_hi; x.hi_flags = h->peer_hi ? (const int *)((const double *)h->peer_hi + ed) : nullptr;
  x.shared_seed = shared_seed; x.lo_boundary = colder_boundary_id; x.hi_boundary = hotter_boundary_id; x.abort = h->d_abort;
  return 0;
}
// n_steps PT iterations in ONE launch; apply_pending: first run the boundary swap trials against what the neighbours published at
// the end of their previous exchange launch; publish: publish this block's edge rungs at the end (exchange index = launches so far)
extern "C" int ptg_step_exchange(ptg_handle *h, int64_t n_steps, int32_t every, int32_t apply_pending, int32_t publish) {
  int rc = xchg_check(h); if (rc) return rc;
  if (!h->xchg.on) return fail(PTG_EINVAL, "ptg_xchg_connect first");
  if (n_steps < 0 || n_steps > 16384) return fail(PTG_EINVAL, "n_steps must be within one launch (0..16384)");
  if (apply_pending && h->xchg.index == 0) return fail(PTG_EINVAL, "nothing has been published yet");
  if (every < 0 || (every > 0 && n_steps % every != 0)) return fail(PTG_EINVAL, "n_steps must be a multiple of `every`");
  const bool in_launch = every > 0 && n_steps > every;
  if (in_launch) {
    // warps wait for the neighbour GPU's warp of the same ladder: every CTA of the grid must be resident
    int W = 0; pick_kernel(h, &W);
    const long long warps = (h->m.n_ladders + (32 / W) - 1) / (32 / W);
    if (!ptg_fstep_grid_is_resident(warps)) return fail(PTG_EINVAL, "in-launch exchange needs the whole grid resident (more ladders than one wave holds): use every = 0 and one launch per exchange");
  }
  if (h->dead) return fail(PTG_EXCHANGE, "handle stopped after an aborted boundary exchange");
  PtgXchg &x = h->xchg_cur; // kept in the handle: nothing points into this frame after an early return
  x = h->xchg;
  x.swap_in = apply_pending ? 1 : 0; x.publish_out = publish ? 1 : 0; x.every = in_launch ? every : 0;
  h->xchg_launch = true;
  if (n_steps == 0) { // prologue / epilogue only
    int W = 0; pick_kernel(h, &W);
    cudaError_t e = cudaErrorInvalidValue;
    if (cudaSetDevice(h->cfg.device) == cudaSuccess) switch (h->m.dim) {
#define X(D) case D: e = ptg_launch_fstep_d##D(h->m, h->s, h->istep, 0, W, x, fstep_streamlined_kind(h), h->stream); break;
      PTG_DIM_LIST(X)
#undef X
    }
    h->xchg_launch = false;
    CUDA_TRY(e);
    h->launches++;
  } else {
    rc = ptg_step(h, n_steps);
    h->xchg_launch = false;
    if (rc) return rc;
  }
  h->xchg.index += (in_launch ? n_steps / every - 1 : 0) + (publish ? 1 : 0);
  return 0;
}

// ------------------------------------------------------------------------------------------------- checkpoint
// One binary file: magic, config, istep, ring geometry, exchange index, then every device array of PtgState in declaration order.
// This is the ENGINE's restart blob (all ladders, all rungs, the whole ring); the reference-layout per-chain files
// (MHchain.cp / PTchain.cp, chain.cc:656-731,1213-1239) are written by the host facade from ptg_get_history.
// Arrays are streamed through one bounded staging buffer (the history ring of a production batch is tens of GB).
struct CkArr { void *p; size_t bytes; };
static std::vector<CkArr> ck_arrays(ptg_handle *h) {
  PtgModel &m = h->m; PtgState &s = h->s;
  const size_t n = (size_t)m.n_chains, d = (size_t)m.dim, cap = (size_t)m.hist_cap;
  std::vector<CkArr> v = {
      {s.cur_x, n * d * 8}, {s.lpost, n * 8}, {s.llike, n * 8}, {s.lprior, n * 8}, {s.beta, n * 8}, {s.map_lpost, n * 8}, {s.map_x, n * d * 8},
      {s.nhist, n * 8}, {s.nsize, n * 8}, {s.ntries, n * 8}, {s.naccept, n * 8}, {s.last_type, n * 4}, {s.hist, n * cap * (size_t)m.hx * 8}, {s.hist_lp, n * cap * 2 * 8},
      {s.swap_count, n * 8}, {s.swap_accept, n * 8}, {s.directions, n * 4}, {s.ups, n * 4}, {s.downs, n * 4}, {s.instances, n * 4},
      {s.u_pos, (n + m.n_ladders) * 8}, {s.z_pos, (n + m.n_ladders) * 8}};
  if (m.record_full) { v.push_back({s.hist_acc, n * cap * 8}); v.push_back({s.hist_beta, n * cap * 8}); v.push_back({s.hist_type, n * cap * 4}); }
  if (s.ad_shares) { // the adapted shares of every chain's proposal set (proposal_distribution_set::checkpoint, proposal_distribution.cc:168-200)
    const size_t np = (size_t)m.n_bins;
    v.push_back({s.ad_shares, n * np * 8}); v.push_back({s.ad_bins, n * np * 8}); v.push_back({s.ad_last, n * 4}); v.push_back({s.ad_count, n * 4});
    v.push_back({s.ad_count2, n * 4});
  }
  return v;
}
static const uint64_t CK_MAGIC = 0x70746763686b3032ull; // "ptgchk02"
static const size_t CK_CHUNK = (size_t)64 << 20;
struct CkFile { // closes on every exit path
  FILE *f;
  explicit CkFile(FILE *f_) : f(f_) {}
  ~CkFile() { if (f) fclose(f); }
  int close() { FILE *g = f; f = nullptr; return g ? fclose(g) : 0; }
};
static int checkpoint_impl(ptg_handle *h, const char *path) {
  int rc = ptg_synchronize(h); if (rc) return rc; // also orders this after every step in flight on h->stream
  CkFile ck(fopen(path, "wb"));
  if (!ck.f) return fail(PTG_EINVAL, "cannot open %s", path);
  const int32_t cap = h->m.hist_cap, hx = h->m.hx;
  const int64_t istep = h->istep, xindex = h->xchg.index;
  bool ok = fwrite(&CK_MAGIC, 8, 1, ck.f) == 1 && fwrite(&h->cfg, sizeof(h->cfg), 1, ck.f) == 1 && fwrite(&istep, 8, 1, ck.f) == 1 &&
            fwrite(&cap, 4, 1, ck.f) == 1 && fwrite(&hx, 4, 1, ck.f) == 1 && fwrite(&xindex, 8, 1, ck.f) == 1;
  std::vector<char> buf(CK_CHUNK);
  for (const CkArr &a : ck_arrays(h)) {
    const uint64_t nb = a.bytes;
    ok = ok && fwrite(&nb, 8, 1, ck.f) == 1;
    for (size_t off = 0; ok && off < a.bytes; off += CK_CHUNK) {
      const size_t n = a.bytes - off < CK_CHUNK ? a.bytes - off : CK_CHUNK;
      CUDA_TRY(cudaMemcpyAsync(buf.data(), (const char *)a.p + off, n, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
      ok = fwrite(buf.data(), 1, n, ck.f) == n;
    }
  }
  if (ck.close() != 0) ok = false;
  if (!ok) { remove(path); return fail(PTG_EINVAL, "short write to %s (disk full?): checkpoint removed", path); }
  return 0;
}
extern "C" int ptg_checkpoint(ptg_handle *h, const char *path) {
  if (!h || !path) return fail(PTG_EINVAL, "null argument");
  if (!h->inited) return fail(PTG_EINVAL, "not initialised");
  try { return checkpoint_impl(h, path); }
  catch (const std::exception &e) { return fail(PTG_ENOMEM, "checkpoint: %s", e.what()); }
}
static int restore_impl(ptg_handle *h, const char *path) {
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(h->stream)); // nothing of this handle may be in flight while its arrays are overwritten
  if (!h->model_uploaded) { int rc = upload_model(h); if (rc) return rc; }
  CkFile ck(fopen(path, "rb"));
  if (!ck.f) return fail(PTG_EINVAL, "cannot open %s", path);
  uint64_t magic = 0; ptg_config c; int64_t istep = 0, xindex = 0; int32_t cap = 0, hx = 0;
  bool ok = fread(&magic, 8, 1, ck.f) == 1 && fread(&c, sizeof(c), 1, ck.f) == 1 && fread(&istep, 8, 1, ck.f) == 1 && fread(&cap, 4, 1, ck.f) == 1 &&
            fread(&hx, 4, 1, ck.f) == 1 && fread(&xindex, 8, 1, ck.f) == 1;
  if (!ok || magic != CK_MAGIC) return fail(PTG_EINVAL, "%s is not a ptg checkpoint (format ptgchk02)", path);
  const ptg_config &g = h->cfg;
  if (c.n_ladders != g.n_ladders || c.n_rungs != g.n_rungs || c.dim != g.dim || cap != h->m.hist_cap || hx != h->m.hx || c.record_level != g.record_level)
    return fail(PTG_EINVAL, "checkpoint shape does not match this engine");
  // the continuation must be the run that was interrupted: same draws, same save cadence, same swap schedule, same ladder ids
#define CK_SAME(f) if (!(c.f == g.f)) return fail(PTG_EINVAL, "checkpoint was written with a different " #f "; create the engine with the saved value")
  CK_SAME(seed); CK_SAME(save_every); CK_SAME(n_init); CK_SAME(swap_rate); CK_SAME(swap_mode); CK_SAME(rng_mode); CK_SAME(evolve_rate);
  CK_SAME(evolve_lpost_cut); CK_SAME(dprior_min); CK_SAME(ladder_offset);
#undef CK_SAME
  std::vector<char> buf(CK_CHUNK);
  for (const CkArr &a : ck_arrays(h)) {
    uint64_t nb = 0;
    if (fread(&nb, 8, 1, ck.f) != 1 || nb != a.bytes) return fail(PTG_EINVAL, "checkpoint array size mismatch");
    for (size_t off = 0; off < a.bytes; off += CK_CHUNK) {
      const size_t n = a.bytes - off < CK_CHUNK ? a.bytes - off : CK_CHUNK;
      if (fread(buf.data(), 1, n, ck.f) != n) return fail(PTG_EINVAL, "checkpoint truncated");
      CUDA_TRY(cudaMemcpyAsync((char *)a.p + off, buf.data(), n, cudaMemcpyHostToDevice, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
  }
  h->istep = istep; h->xchg.index = xindex; h->inited = true;
  return 0;
}
extern "C" int ptg_restore(ptg_handle *h, const char *path) {
  if (!h || !path) return fail(PTG_EINVAL, "null argument");
  try { return restore_impl(h, path); }
  catch (const std::exception &e) { return fail(PTG_ENOMEM, "restore: %s", e.what()); }
}
