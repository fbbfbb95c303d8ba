// ptg_kernels.cuh -- thread-per-chain kernels (dim <= 16): initialisation and the fused PT step.
//
// Layout: one thread = one chain (one rung of one ladder); a CTA holds `lpb` whole ladders
// (blockDim.x = lpb * n_rungs), so the replica-swap phase never leaves the CTA.  A launch runs n_steps
// iterations of parallel_tempering_chains::step (chain.cc:1393-1570) for every ladder with the chain state in
// registers; only history appends and differential-evolution gathers touch HBM.
//
// Per PT step:
//   1. every lane publishes (x, llike, lpost, lprior, beta) to shared memory
//   2. swap phase -- the ladder's lane 0 replays the reference's serial schedule (candidate list chain.cc:1410-1420,
//      trials :1436-1538, pry_temps :1809-1846) on the shared copies and records, per rung, up to two history
//      appends (a rung can sit in two trials); Philox draws of the trials are produced by the
//      ladder's lanes in parallel beforehand
//   3. lanes that took part in a trial reload their state and perform their own appends (chain.cc:1487-1534);
//      all other lanes run MH_chain::step (chain.cc:966-1022)
#pragma once
#include "ptg_device.cuh"


template <int D>
struct Chain {
  double x[D];
  double lpost, llike, lprior, beta, map_lpost;
  long long nhist, nsize, ntries, naccept;
  int last_type;
  int slot;        // nsize % hist_cap
  int since_save;  // nhist % save_every
  long long chain; // global index within this engine
};

template <int D>
__device__ __forceinline__ void chain_load(const PtgModel &m, const PtgState &s, long long c, Chain<D> &ch) {
#pragma unroll
  for (int k = 0; k < D; k++) ch.x[k] = s.cur_x[(long long)k * m.n_chains + c];
  ch.lpost = s.lpost[c]; ch.llike = s.llike[c]; ch.lprior = s.lprior[c]; ch.beta = s.beta[c];
  ch.map_lpost = s.map_lpost[c];
  ch.nhist = s.nhist[c]; ch.nsize = s.nsize[c]; ch.ntries = s.ntries[c]; ch.naccept = s.naccept[c];
  ch.last_type = s.last_type[c];
  ch.slot = (int)(ch.nsize % m.hist_cap);
  ch.since_save = (int)(ch.nhist % m.save_every);
  ch.chain = c;
}
template <int D>
__device__ __forceinline__ void chain_store(const PtgModel &m, const PtgState &s, const Chain<D> &ch) {
  long long c = ch.chain;
#pragma unroll
  for (int k = 0; k < D; k++) s.cur_x[(long long)k * m.n_chains + c] = ch.x[k];
  s.lpost[c] = ch.lpost; s.llike[c] = ch.llike; s.lprior[c] = ch.lprior; s.beta[c] = ch.beta;
  s.map_lpost[c] = ch.map_lpost;
  s.nhist[c] = ch.nhist; s.nsize[c] = ch.nsize; s.ntries[c] = ch.ntries; s.naccept[c] = ch.naccept;
  s.last_type[c] = ch.last_type;
}

// MH_chain::add_state (chain.cc:916-949) for the chain's CURRENT state, with the posterior/beta to record
template <int D>
__device__ __forceinline__ void chain_append(const PtgModel &m, const PtgState &s, Chain<D> &ch, const double x[D],
                                             double llike, double lpost, double beta) {
  if (lpost > ch.map_lpost) {
    ch.map_lpost = lpost;
#pragma unroll
    for (int k = 0; k < D; k++) s.map_x[(long long)k * m.n_chains + ch.chain] = x[k];
  }
  if (ch.since_save == 0) {
    long long rec = ch.chain * m.hist_cap + ch.slot;
    double *h = s.hist + rec * PTG_HX(D);
#pragma unroll
    for (int k = 0; k < D; k++) h[k] = x[k];
    s.hist_lp[2 * rec] = lpost; s.hist_lp[2 * rec + 1] = llike;
    if (m.record_full) {
      s.hist_acc[rec] = ch.naccept / (double)ch.ntries;
      s.hist_beta[rec] = beta;
      s.hist_type[rec] = ch.last_type;
    }
    ch.nsize++;
    ch.slot = (ch.slot + 1 == m.hist_cap) ? 0 : ch.slot + 1;
  }
  ch.since_save = (ch.since_save + 1 == m.save_every) ? 0 : ch.since_save + 1;
  ch.nhist++;
}

// physical address of history element `index` of the eligible window (newest min(nsize,cap) samples)
template <int D>
__device__ __forceinline__ const double *hist_elem(const PtgModel &m, const PtgState &s, const Chain<D> &ch, int index) {
  int p = index;
  if (ch.nsize > m.hist_cap) { p = ch.slot + index; if (p >= m.hist_cap) p -= m.hist_cap; }
  return s.hist + (ch.chain * m.hist_cap + p) * PTG_HX(D);
}
// lpost of that element (differential_evolution's unlikely_alpha test, proposal_distribution.cc:768)
template <int D>
__device__ __forceinline__ double hist_lpost(const PtgModel &m, const PtgState &s, const Chain<D> &ch, int index) {
  int p = index;
  if (ch.nsize > m.hist_cap) { p = ch.slot + index; if (p >= m.hist_cap) p -= m.hist_cap; }
  return s.hist_lp[2 * (ch.chain * m.hist_cap + p)];
}

// differential_evolution::draw_i_from_chain (proposal_distribution.cc:744-778)
template <int D, int MODE>
__device__ __forceinline__ int de_draw_index(const PtgModel &m, const PtgState &s, const Chain<D> &ch, const PtgProp &p,
                                             Stream<MODE> &rs, const uint32_t wA[4], const uint32_t wB[4], int which, int &attempt) {
  const uint32_t w0 = (which == 1) ? wA[3] : (which == 2 ? wB[0] : wB[1]); // PTG_IDX_BLK / PTG_IDX_WORD
  int size = (int)(ch.nsize > m.hist_cap ? (long long)m.hist_cap : ch.nsize);
  int start = 0, mins = D * 10, minc = D * 100;
  if ((size - minc) * (1 - p.ignore_frac) > mins) start = (int)((size - minc) * p.ignore_frac);
  double lpost0 = ch.map_lpost - D;
  double alpha = p.unlikely_alpha;
  while (true) {
    int a = attempt;
    uint32_t wr[4];
    double xrnd;
    if (a == 0 && alpha <= 0) { if constexpr (MODE == PTG_RNG_PHILOX) xrnd = ptg_u32_to_unit(w0); else xrnd = rs.next_u(); }
    else {
      if constexpr (MODE == PTG_RNG_PHILOX) {
        rs.fetch(PTG_BLK_RETRY + which * 0x100 + (a & 0xff), wr);
        xrnd = (a == 0) ? ptg_u32_to_unit(w0) : ptg_u32_to_unit(wr[0]);
      } else xrnd = rs.next_u();
    }
    attempt++;
    int index = (int)(start + (size - start) * xrnd);
    if (alpha > 0) {
      double lpost = hist_lpost<D>(m, s, ch, index);
      if (lpost0 > lpost) {
        double pr = exp(alpha * (lpost - lpost0));
        double x2 = rs.u32(wr, 1);
        if (x2 < pr) return index;
        alpha *= 0.9;
        continue;
      }
    }
    return index;
  }
}

struct MhOut { double lhr; int code; };

// A proposal between its generation and the Metropolis test: what MH_chain::step holds when it calls the likelihood
// (chain.cc:975-981).  The fused kernels evaluate a device functor in between; the host-callback mode (ptg_callback.cuh)
// parks this in global memory while the caller's likelihood runs.
template <int D>
struct MhPending {
  double newx[D];
  double newlprior, prop_lh;
  int type;
  bool valid, gate;    // gate: the likelihood is evaluated (chain.cc:980)
  uint32_t wacc[2];    // the acceptance draw's Philox words
};

// first half of MH_chain::step(prop) (chain.cc:966-980) with proposal_distribution_set::draw (proposal_distribution.cc:99-129):
// draw, enforce, prior, likelihood gate
template <int D, int MODE>
__device__ __forceinline__ void mh_propose(const PtgModel &m, const PtgState &s, Chain<D> &ch, Stream<MODE> &rs,
                                           const double *__restrict__ bins, MhPending<D> &pend) {
  double (&newx)[D] = pend.newx;
  double prop_lh = 0;
  int type = 0;
  bool valid = m.zero_valid != 0;
  const double oldlprior = ch.lpost - ch.beta * ch.llike;

  uint32_t wsel[4], widx[4]; // blocks A and B of this step (include/ptmcmc_b200_rng.h)
  rs.fetch(PTG_BLK_A, wsel);
  rs.fetch(PTG_BLK_B, widx);
  // ---- member selection
  int member = 0;
  if (m.wrap_in_set) {
    const int hsize = (int)(ch.nsize > m.hist_cap ? (long long)m.hist_cap : ch.nsize);
    member = -1;
    for (int count = 0; count <= 100 && member < 0; count++) {
      double x = (m.n_props > 1) ? rs.u32(wsel, 0) : 0.0;
      for (int i = 0; i < m.n_props; i++) {
        bool ready = (m.props[i].kind != PTG_PROP_DE) || hsize >= D * 10; // differential_evolution::is_ready
        if (ready && x < bins[i]) { member = i; break; }
      }
      if constexpr (MODE == PTG_RNG_PHILOX) break;
    }
    if (member < 0) { rs.err = 2; member = 0; }
  }
  const PtgProp &p = m.props[member];
  if (p.kind == PTG_PROP_DE) {
    // differential_evolution::draw (proposal_distribution.cc:790-801)
    double usnk = rs.u32(wsel, 1);
    double ug = rs.u32(wsel, 2);
    if (!(p.snooker > usnk)) {
      // draw_standard (proposal_distribution.cc:489-535)
      double gamma = p.gamma_std;
      if (ug < p.g1frac) gamma = 1;
      int a1 = 0, a2 = 0;
      int i1 = de_draw_index<D, MODE>(m, s, ch, p, rs, wsel, widx, 1, a1);
      int i2 = de_draw_index<D, MODE>(m, s, ch, p, rs, wsel, widx, 2, a2);
      const double *s1 = hist_elem<D>(m, s, ch, i1), *s2 = hist_elem<D>(m, s, ch, i2);
      if constexpr (MODE == PTG_RNG_TAPE) { // the d normals of the discarded jitter are still consumed (H8-1)
        for (int j = 0; j < D; j++) (void)rs.next_z();
      }
#pragma unroll
      for (int i = 0; i < D; i++) {
        double t = ch.x[i] + s1[i] * gamma;
        newx[i] = t + s2[i] * (-gamma);
      }
      type = 0;
    } else {
      // draw_snooker (proposal_distribution.cc:538-591)
      double gamma = (1.2 + ug) / p.reduce_gamma;
      double smznorm2 = 0, minusz[D], smz[D];
      int az = 0, isafe = 0;
      while (smznorm2 == 0) {
        int iz = de_draw_index<D, MODE>(m, s, ch, p, rs, wsel, widx, 0, az);
        const double *z = hist_elem<D>(m, s, ch, iz);
        smznorm2 = 0;
#pragma unroll
        for (int i = 0; i < D; i++) { minusz[i] = z[i] * (-1); smz[i] = ch.x[i] + minusz[i]; }
#pragma unroll
        for (int i = 0; i < D; i++) smznorm2 += smz[i] * smz[i];
        if (++isafe > 1000) break;
      }
      int a1 = 0, a2 = 0;
      int i1 = de_draw_index<D, MODE>(m, s, ch, p, rs, wsel, widx, 1, a1);
      int i2 = de_draw_index<D, MODE>(m, s, ch, p, rs, wsel, widx, 2, a2);
      const double *s1 = hist_elem<D>(m, s, ch, i1), *s2 = hist_elem<D>(m, s, ch, i2);
      double dot = 0;
#pragma unroll
      for (int i = 0; i < D; i++) {
        double ds12 = s1[i] * gamma + s2[i] * (-gamma);
        dot += ds12 * smz[i];
      }
      double fac = dot / smznorm2, pmz2 = 0;
#pragma unroll
      for (int i = 0; i < D; i++) {
        newx[i] = ch.x[i] + smz[i] * fac;
        double pmz = newx[i] + minusz[i];
        pmz2 += pmz * pmz;
      }
      prop_lh = (log(pmz2) - log(smznorm2)) * (D - 1) / 2.0;
      type = 1;
    }
  } else if (p.kind == PTG_PROP_GAUSS) {
    // gaussian_prop::draw (proposal_distribution.hh:194-218)
    double off[D];
    draw_normals<D, MODE>(rs, off);
    const double *__restrict__ sig = m.prop_data + p.sigma_off;
#pragma unroll
    for (int i = 0; i < D; i++) off[i] = off[i] * __ldg(sig + i) + 0.0;
    double x1 = 1;
    if (p.one_d_frac > 0) x1 = rs.u32(wsel, 1);
    if (p.one_d_frac > 0 && x1 < p.one_d_frac) {
      int ia = (int)(D * rs.u32(wsel, 2));
#pragma unroll
      for (int j = 0; j < D; j++) if (j != ia) off[j] = 0.0;
      type = 1;
    } else type = 0;
    if (p.has_transform) {
      const double *__restrict__ M = m.prop_data + p.trans_off;
      double t[D];
      // summation order of Eigen 3.3.7's column-major GEMV, which is what `vec=diagTransform*vec`
      // (proposal_distribution.hh:212) runs: four columns at a time as (a0v0+a1v1)+(a2v2+a3v3) on packet rows,
      // sequentially on an odd trailing row, leftover columns one by one
      constexpr int CB = (D / 4) * 4, EVEN_ROWS = D & ~1;
#pragma unroll
      for (int i = 0; i < D; i++) {
        const double *__restrict__ a = M + i * D;
        double acc = 0;
#pragma unroll
        for (int j = 0; j < CB; j += 4) {
          if (i < EVEN_ROWS) acc = acc + ((__ldg(a + j) * off[j] + __ldg(a + j + 1) * off[j + 1]) + (__ldg(a + j + 2) * off[j + 2] + __ldg(a + j + 3) * off[j + 3]));
          else { acc = __ldg(a + j) * off[j] + acc; acc = __ldg(a + j + 1) * off[j + 1] + acc; acc = __ldg(a + j + 2) * off[j + 2] + acc; acc = __ldg(a + j + 3) * off[j + 3] + acc; }
        }
#pragma unroll
        for (int j = CB; j < D; j++) acc += __ldg(a + j) * off[j];
        t[i] = acc;
      }
#pragma unroll
      for (int i = 0; i < D; i++) off[i] = t[i];
    }
#pragma unroll
    for (int i = 0; i < D; i++) newx[i] = ch.x[i] + off[i];
  } else {
    // draw_from_dist::draw (proposal_distribution.hh:124-129)
    valid = prior_draw<D, MODE>(m, rs, PTG_BLK_PRIOR, newx);
    prop_lh = prior_eval_log<D>(m, ch.x, true) - prior_eval_log<D>(m, newx, valid);
    type = 0;
  }
  if (m.wrap_in_set) type = member + 10 * type;

  // ---- enforce, prior, likelihood gate (chain.cc:976-980)
  if (valid) valid = space_enforce<D>(m, newx);
  const double newlprior = prior_eval_log<D>(m, newx, valid);
  pend.newlprior = newlprior; pend.prop_lh = prop_lh; pend.type = type; pend.valid = valid;
  pend.gate = valid && ((newlprior > -1e200) || (newlprior - oldlprior > m.dprior_min));
  pend.wacc[0] = widx[2]; pend.wacc[1] = widx[3];
}

// second half of MH_chain::step (chain.cc:981-1019): posterior, Metropolis test, add_state.  newlike is ignored unless pend.gate.
template <int D, int MODE>
__device__ __forceinline__ MhOut mh_finish(const PtgModel &m, const PtgState &s, Chain<D> &ch, Stream<MODE> &rs, const MhPending<D> &pend,
                                           double newlike) {
  const double (&newx)[D] = pend.newx;
  const double newlprior = pend.newlprior;
  const int type = pend.type;
  const bool valid = pend.valid;
  double newlpost;
  int code = 0;
  bool accept = true;
  if (pend.gate) newlpost = newlike * ch.beta + newlprior;
  else { newlike = newlpost = -CUDART_INF; code |= PTG_TRACE_NOLIKE; }
  // ---- Metropolis test (chain.cc:989-1001)
  double lhr = pend.prop_lh;
  if (isnan(lhr)) accept = false;
  lhr += newlpost - ch.lpost;
  if (!valid) { accept = false; code |= PTG_TRACE_INVALID; }
  if (accept && lhr < 0) {
    double u;
    if constexpr (MODE == PTG_RNG_PHILOX) u = ptg_u52_to_unit(pend.wacc[0], pend.wacc[1]); else u = rs.next_u();
    accept = (log(u) < lhr);
  }
  ch.ntries++;
  if (accept) {
    ch.naccept++;
    ch.last_type = type;
#pragma unroll
    for (int i = 0; i < D; i++) ch.x[i] = newx[i];
    ch.llike = newlike; ch.lpost = newlpost; ch.lprior = newlprior;
    code |= PTG_TRACE_ACCEPT;
  }
  chain_append<D>(m, s, ch, ch.x, ch.llike, ch.lpost, ch.beta);
  code |= (type & PTG_TRACE_TYPE_MASK);
  MhOut o; o.lhr = lhr; o.code = code;
  return o;
}

// MH_chain::step(prop) (chain.cc:966-1022), fused with a device likelihood functor
template <int D, int MODE>
__device__ __forceinline__ MhOut mh_step(const PtgModel &m, const PtgState &s, Chain<D> &ch, Stream<MODE> &rs,
                                         const double *__restrict__ bins) {
  MhPending<D> pend;
  mh_propose<D, MODE>(m, s, ch, rs, bins, pend);
  double newlike = -CUDART_INF;
  if (pend.gate) newlike = like_eval<D>(m, pend.newx);
  return mh_finish<D, MODE>(m, s, ch, rs, pend, newlike);
}

template <int MODE>
__device__ __forceinline__ void stream_open(const PtgModel &m, const PtgState &s, Stream<MODE> &rs, long long stream_index,
                                            uint64_t id, int domain) {
  rs.seed = m.seed; rs.id = id; rs.step = 0; rs.domain = domain; rs.err = 0;
  if constexpr (MODE == PTG_RNG_TAPE) {
    rs.ut = s.tape_u; rs.zt = s.tape_z;
    rs.upos = s.u_pos[stream_index]; rs.uend = s.u_end[stream_index];
    rs.zpos = s.z_pos[stream_index]; rs.zend = s.z_end[stream_index];
  } else { rs.ut = rs.zt = nullptr; rs.upos = rs.uend = rs.zpos = rs.zend = 0; }
}
// a stream that addresses nothing (ghost lanes of the warp kernel)
template <int MODE>
__device__ __forceinline__ void stream_blank(const PtgModel &m, Stream<MODE> &rs) {
  rs.seed = m.seed; rs.id = 0; rs.step = 0; rs.domain = PTG_DOMAIN_STEP; rs.err = 0;
  rs.ut = rs.zt = nullptr; rs.upos = rs.uend = rs.zpos = rs.zend = 0;
}
template <int MODE>
__device__ __forceinline__ void stream_close(const PtgState &s, Stream<MODE> &rs, long long stream_index) {
  if constexpr (MODE == PTG_RNG_TAPE) { s.u_pos[stream_index] = rs.upos; s.z_pos[stream_index] = rs.zpos; }
  if (rs.err) atomicMax(s.err, rs.err);
}

// ------------------------------------------------------------------------------------------------- init
// MH_chain::initialize(n) (chain.cc:846-876): n_init prior draws per chain (redrawn while invalid or
// llike < -1e100), each appended to the history without thinning; Nhist = 0 afterwards.
// init_x != nullptr: states provided by the caller instead ([n_chains][n_init][D]).
template <int D, int MODE>
__global__ void __launch_bounds__(128) ptg_init_kernel(const __grid_constant__ PtgModel m, PtgState s, const double *__restrict__ init_x) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m.n_chains) return;
  const int R = m.n_rungs;
  long long ladder = c / R; int rung = (int)(c - ladder * R);
  Chain<D> ch;
  chain_load<D>(m, s, c, ch);
  Stream<MODE> rs;
  stream_open<MODE>(m, s, rs, c, (uint64_t)(m.ladder_offset + ladder) * PTG_STREAM_STRIDE + (uint64_t)rung, PTG_DOMAIN_INIT);
  for (int k = 0; k < m.n_init; k++) {
    double x[D];
    if (init_x) {
#pragma unroll
      for (int i = 0; i < D; i++) x[i] = init_x[((long long)c * m.n_init + k) * D + i];
    } else {
      rs.step = (uint64_t)k;
      int icnt = 0;
      bool valid = prior_draw<D, MODE>(m, rs, 0, x);
      while (!valid || like_eval<D>(m, x) < -1e100) {
        icnt++;
        if (icnt >= 100000) { rs.err = 4; break; }
        valid = prior_draw<D, MODE>(m, rs, (uint32_t)icnt * PTG_INIT_ATTEMPT_STRIDE, x);
      }
    }
    double ll = like_eval<D>(m, x);
    double lp = prior_eval_log<D>(m, x, true);
#pragma unroll
    for (int i = 0; i < D; i++) ch.x[i] = x[i];
    ch.llike = ll; ch.lprior = lp; ch.lpost = lp + ch.beta * ll;
    ch.nhist = 0; ch.since_save = 0; // "As long as Nhist remains zero we will add the state regardless of add_every_N"
    chain_append<D>(m, s, ch, ch.x, ch.llike, ch.lpost, ch.beta);
  }
  ch.nhist = 0; ch.since_save = 0;
  chain_store<D>(m, s, ch);
  stream_close<MODE>(s, rs, c);
}

// ------------------------------------------------------------------------------------------------- batch evaluation
// the device likelihood / prior functors applied to caller-provided states x[n][D] (row-major): the batched form of
// bayes_likelihood::evaluate_log (bayesian.hh:553-581) and sampleable_probability_function::evaluate_log after
// state::enforce (probability_function.hh:59, states.cc:161-166)
template <int D>
__global__ void __launch_bounds__(128) ptg_eval_kernel(const __grid_constant__ PtgModel m, const double *__restrict__ x, long long n,
                                                       double *__restrict__ out_ll, double *__restrict__ out_lp) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v[D];
#pragma unroll
  for (int k = 0; k < D; k++) v[k] = x[i * D + k];
  if (out_ll) out_ll[i] = like_eval<D>(m, v);
  if (out_lp) {
    bool valid = space_enforce<D>(m, v);
    out_lp[i] = prior_eval_log<D>(m, v, valid);
  }
}

// ------------------------------------------------------------------------------------------------- host-callback likelihood
// completes the MH step of every chain that parked a proposal (ptg_step_kernel phase 1), with the caller's log-likelihoods
template <int D>
__global__ void __launch_bounds__(128) ptg_cb_finish_kernel(const __grid_constant__ PtgModel m, PtgState s, long long step) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m.n_chains) return;
  const int flags = s.pend_flags[c];
  if (!(flags & 4)) return;
  Chain<D> ch;
  chain_load<D>(m, s, c, ch);
  MhPending<D> pend;
#pragma unroll
  for (int k = 0; k < D; k++) pend.newx[k] = s.pend_x[c * D + k];
  pend.newlprior = s.pend_lprior[c]; pend.prop_lh = s.pend_lh[c]; pend.type = s.pend_type[c];
  pend.valid = (flags & 1) != 0; pend.gate = (flags & 2) != 0;
  pend.wacc[0] = s.pend_w[2 * c]; pend.wacc[1] = s.pend_w[2 * c + 1];
  double newlike = s.pend_like[c];
  if (!isfinite(newlike)) newlike = -CUDART_INF; // bayesian.hh:569-575
  Stream<PTG_RNG_PHILOX> rs;
  stream_blank<PTG_RNG_PHILOX>(m, rs);
  const MhOut o = mh_finish<D, PTG_RNG_PHILOX>(m, s, ch, rs, pend, newlike);
  chain_store<D>(m, s, ch);
  if (step < m.trace_steps) { s.trace_lhr[step * m.n_chains + c] = o.lhr; s.trace_code[step * m.n_chains + c] = o.code; }
}
// MH_chain::initialize with a host likelihood, one prior draw per chain and round: chains whose flag is still 0 draw sample `k`
// (attempt[c]) into pend_x; after the caller evaluated pend_like, ptg_cb_init_accept_kernel appends the accepted ones
template <int D>
__global__ void __launch_bounds__(128) ptg_cb_init_draw_kernel(const __grid_constant__ PtgModel m, PtgState s, int k, const int32_t *__restrict__ attempt) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m.n_chains || s.pend_flags[c] == 8) return; // 8 = sample k already accepted
  const int R = m.n_rungs;
  const long long ladder = c / R; const int rung = (int)(c - ladder * R);
  Stream<PTG_RNG_PHILOX> rs;
  stream_blank<PTG_RNG_PHILOX>(m, rs);
  rs.id = (uint64_t)(m.ladder_offset + ladder) * PTG_STREAM_STRIDE + (uint64_t)rung; rs.domain = PTG_DOMAIN_INIT; rs.step = (uint64_t)k;
  double x[D];
  const bool valid = prior_draw<D, PTG_RNG_PHILOX>(m, rs, (uint32_t)attempt[c] * PTG_INIT_ATTEMPT_STRIDE, x);
#pragma unroll
  for (int i = 0; i < D; i++) s.pend_x[c * D + i] = x[i];
  s.pend_flags[c] = valid ? 3 : 0; // valid + likelihood wanted
}
template <int D>
__global__ void __launch_bounds__(128) ptg_cb_init_accept_kernel(const __grid_constant__ PtgModel m, PtgState s, int32_t *attempt, int32_t *n_open) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m.n_chains || s.pend_flags[c] == 8) return;
  double ll = s.pend_like[c];
  if (!isfinite(ll)) ll = -CUDART_INF;
  if (!(s.pend_flags[c] & 1) || ll < -1e100) { // chain.cc:854-866: redraw
    attempt[c]++;
    if (attempt[c] >= 100000) atomicMax(s.err, 4); else atomicAdd(n_open, 1);
    return;
  }
  Chain<D> ch;
  chain_load<D>(m, s, c, ch);
  double x[D];
#pragma unroll
  for (int i = 0; i < D; i++) { x[i] = s.pend_x[c * D + i]; ch.x[i] = x[i]; }
  const double lp = prior_eval_log<D>(m, x, true);
  ch.llike = ll; ch.lprior = lp; ch.lpost = lp + ch.beta * ll;
  ch.nhist = 0; ch.since_save = 0;
  chain_append<D>(m, s, ch, ch.x, ch.llike, ch.lpost, ch.beta);
  ch.nhist = 0;
  chain_store<D>(m, s, ch);
  s.pend_flags[c] = 8;
  attempt[c] = 0;
}

// ------------------------------------------------------------------------------------------------- PT step
// shared-memory carve-up for one ladder
template <int D>
struct LadderShared {
  double *sx, *sll, *slpost, *slprior, *sbeta;   // published current state      [R][D], [R]...
  double *ax, *all_, *alpost, *abeta;            // append snapshots             [R][D], [R], [R][2], [R][2]
  double *udraw, *split;                          // ladder draws [PTG_SWAP_SLOTS][3]; pry_temps scratch [R]
  int *napp, *iswap, *dir, *ups, *downs, *inst;   // [R], [PTG_SWAP_SLOTS], [R]x4
  long long *scount, *saccept;                    // [R]
  __device__ void carve(unsigned char *base, int R) {
    double *d = reinterpret_cast<double *>(base);
    sx = d; d += (size_t)R * D; sll = d; d += R; slpost = d; d += R; slprior = d; d += R; sbeta = d; d += R;
    ax = d; d += (size_t)R * D; all_ = d; d += R; alpost = d; d += 2 * R; abeta = d; d += 2 * R;
    udraw = d; d += PTG_SWAP_SLOTS * 3; split = d; d += R;
    long long *l = reinterpret_cast<long long *>(d);
    scount = l; l += R; saccept = l; l += R;
    int *i = reinterpret_cast<int *>(l);
    napp = i; i += R; iswap = i; i += PTG_SWAP_SLOTS; dir = i; i += R; ups = i; i += R; downs = i; i += R; inst = i; i += R;
  }
};

// record one history append of rung r with the ladder's shared copies as they are now
template <int D>
__device__ __forceinline__ void swap_record_append(LadderShared<D> &L, int r) {
  int k = L.napp[r];
  if (k == 0) {
#pragma unroll
    for (int i = 0; i < D; i++) L.ax[r * D + i] = L.sx[r * D + i];
    L.all_[r] = L.sll[r];
  }
  L.alpost[2 * r + k] = L.slpost[r];
  L.abeta[2 * r + k] = L.sbeta[r];
  L.napp[r] = k + 1;
}

// the serial swap phase of one ladder, executed by its lane 0 on the shared copies
template <int D, int MODE>
__device__ __forceinline__ void swap_phase_leader(const PtgModel &m, LadderShared<D> &L, Stream<MODE> &ls, int ntrial) {
  const int R = m.n_rungs;
  for (int j = 0; j < ntrial; j++) {
    const int i = L.iswap[j];
    if (i < 0) continue;
    bool accept = true;
    if (i > 0) {
      if (L.dir[i] > 0) L.ups[i]++;
      if (L.dir[i] < 0) L.downs[i]++;
    }
    double lla = L.sll[i]; if (!(lla > -1e200)) lla = -1e200;
    double llb = L.sll[i + 1]; if (!(llb > -1e200)) llb = -1e200;
    double lhr = -(L.sbeta[i + 1] - L.sbeta[i]) * (llb - lla);
    if (lhr < 0) {
      double u;
      if constexpr (MODE == PTG_RNG_PHILOX) u = L.udraw[3 * j + 2];
      else u = ls.next_u();
      accept = (log(u) < lhr);
    }
    if (accept) {
      // exchange (state, llike); lpost recomputed = lprior(x) + beta*llike (chain.cc:1487-1490, 925-928)
#pragma unroll
      for (int k = 0; k < D; k++) { double t = L.sx[i * D + k]; L.sx[i * D + k] = L.sx[(i + 1) * D + k]; L.sx[(i + 1) * D + k] = t; }
      { double t = L.sll[i]; L.sll[i] = L.sll[i + 1]; L.sll[i + 1] = t; }
      { double t = L.slprior[i]; L.slprior[i] = L.slprior[i + 1]; L.slprior[i + 1] = t; }
      L.slpost[i + 1] = L.slprior[i + 1] + L.sbeta[i + 1] * L.sll[i + 1];
      L.slpost[i] = L.slprior[i] + L.sbeta[i] * L.sll[i];
      swap_record_append<D>(L, i + 1);
      swap_record_append<D>(L, i);
      { int t = L.dir[i]; L.dir[i] = L.dir[i + 1]; L.dir[i + 1] = t; }
      { int t = L.inst[i]; L.inst[i] = L.inst[i + 1]; L.inst[i + 1] = t; }
      if (i == 0) L.dir[i] = 1;
      if (i + 1 == R - 1) L.dir[i + 1] = -1;
      L.saccept[i]++;
      if (m.evolve_rate > 0) {
        // pry_temps, vector version with one pried gap (chain.cc:1809-1846) + resetTemp (chain.cc:1088-1091)
        const double rate = m.evolve_rate;
        for (int k = 0; k < R - 1; k++) {
          double sp = L.sbeta[k] - L.sbeta[k + 1];
          if (m.evolve_lpost_cut >= 0 && L.slpost[k] - L.slpost[k + 1] > m.evolve_lpost_cut * L.sbeta[k]) sp *= (1.0 + rate);
          L.split[k] = sp;
        }
        L.split[i] *= 1.0 + rate;
        double sum = 0;
        for (int k = 0; k < R - 1; k++) sum += L.split[k];
        double norm = sum / (1 - L.sbeta[R - 1]);
        double invtemp = 1;
        for (int k = 1; k < R - 1; k++) {
          invtemp -= L.split[k - 1] / norm;
          L.sbeta[k] = invtemp;
          L.slpost[k] = L.slprior[k] + invtemp * L.sll[k];
        }
      }
    } else {
      swap_record_append<D>(L, i);
      swap_record_append<D>(L, i + 1);
    }
    L.scount[i]++;
  }
}

// The fused PT-step kernel.  blockDim.x = lpb * n_rungs; dynamic shared memory = lpb * ptg_ladder_shared_bytes + bins.
template <int D, int MODE>
// phase 0: fused step.  phase 1 (host-callback likelihood, n_steps = 1): swap phase + proposal, the MH lanes park their proposal in
// s.pend_* and ptg_cb_finish_kernel completes the step once the caller's likelihood values are in s.pend_like.
__global__ void __launch_bounds__(256) ptg_step_kernel(const __grid_constant__ PtgModel m, PtgState s, long long step0, int n_steps, int lpb, int phase) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int R = m.n_rungs;
  const int ll_ = threadIdx.x / R;              // local ladder
  const int rung = threadIdx.x - ll_ * R;
  const long long ladder = (long long)blockIdx.x * lpb + ll_;
  const bool active = (ll_ < lpb) && (ladder < m.n_ladders);
  const long long chain = ladder * R + rung;
  const size_t lbytes = (sizeof(double) * ((size_t)R * D * 2 + (size_t)R * 10 + PTG_SWAP_SLOTS * 3) + sizeof(int) * ((size_t)R * 5 + PTG_SWAP_SLOTS) +
                         sizeof(long long) * (size_t)R * 2 + 15) & ~(size_t)15; // = ptg_ladder_shared_bytes(D, R) of ptg_launch.h
  LadderShared<D> L;
  L.carve(smem_raw + (size_t)(ll_ < lpb ? ll_ : 0) * lbytes, R);
  double *sbins = reinterpret_cast<double *>(smem_raw + (size_t)lpb * lbytes); // [R][n_props]
  for (int i = threadIdx.x; i < R * m.n_props; i += blockDim.x) sbins[i] = m.bins[i];

  Chain<D> ch;
  Stream<MODE> rs, ls;
  const int maxswaps = m.maxswaps;
  const int ntrial = (m.swap_mode == PTG_SWAP_REFERENCE) ? maxswaps : PTG_SWAP_SLOTS;
  if (active) {
    chain_load<D>(m, s, chain, ch);
    stream_open<MODE>(m, s, rs, chain, (uint64_t)(m.ladder_offset + ladder) * PTG_STREAM_STRIDE + (uint64_t)rung, PTG_DOMAIN_STEP);
    stream_open<MODE>(m, s, ls, m.n_chains + ladder, (uint64_t)(m.ladder_offset + ladder) * PTG_STREAM_STRIDE + PTG_STREAM_LADDER,
                      PTG_DOMAIN_STEP);
    L.dir[rung] = s.directions[chain]; L.ups[rung] = s.ups[chain]; L.downs[rung] = s.downs[chain]; L.inst[rung] = s.instances[chain];
    L.scount[rung] = 0; L.saccept[rung] = 0;
  }
  const double swap_thresh = (R - 1) * m.swap_rate / maxswaps; // chain.cc:1413
  double ptry = 2 * m.swap_rate; if (ptry > 1) ptry = 1;       // even/odd mode
  __syncthreads();

  for (int it = 0; it < n_steps; it++) {
    const uint64_t step = (uint64_t)(step0 + it);
    if constexpr (MODE == PTG_RNG_TAPE) {
      // re-synchronise the cursors with the recorded run at every step boundary, so that a last-ulp libm difference
      // that changes HOW MANY draws one step consumes (lhr = -1e-21 vs 0: same decision, one draw fewer) stays local
      if (active && s.u_mark && (long long)step < s.n_mark_steps) {
        const long long ns = m.n_chains + m.n_ladders, row = (long long)step * ns;
        rs.upos = s.u_mark[row + chain]; rs.zpos = s.z_mark[row + chain];
        ls.upos = s.u_mark[row + m.n_chains + ladder]; ls.zpos = s.z_mark[row + m.n_chains + ladder];
      }
    }
    if (active) {
      // 1. publish
#pragma unroll
      for (int k = 0; k < D; k++) L.sx[rung * D + k] = ch.x[k];
      L.sll[rung] = ch.llike; L.slpost[rung] = ch.lpost; L.slprior[rung] = ch.lprior; L.sbeta[rung] = ch.beta;
      L.napp[rung] = 0;
      if constexpr (MODE == PTG_RNG_PHILOX) {
        ls.step = step;
        if (m.swap_mode == PTG_SWAP_REFERENCE) {
          for (int j = rung; j < maxswaps; j += R) {
            uint32_t w[4]; ls.fetch((uint32_t)j, w);
            L.udraw[3 * j] = ptg_u32_to_unit(w[0]); L.udraw[3 * j + 1] = ptg_u32_to_unit(w[1]);
            L.udraw[3 * j + 2] = ptg_u52_to_unit(w[2], w[3]);
          }
        } else {
          // even/odd: lane i owns pair (i,i+1); slot index = position in the candidate list
          const int parity = (int)(step & 1);
          if (rung + 1 < R && ((rung & 1) == parity)) {
            uint32_t w[4]; ls.fetch(PTG_BLK_SWAP_EVENODD + (uint32_t)rung, w);
            int slot = rung >> 1;
            L.udraw[3 * slot] = ptg_u52_to_unit(w[0], w[1]); L.udraw[3 * slot + 2] = ptg_u52_to_unit(w[2], w[3]);
          }
        }
      }
    }
    __syncthreads();
    // 2. swap phase (lane 0 of each ladder)
    if (active && rung == 0) {
      if (m.swap_mode == PTG_SWAP_REFERENCE) {
        for (int i = 0; i < maxswaps; i++) { // chain.cc:1410-1420
          int cand = -2;
          double x;
          if constexpr (MODE == PTG_RNG_PHILOX) x = L.udraw[3 * i]; else x = ls.next_u();
          if (R > 1 && x < swap_thresh) {
            if constexpr (MODE == PTG_RNG_PHILOX) x = L.udraw[3 * i + 1]; else x = ls.next_u();
            cand = (int)(x * (R - 1));
            for (int j = 0; j < i; j++)
              if (L.iswap[j] == cand || L.iswap[j] + 1 == cand) cand = -2;
          }
          L.iswap[i] = cand;
        }
        swap_phase_leader<D, MODE>(m, L, ls, maxswaps);
      } else {
        const int parity = (int)(step & 1);
        int n = 0;
        // candidate list in pair order; Philox slots were filled per pair, compact them to trial order
        for (int i = parity; i + 1 < R; i += 2) {
          double x;
          if constexpr (MODE == PTG_RNG_PHILOX) x = L.udraw[3 * (i >> 1)]; else x = ls.next_u();
          if (x < ptry) {
            if constexpr (MODE == PTG_RNG_PHILOX) L.udraw[3 * n + 2] = L.udraw[3 * (i >> 1) + 2]; // n <= i>>1: no clobber of unread slots
            L.iswap[n++] = i;
          }
        }
        for (int i = n; i < PTG_SWAP_SLOTS; i++) L.iswap[i] = -2;
        swap_phase_leader<D, MODE>(m, L, ls, n);
      }
    }
    __syncthreads();
    // 3. per-lane: swap appends or MH step
    if (active) {
      double lhr = 0; int code = PTG_TRACE_SWAPPED;
      const int na = L.napp[rung];
      ch.beta = L.sbeta[rung];
      if (na > 0) {
        double x0[D];
#pragma unroll
        for (int k = 0; k < D; k++) { x0[k] = L.ax[rung * D + k]; ch.x[k] = L.sx[rung * D + k]; }
        ch.llike = L.sll[rung]; ch.lprior = L.slprior[rung]; ch.lpost = L.slpost[rung];
        chain_append<D>(m, s, ch, x0, L.all_[rung], L.alpost[2 * rung], L.abeta[2 * rung]);
        if (na > 1) chain_append<D>(m, s, ch, ch.x, ch.llike, L.alpost[2 * rung + 1], L.abeta[2 * rung + 1]);
      } else {
        if (m.evolve_rate > 0) ch.lpost = L.slpost[rung]; // resetTemp may have touched any interior rung
        rs.step = step;
        if (phase == 0) {
          MhOut o = mh_step<D, MODE>(m, s, ch, rs, sbins + rung * m.n_props);
          lhr = o.lhr; code = o.code;
        } else {
          MhPending<D> pend;
          mh_propose<D, MODE>(m, s, ch, rs, sbins + rung * m.n_props, pend);
#pragma unroll
          for (int k = 0; k < D; k++) s.pend_x[chain * D + k] = pend.newx[k];
          s.pend_lprior[chain] = pend.newlprior; s.pend_lh[chain] = pend.prop_lh; s.pend_type[chain] = pend.type;
          s.pend_w[2 * chain] = pend.wacc[0]; s.pend_w[2 * chain + 1] = pend.wacc[1];
          s.pend_flags[chain] = (pend.valid ? 1 : 0) | (pend.gate ? 2 : 0) | 4;
        }
      }
      if (phase == 1 && na > 0) s.pend_flags[chain] = 0;
      if ((long long)step < m.trace_steps && !(phase == 1 && na == 0)) {
        s.trace_lhr[step * m.n_chains + chain] = lhr;
        s.trace_code[step * m.n_chains + chain] = code;
      }
    }
    __syncthreads();
  }
  if (active) {
    chain_store<D>(m, s, ch);
    stream_close<MODE>(s, rs, chain);
    if (rung == 0) stream_close<MODE>(s, ls, m.n_chains + ladder);
    s.directions[chain] = L.dir[rung]; s.ups[chain] = L.ups[rung]; s.downs[chain] = L.downs[rung]; s.instances[chain] = L.inst[rung];
    s.swap_count[chain] += L.scount[rung]; s.swap_accept[chain] += L.saccept[rung];
  }
}
