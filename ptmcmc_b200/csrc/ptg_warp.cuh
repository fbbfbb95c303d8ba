// ptg_warp.cuh -- second-generation fused PT step kernel: a ladder lives in ONE warp, in registers.
//
// Layout: one thread = one chain; the n_rungs chains of a ladder occupy W consecutive lanes of a warp (W = 8, 16 or 32,
// the smallest that holds n_rungs), 32/W ladders per warp.  Nothing is staged in shared memory and the kernel has no
// block-level barrier inside the step loop:
//   * replica swaps (chain.cc:1410-1538) run on warp shuffles -- the ladder's Philox draws are produced one trial per lane,
//     every lane of the ladder replays the (tiny, serial) candidate list redundantly, and an accepted trial exchanges
//     (state, llike, lprior) between the two lanes with __shfl_sync; temperature evolution (pry_temps, chain.cc:1809-1846)
//     is a shuffle scan in the reference's summation order;
//   * the MH update (chain.cc:966-1022) is written so that the expensive parts are warp-converged: both Philox blocks of
//     a step, the history gathers, boundary / prior / likelihood and the Metropolis test are executed by all lanes
//     together; the Box-Muller normals that only Gaussian-proposal lanes need are produced COOPERATIVELY -- every lane of
//     the warp evaluates one (chain, pair) work item addressed by the owner's Philox stream and the results are shuffled
//     back to their owners -- instead of by the ~20 % of lanes that took the Gaussian branch while the rest idle.
// Arithmetic is the same unfused fp64 expression tree as the first-generation kernel and the oracle (-fmad=false).
#pragma once
#include "ptg_kernels.cuh"

struct WLane {
  int W, R, rung, lane;
  unsigned gmask;        // lanes of this ladder's group
  long long ladder;
  bool ladder_ok, active;
};

__device__ __forceinline__ double wshfl(unsigned mask, double v, int src, int W) { return __shfl_sync(mask, v, src, W); }

// per-lane ladder statistics (chain.hh:238-246), deltas are flushed at the end of the launch
struct WStats {
  int dir, ups, downs, inst;
  int scount, saccept;
};

// ------------------------------------------------------------------------------------------------- swap trial
// one swap trial of pair (c, c+1) of this lane's ladder (chain.cc:1436-1538); executed by all lanes of the group.
// `u_swap` is consulted only if lhr < 0 (PHILOX: already fetched; TAPE: read now from the ladder tape).
template <int D, int MODE>
__device__ __forceinline__ void wswap_trial(const PtgModel &m, const PtgState &s, const WLane &w, Chain<D> &ch, WStats &st,
                                            Stream<MODE> &ls, int c, double u_swap_philox, int &napp) {
  const int W = w.W, R = w.R;
  const unsigned gm = w.gmask;
  const bool is_lo = (w.rung == c), is_hi = (w.rung == c + 1), involved = is_lo || is_hi;
  if (is_lo && c > 0) {
    if (st.dir > 0) st.ups++;
    if (st.dir < 0) st.downs++;
  }
  double lla = wshfl(gm, ch.llike, c, W); if (!(lla > -1e200)) lla = -1e200;
  double llb = wshfl(gm, ch.llike, c + 1, W); if (!(llb > -1e200)) llb = -1e200;
  const double ba = wshfl(gm, ch.beta, c, W), bb = wshfl(gm, ch.beta, c + 1, W);
  const double lhr = -(bb - ba) * (llb - lla);
  bool accept = true;
  if (lhr < 0) {
    double u;
    if constexpr (MODE == PTG_RNG_PHILOX) u = u_swap_philox; else u = ls.next_u();
    accept = (log(u) < lhr);
  }
  if (accept) {
    // exchange (state, llike); lpost recomputed = lprior(x) + beta*llike (chain.cc:1487-1490, 925-928)
    const int partner = is_lo ? c + 1 : (is_hi ? c : w.rung);
#pragma unroll
    for (int k = 0; k < D; k++) { double v = wshfl(gm, ch.x[k], partner, W); ch.x[k] = v; }
    { double v = wshfl(gm, ch.llike, partner, W); ch.llike = v; }
    { double v = wshfl(gm, ch.lprior, partner, W); ch.lprior = v; }
    if (involved) ch.lpost = ch.lprior + ch.beta * ch.llike;
  }
  if (involved && w.active) {
    // reference order: chain i+1 appends, then chain i -- independent chains, so the lanes append concurrently
    chain_append<D>(m, s, ch, ch.x, ch.llike, ch.lpost, ch.beta);
    napp++;
  }
  if (accept) {
    const int partner = is_lo ? c + 1 : (is_hi ? c : w.rung);
    { int v = __shfl_sync(gm, st.dir, partner, W); st.dir = v; }
    { int v = __shfl_sync(gm, st.inst, partner, W); st.inst = v; }
    if (c == 0 && is_lo) st.dir = 1;
    if (c + 1 == R - 1 && is_hi) st.dir = -1;
    if (is_lo) st.saccept++;
    if (m.evolve_rate > 0) {
      // pry_temps, vector version with one pried gap (chain.cc:1809-1846) + resetTemp (chain.cc:1088-1091)
      const double rate = m.evolve_rate;
      const double b_next = __shfl_down_sync(gm, ch.beta, 1, W), lp_next = __shfl_down_sync(gm, ch.lpost, 1, W);
      double sp = ch.beta - b_next;
      if (m.evolve_lpost_cut >= 0 && ch.lpost - lp_next > m.evolve_lpost_cut * ch.beta) sp *= (1.0 + rate);
      if (is_lo) sp *= 1.0 + rate;
      double sum = 0;
      for (int k = 0; k < R - 1; k++) sum += wshfl(gm, sp, k, W);
      const double norm = sum / (1 - wshfl(gm, ch.beta, R - 1, W));
      const double q = sp / norm;
      double invtemp = 1, mine = ch.beta;
      for (int k = 1; k < R - 1; k++) {
        invtemp -= wshfl(gm, q, k - 1, W);
        if (w.rung == k) mine = invtemp;
      }
      if (w.rung >= 1 && w.rung < R - 1) { ch.beta = mine; ch.lpost = ch.lprior + mine * ch.llike; }
    }
  }
  if (is_lo) st.scount++;
}

// ------------------------------------------------------------------------------------------------- swap phase
template <int D, int MODE>
__device__ __forceinline__ void wswap_phase(const PtgModel &m, const PtgState &s, const WLane &w, Chain<D> &ch, WStats &st,
                                            Stream<MODE> &ls, uint64_t step, double swap_thresh, double ptry, int &napp) {
  const int W = w.W, R = w.R;
  const unsigned gm = w.gmask;
  const int maxswaps = m.maxswaps;
  if (m.swap_mode == PTG_SWAP_REFERENCE) {
    // candidate list (chain.cc:1410-1420).  PHILOX: lane j of the ladder fetched block j = (u_try, u_pair, u_swap) of trial j.
    uint32_t wt[4] = {0u, 0u, 0u, 0u};
    if constexpr (MODE == PTG_RNG_PHILOX) {
      ls.step = step;
      if (w.rung < maxswaps) ls.fetch((uint32_t)w.rung, wt);
    }
    unsigned used = 0;   // bit c set: an earlier surviving candidate is pair (c, c+1)
    int mycand = -2;
    for (int i = 0; i < maxswaps; i++) {
      double x, x2 = 0;
      if constexpr (MODE == PTG_RNG_PHILOX) {
        x = ptg_u32_to_unit(__shfl_sync(gm, wt[0], i, W));
        x2 = ptg_u32_to_unit(__shfl_sync(gm, wt[1], i, W));
      } else x = ls.next_u();
      int cand = -2;
      if (x < swap_thresh) {
        if constexpr (MODE == PTG_RNG_TAPE) x2 = ls.next_u();
        cand = (int)(x2 * (R - 1));
        // de-dup: iswaps[j]==cand or iswaps[j]+1==cand for an earlier surviving j (chain.cc:1417-1418)
        if (((used >> cand) & 1u) || (cand > 0 && ((used >> (cand - 1)) & 1u))) cand = -2;
        else used |= 1u << cand;
      }
      if (w.rung == i) mycand = cand;
    }
    for (int j = 0; j < maxswaps; j++) {
      const int c = __shfl_sync(gm, mycand, j, W);
      double usw = 0.5;
      if constexpr (MODE == PTG_RNG_PHILOX) {
        const uint32_t a = __shfl_sync(gm, wt[2], j, W), b = __shfl_sync(gm, wt[3], j, W);
        usw = ptg_u52_to_unit(a, b);
      }
      if (c < 0) continue;
      wswap_trial<D, MODE>(m, s, w, ch, st, ls, c, usw, napp);
    }
  } else {
    // even/odd performance mode: pairs (i,i+1), i = parity, parity+2, ...; each tried with probability min(1, 2 swap_rate)
    const int parity = (int)(step & 1);
    if constexpr (MODE == PTG_RNG_PHILOX) {
      // all pairs are disjoint: one shuffle round decides and exchanges every pair of the ladder at once
      ls.step = step;
      const bool is_lo = ((w.rung & 1) == parity) && (w.rung + 1 < R);
      const bool is_hi = (w.rung >= 1) && (((w.rung - 1) & 1) == parity) && (w.rung < R);
      double utry = 1.0, usw = 0.5;
      if (is_lo) {
        uint32_t q[4]; ls.fetch(PTG_BLK_SWAP_EVENODD + (uint32_t)w.rung, q);
        utry = ptg_u52_to_unit(q[0], q[1]); usw = ptg_u52_to_unit(q[2], q[3]);
      }
      const int partner = is_lo ? w.rung + 1 : (is_hi ? w.rung - 1 : w.rung);
      const double ll_p = wshfl(gm, ch.llike, partner, W), b_p = wshfl(gm, ch.beta, partner, W);
      int flags = 0; // bit0 tried, bit1 accepted (decided by the lower lane)
      if (is_lo && utry < ptry) {
        double lla = ch.llike; if (!(lla > -1e200)) lla = -1e200;
        double llb = ll_p; if (!(llb > -1e200)) llb = -1e200;
        const double lhr = -(b_p - ch.beta) * (llb - lla);
        bool accept = true;
        if (lhr < 0) accept = (log(usw) < lhr);
        flags = 1 | (accept ? 2 : 0);
        if (w.rung > 0) { if (st.dir > 0) st.ups++; if (st.dir < 0) st.downs++; }
        st.scount++;
        if (accept) st.saccept++;
      }
      const int pflags = __shfl_sync(gm, flags, partner, W);
      if (is_hi) flags = pflags;
      const bool tried = (flags & 1) != 0, accept = (flags & 2) != 0;
      // exchange: every lane shuffles, only accepted pairs keep the partner's values
      {
        double nx[D];
#pragma unroll
        for (int k = 0; k < D; k++) nx[k] = wshfl(gm, ch.x[k], partner, W);
        const double nll = wshfl(gm, ch.llike, partner, W), nlp = wshfl(gm, ch.lprior, partner, W);
        const int ndir = __shfl_sync(gm, st.dir, partner, W), ninst = __shfl_sync(gm, st.inst, partner, W);
        if (tried && accept) {
#pragma unroll
          for (int k = 0; k < D; k++) ch.x[k] = nx[k];
          ch.llike = nll; ch.lprior = nlp;
          ch.lpost = ch.lprior + ch.beta * ch.llike;
          st.dir = ndir; st.inst = ninst;
          if (w.rung == 0) st.dir = 1;
          if (w.rung == R - 1) st.dir = -1;
        }
      }
      if (tried && w.active) { chain_append<D>(m, s, ch, ch.x, ch.llike, ch.lpost, ch.beta); napp++; }
    } else {
      // TAPE: the oracle consumes u_try for every pair first, then u_swap per tried pair in order
      unsigned trymask = 0;
      for (int i = parity; i + 1 < R; i += 2) { double x = ls.next_u(); if (x < ptry) trymask |= 1u << i; }
      for (int i = parity; i + 1 < R; i += 2)
        if ((trymask >> i) & 1u) wswap_trial<D, MODE>(m, s, w, ch, st, ls, i, 0.5, napp);
    }
  }
}

// ------------------------------------------------------------------------------------------------- DE index
// differential_evolution::draw_i_from_chain (proposal_distribution.cc:744-778); w0 = the attempt-0 Philox word
template <int D, int MODE>
__device__ __forceinline__ int wde_index(const PtgModel &m, const PtgState &s, const Chain<D> &ch, const PtgProp &p, Stream<MODE> &rs,
                                         uint32_t w0, int which, int hsize, int &attempt) {
  int start = 0;
  const int mins = D * 10, minc = D * 100;
  if ((hsize - minc) * (1 - p.ignore_frac) > mins) start = (int)((hsize - minc) * p.ignore_frac);
  double alpha = p.unlikely_alpha;
  const double lpost0 = ch.map_lpost - D;
  while (true) {
    const int a = attempt;
    uint32_t wr[4] = {0u, 0u, 0u, 0u};
    double xrnd;
    if constexpr (MODE == PTG_RNG_PHILOX) {
      if (a > 0 || alpha > 0) rs.fetch(PTG_BLK_RETRY + which * 0x100 + (a & 0xff), wr);
      xrnd = (a == 0) ? ptg_u32_to_unit(w0) : ptg_u32_to_unit(wr[0]);
    } else xrnd = rs.next_u();
    attempt++;
    const int index = (int)(start + (hsize - start) * xrnd);
    if (alpha > 0) {
      const double lpost = hist_lpost<D>(m, s, ch, index);
      if (lpost0 > lpost) {
        const double pr = exp(alpha * (lpost - lpost0));
        double x2;
        if constexpr (MODE == PTG_RNG_PHILOX) x2 = ptg_u32_to_unit(wr[1]); else x2 = rs.next_u();
        if (x2 < pr) return index;
        alpha *= 0.9;
        continue;
      }
    }
    return index;
  }
}


// ------------------------------------------------------------------------------------------------- temperature mixing
// differential_evolution::draw_from_chain with support_mixing (proposal_distribution.cc:594-741): a bare DE proposal on a ladder weighs
// the rungs by Nmean = 10 log-likelihoods of the CALLER's history (at indices drawn against rung i's size) and the median of Nmedian = 10
// of rung i's own, picks a rung and draws the state from its history.  Sizes are the frozen sizes (history_freeze, chain.cc:1552): the
// ladder's lanes published (window size, ring offset, beta) of their rung to `mx` before the MH phase.  unlikely_alpha = 0 (checked on
// the host), so draw_i_from_chain (:744-778) is one uniform: the k-th of history draw `which` has the Philox address
// PTG_BLK_MIX + which * 0x1000 + k / 4, word k % 4; tape runs consume the chain's tape in the same order.
struct WMix { int size[32], off[32]; double beta[32]; };   // per warp
template <int D, int MODE>
__device__ __noinline__ const double *wde_state_mixed(const PtgModel &m, const PtgState &s, const WLane &w, const Chain<D> &ch, const PtgProp &p,
                                                     Stream<MODE> &rs, int which, int &kcount, const WMix &mx) {
  const int R = w.R, g0 = w.lane - w.rung;          // lane of rung 0 of this ladder
  const long long chain0 = ch.chain - w.rung;
  uint32_t q0 = 0, q1 = 0, q2 = 0, q3 = 0; int blk = -1;
  auto nextu = [&]() -> double {
    const int kk = kcount++;
    if constexpr (MODE == PTG_RNG_PHILOX) {
      if ((kk >> 2) != blk) { uint32_t q[4]; blk = kk >> 2; rs.fetch(PTG_BLK_MIX + (uint32_t)which * 0x1000u + (uint32_t)blk, q); q0 = q[0]; q1 = q[1]; q2 = q[2]; q3 = q[3]; }
      const int wd = kk & 3;
      return ptg_u32_to_unit(wd == 0 ? q0 : (wd == 1 ? q1 : (wd == 2 ? q2 : q3)));
    } else return rs.next_u();
  };
  auto draw_i = [&](int j) -> int {
    const int size = mx.size[g0 + j];
    int start = 0;
    const int mins = D * 10, minc = D * 100;
    if ((size - minc) * (1 - p.ignore_frac) > mins) start = (int)((size - minc) * p.ignore_frac);
    return (int)(start + (size - start) * nextu());
  };
  auto llike_at = [&](int j, int index, double current) -> double { // MH_chain::getLogLike(index, true), chain.cc:1078-1085
    if (index < 0 || index >= mx.size[g0 + j]) return current;
    int pp = mx.off[g0 + j] + index; if (pp >= m.hist_cap) pp -= m.hist_cap;
    return s.hist_lp[2 * ((chain0 + j) * m.hist_cap + pp) + 1];
  };
  constexpr int Nmean = 10, Nmedian = 10;
  const double pmix = m.de_Tmix, beta = ch.beta;
  double ksum = 0, kthis_lo = 0; // running k[i]; the pick needs k[] again, so the weights are kept
  double kw[32];
  int ithis = 0, guard = 0;
  for (int i = 0; i < R; i++) {
    double l0[Nmean], l[Nmedian];
    double l0max = -1e100, l0min = 1e100;
    for (int j = 0; j < Nmean; j++) {
      const int index = draw_i(i);
      const double dl = llike_at(w.rung, index, ch.llike);
      if (isfinite(dl)) {
        if (dl > l0max) l0max = dl;
        if (dl < l0min) l0min = dl;
        l0[j] = dl;
      } else { j--; if (++guard > 100000) { rs.err = 3; break; } }
    }
    const double alpha = mx.beta[g0 + i];
    double amb = alpha - beta;
    amb = -amb;
    if (amb == 0) ithis = i;
    double sum = 0;
    const double l0scale = amb < 0 ? l0min : l0max;
    for (int ii = 0; ii < Nmean; ii++) sum += exp((l0[ii] - l0scale) * amb);
    const double ll0 = log(sum / Nmean) + l0scale * amb;
    for (int j = 0; j < Nmedian; j++) { const int index = draw_i(i); l[j] = llike_at(i, index, 0.0); }
    for (int a = 1; a < Nmedian; a++) { // sort(l.begin(), l.end())
      const double v = l[a]; int b = a - 1;
      while (b >= 0 && l[b] > v) { l[b + 1] = l[b]; b--; }
      l[b + 1] = v;
    }
    const double ll = l[Nmedian / 2];
    double lk = ll0 - ll * amb;
    lk = -lk;
    lk /= pmix;
    if (lk > 0) lk = 0;
    ksum = ksum + exp(lk);
    kw[i] = ksum;
  }
  (void)kthis_lo;
  int ipick = ithis;
  const double xrnd = nextu() * ksum;
  for (int i = 0; i < R; i++) if (xrnd <= kw[i]) { ipick = i; break; }
  const int index = draw_i(ipick);
  int pp = mx.off[g0 + ipick] + index; if (pp >= m.hist_cap) pp -= m.hist_cap;
  return s.hist + ((chain0 + ipick) * m.hist_cap + pp) * PTG_HX(D);
}

// ------------------------------------------------------------------------------------------------- cooperative normals
// PHILOX: standard normals z[0..D) for every lane with `need`; each (owner lane, Box-Muller pair) is one work item
// evaluated by some lane of the warp from the OWNER's stream address, so the values equal per-lane draw_normals().
// Must be called by all 32 lanes.
template <int D>
__device__ __forceinline__ void wcoop_normals(uint64_t seed, uint64_t step, uint64_t stream_base, int W, bool need, double z[D]) {
  constexpr int NP = (D + 1) / 2;
  const int lane = threadIdx.x & 31;
  const unsigned mask = __ballot_sync(0xffffffffu, need);
  const int total = __popc(mask) * NP;
  const int myrank = __popc(mask & ((1u << lane) - 1u));
#pragma unroll
  for (int i = 0; i < D; i++) z[i] = 0;
  for (int base = 0; base < total; base += 32) {
    const int t = base + lane;
    double z0 = 0, z1 = 0;
    if (t < total) {
      const int orank = t / NP, pr = t - orank * NP;
      const int owner = __fns(mask, 0, orank + 1);
      // stream of the owner's chain: (global ladder)*128 + rung, with ladder = warp's first ladder + owner / W
      const uint64_t stream = stream_base + (uint64_t)(owner / W) * PTG_STREAM_STRIDE + (uint64_t)(owner % W);
      uint32_t q[4];
      ptg_philox_draw(seed, stream, PTG_DOMAIN_STEP, step, PTG_BLK_NORMAL + pr, q);
      box_muller(q, z0, z1);
    }
#pragma unroll
    for (int pr = 0; pr < NP; pr++) {
      const int t_mine = myrank * NP + pr;
      const int src = t_mine & 31;
      const double v0 = __shfl_sync(0xffffffffu, z0, src), v1 = __shfl_sync(0xffffffffu, z1, src);
      if (need && t_mine >= base && t_mine < base + 32) {
        z[2 * pr] = v0;
        if (2 * pr + 1 < D) z[2 * pr + 1] = v1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------- MH step
// MH_chain::step(prop) (chain.cc:966-1022) with proposal_distribution_set::draw (proposal_distribution.cc:99-129).
// Called by ALL lanes of the warp; lanes with do_mh == false take part in the cooperative work only.
template <int D, int MODE>
__device__ __forceinline__ MhOut wmh_step(const PtgModel &m, const PtgState &s, const WLane &w, Chain<D> &ch, Stream<MODE> &rs,
                                          const double *__restrict__ bins, bool do_mh, uint64_t stream_base, const WMix &mx) {
  double newx[D];
#pragma unroll
  for (int i = 0; i < D; i++) newx[i] = ch.x[i];
  double prop_lh = 0;
  int type = 0;
  bool valid = m.zero_valid != 0;
  const double oldlprior = ch.lpost - ch.beta * ch.llike;
  const int hsize = (int)(ch.nsize > m.hist_cap ? (long long)m.hist_cap : ch.nsize);

  uint32_t wA[4] = {0u, 0u, 0u, 0u}, wB[4] = {0u, 0u, 0u, 0u};
  if constexpr (MODE == PTG_RNG_PHILOX) { rs.fetch(PTG_BLK_A, wA); rs.fetch(PTG_BLK_B, wB); }

  // ---- member selection: first ready member with u < bin_max (proposal_distribution.cc:105-112); top-level SLOTS map to members, one
  // slot may hold a nested set (ptg_set_nested_set) that repeats the selection over its own members with a second uniform
  int member = 0, slot = 0, nested = -1;
  const bool adaptive = (m.adapt_rate != 0 || m.nest_adapt != 0) && m.wrap_in_set != 0;   // this chain's own, adapting bins
  if (adaptive && do_mh) bins = s.ad_bins + ch.chain * m.n_bins;
  const bool mixing = m.de_mixing != 0 && w.R > 1;
  if (m.wrap_in_set && do_mh) {
    member = -1;
    for (int count = 0; count <= 100 && member < 0; count++) {
      double x = 0.0;
      if (m.n_slots > 1) { if constexpr (MODE == PTG_RNG_PHILOX) x = ptg_u32_to_unit(wA[0]); else x = rs.next_u(); }
      slot = -1;
      for (int sl = 0; sl < m.n_slots; sl++) {
        const int mem = (m.nest_count == 0 || sl < m.nest_first) ? sl : (sl == m.nest_first ? -1 : sl + m.nest_count - 1);
        const bool ready = mem < 0 || (m.props[mem].kind != PTG_PROP_DE) || hsize >= D * 10; // differential_evolution::is_ready; a set is always ready
        if (slot < 0 && ready && x < bins[sl]) { slot = sl; member = mem; }
      }
      if (slot >= 0 && member < 0) { // the nested set's own draw
        const double *nb = bins + m.n_slots;
        for (int ncount = 0; ncount <= 100 && member < 0; ncount++) {
          double x2 = 0.0;
          if (m.nest_count > 1) {
            if constexpr (MODE == PTG_RNG_PHILOX) { uint32_t q[4]; rs.fetch(PTG_BLK_NEST, q); x2 = ptg_u32_to_unit(q[0]); } else x2 = rs.next_u();
          }
          for (int j = 0; j < m.nest_count; j++) {
            const bool ready = (m.props[m.nest_first + j].kind != PTG_PROP_DE) || hsize >= D * 10;
            if (member < 0 && ready && x2 < nb[j]) { member = m.nest_first + j; nested = j; }
          }
          if constexpr (MODE == PTG_RNG_PHILOX) break;
        }
        break; // (a nested set whose members are all unready ends the run in the reference, proposal_distribution.cc:121-127)
      }
      if constexpr (MODE == PTG_RNG_PHILOX) break;
    }
    if (member < 0) { rs.err = 2; member = 0; slot = 0; nested = -1; }
  }
  const PtgProp &p = m.props[member];
  const int kind = do_mh ? p.kind : 0;

  // ---- normals for Gaussian-proposal lanes
  double z[D];
  if constexpr (MODE == PTG_RNG_PHILOX) wcoop_normals<D>(m.seed, rs.step, stream_base, w.W, kind == PTG_PROP_GAUSS, z);

  if (kind == PTG_PROP_DE) {
    // differential_evolution::draw (proposal_distribution.cc:790-801)
    double usnk, ug;
    if constexpr (MODE == PTG_RNG_PHILOX) { usnk = ptg_u32_to_unit(wA[1]); ug = ptg_u32_to_unit(wA[2]); }
    else { usnk = rs.next_u(); ug = rs.next_u(); }
    if (!(p.snooker > usnk)) {
      // draw_standard (proposal_distribution.cc:489-535)
      double gamma = p.gamma_std;
      if (ug < p.g1frac) gamma = 1;
      int a1 = 0, a2 = 0;
      const double *s1, *s2;
      if (mixing) { s1 = wde_state_mixed<D, MODE>(m, s, w, ch, p, rs, 1, a1, mx); s2 = wde_state_mixed<D, MODE>(m, s, w, ch, p, rs, 2, a2, mx); }
      else {
        const int i1 = wde_index<D, MODE>(m, s, ch, p, rs, wA[3], 1, hsize, a1);
        const int i2 = wde_index<D, MODE>(m, s, ch, p, rs, wB[0], 2, hsize, a2);
        s1 = hist_elem<D>(m, s, ch, i1); s2 = hist_elem<D>(m, s, ch, i2);
      }
      if constexpr (MODE == PTG_RNG_TAPE) { // the d normals of the discarded jitter are still consumed (H8-1)
        for (int j = 0; j < D; j++) (void)rs.next_z();
      }
      double a[D], b[D];
#pragma unroll
      for (int i = 0; i < D; i++) { a[i] = s1[i]; b[i] = s2[i]; }
#pragma unroll
      for (int i = 0; i < D; i++) {
        const double t = ch.x[i] + a[i] * gamma;
        newx[i] = t + b[i] * (-gamma);
      }
      type = 0;
    } else {
      // draw_snooker (proposal_distribution.cc:538-591)
      const double gamma = (1.2 + ug) / p.reduce_gamma;
      double smznorm2 = 0, minusz[D], smz[D];
      int az = 0, isafe = 0;
      while (smznorm2 == 0) {
        const double *zz;
        if (mixing) zz = wde_state_mixed<D, MODE>(m, s, w, ch, p, rs, 0, az, mx);
        else { const int iz = wde_index<D, MODE>(m, s, ch, p, rs, wB[1], 0, hsize, az); zz = hist_elem<D>(m, s, ch, iz); }
        smznorm2 = 0;
#pragma unroll
        for (int i = 0; i < D; i++) { minusz[i] = zz[i] * (-1); smz[i] = ch.x[i] + minusz[i]; }
#pragma unroll
        for (int i = 0; i < D; i++) smznorm2 += smz[i] * smz[i];
        if (++isafe > 1000) break;
      }
      int a1 = 0, a2 = 0;
      const double *s1, *s2;
      if (mixing) { s1 = wde_state_mixed<D, MODE>(m, s, w, ch, p, rs, 1, a1, mx); s2 = wde_state_mixed<D, MODE>(m, s, w, ch, p, rs, 2, a2, mx); }
      else {
        const int i1 = wde_index<D, MODE>(m, s, ch, p, rs, wA[3], 1, hsize, a1);
        const int i2 = wde_index<D, MODE>(m, s, ch, p, rs, wB[0], 2, hsize, a2);
        s1 = hist_elem<D>(m, s, ch, i1); s2 = hist_elem<D>(m, s, ch, i2);
      }
      double dot = 0;
#pragma unroll
      for (int i = 0; i < D; i++) {
        const double ds12 = s1[i] * gamma + s2[i] * (-gamma);
        dot += ds12 * smz[i];
      }
      const double fac = dot / smznorm2;
      double pmz2 = 0;
#pragma unroll
      for (int i = 0; i < D; i++) {
        newx[i] = ch.x[i] + smz[i] * fac;
        const double pmz = newx[i] + minusz[i];
        pmz2 += pmz * pmz;
      }
      prop_lh = (log(pmz2) - log(smznorm2)) * (D - 1) / 2.0;
      type = 1;
    }
  } else if (kind == PTG_PROP_GAUSS) {
    // gaussian_prop::draw (proposal_distribution.hh:194-218)
    double off[D];
    if constexpr (MODE == PTG_RNG_TAPE) {
#pragma unroll
      for (int j = 0; j < D; j++) z[j] = rs.next_z();
    }
    const double *__restrict__ sig = m.prop_data + p.sigma_off;
#pragma unroll
    for (int i = 0; i < D; i++) off[i] = z[i] * __ldg(sig + i) + 0.0;
    double x1 = 1;
    if (p.one_d_frac > 0) { if constexpr (MODE == PTG_RNG_PHILOX) x1 = ptg_u32_to_unit(wA[1]); else x1 = rs.next_u(); }
    if (p.one_d_frac > 0 && x1 < p.one_d_frac) {
      double ua;
      if constexpr (MODE == PTG_RNG_PHILOX) ua = ptg_u32_to_unit(wA[2]); else ua = rs.next_u();
      const int ia = (int)(D * ua);
#pragma unroll
      for (int j = 0; j < D; j++) if (j != ia) off[j] = 0.0;
      type = 1;
    } else type = 0;
    if (p.has_transform) {
      const double *__restrict__ M = m.prop_data + p.trans_off;
      double t[D];
      // summation order of Eigen 3.3.7's column-major GEMV (`vec=diagTransform*vec`, proposal_distribution.hh:212)
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
  } else if (kind == PTG_PROP_PRIOR_DRAW) {
    // draw_from_dist::draw (proposal_distribution.hh:124-129)
    valid = prior_draw<D, MODE>(m, rs, PTG_BLK_PRIOR, newx);
    prop_lh = prior_eval_log<D>(m, ch.x, true) - prior_eval_log<D>(m, newx, valid);
    type = 0;
  }
  if (m.wrap_in_set) {
    if (nested >= 0) type = nested + 10 * type;   // the nested set's type() first (proposal_distribution.cc:113)
    type = slot + 10 * type;
  }

  // ---- enforce, prior, gated likelihood (chain.cc:976-987)
  if (valid) valid = space_enforce<D>(m, newx);
  const double newlprior = prior_eval_log<D>(m, newx, valid);
  double newlike = -CUDART_INF, newlpost = -CUDART_INF;
  int code = 0;
  bool accept = true;
  const bool gate = valid && ((newlprior > -1e200) || (newlprior - oldlprior > m.dprior_min));
  if (gate && do_mh) {
    newlike = like_eval<D>(m, newx);
    newlpost = newlike * ch.beta + newlprior;
  } else code |= PTG_TRACE_NOLIKE;
  // ---- Metropolis test (chain.cc:989-1001)
  double lhr = prop_lh;
  if (isnan(lhr)) accept = false;
  lhr += newlpost - ch.lpost;
  if (!valid) { accept = false; code |= PTG_TRACE_INVALID; }
  if (accept && lhr < 0 && do_mh) {
    double u;
    if constexpr (MODE == PTG_RNG_PHILOX) u = ptg_u52_to_unit(wB[2], wB[3]); else u = rs.next_u();
    accept = (log(u) < lhr);
  }
  if (do_mh) {
    ch.ntries++;
    if (adaptive) set_adapt(m, s, ch.chain, slot, nested, accept, ch.beta);
    if (accept) {
      ch.naccept++;
      ch.last_type = type;
#pragma unroll
      for (int i = 0; i < D; i++) ch.x[i] = newx[i];
      ch.llike = newlike; ch.lpost = newlpost; ch.lprior = newlprior;
      code |= PTG_TRACE_ACCEPT;
    }
    chain_append<D>(m, s, ch, ch.x, ch.llike, ch.lpost, ch.beta);
  }
  code |= (type & PTG_TRACE_TYPE_MASK);
  MhOut o; o.lhr = lhr; o.code = code;
  return o;
}

// ------------------------------------------------------------------------------------------------- kernel
// blockDim.x = 128 (4 warps); dynamic shared memory = n_rungs * n_bins doubles (the proposal bins, read-only)
#ifndef PTG_WSTEP_MINB
#define PTG_WSTEP_MINB 1
#endif
template <int D, int MODE>
__global__ void __launch_bounds__(128, PTG_WSTEP_MINB) ptg_wstep_kernel(const __grid_constant__ PtgModel m, PtgState s, long long step0, int n_steps, int W) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sbins = reinterpret_cast<double *>(smem_raw); // [R][n_bins]
  __shared__ WMix smix[4];
  const int R = m.n_rungs;
  for (int i = threadIdx.x; i < R * m.n_bins; i += blockDim.x) sbins[i] = m.bins[i];
  __syncthreads();
  WMix &mx = smix[threadIdx.x >> 5];

  WLane w;
  w.W = W; w.R = R; w.lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int gpw = 32 / W, g = w.lane / W;
  w.rung = w.lane - g * W;
  w.ladder = warp * gpw + g;
  w.ladder_ok = w.ladder < m.n_ladders;
  w.active = w.ladder_ok && w.rung < R;
  w.gmask = (W == 32) ? 0xffffffffu : (((1u << W) - 1u) << (g * W));
  if (warp * gpw >= m.n_ladders) return; // whole warp beyond the batch
  const long long chain = w.ladder * R + w.rung;
  const uint64_t stream_base = (uint64_t)(m.ladder_offset + warp * gpw) * PTG_STREAM_STRIDE; // stream of (first ladder of the warp, rung 0)

  Chain<D> ch;
  WStats st;
  Stream<MODE> rs, ls;
  if (w.active) {
    chain_load<D>(m, s, chain, ch);
    stream_open<MODE>(m, s, rs, chain, (uint64_t)(m.ladder_offset + w.ladder) * PTG_STREAM_STRIDE + (uint64_t)w.rung, PTG_DOMAIN_STEP);
    st.dir = s.directions[chain]; st.ups = s.ups[chain]; st.downs = s.downs[chain]; st.inst = s.instances[chain];
  } else {
    // ghost lane (rung >= n_rungs or ladder beyond the batch): takes part in shuffles, never touches memory
#pragma unroll
    for (int k = 0; k < D; k++) ch.x[k] = 0;
    ch.lpost = ch.llike = ch.lprior = 0; ch.beta = 1; ch.map_lpost = 0;
    ch.nhist = ch.nsize = 0; ch.ntries = ch.naccept = 1; ch.last_type = -1; ch.slot = 0; ch.since_save = 0; ch.chain = 0;
    stream_blank<MODE>(m, rs);
    st.dir = st.ups = st.downs = st.inst = 0;
  }
  st.scount = st.saccept = 0;
  if (w.ladder_ok)
    stream_open<MODE>(m, s, ls, m.n_chains + w.ladder, (uint64_t)(m.ladder_offset + w.ladder) * PTG_STREAM_STRIDE + PTG_STREAM_LADDER, PTG_DOMAIN_STEP);
  else
    stream_blank<MODE>(m, ls);
  const double swap_thresh = (R - 1) * m.swap_rate / m.maxswaps; // chain.cc:1413
  double ptry = 2 * m.swap_rate; if (ptry > 1) ptry = 1;
  const double *bins = sbins + (w.rung < R ? w.rung : 0) * m.n_bins;

  for (int it = 0; it < n_steps; it++) {
    const uint64_t step = (uint64_t)(step0 + it);
    if constexpr (MODE == PTG_RNG_TAPE) {
      // re-synchronise the cursors with the recorded run at every step boundary (see ptg_kernels.cuh)
      if (s.u_mark && (long long)step < s.n_mark_steps) {
        const long long ns = m.n_chains + m.n_ladders, row = (long long)step * ns;
        if (w.active) { rs.upos = s.u_mark[row + chain]; rs.zpos = s.z_mark[row + chain]; }
        if (w.ladder_ok) { ls.upos = s.u_mark[row + m.n_chains + w.ladder]; ls.zpos = s.z_mark[row + m.n_chains + w.ladder]; }
      }
    }
    int napp = 0;
    if (w.ladder_ok && R > 1) wswap_phase<D, MODE>(m, s, w, ch, st, ls, step, swap_thresh, ptry, napp);
    __syncwarp();
    rs.step = step;
    const bool do_mh = w.active && napp == 0;
    if (m.de_mixing) { // history_freeze (chain.cc:1552): what the other rungs of the ladder see of this chain during the MH phase
      const bool wrapped = ch.nsize > m.hist_cap;
      mx.size[w.lane] = (int)(wrapped ? (long long)m.hist_cap : ch.nsize); mx.off[w.lane] = wrapped ? ch.slot : 0; mx.beta[w.lane] = ch.beta;
      __syncwarp();
    }
    MhOut o = wmh_step<D, MODE>(m, s, w, ch, rs, bins, do_mh, stream_base, mx);
    if (m.de_mixing) __syncwarp();
    if (w.active && (long long)step < m.trace_steps) {
      s.trace_lhr[step * m.n_chains + chain] = do_mh ? o.lhr : 0.0;
      s.trace_code[step * m.n_chains + chain] = do_mh ? o.code : PTG_TRACE_SWAPPED;
    }
  }
  if (w.active) {
    chain_store<D>(m, s, ch);
    stream_close<MODE>(s, rs, chain);
    s.directions[chain] = st.dir; s.ups[chain] = st.ups; s.downs[chain] = st.downs; s.instances[chain] = st.inst;
    s.swap_count[chain] += st.scount; s.swap_accept[chain] += st.saccept;
  }
  if (w.ladder_ok && w.rung == 0) stream_close<MODE>(s, ls, m.n_chains + w.ladder);
}
