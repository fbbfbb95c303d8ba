// ptg_fast.cuh -- the production step kernel (Philox draws): ladder-in-a-warp like ptg_warp.cuh, restructured around the
// ncu findings of round 1 (profiles/README.md):
//   * one wave: chain state is slimmed to <= 72 registers/thread (32-bit launch-local counters, packed ladder statistics)
//     so that all 4096 warps of BASELINE config C1 are resident at once (7 CTAs x 4 warps per SM);
//   * swap phase: every lane prepares ITS trial in parallel (candidate pair from its Philox block, log of its
//     swap draw), the serial part is a 7-iteration bit-mask de-dup plus one shuffle round per surviving trial;
//   * MH update: member selection, both Philox blocks, all three DE history indices, gathers, prior, likelihood and the
//     Metropolis test are straight-line code executed by the whole warp; only the snooker override (~8 % of lanes) and
//     rare paths (unlikely_alpha, prior-draw member, bounded spaces, non-uniform priors) branch;
//   * Box-Muller normals for the Gaussian-proposal lanes are produced cooperatively by all 32 lanes;
//   * proposal-member parameters live in shared memory (divergent member indices would serialise constant-bank loads).
// The arithmetic is expression-for-expression that of ptg_warp.cuh / ptg_kernels.cuh / the oracle: in Philox mode the
// three kernels produce bit-identical chains (tests/test_gpu_properties.py::test_kernels_agree_bitwise).
#pragma once
#include "ptg_warp.cuh"

template <int D>
struct FChain {
  double x[D];
  double lpost, llike, lprior, beta;
  int slot, hfill, since_save;           // ring write position, min(nsize, capacity), nhist % save_every
};
// Launch-local counters that are touched once per step live in shared memory, one word per thread and counter
// ([counter][thread]: conflict-free), to keep the register-resident state within the one-wave budget.
// largest CTA the production kernel is launched with (28 warps = one CTA per SM at 72 registers); smaller batches use smaller CTAs
#define PTG_FSTEP_MAX_THREADS 896
#define PTG_FC_STRIDE 7 // per-thread counter row in shared memory, odd stride: conflict-free
enum { FC_NHIST = 0, FC_NTRIES, FC_NACCEPT, FC_LAST_TYPE, FC_UD, FC_SC, FC_COUNT };

// proposal member parameters staged in shared memory
struct FProp {
  double snooker, g1frac, gamma_std, reduce_gamma, ignore_frac, unlikely_alpha, one_d_frac;
  int kind, has_transform, sigma_off, trans_off;
};

// MH_chain::add_state (chain.cc:916-949) of the chain's current state
template <int D>
__device__ __forceinline__ void fappend(const PtgModel &m, const PtgState &s, FChain<D> &ch, long long chain, double *__restrict__ hbase, int *cnt) {
  if (ch.lpost > s.map_lpost[chain]) { // MAP update (chain.cc:931-934); the running maximum stays in global memory (L1/L2-resident)
    s.map_lpost[chain] = ch.lpost;
#pragma unroll
    for (int k = 0; k < D; k++) s.map_x[(long long)k * m.n_chains + chain] = ch.x[k];
  }
  if (ch.since_save == 0) {
    double *h = hbase + ch.slot * (D + 2);
#pragma unroll
    for (int k = 0; k < D; k++) h[k] = ch.x[k];
    h[D] = ch.lpost; h[D + 1] = ch.llike;
    if (m.record_full) {
      const long long rec = chain * m.hist_cap + ch.slot;
      s.hist_acc[rec] = (s.naccept[chain] + cnt[FC_NACCEPT]) / (double)(s.ntries[chain] + cnt[FC_NTRIES]);
      s.hist_beta[rec] = ch.beta;
      s.hist_type[rec] = cnt[FC_LAST_TYPE];
    }
    if (ch.hfill < m.hist_cap) ch.hfill++;
    ch.slot = (ch.slot + 1 == m.hist_cap) ? 0 : ch.slot + 1;
  }
  ch.since_save = (ch.since_save + 1 == m.save_every) ? 0 : ch.since_save + 1;
  cnt[FC_NHIST]++;
}

// element `index` of the eligible window (newest min(nsize,cap) samples): once the ring is full the oldest sits at `slot`
template <int D>
__device__ __forceinline__ const double *fhist(const PtgModel &m, const FChain<D> &ch, const double *__restrict__ hbase, int index) {
  int p = index;
  if (ch.hfill == m.hist_cap) { p = ch.slot + index; if (p >= m.hist_cap) p -= m.hist_cap; }
  return hbase + p * (D + 2);
}

// differential_evolution::draw_i_from_chain with unlikely_alpha > 0 or a retry (proposal_distribution.cc:744-778): rare path
template <int D>
__device__ __noinline__ int2 fde_index_slow(const PtgModel &m, int hfill, int slot, const double *hbase, double map_lpost, uint64_t seed, uint64_t stream,
                                             uint64_t step, double ignore_frac, double alpha, uint32_t w0, int which, int attempt) {
  const int hsize = hfill; // returns (index, attempts consumed so far)
  int start = 0;
  const int mins = D * 10, minc = D * 100;
  if ((hsize - minc) * (1 - ignore_frac) > mins) start = (int)((hsize - minc) * ignore_frac);
  const double lpost0 = map_lpost - D;
  while (true) {
    const int a = attempt;
    uint32_t wr[4] = {0u, 0u, 0u, 0u};
    if (a > 0 || alpha > 0) ptg_philox_draw(seed, stream, PTG_DOMAIN_STEP, step, PTG_BLK_RETRY + which * 0x100 + (a & 0xff), wr);
    const double xrnd = (a == 0) ? ptg_u32_to_unit(w0) : ptg_u32_to_unit(wr[0]);
    attempt++;
    const int index = (int)(start + (hsize - start) * xrnd);
    if (alpha > 0) {
      int p = index;
      if (hfill == m.hist_cap) { p = slot + index; if (p >= m.hist_cap) p -= m.hist_cap; }
      const double lpost = hbase[p * (D + 2) + D];
      if (lpost0 > lpost) {
        const double pr = exp(alpha * (lpost - lpost0));
        if (ptg_u32_to_unit(wr[1]) < pr) return make_int2(index, attempt);
        alpha *= 0.9;
        continue;
      }
    }
    return make_int2(index, attempt);
  }
}

// Rare paths are out-of-line (they would only bloat the hot loop's instruction footprint) and exchange vectors BY VALUE,
// so that taking no address keeps the caller's proposal vector in registers.
template <int D> struct FVec { double v[D]; };
template <int D> struct FVecFlag { double v[D]; double aux; int ok; };

// the prior-draw member (draw_from_dist::draw, proposal_distribution.hh:124-129)
template <int D>
__device__ __noinline__ FVecFlag<D> fprior_member(const PtgModel &m, uint64_t stream, uint64_t step, FVec<D> x) {
  Stream<PTG_RNG_PHILOX> rs;
  stream_blank<PTG_RNG_PHILOX>(m, rs);
  rs.id = stream; rs.step = step;
  double nx[D];
  FVecFlag<D> r;
  const bool valid = prior_draw<D, PTG_RNG_PHILOX>(m, rs, PTG_BLK_PRIOR, nx);
  r.aux = prior_eval_log<D>(m, x.v, true) - prior_eval_log<D>(m, nx, valid);
#pragma unroll
  for (int i = 0; i < D; i++) r.v[i] = nx[i];
  r.ok = valid ? 1 : 0;
  return r;
}

// state::enforce on a bounded space (states.cc:11-58)
template <int D>
__device__ __noinline__ FVecFlag<D> fenforce(const PtgModel &m, FVec<D> x) {
  FVecFlag<D> r;
  r.ok = space_enforce<D>(m, x.v) ? 1 : 0;
#pragma unroll
  for (int i = 0; i < D; i++) r.v[i] = x.v[i];
  r.aux = 0;
  return r;
}

template <int D>
__device__ __noinline__ double fprior_general(const PtgModel &m, FVec<D> x, bool valid) {
  return prior_eval_log<D>(m, x.v, valid);
}

// eigen-rotation of a Gaussian-proposal offset (see ptg_warp.cuh for the summation order)
template <int D>
__device__ __noinline__ FVec<D> ftransform(const double *__restrict__ M, FVec<D> o) {
  FVec<D> t;
  constexpr int CB = (D / 4) * 4, EVEN_ROWS = D & ~1;
#pragma unroll
  for (int i = 0; i < D; i++) {
    const double *__restrict__ a = M + i * D;
    double acc = 0;
#pragma unroll
    for (int j = 0; j < CB; j += 4) {
      if (i < EVEN_ROWS) acc = acc + ((__ldg(a + j) * o.v[j] + __ldg(a + j + 1) * o.v[j + 1]) + (__ldg(a + j + 2) * o.v[j + 2] + __ldg(a + j + 3) * o.v[j + 3]));
      else { acc = __ldg(a + j) * o.v[j] + acc; acc = __ldg(a + j + 1) * o.v[j + 1] + acc; acc = __ldg(a + j + 2) * o.v[j + 2] + acc; acc = __ldg(a + j + 3) * o.v[j + 3] + acc; }
    }
#pragma unroll
    for (int j = CB; j < D; j++) acc += __ldg(a + j) * o.v[j];
    t.v[i] = acc;
  }
  return t;
}


// packed ladder statistics: st_di = dir (low 2 bits, biased by 1) | inst << 2 ; st_ud = ups delta | downs delta << 16 ;
// st_sc = swap_count delta | swap_accept delta << 16   (deltas per launch; the host keeps launches <= 32767 steps)
#define PTG_FAST_MAX_STEPS 16384

// ---- rung-sharded ladders, exchange over peer memory (PtgXchg, ptg_types.h) ---------------------------------------------------
// publish p: the lanes that hold this block's edge rungs write (x, llike, lprior, beta) into parity p & 1 of this rank's area, then
// raise the ladder's flag to p + 1
template <int D>
__device__ __forceinline__ void fx_publish(const PtgModel &m, const PtgXchg &xc, long long p, const FChain<D> &ch, long long ladder, int rung, int R) {
  if (rung != 0 && rung != R - 1) return;
#pragma unroll 1
  for (int e = 0; e < 2; e++) {
    if (rung != (e == 0 ? 0 : R - 1)) continue; // R == 1: the one rung is both edges
    double *rec = xc.my_edges + (((size_t)(p & 1) * 2 + e) * m.n_ladders + ladder) * (D + 3);
#pragma unroll
    for (int k = 0; k < D; k++) rec[k] = ch.x[k];
    rec[D] = ch.llike; rec[D + 1] = ch.lprior; rec[D + 2] = ch.beta;
    __threadfence_system();
    *((volatile int *)(xc.my_flags + (size_t)e * m.n_ladders + ladder)) = (int)(p + 1);
  }
}
// boundary trial of exchange p (ptg_boundary_swap_kernel: chain.cc:1459-1490 + add_state) against the neighbour's record, read from
// ITS memory over NVLink once its per-ladder flag says publish p is complete
template <int D>
__device__ __forceinline__ void fx_swap(const PtgModel &m, const PtgState &s, const PtgXchg &xc, long long p, FChain<D> &ch, int chain, long long ladder, int gl,
                                     int rung, int R, int *cnt, int &err) {
  const bool edge_lo = (rung == 0) && xc.has_lo, edge_hi = (rung == R - 1) && xc.has_hi;
  if (!edge_lo && !edge_hi) return;
  double *__restrict__ hb = s.hist + chain * ((long long)m.hist_cap * (D + 2));
#pragma unroll 1
  for (int e = 0; e < 2; e++) {
    if (!(e == 0 ? edge_lo : edge_hi)) continue;
    // e = 0: my coldest rung against the colder neighbour's hottest (I am the pair's upper rung); e = 1: the mirror image
    const int their_edge = (e == 0) ? 1 : 0;
    const volatile int *flag = (e == 0 ? xc.lo_flags : xc.hi_flags) + (size_t)their_edge * m.n_ladders + ladder;
    long long spins = 0;
    while (*flag < (int)(p + 1)) { if (++spins > (1ll << 24)) { err = 6; break; } __nanosleep(64); }
    if (err) break;
    __threadfence_system();
    const volatile double *rec = (e == 0 ? xc.lo_edges : xc.hi_edges) + (((size_t)(p & 1) * 2 + their_edge) * m.n_ladders + ladder) * (D + 3);
    const double nb_ll = rec[D], nb_lprior = rec[D + 1], nb_beta = rec[D + 2];
    const bool i_am_lower = (e == 1);
    double lla = i_am_lower ? ch.llike : nb_ll; if (!(lla > -1e200)) lla = -1e200;
    double llb = i_am_lower ? nb_ll : ch.llike; if (!(llb > -1e200)) llb = -1e200;
    const double ba = i_am_lower ? ch.beta : nb_beta, bb = i_am_lower ? nb_beta : ch.beta;
    const double lhr = -(bb - ba) * (llb - lla);
    bool accept = true;
    if (lhr < 0) {
      uint32_t q[4];
      ptg_philox_draw(xc.shared_seed, (uint64_t)gl * PTG_STREAM_STRIDE + PTG_STREAM_LADDER, PTG_DOMAIN_BOUNDARY, (uint64_t)p,
                      (uint32_t)(e == 0 ? xc.lo_boundary : xc.hi_boundary), q);
      accept = (log(ptg_u52_to_unit(q[0], q[1])) < lhr);
    }
    if (accept) {
#pragma unroll
      for (int k = 0; k < D; k++) ch.x[k] = rec[k];
      ch.llike = nb_ll; ch.lprior = nb_lprior;
      ch.lpost = nb_lprior + ch.beta * nb_ll; // lpost recomputed = lprior(x) + beta*llike (chain.cc:925-928)
    }
    fappend<D>(m, s, ch, chain, hb, cnt);
    if (i_am_lower) cnt[FC_SC] += 1 + (accept ? (1 << 16) : 0); // counted at the pair's lower rung
  }
}

// XCHG: 0 = plain; 1 = with the rung-boundary exchange in the prologue / epilogue (ptg_step_exchange); 2 = also inside the iteration loop.
// Separate instantiations: each form's extra code costs registers and stack, and the plain kernel carries none of it
template <int D, int XCHG>
__global__ void __launch_bounds__(PTG_FSTEP_MAX_THREADS, 1) ptg_fstep_kernel(const __grid_constant__ PtgModel m, PtgState s, long long step0, int n_steps, int W, const __grid_constant__ PtgXchg xc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int R = m.n_rungs, NP = m.n_props;
  double *sbins = reinterpret_cast<double *>(smem_raw);                   // [R][NP]
  FProp *sprop = reinterpret_cast<FProp *>(sbins + R * NP);               // [NP]
  int *cnt = reinterpret_cast<int *>(sprop + NP) + threadIdx.x * PTG_FC_STRIDE; // [blockDim][PTG_FC_STRIDE], this thread's row
  for (int i = threadIdx.x; i < R * NP; i += blockDim.x) sbins[i] = m.bins[i];
  for (int i = threadIdx.x; i < NP; i += blockDim.x) {
    const PtgProp &p = m.props[i];
    FProp q;
    q.snooker = p.snooker; q.g1frac = p.g1frac; q.gamma_std = p.gamma_std; q.reduce_gamma = p.reduce_gamma; q.ignore_frac = p.ignore_frac;
    q.unlikely_alpha = p.unlikely_alpha; q.one_d_frac = p.one_d_frac; q.kind = p.kind; q.has_transform = p.has_transform;
    q.sigma_off = p.sigma_off; q.trans_off = p.trans_off;
    sprop[i] = q;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int gpw = 32 / W, g = lane / W, rung = lane - g * W;
  if (warp * gpw >= m.n_ladders) return; // whole warp beyond the batch
  const long long ladder = warp * gpw + g;
  const bool ladder_ok = ladder < m.n_ladders;
  const bool active = ladder_ok && rung < R;
  const unsigned gm = (W == 32) ? 0xffffffffu : (((1u << W) - 1u) << (g * W));
  const int chain = active ? (int)(ladder * R + rung) : 0; // ptg_create guarantees n_chains < 2^31
  // Long-lived values are kept to 32 bits and widened at the point of use (register budget): global ladder id, chain id.
  const int gl = (int)(m.ladder_offset + warp * gpw) + g; // ptg_create guarantees ladder ids < 2^31
#ifndef PTG_NO_NARROW_IDS
#define stream_base ((uint64_t)(gl - g) * PTG_STREAM_STRIDE)
#define my_stream ((uint64_t)gl * PTG_STREAM_STRIDE + (uint64_t)rung)
#define ladder_stream ((uint64_t)gl * PTG_STREAM_STRIDE + PTG_STREAM_LADDER)
#define hbase (s.hist + chain * ((long long)m.hist_cap * (D + 2)))
#define step ((uint64_t)(step0 + it))
#else
  const uint64_t stream_base = ((uint64_t)(gl - g) * PTG_STREAM_STRIDE);
  const uint64_t my_stream = ((uint64_t)gl * PTG_STREAM_STRIDE + (uint64_t)rung);
  const uint64_t ladder_stream = ((uint64_t)gl * PTG_STREAM_STRIDE + PTG_STREAM_LADDER);
  double *__restrict__ hbase = (s.hist + chain * ((long long)m.hist_cap * (D + 2)));
#define step ((uint64_t)(step0 + it))
#endif
  const double *bins = sbins + (rung < R ? rung : 0) * NP;

  FChain<D> ch;
  int st_di = 1;
#pragma unroll
  for (int k = 0; k < FC_COUNT; k++) cnt[k] = 0;
  if (active) {
#pragma unroll
    for (int k = 0; k < D; k++) ch.x[k] = s.cur_x[(long long)k * m.n_chains + chain];
    ch.lpost = s.lpost[chain]; ch.llike = s.llike[chain]; ch.lprior = s.lprior[chain]; ch.beta = s.beta[chain];
    const long long nsize = s.nsize[chain];
    ch.slot = (int)(nsize % m.hist_cap);
    ch.hfill = (int)(nsize > m.hist_cap ? (long long)m.hist_cap : nsize);
    ch.since_save = (int)(s.nhist[chain] % m.save_every);
    cnt[FC_LAST_TYPE] = s.last_type[chain];
    st_di = (s.directions[chain] + 1) | (s.instances[chain] << 2);
  } else {
#pragma unroll
    for (int k = 0; k < D; k++) ch.x[k] = 0;
    ch.lpost = ch.llike = ch.lprior = 0; ch.beta = 1;
    ch.slot = 0; ch.hfill = 0; ch.since_save = 0;
  }
  const int since_save0 = ch.since_save;
  int err = 0;

  // ================================================================= rung-sharded ladders: pending cross-GPU boundary swap
  long long xp = xc.index; // running publish index (XCHG instantiation only)
  if (XCHG && xc.on && xc.swap_in && active) fx_swap<D>(m, s, xc, xp - 1, ch, chain, ladder, gl, rung, R, cnt, err);

  const int maxswaps = m.maxswaps;
  const double swap_thresh = (R - 1) * m.swap_rate / maxswaps; // chain.cc:1413
  double ptry = 2 * m.swap_rate; if (ptry > 1) ptry = 1;
  const bool zero_valid = m.zero_valid != 0;

  for (int it = 0; it < n_steps; it++) {
    bool swapped = false;

    // ================================================================= swap phase (chain.cc:1410-1538)
    if (ladder_ok && R > 1) {
      if (m.swap_mode == PTG_SWAP_REFERENCE) {
        // lane j prepares trial j: candidate pair and log of its swap draw, from block j of the ladder's stream
        int raw = -2;
        double logu = 0;
        if (rung < maxswaps) {
          uint32_t q[4];
          ptg_philox_draw(m.seed, ladder_stream, PTG_DOMAIN_STEP, step, (uint32_t)rung, q);
          if (ptg_u32_to_unit(q[0]) < swap_thresh) {
            raw = (int)(ptg_u32_to_unit(q[1]) * (R - 1));
            logu = log(ptg_u52_to_unit(q[2], q[3]));
          }
        }
        // serial de-dup (iswaps[j]==cand or iswaps[j]+1==cand for an earlier surviving j) over the trials that drew a pair
        unsigned used = 0, live = 0; // live: bit j = trial j survives
        unsigned cand = (__ballot_sync(gm, raw >= 0) & gm) >> (g * W);
        while (cand) {
          const int i = __ffs(cand) - 1;
          cand &= cand - 1;
          const int c = __shfl_sync(gm, raw, i, W);
          if (!((used >> c) & 1u) && !(c > 0 && ((used >> (c - 1)) & 1u))) { used |= 1u << c; live |= 1u << i; }
        }
        while (live) {
          const int j = __ffs(live) - 1;
          live &= live - 1;
          const int c = __shfl_sync(gm, raw, j, W);
          const double lu = __shfl_sync(gm, logu, j, W);
          // every lane r evaluates pair (r, r+1); the decision of pair c is broadcast
          const double ll_up = __shfl_down_sync(gm, ch.llike, 1, W), b_up = __shfl_down_sync(gm, ch.beta, 1, W);
          double lla = ch.llike; if (!(lla > -1e200)) lla = -1e200;
          double llb = ll_up; if (!(llb > -1e200)) llb = -1e200;
          const double lhr = -(b_up - ch.beta) * (llb - lla);
          int acc_mine = 1;
          if (lhr < 0) acc_mine = (lu < lhr) ? 1 : 0;
          const bool accept = __shfl_sync(gm, acc_mine, c, W) != 0;
          const bool is_lo = (rung == c), is_hi = (rung == c + 1), involved = is_lo || is_hi;
          if (is_lo && c > 0) { // ups / downs of the lower rung before the exchange (chain.cc:1440-1443)
            const int dir = (st_di & 3) - 1;
            if (dir > 0) cnt[FC_UD] += 1;
            if (dir < 0) cnt[FC_UD] += 1 << 16;
          }
          if (accept) {
            const int partner = is_lo ? c + 1 : (is_hi ? c : rung);
#pragma unroll
            for (int k = 0; k < D; k++) { const double v = __shfl_sync(gm, ch.x[k], partner, W); ch.x[k] = v; }
            { const double v = __shfl_sync(gm, ch.llike, partner, W); ch.llike = v; }
            { const double v = __shfl_sync(gm, ch.lprior, partner, W); ch.lprior = v; }
            if (involved) ch.lpost = ch.lprior + ch.beta * ch.llike; // lpost recomputed = lprior(x) + beta*llike (chain.cc:925-928)
          }
          if (involved) {
            if (active) fappend<D>(m, s, ch, chain, hbase, cnt);
            swapped = true;
          }
          if (accept) {
            const int partner = is_lo ? c + 1 : (is_hi ? c : rung);
            { const int v = __shfl_sync(gm, st_di, partner, W); st_di = v; }
            if (c == 0 && is_lo) st_di = (st_di & ~3) | 2;          // directions[0] = +1
            if (c + 1 == R - 1 && is_hi) st_di = (st_di & ~3) | 0;  // directions[R-1] = -1
            if (is_lo) cnt[FC_SC] += 1 << 16;
            if (m.evolve_rate > 0) {
              // pry_temps, one pried gap (chain.cc:1809-1846) + resetTemp (chain.cc:1088-1091), reference summation order
              const double rate = m.evolve_rate;
              const double b_next = __shfl_down_sync(gm, ch.beta, 1, W), lp_next = __shfl_down_sync(gm, ch.lpost, 1, W);
              double sp = ch.beta - b_next;
              if (m.evolve_lpost_cut >= 0 && ch.lpost - lp_next > m.evolve_lpost_cut * ch.beta) sp *= (1.0 + rate);
              if (is_lo) sp *= 1.0 + rate;
              double sum = 0;
              for (int k = 0; k < R - 1; k++) sum += __shfl_sync(gm, sp, k, W);
              const double norm = sum / (1 - __shfl_sync(gm, ch.beta, R - 1, W));
              const double qn = sp / norm;
              double invtemp = 1, mine = ch.beta;
              for (int k = 1; k < R - 1; k++) {
                invtemp -= __shfl_sync(gm, qn, k - 1, W);
                if (rung == k) mine = invtemp;
              }
              if (rung >= 1 && rung < R - 1) { ch.beta = mine; ch.lpost = ch.lprior + mine * ch.llike; }
            }
          }
          if (is_lo) cnt[FC_SC] += 1;
        }
      } else {
        // even/odd performance mode: all pairs (i,i+1), i = parity, parity+2, ... are disjoint -> one shuffle round
        const int parity = (int)(step & 1);
        const bool is_lo = ((rung & 1) == parity) && (rung + 1 < R);
        const bool is_hi = (rung >= 1) && (((rung - 1) & 1) == parity) && (rung < R);
        const int partner = is_lo ? rung + 1 : (is_hi ? rung - 1 : rung);
        const double ll_p = __shfl_sync(gm, ch.llike, partner, W), b_p = __shfl_sync(gm, ch.beta, partner, W);
        int flags = 0; // bit0 tried, bit1 accepted (decided by the lower lane)
        if (is_lo) {
          uint32_t q[4];
          ptg_philox_draw(m.seed, ladder_stream, PTG_DOMAIN_STEP, step, PTG_BLK_SWAP_EVENODD + (uint32_t)rung, q);
          if (ptg_u52_to_unit(q[0], q[1]) < ptry) {
            double lla = ch.llike; if (!(lla > -1e200)) lla = -1e200;
            double llb = ll_p; if (!(llb > -1e200)) llb = -1e200;
            const double lhr = -(b_p - ch.beta) * (llb - lla);
            bool accept = true;
            if (lhr < 0) accept = (log(ptg_u52_to_unit(q[2], q[3])) < lhr);
            flags = 1 | (accept ? 2 : 0);
            if (rung > 0) { const int dir = (st_di & 3) - 1; if (dir > 0) cnt[FC_UD] += 1; if (dir < 0) cnt[FC_UD] += 1 << 16; }
            cnt[FC_SC] += 1 + (accept ? (1 << 16) : 0);
          }
        }
        const int pflags = __shfl_sync(gm, flags, partner, W);
        if (is_hi) flags = pflags;
        const bool tried = (flags & 1) != 0, accept = (flags & 2) != 0;
        double nx[D];
#pragma unroll
        for (int k = 0; k < D; k++) nx[k] = __shfl_sync(gm, ch.x[k], partner, W);
        const double nll = __shfl_sync(gm, ch.llike, partner, W), nlp = __shfl_sync(gm, ch.lprior, partner, W);
        const int ndi = __shfl_sync(gm, st_di, partner, W);
        if (tried && accept) {
#pragma unroll
          for (int k = 0; k < D; k++) ch.x[k] = nx[k];
          ch.llike = nll; ch.lprior = nlp;
          ch.lpost = ch.lprior + ch.beta * ch.llike;
          st_di = ndi;
          if (rung == 0) st_di = (st_di & ~3) | 2;
          if (rung == R - 1) st_di = (st_di & ~3) | 0;
        }
        if (tried) { if (active) fappend<D>(m, s, ch, chain, hbase, cnt); swapped = true; }
      }
    }
    __syncwarp();

    // ================================================================= MH update (chain.cc:966-1022)
    const bool do_mh = active && !swapped;
    uint32_t wA[4], wB[4];
    ptg_philox_draw(m.seed, my_stream, PTG_DOMAIN_STEP, step, PTG_BLK_A, wA);
    {
      // block B is issued AFTER block A (an opaque zero ties its counter to A's output): the two evaluations would otherwise
      // be interleaved for ILP, doubling the live Philox state; with 28 resident warps per SM latency is hidden by TLP
      uint32_t zero = 0;
#ifndef PTG_NO_SERIAL_PHILOX
      asm volatile("and.b32 %0, %1, 0;" : "=r"(zero) : "r"(wA[0]));
#endif
      ptg_philox_draw(m.seed, my_stream, PTG_DOMAIN_STEP, step, PTG_BLK_B + zero, wB);
    }

    // ---- member selection: first ready member with u < bin_max (proposal_distribution.cc:105-112)
    int member = 0;
    if (m.wrap_in_set) {
      const double x = (NP > 1) ? ptg_u32_to_unit(wA[0]) : 0.0;
      if (ch.hfill >= D * 10 || !do_mh) { // every member ready (differential_evolution::is_ready, proposal_distribution.hh:407)
        member = NP - 1;
        for (int i = NP - 2; i >= 0; i--) if (x < bins[i]) member = i;
      } else {
        member = -1;
        for (int i = 0; i < NP; i++) if (member < 0 && sprop[i].kind != PTG_PROP_DE && x < bins[i]) member = i;
        if (member < 0) { err = 2; member = 0; }
      }
    }
    const FProp &p = sprop[member];
    const int kind = do_mh ? p.kind : 0;
    const double oldlprior = ch.lpost - ch.beta * ch.llike;
    double newx[D];
    double prop_lh = 0;
    int type = 0;
    bool valid = zero_valid;

    // ---- Gaussian members: cooperative normals, then x' = x + M (z o sigma)  (proposal_distribution.hh:194-218)
    {
      double z[D];
      wcoop_normals<D>(m.seed, step, stream_base, W, kind == PTG_PROP_GAUSS, z);
#pragma unroll
      for (int i = 0; i < D; i++) newx[i] = ch.x[i];
      if (kind == PTG_PROP_GAUSS) {
        const double *__restrict__ sig = m.prop_data + p.sigma_off;
#pragma unroll
        for (int i = 0; i < D; i++) z[i] = z[i] * __ldg(sig + i) + 0.0;
        if (p.one_d_frac > 0 && ptg_u32_to_unit(wA[1]) < p.one_d_frac) {
          const int ia = (int)(D * ptg_u32_to_unit(wA[2]));
#pragma unroll
          for (int j = 0; j < D; j++) if (j != ia) z[j] = 0.0;
          type = 1;
        }
        if (p.has_transform) {
          FVec<D> o;
#pragma unroll
          for (int i = 0; i < D; i++) o.v[i] = z[i];
          o = ftransform<D>(m.prop_data + p.trans_off, o);
#pragma unroll
          for (int i = 0; i < D; i++) z[i] = o.v[i];
        }
#pragma unroll
        for (int i = 0; i < D; i++) newx[i] = ch.x[i] + z[i];
      }
    }
    // ---- differential evolution (proposal_distribution.cc:489-591,744-801)
    if (kind == PTG_PROP_DE) {
      const int hsize = ch.hfill;
      int start = 0;
      if ((hsize - D * 100) * (1 - p.ignore_frac) > D * 10) start = (int)((hsize - D * 100) * p.ignore_frac);
      const bool snooker = p.snooker > ptg_u32_to_unit(wA[1]);
      const double ug = ptg_u32_to_unit(wA[2]);
      int i1, i2, iz = 0, az = 0;
      const bool slow = p.unlikely_alpha > 0;
      if (!slow) {
        i1 = (int)(start + (hsize - start) * ptg_u32_to_unit(wA[3]));
        i2 = (int)(start + (hsize - start) * ptg_u32_to_unit(wB[0]));
        iz = (int)(start + (hsize - start) * ptg_u32_to_unit(wB[1]));
        az = 1;
      } else {
        const double mapl = s.map_lpost[chain];
        if (snooker) { const int2 r = fde_index_slow<D>(m, ch.hfill, ch.slot, hbase, mapl, m.seed, my_stream, step, p.ignore_frac, p.unlikely_alpha, wB[1], 0, 0); iz = r.x; az = r.y; }
        i1 = fde_index_slow<D>(m, ch.hfill, ch.slot, hbase, mapl, m.seed, my_stream, step, p.ignore_frac, p.unlikely_alpha, wA[3], 1, 0).x;
        i2 = fde_index_slow<D>(m, ch.hfill, ch.slot, hbase, mapl, m.seed, my_stream, step, p.ignore_frac, p.unlikely_alpha, wB[0], 2, 0).x;
      }
      const double *s1 = fhist<D>(m, ch, hbase, i1), *s2 = fhist<D>(m, ch, hbase, i2);
      double a[D], b[D];
#pragma unroll
      for (int i = 0; i < D; i++) { a[i] = __ldcg(s1 + i); b[i] = __ldcg(s2 + i); } // random records: no reuse, keep them out of L1
      if (!snooker) {
        // draw_standard: prop = (s + gamma s1) + (-gamma s2); the jitter drawn by the reference is discarded (H8-1)
        double gamma = p.gamma_std;
        if (ug < p.g1frac) gamma = 1;
#pragma unroll
        for (int i = 0; i < D; i++) {
          const double t = ch.x[i] + a[i] * gamma;
          newx[i] = t + b[i] * (-gamma);
        }
      } else {
        // draw_snooker (proposal_distribution.cc:538-591)
        const double gamma = (1.2 + ug) / p.reduce_gamma;
        double smznorm2 = 0, minusz[D], smz[D];
        int isafe = 0;
        while (true) {
          const double *zz = fhist<D>(m, ch, hbase, iz);
          smznorm2 = 0;
#pragma unroll
          for (int i = 0; i < D; i++) { minusz[i] = zz[i] * (-1); smz[i] = ch.x[i] + minusz[i]; }
#pragma unroll
          for (int i = 0; i < D; i++) smznorm2 += smz[i] * smz[i];
          if (smznorm2 != 0 || ++isafe > 1000) break;
          const int2 r = fde_index_slow<D>(m, ch.hfill, ch.slot, hbase, s.map_lpost[chain], m.seed, my_stream, step, p.ignore_frac, p.unlikely_alpha, wB[1], 0, az);
          iz = r.x; az = r.y;
        }
        double dot = 0;
#pragma unroll
        for (int i = 0; i < D; i++) {
          const double ds12 = a[i] * gamma + b[i] * (-gamma);
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
    } else if (kind == PTG_PROP_PRIOR_DRAW) {
      FVec<D> xv;
#pragma unroll
      for (int i = 0; i < D; i++) xv.v[i] = ch.x[i];
      const FVecFlag<D> r = fprior_member<D>(m, my_stream, step, xv);
#pragma unroll
      for (int i = 0; i < D; i++) newx[i] = r.v[i];
      prop_lh = r.aux; valid = r.ok != 0;
    }
    if (m.wrap_in_set) type = member + 10 * type;

    // ---- enforce, prior, gated likelihood (chain.cc:976-987)
    if (m.any_bound && valid) {
      FVec<D> xv;
#pragma unroll
      for (int i = 0; i < D; i++) xv.v[i] = newx[i];
      const FVecFlag<D> r = fenforce<D>(m, xv);
#pragma unroll
      for (int i = 0; i < D; i++) newx[i] = r.v[i];
      valid = r.ok != 0;
    }
    double newlprior;
    if (m.all_uniform_prior) {
      bool in = valid;
#pragma unroll
      for (int i = 0; i < D; i++) in = in && !(newx[i] < m.prior[i].a) && !(newx[i] > m.prior[i].b);
      newlprior = in ? m.uniform_lprior : -CUDART_INF;
    } else {
      FVec<D> xv;
#pragma unroll
      for (int i = 0; i < D; i++) xv.v[i] = newx[i];
      newlprior = fprior_general<D>(m, xv, valid);
    }
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
    if (accept && lhr < 0) accept = (log(ptg_u52_to_unit(wB[2], wB[3])) < lhr);
    if (do_mh) {
      cnt[FC_NTRIES]++;
      if (accept) {
        cnt[FC_NACCEPT]++;
        cnt[FC_LAST_TYPE] = type;
#pragma unroll
        for (int i = 0; i < D; i++) ch.x[i] = newx[i];
        ch.llike = newlike; ch.lpost = newlpost; ch.lprior = newlprior;
        code |= PTG_TRACE_ACCEPT;
      }
      fappend<D>(m, s, ch, chain, hbase, cnt);
    }
    if (active && (long long)step < m.trace_steps) {
      s.trace_lhr[step * m.n_chains + chain] = do_mh ? lhr : 0.0;
      s.trace_code[step * m.n_chains + chain] = do_mh ? (code | (type & PTG_TRACE_TYPE_MASK)) : PTG_TRACE_SWAPPED;
    }
    // in-launch exchange: after every xc.every-th iteration (not the launch's last, whose publish the epilogue does) publish the
    // edges and run the boundary trial as soon as the neighbour's same ladder has published -- warps wait one by one, the rest of
    // the SM keeps stepping.  The host only uses this when every CTA of the grid is resident (one wave on each GPU).
    if (XCHG == 2 && xc.on && xc.every > 0 && it + 1 < n_steps && (it + 1) % xc.every == 0) {
      if (active) {
        fx_publish<D>(m, xc, xp, ch, ladder, rung, R);
        fx_swap<D>(m, s, xc, xp, ch, chain, ladder, gl, rung, R, cnt, err);
      }
      xp++;
      __syncwarp();
    }
  }

  // ================================================================= rung-sharded ladders: publish this block's edge rungs
  if (XCHG && xc.on && xc.publish_out && active) fx_publish<D>(m, xc, xp, ch, ladder, rung, R);

  if (active) {
#pragma unroll
    for (int k = 0; k < D; k++) s.cur_x[(long long)k * m.n_chains + chain] = ch.x[k];
    s.lpost[chain] = ch.lpost; s.llike[chain] = ch.llike; s.lprior[chain] = ch.lprior; s.beta[chain] = ch.beta;
    // saves = appends k in [0, dnhist) with (since_save0 + k) % save_every == 0
    const int dnhist = cnt[FC_NHIST], se = m.save_every;
    const int dnsize = (since_save0 + dnhist + se - 1) / se - (since_save0 + se - 1) / se;
    s.nhist[chain] += dnhist; s.nsize[chain] += dnsize; s.ntries[chain] += cnt[FC_NTRIES]; s.naccept[chain] += cnt[FC_NACCEPT];
    s.last_type[chain] = cnt[FC_LAST_TYPE];
    s.directions[chain] = (st_di & 3) - 1; s.instances[chain] = st_di >> 2;
    const int st_ud = cnt[FC_UD], st_sc = cnt[FC_SC];
    s.ups[chain] += st_ud & 0xffff; s.downs[chain] += st_ud >> 16;
    s.swap_count[chain] += st_sc & 0xffff; s.swap_accept[chain] += st_sc >> 16;
    if (err) atomicMax(s.err, err);
  }
}
#undef stream_base
#undef my_stream
#undef ladder_stream
#undef hbase
#undef step
