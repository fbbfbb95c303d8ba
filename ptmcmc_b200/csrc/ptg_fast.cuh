// ptg_fast.cuh -- the production step kernel (Philox draws): ladder-in-a-warp like ptg_warp.cuh, restructured around the
// ncu findings of rounds 1 and 2 (profiles/README.md):
//   * one wave: chain state is slimmed to <= 72 registers/thread (launch-local counters and the per-chain MAP in shared
//     memory, packed ladder statistics) so that all 4096 warps of BASELINE config C1 are resident at once;
//   * compile-time specialisation: LK >= 0 instantiates the kernel for the common production configuration (reference swap
//     schedule without temperature evolution, open state space, uniform priors, DE + plain Gaussian members, likelihood LK);
//     every rare-path branch folds away (13.5 k -> ~3 k SASS instructions, +22 % measured).  LK = -1 keeps every feature;
//   * swap phase: every lane prepares ITS trial in parallel; the serial part is a bit-mask de-dup plus one shuffle round per
//     surviving trial; the history appends of swapped rungs are deferred to the ONE append at the end of the iteration
//     that MH lanes execute too (a rung in two trials of one step, SURVEY.md H4, flushes the first before the second);
//   * pooled work rounds: the few-lane sections of a step share ONE converged call site each --
//       round 1: Philox block + log of (a) every Box-Muller pair a Gaussian-member lane needs (1-D steps only need the
//                pair of their axis) and (b) the NEXT iteration's swap-trial draws, spread over all 32 lanes;
//       round 2: log of the Metropolis draw of every MH lane and, on the lanes that do no MH update this iteration, the
//                two logarithms of the snooker Hastings term;
//     values travel through per-warp shared-memory slots, not registers;
//   * history x-ring in whole 32-byte sectors (PTG_HX): a DE gather at dim <= 4 is one 256-bit load of exactly one sector.
// The arithmetic is expression-for-expression that of ptg_warp.cuh / ptg_kernels.cuh / the oracle: in Philox mode the
// three kernels produce bit-identical chains (tests/test_gpu_properties.py::test_kernels_agree_bitwise).
#pragma once
#include "ptg_warp.cuh"

// register-resident part of a chain's state; lpost, lprior and the running MAP live in per-thread shared-memory slots (FShared):
// they are touched a few times per iteration and would otherwise be spilled under the one-wave register budget
template <int D>
struct FChain {
  double x[D];
  double llike, beta;
};
// largest CTA the production kernel is launched with (28 warps = one CTA per SM at 72 registers); smaller batches use smaller CTAs
#define PTG_FSTEP_MAX_THREADS 896
// Launch-local per-thread words in shared memory ([thread][PTG_FC_STRIDE], odd stride: conflict-free): counters that change rarely
// the swap-trial draw this lane prepared for the next iteration, and the chain's ring position (FC_SLOT write position, FC_HFILL = min(nsize,
// capacity), FC_SSAVE = nhist % save_every) and packed replica direction / instance (FC_DI): read a few times per iteration, kept out of registers
enum { FC_XAPP = 0, FC_NSWAPPED, FC_NACCEPT, FC_LAST_TYPE, FC_UD, FC_SC, FC_RAW, FC_LOGU_LO, FC_LOGU_HI, FC_SLOT, FC_HFILL, FC_SSAVE, FC_DI, FC_K0, FC_K1, FC_K2, FC_K3, FC_K4, FC_K5, FC_SS0, FC_PAD, FC_COUNT };
#define PTG_FC_STRIDE 21
static_assert(FC_COUNT == PTG_FC_STRIDE, "counter row");

// proposal member parameters staged in shared memory
struct FProp {
  double snooker, g1frac, gamma_std, reduce_gamma, ignore_frac, unlikely_alpha, one_d_frac;
  int kind, has_transform, sigma_off, trans_off;
};

// Dynamic shared memory of a CTA.  Every offset is a compile-time constant of D (fixed-size tables first, then one block per warp), so an
// address is `window base + constant (+ warp * W_BYTES) (+ lane * stride)`: with run-time offsets the compiler re-derived the bases
// inside the step loop (~10 % of the issued instructions, profiles/README.md).
template <int D>
struct FShared {
  static constexpr int NPAIR = (D + 1) / 2;
  static constexpr int OFF_SPAR = 0;                                        // [4 D] doubles   streamlined sines functor: per dimension (min, width or 1/width, k pi, k)
  static constexpr int OFF_PROP = OFF_SPAR + 8 * 4 * D;                     // [16] FProp      proposal members
  static constexpr int OFF_BINS = OFF_PROP + (int)sizeof(FProp) * PTG_MAX_PROPOSALS; // [32][16] doubles bins[rung][member] (n_rungs <= 32)
  static constexpr int OFF_WARP = OFF_BINS + 8 * 32 * PTG_MAX_PROPOSALS;    // per-warp blocks
  static constexpr int W_MAP = 0;                                           // [32][3] doubles  per lane: running MAP posterior, current lpost, current lprior (stride 3: conflict-free)
  static constexpr int W_CNT = W_MAP + 8 * 3 * 32;                          // [32][PTG_FC_STRIDE] ints  per lane: launch-local counters, prepared swap-trial draw, ring position
  static constexpr int W_LUS = W_CNT + 4 * PTG_FC_STRIDE * 32;              // [32] doubles     swap phase: log u of the trial on pair c, handed to the lane of rung c
  static constexpr int W_POOL = W_LUS + 8 * 32;                             // [2 NPAIR][32] doubles  round 1: normals of lane l at [j][l]; round 2 (aliased): [64] arguments / logarithms
  static constexpr int W_ITEMS = W_POOL + 8 * 2 * NPAIR * 32;               // [32 NPAIR] bytes  Box-Muller work list: owner lane << 3 | pair
  static constexpr int W_BYTES = (W_ITEMS + 32 * NPAIR + 15) & ~15;
  __host__ __device__ static size_t bytes(int threads) { return (size_t)OFF_WARP + (size_t)(threads / 32) * W_BYTES; }
};

// ---- history ring access: x-records of PTG_HX(D) doubles (whole 32-byte sectors), (lpost, llike) pairs in their own ring
// Ring traffic never re-uses a line soon (a record is read ~1.7 times in its life, thousands of iterations apart): it is marked
// evict-first in L2 so that it does not push out what IS re-used every iteration (the step loop's local-memory lines, the tables).
struct __align__(32) FRec4 { double a, b, c, d; };
__device__ __forceinline__ uint64_t fring_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
template <int D>
__device__ __forceinline__ void fload_rec(const double *__restrict__ p, double v[D]) {
  constexpr int HX = PTG_HX(D);
#ifndef PTG_F_NO_EVICT_FIRST
  const uint64_t pol = fring_policy();
#endif
#pragma unroll
  for (int k = 0; k < HX; k += 4) {
    double a, b, c, d;
    // random records with no reuse: one 256-bit load per sector, not allocated in L1
#if defined(PTG_F_PREFETCH_L1)
    asm volatile("ld.global.L1::evict_first.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p + k), "l"(pol) : "memory");
#elif !defined(PTG_F_NO_EVICT_FIRST)
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p + k), "l"(pol) : "memory");
#else
    asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p + k) : "memory");
#endif
    v[k] = a;
    if (k + 1 < D) v[k + 1] = b;
    if (k + 2 < D) v[k + 2] = c;
    if (k + 3 < D) v[k + 3] = d;
  }
}
template <int D>
__device__ __forceinline__ void fstore_rec(double *__restrict__ p, const double v[D]) {
  constexpr int HX = PTG_HX(D);
#ifndef PTG_F_NO_EVICT_FIRST
  const uint64_t pol = fring_policy();
#endif
#pragma unroll
  for (int k = 0; k < HX; k += 4) {
    const double b = (k + 1 < D) ? v[k + 1] : 0.0, c = (k + 2 < D) ? v[k + 2] : 0.0, d = (k + 3 < D) ? v[k + 3] : 0.0;
#ifndef PTG_F_NO_EVICT_FIRST
    asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" :: "l"(p + k), "d"(v[k]), "d"(b), "d"(c), "d"(d), "l"(pol) : "memory");
#else
    FRec4 r; r.a = v[k]; r.b = b; r.c = c; r.d = d;
    *reinterpret_cast<FRec4 *>(p + k) = r;
#endif
  }
}
__device__ __forceinline__ void fstore_lp(double *__restrict__ p, double lpost, double llike) {
#ifndef PTG_F_NO_EVICT_FIRST
  const uint64_t pol = fring_policy();
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" :: "l"(p), "d"(lpost), "d"(llike), "l"(pol) : "memory");
#else
  *reinterpret_cast<double2 *>(p) = make_double2(lpost, llike);
#endif
}

// MH_chain::add_state (chain.cc:916-949) of the chain's current state
template <int D, bool RECORD_FULL_POSSIBLE>
__device__ __forceinline__ void fappend(const PtgModel &m, const PtgState &s, FChain<D> &ch, long long chain, double *__restrict__ hbase, int *cnt, double *mapp, int n_mh_so_far) {
  const double lpost = mapp[1]; // per-thread slots: [0] running MAP, [1] current lpost, [2] current lprior
  if (lpost > *mapp) { // MAP update (chain.cc:931-934)
    *mapp = lpost;
#pragma unroll
    for (int k = 0; k < D; k++) s.map_x[(long long)k * m.n_chains + chain] = ch.x[k];
  }
  if (cnt[FC_SSAVE] == 0) {
    fstore_rec<D>(hbase + (long long)cnt[FC_SLOT] * PTG_HX(D), ch.x);
    const long long rec = chain * m.hist_cap + cnt[FC_SLOT];
    fstore_lp(s.hist_lp + 2 * rec, lpost, ch.llike);
    if (RECORD_FULL_POSSIBLE && m.record_full) {
      s.hist_acc[rec] = (s.naccept[chain] + cnt[FC_NACCEPT]) / (double)(s.ntries[chain] + n_mh_so_far);
      s.hist_beta[rec] = ch.beta;
      s.hist_type[rec] = cnt[FC_LAST_TYPE];
    }
    if (cnt[FC_HFILL] < m.hist_cap) cnt[FC_HFILL]++;
    cnt[FC_SLOT] = (cnt[FC_SLOT] + 1 == m.hist_cap) ? 0 : cnt[FC_SLOT] + 1;
  }
  cnt[FC_SSAVE] = (cnt[FC_SSAVE] + 1 == m.save_every) ? 0 : cnt[FC_SSAVE] + 1;
}

// element `index` of the eligible window (newest min(nsize,cap) samples): once the ring is full the oldest sits at `slot`
template <int D>
__device__ __forceinline__ int fslot(const PtgModel &m, const int *cnt, int index) {
  int p = index;
  if (cnt[FC_HFILL] == m.hist_cap) { p = cnt[FC_SLOT] + index; if (p >= m.hist_cap) p -= m.hist_cap; }
  return p;
}

// differential_evolution::draw_i_from_chain with unlikely_alpha > 0 or a retry (proposal_distribution.cc:744-778): rare path
template <int D>
__device__ __noinline__ int2 fde_index_slow(const PtgModel &m, int hfill, int slot, const double *lpbase, double map_lpost, uint64_t seed, uint64_t stream,
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
      const double lpost = lpbase[2 * p];
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

// ---- data chi-squared functors of the PRODUCTION instantiations (Philox runs under automatic kernel selection; bayesian.hh:595-622).
// One thread evaluates its chain's whole data sum with explicit fused multiply-adds and four independent partial sums; the data block is
// read with warp-uniform (broadcast) loads, 1/S_i is precomputed (ptg_api.cu).  The tape-capable kernels and the general instantiation
// keep the reference's unfused arithmetic (like_eval_kind); these agree with it to ~1e-15 relative (1e-12 gate in the tests).
//   polynomial (poly_example.cc:85-106): Horner's rule, D + 2 fp64 instructions per point
// sum over the points [lo, hi) of (poly(x_i) - y_i)^2 / S_i
template <int D>
__device__ __forceinline__ double flike_poly_partial(const PtgModel &m, const double x[D], long long lo, long long hi) {
  const long long N = m.n_ldata / 3;
  const double *__restrict__ xs = m.ldata, *__restrict__ ys = m.ldata + N, *__restrict__ iS = m.ldata + 3 * N;
  // eight points per trip: all 24 loads are issued before the eight independent Horner chains start (a chain that runs alone on its
  // scheduler -- config B puts one or two warps on each -- has nothing else to cover load and fp64 latency with)
  double p4[4] = {0, 0, 0, 0};
  long long i = lo;
  for (; i + 8 <= hi; i += 8) {
    double xi[8], yi[8], wi[8];
#pragma unroll
    for (int u = 0; u < 8; u++) { xi[u] = __ldg(xs + i + u); yi[u] = __ldg(ys + i + u); wi[u] = __ldg(iS + i + u); }
    double y[8];
#pragma unroll
    for (int u = 0; u < 8; u++) y[u] = x[D - 1];
#pragma unroll
    for (int j = D - 2; j >= 0; j--) {
#pragma unroll
      for (int u = 0; u < 8; u++) y[u] = fma(y[u], xi[u], x[j]);
    }
#pragma unroll
    for (int u = 0; u < 8; u++) { const double dd = y[u] - yi[u]; p4[u & 3] = fma(dd * dd, wi[u], p4[u & 3]); }
  }
  for (; i < hi; i++) {
    const double xi = __ldg(xs + i);
    double y = x[D - 1];
#pragma unroll
    for (int j = D - 2; j >= 0; j--) y = fma(y, xi, x[j]);
    const double dd = y - __ldg(ys + i);
    p4[0] = fma(dd * dd, __ldg(iS + i), p4[0]);
  }
  return (p4[0] + p4[1]) + (p4[2] + p4[3]);
}
__device__ __forceinline__ double flike_chi2_finish(const PtgModel &m, double part) {
  double sum = part + m.like_nsum;
  sum /= -2;
  double result = sum - __ldg(m.lparams);
  if (!isfinite(result)) result = -CUDART_INF;
  return result;
}
template <int D>
__device__ __forceinline__ double flike_poly_fused(const PtgModel &m, const double x[D]) {
  return flike_chi2_finish(m, flike_poly_partial<D>(m, x, 0, m.n_ldata / 3));
}
//   sum of sinusoids y(t) = sum_k A_k sin(2 pi f_k t + phi_k): on a uniform time grid sin / cos of every component
//   advance by one rotation per sample (4 fused multiply-adds) and are re-anchored with sincos of the reference's own phase expression
//   every 128 samples (drift < 1e-13); an irregular grid evaluates sin at every sample
// sum over the samples [lo, hi) (lo a multiple of 128 on a uniform grid: the re-anchoring blocks) of (y(t_i) - y_i)^2 / S_i
template <int D>
__device__ __forceinline__ double flike_sinusoid_partial(const PtgModel &m, const double x[D], long long lo, long long hi) {
  constexpr int NS = D / 3;
  const long long N = m.n_ldata / 3;
  const double *__restrict__ xs = m.ldata, *__restrict__ ys = m.ldata + N, *__restrict__ iS = m.ldata + 3 * N;
  double part = 0;
  if (m.like_uniform_t) {
    double sd[NS > 0 ? NS : 1], cd[NS > 0 ? NS : 1];
#pragma unroll
    for (int k = 0; k < NS; k++) sincos(2 * PTG_PI * x[3 * k + 1] * m.like_dt, &sd[k], &cd[k]);
    for (long long base = lo; base < hi; base += 128) {
      double sn[NS > 0 ? NS : 1], cs[NS > 0 ? NS : 1];
      const double tb = __ldg(xs + base);
#pragma unroll
      for (int k = 0; k < NS; k++) sincos(2 * PTG_PI * x[3 * k + 1] * tb + x[3 * k + 2], &sn[k], &cs[k]);
      const int nq = (int)((hi - base) < 128 ? (hi - base) : 128);
      // y_i and 1 / S_i are loaded TWO samples ahead of their use (ncu: the subtraction behind the y_i load held 24 % of all stall
      // samples -- L1-hit latency that 3.5 warps per scheduler cannot cover); the index is clamped to the slice, so no branch
      const double *__restrict__ yblk = ys + base, *__restrict__ wblk = iS + base;
      const int lim = nq - 1;                      // the block's last sample
      double ya = __ldg(yblk), wa = __ldg(wblk);
      double yb = __ldg(yblk + (1 < lim ? 1 : lim)), wb = __ldg(wblk + (1 < lim ? 1 : lim));
#pragma unroll 2
      for (int q = 0; q < nq; q++) {
        const double yq = ya, wq = wa;
        ya = yb; wa = wb;
        const int qn = q + 2 < lim ? q + 2 : lim;
        yb = __ldg(yblk + qn); wb = __ldg(wblk + qn);
        double y = 0;
#pragma unroll
        for (int k = 0; k < NS; k++) y = fma(x[3 * k], sn[k], y);
        const double dd = y - yq;
        part = fma(dd * dd, wq, part);
#pragma unroll
        for (int k = 0; k < NS; k++) {
          const double ns = fma(sn[k], cd[k], cs[k] * sd[k]);
          const double nc = fma(cs[k], cd[k], -(sn[k] * sd[k]));
          sn[k] = ns; cs[k] = nc;
        }
      }
    }
  } else {
    for (long long i = lo; i < hi; i++) {
      const double ti = __ldg(xs + i);
      double y = 0;
#pragma unroll
      for (int k = 0; k < NS; k++) y = fma(x[3 * k], sin(2 * PTG_PI * x[3 * k + 1] * ti + x[3 * k + 2]), y);
      const double dd = y - __ldg(ys + i);
      part = fma(dd * dd, __ldg(iS + i), part);
    }
  }
  return part;
}
template <int D>
__device__ __forceinline__ double flike_sinusoid_fused(const PtgModel &m, const double x[D]) {
  return flike_chi2_finish(m, flike_sinusoid_partial<D>(m, x, 0, m.n_ldata / 3));
}

// The data sums of a WARP'S LADDER, compacted: of the 32 lanes only the chains whose proposal passed the prior gate evaluate a likelihood
// (ncu: 15-17 of 32 lanes active in the data loops of configs B and C2), so the n wanting chains of the warp's ladder are given S = 32 / n
// lanes each -- lane g S + h evaluates slice h of the g-th wanting chain's data (its proposal arrives by shuffle) and the slices are added
// up in slice order.  The data-likelihood instantiations always run ONE ladder per warp (the host launches them with W = 32), so n and
// with it the summation order depend on the ladder's own state only: a shard and the full batch still produce the same chains.
// Slices are whole groups of 8 points (polynomial) / whole re-anchoring blocks of 128 samples (sinusoids on a uniform grid).
// Called by all 32 lanes, converged; returns the log-likelihood to the wanting lanes.
// (Measured and rejected: cutting every sum into K = max(8, 32 / n) slices dealt to the lanes 32 at a time, which also fills the warp when
// 17 ... 31 chains want a likelihood -- config B 7.5e8 -> 1.9e8, C2 5.8e7 -> 4.2e7: short slices lose the eight-point software pipeline
// and every pass pays the proposal shuffles and the recurrence set-up again.)
// One pass of the compacted data sums: the wanting chains of rank [c0, c0 + nc) get S = 32 / nc lanes each.  Returns the chain's data sum
// to the wanting lanes of those ranks (garbage elsewhere).  Called by all 32 lanes, converged.
template <int D, int LK>
__device__ __forceinline__ double flike_data_pass(const PtgModel &m, const double x[D], unsigned mask, int r, int c0, int nc) {
  const int lane = threadIdx.x & 31;
  const int S = 32 / nc;
  const int g = lane / S, h = lane - g * S;
  const bool valid = g < nc;
  const int owner = valid ? (int)__fns(mask, 0, c0 + g + 1) : 0;
  double hx[D];
#pragma unroll
  for (int i = 0; i < D; i++) hx[i] = __shfl_sync(0xffffffffu, x[i], owner);
  const long long N = m.n_ldata / 3;
  const long long unit = (LK == PTG_LIKE_SINUSOID_CHI2 && m.like_uniform_t) ? 128 : 8;
  const long long nunit = (N + unit - 1) / unit, per = (nunit + S - 1) / S;
  long long lo = (long long)h * per * unit, hi = lo + per * unit;
  if (lo > N) lo = N;
  if (hi > N) hi = N;
  double part = 0;
  if (valid) {
    if constexpr (LK == PTG_LIKE_POLY_CHI2) part = flike_poly_partial<D>(m, hx, lo, hi);
    else part = flike_sinusoid_partial<D>(m, hx, lo, hi);
  }
  double tot = 0;
  for (int k = 0; k < S; k++) tot += __shfl_sync(0xffffffffu, part, ((r - c0) * S + k) & 31);
  return tot;
}
// One pass gives each of the n wanting chains 32 / n lanes (integer division): 11 ... 15 chains get two lanes and leave up to 10 idle, 17 ... 31
// get ONE lane each (config C2: 32 rungs, 17-20 proposals pass the gate most of the time).  TWO passes can do better: the first a chains
// (a a power of two) on 32 / a lanes each, the other n - a on 32 / (n - a) lanes each -- e.g. 18 chains: N / 2 + N / 16 samples per lane
// instead of N; 11 chains: N / 4 + N / 10 instead of N / 2.  The split with the fewest slice units per lane is taken when it saves at least a
// tenth (a pass repeats the proposal shuffles and the recurrence set-up); it depends on n and the data size alone, like the slice width.
template <int D, int LK>
__device__ __forceinline__ double flike_data_compact(const PtgModel &m, const double x[D], bool want) {
  const int lane = threadIdx.x & 31;
  const unsigned mask = __ballot_sync(0xffffffffu, want);
  const int n = __popc(mask);
  if (n == 0) return 0.0;
  const int r = __popc(mask & ((1u << lane) - 1u));
  const long long N = m.n_ldata / 3;
  const int unit = (LK == PTG_LIKE_SINUSOID_CHI2 && m.like_uniform_t) ? 128 : 8;
  const int nunit = (int)((N + unit - 1) / unit);
  auto per = [&](int nc) { const int S = 32 / nc; return (nunit + S - 1) / S; };   // slice units per lane of a pass over nc chains
  int first = n, best = per(n);
  const int single = best;
  // sinusoids only: the polynomial's short dependent stretch per point makes a second pass cost more than the idle lanes (config B,
  // 1024 warps on 592 schedulers: 7.1e8 -> 4.2e8 chain-steps/s with the split, measured)
  if constexpr (LK == PTG_LIKE_SINUSOID_CHI2) {
#pragma unroll
    for (int a = 16; a >= 1; a >>= 1) {
      if (a < n) {
        const int c = per(a) + per(n - a);
        if (c < best && 10 * c <= 9 * single) { best = c; first = a; }
      }
    }
  }
  double tot = 0;
  int c0 = 0;
#pragma unroll 1
  for (int pass = 0; pass < (first < n ? 2 : 1); pass++) {   // one body for both passes (separate inlined copies spilled)
    const int nc = (pass == 0) ? first : n - first;
    const double t = flike_data_pass<D, LK>(m, x, mask, r, c0, nc);
    if (r >= c0 && r < c0 + nc) tot = t;
    c0 += nc;
  }
  return flike_chi2_finish(m, tot);
}

// likelihood of the specialised instantiations: the one functor, no switch
template <int D, int LK>
__device__ __forceinline__ double flike(const PtgModel &m, const double x[D]) {
  if constexpr (LK < 0) return like_eval<D>(m, x);
  else if constexpr (LK == PTG_LIKE_POLY_CHI2) return flike_poly_fused<D>(m, x);
  else if constexpr (LK == PTG_LIKE_SINUSOID_CHI2) return flike_sinusoid_fused<D>(m, x);
  else return like_eval_kind<D, LK>(m, x);
}

// sines.hh:22-54 with the per-dimension constants staged in shared memory (spar: min, width or 1/width, k pi, k).  Expression for
// expression like_eval_kind<D, PTG_LIKE_SINES>; `pow2` (every width a power of two) replaces the division by the exact reciprocal product.
template <int D>
__device__ __forceinline__ double flike_sines_staged(const double *__restrict__ spar, bool pow2, const double *__restrict__ P, const double x[D]) {
  const double height = __ldg(P), step_scale = __ldg(P + 1);
  double lprod = 0; int isum = 0;
#pragma unroll 1
  for (int i = 0; i < D; i++) {
    double xi = x[0];
#pragma unroll
    for (int j = 1; j < D; j++) if (i == j) xi = x[j];
    const double mn = spar[4 * i], w = spar[4 * i + 1], kpi = spar[4 * i + 2], kd = spar[4 * i + 3];
    const double t = xi - mn;
    const double xx = pow2 ? t * w : t / w;
    double sv = sin(kpi * xx);
    sv = sv * sv;
    lprod += (sv * sv - 1) * height;
    isum += (int)(xx * kd);
  }
  const double result = lprod + (-isum * step_scale);
  return result;
}

// packed ladder statistics: st_di = dir (low 2 bits, biased by 1) | inst << 2 ; cnt[FC_UD] = ups delta | downs delta << 16 ;
// cnt[FC_SC] = swap_count delta | swap_accept delta << 16   (deltas per launch; the host keeps launches <= 16384 steps)
#define PTG_FAST_MAX_STEPS 16384

// ---- rung-sharded ladders, exchange over peer memory (PtgXchg, ptg_types.h) ---------------------------------------------------
// publish p: the lanes that hold this block's edge rungs write (x, llike, lprior, beta) into parity p & 1 of this rank's area, then
// raise the ladder's flag to p + 1
template <int D>
__device__ __forceinline__ void fx_publish(const PtgModel &m, const PtgXchg &xc, long long p, const FChain<D> &ch, double lprior, long long ladder, int rung, int R) {
  if (rung != 0 && rung != R - 1) return;
#pragma unroll 1
  for (int e = 0; e < 2; e++) {
    if (rung != (e == 0 ? 0 : R - 1)) continue; // R == 1: the one rung is both edges
    double *rec = xc.my_edges + (((size_t)(p & 1) * 2 + e) * m.n_ladders + ladder) * (D + 3);
#pragma unroll
    for (int k = 0; k < D; k++) rec[k] = ch.x[k];
    rec[D] = ch.llike; rec[D + 1] = lprior; rec[D + 2] = ch.beta;
    __threadfence_system();
    *((volatile int *)(xc.my_flags + (size_t)e * m.n_ladders + ladder)) = (int)(p + 1);
  }
}
// boundary trial of exchange p (ptg_boundary_swap_kernel: chain.cc:1459-1490 + add_state) against the neighbour's record, read from
// ITS memory over NVLink once its per-ladder flag says publish p is complete.  Returns the number of appends made (0 or 1 per edge).
// The wait is bounded by a WATCHDOG the host arms (xc.abort: a word in this rank's mapped memory the host sets when a launch overstays
// its deadline): an aborted wait raises PTG_ERR_XCHG_TIMEOUT, a sticky device error that stops this handle (ptg_step refuses further work).
template <int D, bool RF>
__device__ __forceinline__ int fx_swap(const PtgModel &m, const PtgState &s, const PtgXchg &xc, long long p, FChain<D> &ch, int chain, long long ladder, int gl,
                                       int rung, int R, int *cnt, double *mapp, int n_mh, int &err) {
  const bool edge_lo = (rung == 0) && xc.has_lo, edge_hi = (rung == R - 1) && xc.has_hi;
  if (!edge_lo && !edge_hi) return 0;
  int napp = 0;
  double *__restrict__ hb = s.hist + chain * ((long long)m.hist_cap * PTG_HX(D));
#pragma unroll 1
  for (int e = 0; e < 2; e++) {
    if (!(e == 0 ? edge_lo : edge_hi)) continue;
    // e = 0: my coldest rung against the colder neighbour's hottest (I am the pair's upper rung); e = 1: the mirror image
    const int their_edge = (e == 0) ? 1 : 0;
    const volatile int *flag = (e == 0 ? xc.lo_flags : xc.hi_flags) + (size_t)their_edge * m.n_ladders + ladder;
    long long spins = 0;
    while (*flag < (int)(p + 1)) {
      if ((++spins & 1023) == 0 && xc.abort && *((volatile int *)xc.abort)) { err = PTG_ERR_XCHG_TIMEOUT; break; }
      __nanosleep(64);
    }
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
      ch.llike = nb_ll; mapp[2] = nb_lprior;
      mapp[1] = nb_lprior + ch.beta * nb_ll; // lpost recomputed = lprior(x) + beta*llike (chain.cc:925-928)
    }
    fappend<D, RF>(m, s, ch, chain, hb, cnt, mapp, n_mh);
    napp++;
    if (i_am_lower) cnt[FC_SC] += 1 + (accept ? (1 << 16) : 0); // counted at the pair's lower rung
  }
  return napp;
}

// XCHG: 0 = plain; 1 = with the rung-boundary exchange in the prologue / epilogue (ptg_step_exchange); 2 = also inside the iteration loop.
// LK  : -1 = every feature at run time; >= 0 = streamlined configuration (see the head of this file) with likelihood kind LK.
// Separate instantiations: each form's extra code costs registers and stack, and the plain kernel carries none of it
// MAXT: largest CTA of the instantiation = its register budget (896 threads: 72 registers, one wave of 28 warps per SM; 448 threads: 144
// registers and no spills, for batches that put at most 14 warps on an SM, e.g. BASELINE config 5's 2048 ladders per GPU)
template <int D, int XCHG, int LK, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) ptg_fstep_kernel(const __grid_constant__ PtgModel m, PtgState s, long long step0, int n_steps, int W, const __grid_constant__ PtgXchg xc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // streamlined: the FS(run-time expression, folded value) flags below are constants.  The data chi-squared kinds only swap the functor
  // (their evaluation dwarfs every flag test) and keep all features at run time.
  constexpr bool SL = (LK >= 0) && LK != PTG_LIKE_POLY_CHI2 && LK != PTG_LIKE_SINUSOID_CHI2;
#define FS(expr, val) (SL ? (val) : (expr))
  constexpr int NPAIR = (D + 1) / 2;
  const int R = m.n_rungs, NP = m.n_props;
  typedef FShared<D> SH;
  double *const sbins = reinterpret_cast<double *>(smem_raw + SH::OFF_BINS);      // [R][NP]
  FProp *const sprop = reinterpret_cast<FProp *>(smem_raw + SH::OFF_PROP);
  double *const spar = reinterpret_cast<double *>(smem_raw + SH::OFF_SPAR);
  for (int i = threadIdx.x; i < R * NP; i += blockDim.x) sbins[i] = m.bins[i];
  for (int i = threadIdx.x; i < NP; i += blockDim.x) {
    const PtgProp &p = m.props[i];
    FProp q;
    q.snooker = p.snooker; q.g1frac = p.g1frac; q.gamma_std = p.gamma_std; q.reduce_gamma = p.reduce_gamma; q.ignore_frac = p.ignore_frac;
    q.unlikely_alpha = p.unlikely_alpha; q.one_d_frac = p.one_d_frac; q.kind = p.kind; q.has_transform = p.has_transform;
    q.sigma_off = p.sigma_off; q.trans_off = p.trans_off;
    sprop[i] = q;
  }
  bool sines_pow2 = true; // every width of the sines surface is a power of two: x / w == x * (1 / w) exactly
  if (LK == PTG_LIKE_SINES) {
    // per-dimension constants of sines.hh:22-54 in the order the functor uses them (like_eval_kind<D, PTG_LIKE_SINES>)
    const double *__restrict__ P = m.lparams;
#pragma unroll
    for (int i = 0; i < D; i++) {
      const double w = __ldg(P + 2 + 2 * D + i) - __ldg(P + 2 + D + i);
      int e; const double mant = frexp(w, &e);
      if (!(mant == 0.5 && e > -1000 && e < 1000)) sines_pow2 = false;
    }
    if (threadIdx.x < D) {
      const int i = threadIdx.x;
      const int k = (int)__ldg(P + 2 + i);
      const double mn = __ldg(P + 2 + D + i), w = __ldg(P + 2 + 2 * D + i) - mn;
      spar[4 * i] = mn; spar[4 * i + 1] = sines_pow2 ? 1.0 / w : w; spar[4 * i + 2] = k * PTG_PI; spar[4 * i + 3] = (double)k;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  unsigned char *const wblock = smem_raw + SH::OFF_WARP + (threadIdx.x >> 5) * SH::W_BYTES;     // this warp's block
  double *const mapp = reinterpret_cast<double *>(wblock + SH::W_MAP) + 3 * lane;               // [0] running MAP, [1] current lpost, [2] current lprior
  int *const cnt = reinterpret_cast<int *>(wblock + SH::W_CNT) + PTG_FC_STRIDE * lane;          // this thread's row
#define LPOST (mapp[1])
#define LPRIOR (mapp[2])
  double *const wlus = reinterpret_cast<double *>(wblock + SH::W_LUS);
  double *const wpool = reinterpret_cast<double *>(wblock + SH::W_POOL);
  unsigned char *const witems = wblock + SH::W_ITEMS;

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
#define stream_base ((uint64_t)(gl - g) * PTG_STREAM_STRIDE)
#define my_stream ((uint64_t)gl * PTG_STREAM_STRIDE + (uint64_t)rung)
#define ladder_stream ((uint64_t)gl * PTG_STREAM_STRIDE + PTG_STREAM_LADDER)
#define hbase (s.hist + chain * ((long long)m.hist_cap * PTG_HX(D)))
#define step ((uint64_t)(step0 + it))
  const double *bins = sbins + (rung < R ? rung : 0) * NP;

  FChain<D> ch;
#define st_di (cnt[FC_DI])
#pragma unroll
  for (int k = 0; k < FC_COUNT; k++) cnt[k] = 0;
  st_di = 1;
  if (active) {
#pragma unroll
    for (int k = 0; k < D; k++) ch.x[k] = s.cur_x[(long long)k * m.n_chains + chain];
    LPOST = s.lpost[chain]; ch.llike = s.llike[chain]; LPRIOR = s.lprior[chain]; ch.beta = s.beta[chain];
    const long long nsize = s.nsize[chain];
    cnt[FC_SLOT] = (int)(nsize % m.hist_cap);
    cnt[FC_HFILL] = (int)(nsize > m.hist_cap ? (long long)m.hist_cap : nsize);
    cnt[FC_SSAVE] = (int)(s.nhist[chain] % m.save_every);
    cnt[FC_LAST_TYPE] = s.last_type[chain];
    st_di = (s.directions[chain] + 1) | (s.instances[chain] << 2);
    mapp[0] = s.map_lpost[chain];
  } else {
#pragma unroll
    for (int k = 0; k < D; k++) ch.x[k] = 0;
    ch.llike = 0; ch.beta = 1;
    cnt[FC_SLOT] = 0; cnt[FC_HFILL] = 0; cnt[FC_SSAVE] = 0;
    mapp[0] = 0; LPOST = 0; LPRIOR = 0;
  }
  cnt[FC_SS0] = cnt[FC_SSAVE]; // nhist % save_every at the start of the launch
  int err = 0;

  // ================================================================= rung-sharded ladders: pending cross-GPU boundary swap
  long long xp = xc.index; // running publish index (XCHG instantiation only)
  if (XCHG && xc.on && xc.swap_in && active) cnt[FC_XAPP] += fx_swap<D, !SL>(m, s, xc, xp - 1, ch, chain, ladder, gl, rung, R, cnt, mapp, 0, err);

  const int maxswaps = m.maxswaps;
  const double swap_thresh = (R - 1) * m.swap_rate / maxswaps; // chain.cc:1413
  double ptry = 2 * m.swap_rate; if (ptry > 1) ptry = 1;
  const bool zero_valid = FS(m.zero_valid != 0, true);
  const bool ref_mode = FS(m.swap_mode == PTG_SWAP_REFERENCE, true);
  // Without temperature evolution the trials of one iteration that share no rung commute: they run as ONE batch and their history
  // appends are deferred to the iteration's single append.  pry_temps (chain.cc:1809-1846) changes every rung's beta / lpost after
  // each accepted trial, so with evolution the trials run one by one and append at once, as the reference does.
  const bool batched = FS(!(m.evolve_rate > 0), true);
  // lanes that own a swap-trial draw of their ladder (trial j = lane j of the ladder's group)
  const bool has_swap_item = ref_mode && ladder_ok && R > 1 && rung < maxswaps;
  const unsigned swapmask = __ballot_sync(0xffffffffu, has_swap_item);
  const int nfree = 32 - __popc(swapmask), frank = __popc(~swapmask & ((1u << lane) - 1u));
  // the trial draws of iteration 0 (later iterations get theirs from pool round 1 of the iteration before)
  if (has_swap_item) {
    uint32_t q[4];
    ptg_philox_draw(m.seed, ladder_stream, PTG_DOMAIN_STEP, (uint64_t)step0, (uint32_t)rung, q);
    int raw = -2; double logu = 0;
    if (ptg_u32_to_unit(q[0]) < swap_thresh) { raw = (int)(ptg_u32_to_unit(q[1]) * (R - 1)); logu = log(ptg_u52_to_unit(q[2], q[3])); }
    cnt[FC_RAW] = raw; cnt[FC_LOGU_LO] = __double2loint(logu); cnt[FC_LOGU_HI] = __double2hiint(logu);
  } else cnt[FC_RAW] = -2;

  for (int it = 0; it < n_steps; it++) {
    bool swapped = false;   // took part in a swap trial: no MH update this iteration (chain.cc:1554-1557)
    bool pending = false;   // owes the history append of a swap trial (made at the end of the iteration)
    const int n_mh = it - cnt[FC_NSWAPPED]; // MH updates of this chain so far in this launch

    // ================================================================= swap phase (chain.cc:1410-1538)
    if (ladder_ok && R > 1) {
      if (ref_mode) {
        // lane j holds trial j: candidate pair and log of its swap draw, from block j of the ladder's stream
        const int raw = cnt[FC_RAW];
        const double logu = __hiloint2double(cnt[FC_LOGU_HI], cnt[FC_LOGU_LO]);
        // serial de-dup (iswaps[j]==cand or iswaps[j]+1==cand for an earlier surviving j) over the trials that drew a pair
        unsigned used = 0, live = 0; // live: bit j = trial j survives
        unsigned cand = (__ballot_sync(gm, raw >= 0) & gm) >> (g * W);
        while (cand) {
          const int i = __ffs(cand) - 1;
          cand &= cand - 1;
          const int c = __shfl_sync(gm, raw, i, W);
          if (!((used >> c) & 1u) && !(c > 0 && ((used >> (c - 1)) & 1u))) { used |= 1u << c; live |= 1u << i; }
        }
        if (batched) {
          // A surviving trial on pair c can only share a rung with an EARLIER surviving trial on pair c+1 (the de-dup removed the other
          // overlaps).  Round by round, every live trial whose pair c has no live trial on c+1 runs; all pairs of a round are disjoint.
          while (live) {
            const bool mine = ((live >> rung) & 1u) != 0;                            // this lane's trial is live
            const unsigned pairs_live = __reduce_or_sync(gm, mine ? (1u << raw) : 0u);
            const bool now = mine && !((pairs_live >> (raw + 1)) & 1u);
            const unsigned pairs = __reduce_or_sync(gm, now ? (1u << raw) : 0u);     // bit c: pair (c, c+1) is tried in this round
            live &= ~((__ballot_sync(gm, now) & gm) >> (g * W));
            if (now) wlus[g * W + raw] = logu;                                       // hand log u to the pair's lower rung
            __syncwarp(gm);                                                          // (ladder groups of one warp loop independently)
            const bool is_lo = ((pairs >> rung) & 1u) != 0, is_hi = rung > 0 && ((pairs >> (rung - 1)) & 1u) != 0, involved = is_lo || is_hi;
            const double lu = is_lo ? wlus[g * W + rung] : 0.0;
            // every lane r evaluates pair (r, r+1); the lower rung decides
            const double ll_up = __shfl_down_sync(gm, ch.llike, 1, W), b_up = __shfl_down_sync(gm, ch.beta, 1, W);
            double lla = ch.llike; if (!(lla > -1e200)) lla = -1e200;
            double llb = ll_up; if (!(llb > -1e200)) llb = -1e200;
            const double lhr = -(b_up - ch.beta) * (llb - lla);
            int acc_mine = is_lo ? 1 : 0;
            if (is_lo && lhr < 0) acc_mine = (lu < lhr) ? 1 : 0;
            const int acc_below = __shfl_up_sync(gm, acc_mine, 1, W);
            const bool accept = is_lo ? (acc_mine != 0) : (is_hi && acc_below != 0);
            if (is_lo && rung > 0) { // ups / downs of the lower rung before the exchange (chain.cc:1440-1443)
              const int dir = (st_di & 3) - 1;
              if (dir > 0) cnt[FC_UD] += 1;
              if (dir < 0) cnt[FC_UD] += 1 << 16;
            }
            // a rung already in an earlier trial of this iteration (H4) stores that trial's state before it changes again
            if (involved && pending && active) { fappend<D, !SL>(m, s, ch, chain, hbase, cnt, mapp, n_mh); cnt[FC_XAPP]++; pending = false; }
            const int partner = is_lo ? rung + 1 : (is_hi ? rung - 1 : rung);
            {
#pragma unroll
              for (int k = 0; k < D; k++) { const double v = __shfl_sync(gm, ch.x[k], partner, W); if (accept) ch.x[k] = v; }
              { const double v = __shfl_sync(gm, ch.llike, partner, W); if (accept) ch.llike = v; }
              { const int v = __shfl_sync(gm, st_di, partner, W); if (accept) st_di = v; }
              double nlp = 0;
              if (accept) nlp = mapp[3 * (partner - rung) + 2];                       // the partner's lprior slot
              __syncwarp(gm);
              if (accept) { LPRIOR = nlp; LPOST = nlp + ch.beta * ch.llike; }         // lpost recomputed = lprior(x) + beta*llike (chain.cc:925-928)
            }
            if (accept) {
              if (is_lo && rung == 0) st_di = (st_di & ~3) | 2;          // directions[0] = +1
              if (is_hi && rung == R - 1) st_di = (st_di & ~3) | 0;      // directions[R-1] = -1
            }
            if (is_lo) cnt[FC_SC] += 1 + (accept ? (1 << 16) : 0);
            if (involved) { if (active) pending = true; swapped = true; }
          }
        } else
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
            { const double v = __shfl_sync(gm, LPRIOR, partner, W); LPRIOR = v; }
            if (involved) LPOST = LPRIOR + ch.beta * ch.llike; // lpost recomputed = lprior(x) + beta*llike (chain.cc:925-928)
          }
          if (involved) {
            if (active) { fappend<D, !SL>(m, s, ch, chain, hbase, cnt, mapp, n_mh); if (swapped) cnt[FC_XAPP]++; }
            swapped = true;
          }
          if (accept) {
            const int partner = is_lo ? c + 1 : (is_hi ? c : rung);
            { const int v = __shfl_sync(gm, st_di, partner, W); st_di = v; }
            if (c == 0 && is_lo) st_di = (st_di & ~3) | 2;          // directions[0] = +1
            if (c + 1 == R - 1 && is_hi) st_di = (st_di & ~3) | 0;  // directions[R-1] = -1
            if (is_lo) cnt[FC_SC] += 1 << 16;
            {
              // pry_temps, one pried gap (chain.cc:1809-1846) + resetTemp (chain.cc:1088-1091), reference summation order
              const double rate = m.evolve_rate;
              const double lp_mine = LPOST;
              const double b_next = __shfl_down_sync(gm, ch.beta, 1, W), lp_next = __shfl_down_sync(gm, lp_mine, 1, W);
              double sp = ch.beta - b_next;
              if (m.evolve_lpost_cut >= 0 && lp_mine - lp_next > m.evolve_lpost_cut * ch.beta) sp *= (1.0 + rate);
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
              if (rung >= 1 && rung < R - 1) { ch.beta = mine; LPOST = LPRIOR + mine * ch.llike; }
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
        const double nll = __shfl_sync(gm, ch.llike, partner, W), nlp = __shfl_sync(gm, LPRIOR, partner, W);
        const int ndi = __shfl_sync(gm, st_di, partner, W);
        if (tried && accept) {
#pragma unroll
          for (int k = 0; k < D; k++) ch.x[k] = nx[k];
          ch.llike = nll; LPRIOR = nlp;
          LPOST = nlp + ch.beta * ch.llike;
          st_di = ndi;
          if (rung == 0) st_di = (st_di & ~3) | 2;
          if (rung == R - 1) st_di = (st_di & ~3) | 0;
        }
        if (tried) { if (active) pending = true; swapped = true; }
      }
    }
    __syncwarp();
    if (swapped) cnt[FC_NSWAPPED]++;

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
    if (FS(m.wrap_in_set, 1)) {
      const double x = (NP > 1) ? ptg_u32_to_unit(wA[0]) : 0.0;
      if (cnt[FC_HFILL] >= D * 10 || !do_mh) { // every member ready (differential_evolution::is_ready, proposal_distribution.hh:407)
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
    double newx[D];
    double prop_lh = 0;
    int type = 0;
    bool valid = zero_valid;

    // ---- differential evolution, part 1: history indices (proposal_distribution.cc:744-778); the records are requested from L2 now
    //      and read after pool round 1, whose work covers their DRAM latency
    const bool is_de = (kind == PTG_PROP_DE);
    bool snooker = false;
    int i1 = 0, i2 = 0, iz = 0, az = 0;
    if (is_de) {
      const int hsize = cnt[FC_HFILL];
      int start = 0;
      if ((hsize - D * 100) * (1 - p.ignore_frac) > D * 10) start = (int)((hsize - D * 100) * p.ignore_frac);
      snooker = p.snooker > ptg_u32_to_unit(wA[1]);
      const bool slow = FS(p.unlikely_alpha > 0, false);
      if (!slow) {
        i1 = (int)(start + (hsize - start) * ptg_u32_to_unit(wA[3]));
        i2 = (int)(start + (hsize - start) * ptg_u32_to_unit(wB[0]));
        iz = (int)(start + (hsize - start) * ptg_u32_to_unit(wB[1]));
        az = 1;
      } else {
        const double mapl = mapp[0];
        const double *lpb = s.hist_lp + 2 * (chain * (long long)m.hist_cap);
        if (snooker) { const int2 r = fde_index_slow<D>(m, cnt[FC_HFILL], cnt[FC_SLOT], lpb, mapl, m.seed, my_stream, step, p.ignore_frac, p.unlikely_alpha, wB[1], 0, 0); iz = r.x; az = r.y; }
        i1 = fde_index_slow<D>(m, cnt[FC_HFILL], cnt[FC_SLOT], lpb, mapl, m.seed, my_stream, step, p.ignore_frac, p.unlikely_alpha, wA[3], 1, 0).x;
        i2 = fde_index_slow<D>(m, cnt[FC_HFILL], cnt[FC_SLOT], lpb, mapl, m.seed, my_stream, step, p.ignore_frac, p.unlikely_alpha, wB[0], 2, 0).x;
      }
      i1 = fslot<D>(m, cnt, i1); i2 = fslot<D>(m, cnt, i2); // physical slots from here on
#if defined(PTG_F_PREFETCH_L1)
      asm volatile("prefetch.global.L1 [%0];" :: "l"(hbase + (long long)i1 * PTG_HX(D)));
      asm volatile("prefetch.global.L1 [%0];" :: "l"(hbase + (long long)i2 * PTG_HX(D)));
      if (snooker) asm volatile("prefetch.global.L1 [%0];" :: "l"(hbase + (long long)fslot<D>(m, cnt, iz) * PTG_HX(D)));
#elif !defined(PTG_F_NO_PREFETCH)
      asm volatile("prefetch.global.L2 [%0];" :: "l"(hbase + (long long)i1 * PTG_HX(D)));
      asm volatile("prefetch.global.L2 [%0];" :: "l"(hbase + (long long)i2 * PTG_HX(D)));
      if (snooker) asm volatile("prefetch.global.L2 [%0];" :: "l"(hbase + (long long)fslot<D>(m, cnt, iz) * PTG_HX(D)));
#endif
    }

#ifndef PTG_F_NO_STASH
    // values that are only needed again after pool round 1 wait in the thread's shared-memory row, not in registers (the round is the
    // register-hungriest stretch of the iteration: Philox + log + sincospi)
    {
      volatile int *vk = cnt;
      vk[FC_K0] = i1; vk[FC_K1] = i2; vk[FC_K2] = iz; vk[FC_K3] = (int)wA[2]; vk[FC_K4] = (int)wB[2]; vk[FC_K5] = (int)wB[3];
    }
#endif
    // ---- pool round 1: Box-Muller pairs of the Gaussian-member lanes + the next iteration's swap-trial draws
    // A Gaussian lane needs all NPAIR pairs, or -- for a single-axis step (proposal_distribution.hh:197-205) -- only the pair of its axis.
    const bool is_gauss = (kind == PTG_PROP_GAUSS);
    int axis = -1; // >= 0: single-axis step
    if (is_gauss && p.one_d_frac > 0 && ptg_u32_to_unit(wA[1]) < p.one_d_frac) axis = (int)(D * ptg_u32_to_unit(wA[2]));
    {
      const unsigned many = __ballot_sync(0xffffffffu, is_gauss), mfull = __ballot_sync(0xffffffffu, is_gauss && axis < 0);
      const unsigned lt = (1u << lane) - 1u;
      const int total = __popc(many) + (NPAIR - 1) * __popc(mfull);
      if (is_gauss) {
        const int off = __popc(many & lt) + (NPAIR - 1) * __popc(mfull & lt);
        if (axis >= 0) witems[off] = (unsigned char)((lane << 3) | (axis >> 1));
        else {
#pragma unroll
          for (int q = 0; q < NPAIR; q++) witems[off + q] = (unsigned char)((lane << 3) | q);
        }
      }
      __syncwarp();
      // round 0: trial lanes take their own draw, the other lanes the first `nfree` pairs; later rounds: 32 pairs each
      for (int base = 0, first = 1; first || base < total; first = 0) {
        int typ = 0, k = -1;
        if (first) { if (has_swap_item) typ = 1; else k = frank; }
        else k = base + lane;
        if (k >= 0 && k < total) typ = 2;
        if (typ) {
          uint64_t stream = ladder_stream; uint64_t stp = step + 1; uint32_t blk = (uint32_t)rung;
          int owner = 0, pr = 0;
          if (typ == 2) {
            const int code = witems[k];
            owner = code >> 3; pr = code & 7;
            // stream of the owner's chain: (global ladder)*128 + rung, with ladder = warp's first ladder + owner / W
            stream = stream_base + (uint64_t)(owner / W) * PTG_STREAM_STRIDE + (uint64_t)(owner % W);
            stp = step; blk = PTG_BLK_NORMAL + pr;
          }
          uint32_t q[4];
          ptg_philox_draw(m.seed, stream, PTG_DOMAIN_STEP, stp, blk, q);
          bool want = true;
          double ul = ptg_u52_to_unit(q[0], q[1]);
          if (typ == 1) {
            want = ptg_u32_to_unit(q[0]) < swap_thresh;
            ul = ptg_u52_to_unit(q[2], q[3]);
            cnt[FC_RAW] = want ? (int)(ptg_u32_to_unit(q[1]) * (R - 1)) : -2;
          }
          if (want) {
            const double lg = log(ul);
            if (typ == 1) { cnt[FC_LOGU_LO] = __double2loint(lg); cnt[FC_LOGU_HI] = __double2hiint(lg); }
            else { // box_muller (ptg_device.cuh): z0 = r cos, z1 = r sin
              const double r = sqrt(-2.0 * lg);
              double sn, cs;
              sincospi(2.0 * ptg_u52_to_unit(q[2], q[3]), &sn, &cs);
              wpool[(2 * pr) * 32 + owner] = r * cs;
              wpool[(2 * pr + 1) * 32 + owner] = r * sn;
            }
          }
        }
        base += first ? nfree : 32;
      }
      __syncwarp();
    }

#ifndef PTG_F_NO_STASH
    {
      volatile int *vk = cnt;
      i1 = vk[FC_K0]; i2 = vk[FC_K1]; iz = vk[FC_K2]; wA[2] = (uint32_t)vk[FC_K3]; wB[2] = (uint32_t)vk[FC_K4]; wB[3] = (uint32_t)vk[FC_K5];
    }
#endif
    // ---- Gaussian members: x' = x + M (z o sigma)  (proposal_distribution.hh:194-218)
#pragma unroll
    for (int i = 0; i < D; i++) newx[i] = ch.x[i];
    if (is_gauss) {
      double z[D];
      const double *__restrict__ sig = m.prop_data + p.sigma_off;
#pragma unroll
      for (int i = 0; i < D; i++) {
        z[i] = 0.0;
        if (axis < 0 || axis == i) z[i] = wpool[i * 32 + lane] * __ldg(sig + i) + 0.0;
      }
      if (axis >= 0) type = 1;
      if (FS(p.has_transform, 0)) {
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
    __syncwarp(); // the pool slots are reused by round 2
    // ---- differential evolution, part 2 (proposal_distribution.cc:489-591)
    double sn_a = 1, sn_b = 1; // arguments of the two snooker logarithms
    if (is_de) {
      const double ug = ptg_u32_to_unit(wA[2]);
      // gamma of draw_standard (proposal_distribution.cc:495-499) or draw_snooker (:541-543); both proposals use the two products
      // s1 * gamma and s2 * (-gamma), formed as soon as each record arrives
      double gamma = p.gamma_std;
      if (ug < p.g1frac) gamma = 1;
      if (snooker) gamma = (1.2 + ug) / p.reduce_gamma;
      double pa[D], pb[D];
      {
        double a[D];
        fload_rec<D>(hbase + (long long)i1 * PTG_HX(D), a);
#pragma unroll
        for (int i = 0; i < D; i++) pa[i] = a[i] * gamma;
      }
      {
        double b[D];
        fload_rec<D>(hbase + (long long)i2 * PTG_HX(D), b);
#pragma unroll
        for (int i = 0; i < D; i++) pb[i] = b[i] * (-gamma);
      }
      if (!snooker) {
        // draw_standard: prop = (s + gamma s1) + (-gamma s2); the jitter drawn by the reference is discarded (H8-1)
#pragma unroll
        for (int i = 0; i < D; i++) {
          const double t = ch.x[i] + pa[i];
          newx[i] = t + pb[i];
        }
      } else {
        // draw_snooker (proposal_distribution.cc:538-591)
        double smznorm2 = 0, minusz[D], smz[D];
        int isafe = 0;
        while (true) {
          double zz[D];
          fload_rec<D>(hbase + (long long)fslot<D>(m, cnt, iz) * PTG_HX(D), zz);
          smznorm2 = 0;
#pragma unroll
          for (int i = 0; i < D; i++) { minusz[i] = zz[i] * (-1); smz[i] = ch.x[i] + minusz[i]; }
#pragma unroll
          for (int i = 0; i < D; i++) smznorm2 += smz[i] * smz[i];
          if (smznorm2 != 0 || ++isafe > 1000) break;
          const int2 r = fde_index_slow<D>(m, cnt[FC_HFILL], cnt[FC_SLOT], s.hist_lp + 2 * (chain * (long long)m.hist_cap), mapp[0], m.seed, my_stream, step, p.ignore_frac, p.unlikely_alpha, wB[1], 0, az);
          iz = r.x; az = r.y;
        }
        double dot = 0;
#pragma unroll
        for (int i = 0; i < D; i++) {
          const double ds12 = pa[i] + pb[i];
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
        sn_a = pmz2; sn_b = smznorm2; // prop_lh = (log(pmz2) - log(smznorm2)) (D - 1) / 2, logarithms taken in pool round 2
        type = 1;
      }
    } else if (FS(kind == PTG_PROP_PRIOR_DRAW, false)) {
      FVec<D> xv;
#pragma unroll
      for (int i = 0; i < D; i++) xv.v[i] = ch.x[i];
      const FVecFlag<D> r = fprior_member<D>(m, my_stream, step, xv);
#pragma unroll
      for (int i = 0; i < D; i++) newx[i] = r.v[i];
      prop_lh = r.aux; valid = r.ok != 0;
    }
    if (FS(m.wrap_in_set, 1)) type = member + 10 * type;

    // ---- pool round 2: log of the Metropolis draw on every MH lane; the snooker lanes' two logarithms on the lanes without an MH update
    double log_uacc = 0;
    {
      const unsigned msn = __ballot_sync(0xffffffffu, snooker), midle = __ballot_sync(0xffffffffu, !do_mh);
      const unsigned lt = (1u << lane) - 1u;
      const int nitem = 2 * __popc(msn), nidle = __popc(midle), irank = __popc(midle & lt);
      if (snooker) { const int o = 2 * __popc(msn & lt); wpool[o] = sn_a; wpool[o + 1] = sn_b; }
      __syncwarp();
      for (int base = 0, first = 1; first || base < nitem; first = 0) {
        int k = -1;
        if (first) { if (!do_mh) k = irank; }
        else k = base + lane;
        const bool item = (k >= 0 && k < nitem);
        if (item || (first && do_mh)) {
          const double arg = item ? wpool[k] : ptg_u52_to_unit(wB[2], wB[3]);
          const double lg = log(arg);
          if (item) wpool[k] = lg; else log_uacc = lg;
        }
        base += first ? nidle : 32;
        __syncwarp();
      }
      if (snooker) { const int o = 2 * __popc(msn & lt); prop_lh = (wpool[o] - wpool[o + 1]) * (D - 1) / 2.0; }
      __syncwarp();
    }

    // ---- enforce, prior, gated likelihood (chain.cc:976-987)
    if (FS(m.any_bound, 0) && valid) {
      FVec<D> xv;
#pragma unroll
      for (int i = 0; i < D; i++) xv.v[i] = newx[i];
      const FVecFlag<D> r = fenforce<D>(m, xv);
#pragma unroll
      for (int i = 0; i < D; i++) newx[i] = r.v[i];
      valid = r.ok != 0;
    }
    double newlprior;
    if (FS(m.all_uniform_prior, 1)) {
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
    const double cur_lpost = LPOST;
    // chain.cc:980.  A proposal outside the prior's support has newlprior = -inf and newlprior - oldlprior > dprior_min is false for every
    // finite oldlprior, so with an all-uniform prior (every current state inside the box) the second clause never opens the gate
    const bool gate = valid && ((newlprior > -1e200) || FS(newlprior - (cur_lpost - ch.beta * ch.llike) > m.dprior_min, false));
    if constexpr (LK == PTG_LIKE_POLY_CHI2 || LK == PTG_LIKE_SINUSOID_CHI2) {
      // every lane takes part: the data sums of the chains that passed the gate are spread over all 32 lanes (one ladder per warp)
      __syncwarp();
      double v;
      if (W == 32) v = flike_data_compact<D, LK>(m, newx, gate && do_mh);
      else v = (gate && do_mh) ? flike<D, LK>(m, newx) : 0.0;                 // ladders share the warp (fused exchange launches): one lane per chain
      if (gate && do_mh) { newlike = v; newlpost = newlike * ch.beta + newlprior; } else code |= PTG_TRACE_NOLIKE;
    } else if (gate && do_mh) {
      if constexpr (LK == PTG_LIKE_SINES) newlike = flike_sines_staged<D>(spar, sines_pow2, m.lparams, newx);
      else newlike = flike<D, LK>(m, newx);
      newlpost = newlike * ch.beta + newlprior;
    } else code |= PTG_TRACE_NOLIKE;
    // ---- Metropolis test (chain.cc:989-1001)
    double lhr = prop_lh;
    if (isnan(lhr)) accept = false;
    lhr += newlpost - cur_lpost;
    if (!valid) { accept = false; code |= PTG_TRACE_INVALID; }
    if (accept && lhr < 0) accept = (log_uacc < lhr);
    if (do_mh) {
      if (accept) {
        cnt[FC_NACCEPT]++;
        cnt[FC_LAST_TYPE] = type;
#pragma unroll
        for (int i = 0; i < D; i++) ch.x[i] = newx[i];
        ch.llike = newlike; LPOST = newlpost; LPRIOR = newlprior;
        code |= PTG_TRACE_ACCEPT;
      }
    }
    // the iteration's history append: the MH lanes' new state, the swapped lanes' pending trial state
    if (do_mh || pending) fappend<D, !SL>(m, s, ch, chain, hbase, cnt, mapp, n_mh + (do_mh ? 1 : 0));
    if (FS(active && (long long)step < m.trace_steps, false)) {
      s.trace_lhr[step * m.n_chains + chain] = do_mh ? lhr : 0.0;
      s.trace_code[step * m.n_chains + chain] = do_mh ? (code | (type & PTG_TRACE_TYPE_MASK)) : PTG_TRACE_SWAPPED;
    }
    // in-launch exchange: after every xc.every-th iteration (not the launch's last, whose publish the epilogue does) publish the
    // edges and run the boundary trial as soon as the neighbour's same ladder has published -- warps wait one by one, the rest of
    // the SM keeps stepping.  The host only uses this when every CTA of the grid is resident (one wave on each GPU).
    if (XCHG == 2 && xc.on && xc.every > 0 && it + 1 < n_steps && (it + 1) % xc.every == 0) {
      if (active) {
        fx_publish<D>(m, xc, xp, ch, LPRIOR, ladder, rung, R);
        cnt[FC_XAPP] += fx_swap<D, !SL>(m, s, xc, xp, ch, chain, ladder, gl, rung, R, cnt, mapp, it + 1 - cnt[FC_NSWAPPED], err);
      }
      xp++;
      __syncwarp();
    }
  }

  // ================================================================= rung-sharded ladders: publish this block's edge rungs
  if (XCHG && xc.on && xc.publish_out && active) fx_publish<D>(m, xc, xp, ch, LPRIOR, ladder, rung, R);

  if (active) {
#pragma unroll
    for (int k = 0; k < D; k++) s.cur_x[(long long)k * m.n_chains + chain] = ch.x[k];
    s.lpost[chain] = LPOST; s.llike[chain] = ch.llike; s.lprior[chain] = LPRIOR; s.beta[chain] = ch.beta;
    s.map_lpost[chain] = mapp[0];
    // appends of this launch: one per iteration + the extra ones (second trial of a step, boundary exchanges);
    // saves = appends k in [0, dnhist) with (since_save0 + k) % save_every == 0
    const int dnhist = n_steps + cnt[FC_XAPP], se = m.save_every;
    const int since_save0 = cnt[FC_SS0];
    const int dnsize = (since_save0 + dnhist + se - 1) / se - (since_save0 + se - 1) / se;
    s.nhist[chain] += dnhist; s.nsize[chain] += dnsize; s.ntries[chain] += n_steps - cnt[FC_NSWAPPED]; s.naccept[chain] += cnt[FC_NACCEPT];
    s.last_type[chain] = cnt[FC_LAST_TYPE];
    s.directions[chain] = (st_di & 3) - 1; s.instances[chain] = st_di >> 2;
    const int st_ud = cnt[FC_UD], st_sc = cnt[FC_SC];
    s.ups[chain] += st_ud & 0xffff; s.downs[chain] += st_ud >> 16;
    s.swap_count[chain] += st_sc & 0xffff; s.swap_accept[chain] += st_sc >> 16;
    if (err) atomicMax(s.err, err);
  }
#undef FS
#undef st_di
#undef LPOST
#undef LPRIOR
}
#undef stream_base
#undef my_stream
#undef ladder_stream
#undef hbase
#undef step
