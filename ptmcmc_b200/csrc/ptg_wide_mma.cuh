// ptg_wide_mma.cuh -- production variant of the warp-per-chain step kernel (Philox draws): the two dense contractions of
// the d = 100 path are batched over the ladder's chains and run on the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64):
//     T = Off * M^T     eigen-rotation of the Gaussian proposal offsets   (proposal_distribution.hh:212)
//     Y = X'  * C^T     y_i = sum_j Cinv[i][j] x'_j of the quadratic form (cython/exampleGaussian.py:103-109)
// with the CTA's chains as the rows of an [RP x dim] operand in shared memory (RP = n_rungs rounded up to 8) and the
// dim x dim matrix streamed once per CTA per step from L1/L2 instead of once per chain.  tcgen05 has no fp64 kind:
// on sm_100a the fp64 tensor path IS mma.sync (SASS: DMMA).  Warp-level sums use shuffle trees here (Philox runs are
// compared with the oracle to 1e-8, not bit for bit; the exact-order kernel of ptg_wide.cuh is the tape-parity path).
#pragma once
#include "ptg_wide.cuh"

// Out[r][n] = sum_k In[r][k] * Cm[n*D + k]  for r < RP, n < D ; In / Out are [RP][DP] shared-memory matrices whose
// columns >= D are zero.  Called by every warp of the CTA.
__device__ __forceinline__ void xcta_dmma(const double *In, const double *__restrict__ Cm, double *Out, int RP, int D, int DP, int warp, int nwarps, int lane) {
  const int mt_n = RP >> 3, nt_n = (D + 7) >> 3;
  const int g = lane >> 2, t = lane & 3;
  // Columns n >= D of the last tile are computed from a clamped matrix row and never read (consumers stop at D), which keeps
  // the inner loop free of predicates.  When D is a multiple of 4 the k index is consumed in a permuted order -- within each
  // group of 8, lane t takes elements (2t, 2t+1) for two consecutive MMAs instead of (t, 4+t) -- so that both operands come
  // from one 16-byte load each; A and B use the same permutation, so the contraction is unchanged.
  const bool vec = (D & 3) == 0;
  const int K8 = vec ? (D & ~7) : 0;
  for (int tile = warp; tile < mt_n * nt_n; tile += nwarps) {
    const int mt = tile / nt_n, nt = tile - mt * nt_n;
    const double *arow = In + (size_t)(mt * 8 + g) * DP;
    int n = nt * 8 + g;
    if (n >= D) n = D - 1;
    const double *__restrict__ brow = Cm + (size_t)n * D;
    double c0 = 0, c1 = 0;
    for (int k0 = 0; k0 < K8; k0 += 8) {
      const double2 a2 = *reinterpret_cast<const double2 *>(arow + k0 + 2 * t);
      const double2 b2 = __ldg(reinterpret_cast<const double2 *>(brow + k0 + 2 * t));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a2.x), "d"(b2.x));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a2.y), "d"(b2.y));
    }
    for (int k0 = K8; k0 < D; k0 += 4) { // tail (and the whole contraction when D is not a multiple of 4): scalar operands, guarded
      const int k = k0 + t;
      const double a = arow[k];                       // shared-memory columns >= D are zero
      const double b = (k < D) ? __ldg(brow + k) : 0.0;
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    }
    double *orow = Out + (size_t)(mt * 8 + g) * DP + nt * 8 + 2 * t;
    orow[0] = c0; orow[1] = c1;
  }
}

// Swap phase of one ladder on the 32 lanes of warp 0 (Philox draws): lane r holds rung r's scalars and the INDEX of the state
// it carries; the trial logic is the shuffle version of ptg_fast.cuh (every lane prepares its own trial, bit-mask de-dup,
// one shuffle round per surviving trial, shuffle-scan pry_temps).  Reads the published scalars, writes the same outcome
// arrays as the serial leader of ptg_wide.cuh (perm, napp, app_*, n_lpost, n_beta, ladder statistics).
__device__ __forceinline__ void xswap_warp(const PtgModel &m, XShared &L, uint64_t ladder_stream, uint64_t step, int lane, int maxswaps, double swap_thresh,
                                           double ptry) {
  const int R = m.n_rungs;
  const unsigned gm = 0xffffffffu;
  const bool act = lane < R;
  double ll = act ? L.sll[lane] : 0.0, lprior = act ? L.slprior[lane] : 0.0, beta = act ? L.sbeta[lane] : 1.0, lpost = act ? L.slpost[lane] : 0.0;
  int perm = lane, napp = 0, as0 = 0, as1 = 0;
  double al0 = 0, al1 = 0, ab0 = 0, ab1 = 0;
  int dir = act ? L.dir[lane] : 0, inst = act ? L.inst[lane] : 0, ups = 0, downs = 0, sc = 0, sa = 0;
  auto record = [&]() {
    if (napp == 0) { as0 = perm; al0 = lpost; ab0 = beta; } else if (napp == 1) { as1 = perm; al1 = lpost; ab1 = beta; }
    napp++;
  };
  if (m.swap_mode == PTG_SWAP_REFERENCE) {
    int raw = -2;
    double logu = 0;
    if (lane < maxswaps) {
      uint32_t q[4];
      ptg_philox_draw(m.seed, ladder_stream, PTG_DOMAIN_STEP, step, (uint32_t)lane, q);
      if (ptg_u32_to_unit(q[0]) < swap_thresh) {
        raw = (int)(ptg_u32_to_unit(q[1]) * (R - 1));
        logu = log(ptg_u52_to_unit(q[2], q[3]));
      }
    }
    unsigned used = 0, live = 0;
    unsigned cand = __ballot_sync(gm, raw >= 0);
    while (cand) {
      const int i = __ffs(cand) - 1;
      cand &= cand - 1;
      const int c = __shfl_sync(gm, raw, i);
      if (!((used >> c) & 1u) && !(c > 0 && ((used >> (c - 1)) & 1u))) { used |= 1u << c; live |= 1u << i; }
    }
    while (live) {
      const int j = __ffs(live) - 1;
      live &= live - 1;
      const int c = __shfl_sync(gm, raw, j);
      const double lu = __shfl_sync(gm, logu, j);
      const double ll_up = __shfl_down_sync(gm, ll, 1), b_up = __shfl_down_sync(gm, beta, 1);
      double lla = ll; if (!(lla > -1e200)) lla = -1e200;
      double llb = ll_up; if (!(llb > -1e200)) llb = -1e200;
      const double lhr = -(b_up - beta) * (llb - lla);
      int acc_mine = 1;
      if (lhr < 0) acc_mine = (lu < lhr) ? 1 : 0;
      const bool accept = __shfl_sync(gm, acc_mine, c) != 0;
      const bool is_lo = (lane == c), is_hi = (lane == c + 1), involved = is_lo || is_hi;
      if (is_lo && c > 0) { if (dir > 0) ups++; if (dir < 0) downs++; }
      const int partner = is_lo ? c + 1 : (is_hi ? c : lane);
      if (accept) {
        { const double v = __shfl_sync(gm, ll, partner); ll = v; }
        { const double v = __shfl_sync(gm, lprior, partner); lprior = v; }
        { const int v = __shfl_sync(gm, perm, partner); perm = v; }
        if (involved) lpost = lprior + beta * ll;
      }
      if (involved) record();
      if (accept) {
        { const int v = __shfl_sync(gm, dir, partner); dir = v; }
        { const int v = __shfl_sync(gm, inst, partner); inst = v; }
        if (c == 0 && is_lo) dir = 1;
        if (c + 1 == R - 1 && is_hi) dir = -1;
        if (is_lo) sa++;
        if (m.evolve_rate > 0) {
          const double rate = m.evolve_rate;
          const double b_next = __shfl_down_sync(gm, beta, 1), lp_next = __shfl_down_sync(gm, lpost, 1);
          double sp = beta - b_next;
          if (m.evolve_lpost_cut >= 0 && lpost - lp_next > m.evolve_lpost_cut * beta) sp *= (1.0 + rate);
          if (is_lo) sp *= 1.0 + rate;
          double sum = 0;
          for (int k = 0; k < R - 1; k++) sum += __shfl_sync(gm, sp, k);
          const double norm = sum / (1 - __shfl_sync(gm, beta, R - 1));
          const double qn = sp / norm;
          double invtemp = 1, mine = beta;
          for (int k = 1; k < R - 1; k++) {
            invtemp -= __shfl_sync(gm, qn, k - 1);
            if (lane == k) mine = invtemp;
          }
          if (lane >= 1 && lane < R - 1) { beta = mine; lpost = lprior + mine * ll; }
        }
      }
      if (is_lo) sc++;
    }
  } else {
    const int parity = (int)(step & 1);
    const bool is_lo = ((lane & 1) == parity) && (lane + 1 < R);
    const bool is_hi = (lane >= 1) && (((lane - 1) & 1) == parity) && (lane < R);
    const int partner = is_lo ? lane + 1 : (is_hi ? lane - 1 : lane);
    const double ll_p = __shfl_sync(gm, ll, partner), b_p = __shfl_sync(gm, beta, partner);
    int flags = 0;
    if (is_lo) {
      uint32_t q[4];
      ptg_philox_draw(m.seed, ladder_stream, PTG_DOMAIN_STEP, step, PTG_BLK_SWAP_EVENODD + (uint32_t)lane, q);
      if (ptg_u52_to_unit(q[0], q[1]) < ptry) {
        double lla = ll; if (!(lla > -1e200)) lla = -1e200;
        double llb = ll_p; if (!(llb > -1e200)) llb = -1e200;
        const double lhr = -(b_p - beta) * (llb - lla);
        bool accept = true;
        if (lhr < 0) accept = (log(ptg_u52_to_unit(q[2], q[3])) < lhr);
        flags = 1 | (accept ? 2 : 0);
        if (lane > 0) { if (dir > 0) ups++; if (dir < 0) downs++; }
        sc++; if (accept) sa++;
      }
    }
    const int pflags = __shfl_sync(gm, flags, partner);
    if (is_hi) flags = pflags;
    const bool tried = (flags & 1) != 0, accept = (flags & 2) != 0;
    const double nll = __shfl_sync(gm, ll, partner), nlp = __shfl_sync(gm, lprior, partner);
    const int nperm = __shfl_sync(gm, perm, partner), ndir = __shfl_sync(gm, dir, partner), ninst = __shfl_sync(gm, inst, partner);
    if (tried && accept) {
      ll = nll; lprior = nlp; perm = nperm; lpost = lprior + beta * ll; dir = ndir; inst = ninst;
      if (lane == 0) dir = 1;
      if (lane == R - 1) dir = -1;
    }
    if (tried) record();
  }
  if (act) {
    L.perm[lane] = perm; L.napp[lane] = napp; L.n_lpost[lane] = lpost; L.n_beta[lane] = beta;
    L.app_src[2 * lane] = as0; L.app_src[2 * lane + 1] = as1;
    L.app_lpost[2 * lane] = al0; L.app_lpost[2 * lane + 1] = al1; L.app_beta[2 * lane] = ab0; L.app_beta[2 * lane + 1] = ab1;
    L.dir[lane] = dir; L.inst[lane] = inst; L.ups[lane] += ups; L.downs[lane] += downs; L.scount[lane] += sc; L.saccept[lane] += sa;
  }
}

template <int CPL>
__device__ __forceinline__ double xsum_tree(const double v[CPL]) {
  double a = 0;
#pragma unroll
  for (int k = 0; k < CPL; k++) a += v[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  return a;
}

// chi-squared likelihoods over data (bayesian.hh:595-622; polynomial model poly_example.cc:85-106, sum of sinusoids) with the
// DATA loop spread over the lanes of the chain's warp: lane l takes points l, l+32, ...; the partial sums meet in a shuffle
// tree.  Used when the batch has too few chains to fill the GPU with one thread per chain (BASELINE config B: 16 384 chains).
// dim <= 32 (CPL = 1): parameter j lives in lane j.
// sum over this lane's points of (poly(x_i) - y_i)^2 / S_i for a polynomial with DD coefficients; four independent partial sums per
// lane; 1/S_i precomputed.  This is the PRODUCTION (Philox-only) functor: Horner's rule on explicit fused multiply-adds -- DD + 2 fp64
// instructions per point instead of the 3 DD + 4 of the reference's unfused `y += xn*c_j; xn *= x_i` (poly_example.cc:97-101), which
// the tape-capable kernels keep.  Values agree with the reference's order to ~1e-15 relative (tests: 1e-12 gate on stored samples).
template <int DD>
__device__ __forceinline__ double xpoly_partial(const double *__restrict__ xs, const double *__restrict__ ys, const double *__restrict__ iS, long long N, const double *cj, int lane) {
  double p4[4] = {0, 0, 0, 0};
  for (long long i0 = lane; i0 < N; i0 += 128) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long i = i0 + 32 * u;
      if (i < N) {
        const double xi = __ldg(xs + i);
        double y = cj[DD - 1];
#pragma unroll
        for (int j = DD - 2; j >= 0; j--) y = fma(y, xi, cj[j]);
        const double dd = y - __ldg(ys + i);
        p4[u] = fma(dd * dd, __ldg(iS + i), p4[u]);
      }
    }
  }
  return (p4[0] + p4[1]) + (p4[2] + p4[3]);
}

// sum over this lane's points of (sum_k A_k sin(2 pi f_k t_i + phi_k) - y_i)^2 / S_i on a UNIFORM time grid t_i = t0 + i dt: the lane's
// points are 32 dt apart, so sin / cos of every component advance by a fixed rotation (4 fused multiply-adds) instead of a libm call;
// they are re-anchored with sincos of the reference's own phase expression every 64 points, which bounds the drift to ~1e-14.
template <int NS>
__device__ __forceinline__ double xsinusoid_partial_uniform(const double *__restrict__ xs, const double *__restrict__ ys, const double *__restrict__ iS, long long N,
                                                            double dt, const double *cj, int lane) {
  double sd[NS], cd[NS];
#pragma unroll
  for (int k = 0; k < NS; k++) sincos(2 * PTG_PI * cj[3 * k + 1] * (32 * dt), &sd[k], &cd[k]);
  double part = 0;
  for (long long base = lane; base < N; base += 32 * 64) {
    double sn[NS], cs[NS];
    const double tb = __ldg(xs + base);
#pragma unroll
    for (int k = 0; k < NS; k++) sincos(2 * PTG_PI * cj[3 * k + 1] * tb + cj[3 * k + 2], &sn[k], &cs[k]);
#pragma unroll 4
    for (int q = 0; q < 64; q++) {
      const long long i = base + 32 * q;
      if (i >= N) break;
      double y = 0;
#pragma unroll
      for (int k = 0; k < NS; k++) y = fma(cj[3 * k], sn[k], y);
      const double dd = y - __ldg(ys + i);
      part = fma(dd * dd, __ldg(iS + i), part);
#pragma unroll
      for (int k = 0; k < NS; k++) {
        const double ns = fma(sn[k], cd[k], cs[k] * sd[k]);
        const double nc = fma(cs[k], cd[k], -(sn[k] * sd[k]));
        sn[k] = ns; cs[k] = nc;
      }
    }
  }
  return part;
}

__device__ __forceinline__ double xlike_data_parallel(const PtgModel &m, double xmine, int lane) {
  const int D = m.dim;
  const long long N = m.n_ldata / 3;
  const double *__restrict__ xs = m.ldata, *__restrict__ ys = m.ldata + N, *__restrict__ S = m.ldata + 2 * N;
  double part = 0;
  if (m.like_kind == PTG_LIKE_POLY_CHI2) {
    // D is small (5 in config B): coefficients are pulled once per step; the inner loop is the reference's `y+=xn*c_j; xn*=x_k`
    double cj[16];
#pragma unroll
    for (int j = 0; j < 16; j++) cj[j] = __shfl_sync(0xffffffffu, xmine, j);
    const double *__restrict__ iS = m.ldata + 3 * N; // reciprocals of the variances (uploaded behind the data block)
    switch (D) { // the coefficient count as a compile-time constant: no predicated-off multiply-adds in the inner loop
#define X(DD) case DD: part = xpoly_partial<DD>(xs, ys, iS, N, cj, lane); break;
      X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
#undef X
    }
  } else {
    double cj[18];
#pragma unroll
    for (int j = 0; j < 18; j++) cj[j] = __shfl_sync(0xffffffffu, xmine, j);
    const double *__restrict__ iS = m.ldata + 3 * N;
    if (m.like_uniform_t) {
      switch (D / 3) {
#define X(NS) case NS: part = xsinusoid_partial_uniform<NS>(xs, ys, iS, N, m.like_dt, cj, lane); break;
        X(1) X(2) X(3) X(4) X(5) X(6)
#undef X
      }
    } else
    for (long long i = lane; i < N; i += 32) {
      const double ti = __ldg(xs + i);
      double y = 0;
#pragma unroll
      for (int k = 0; k + 2 < 18; k += 3) if (k + 2 < D) y += cj[k] * sin(2 * PTG_PI * cj[k + 1] * ti + cj[k + 2]);
      const double dd = y - __ldg(ys + i);
      part += dd * dd * __ldg(iS + i);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  double sum = part + m.like_nsum;
  sum /= -2;
  double result = sum - __ldg(m.lparams);
  if (!isfinite(result)) result = -CUDART_INF;
  return result;
}

template <int CPL, int MAXT>
__global__ void __launch_bounds__(MAXT) ptg_xmstep_kernel(const __grid_constant__ PtgModel m, PtgState s, long long step0, int n_steps, int trans_off) {
  constexpr int MODE = PTG_RNG_PHILOX;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int R = m.n_rungs, D = m.dim, DP = 32 * CPL, NP = m.n_props, RP = (R + 7) & ~7;
  XShared L;
  L.carve(smem_raw, R, DP, NP);
  for (int i = threadIdx.x; i < R * NP; i += blockDim.x) L.sbins[i] = m.bins[i];
  for (int i = threadIdx.x; i < RP * DP; i += blockDim.x) { L.rowA[i] = 0; L.rowB[i] = 0; }
  const int lane = threadIdx.x & 31, rung = threadIdx.x >> 5;
  const long long ladder = blockIdx.x;
  const long long chain = ladder * R + rung;
  double *rowA = L.rowA + (size_t)rung * DP, *rowB = L.rowB + (size_t)rung * DP, *myx = L.sx + (size_t)rung * DP;

  XChain<CPL> ch;
  ch.chain = chain;
#pragma unroll
  for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; ch.x[k] = (c < D) ? s.cur_x[(long long)c * m.n_chains + chain] : 0.0; }
  ch.lpost = s.lpost[chain]; ch.llike = s.llike[chain]; ch.lprior = s.lprior[chain]; ch.beta = s.beta[chain]; ch.map_lpost = s.map_lpost[chain];
  ch.nhist = s.nhist[chain]; ch.nsize = s.nsize[chain]; ch.ntries = s.ntries[chain]; ch.naccept = s.naccept[chain]; ch.last_type = s.last_type[chain];
  ch.slot = (int)(ch.nsize % m.hist_cap); ch.since_save = (int)(ch.nhist % m.save_every);
  Stream<MODE> rs, ls;
  stream_open<MODE>(m, s, rs, chain, (uint64_t)(m.ladder_offset + ladder) * PTG_STREAM_STRIDE + (uint64_t)rung, PTG_DOMAIN_STEP);
  stream_open<MODE>(m, s, ls, m.n_chains + ladder, (uint64_t)(m.ladder_offset + ladder) * PTG_STREAM_STRIDE + PTG_STREAM_LADDER, PTG_DOMAIN_STEP);
  if (lane == 0) {
    L.dir[rung] = s.directions[chain]; L.ups[rung] = s.ups[chain]; L.downs[rung] = s.downs[chain]; L.inst[rung] = s.instances[chain];
    L.scount[rung] = 0; L.saccept[rung] = 0;
  }
  const int maxswaps = m.maxswaps;
  const double swap_thresh = (R - 1) * m.swap_rate / maxswaps;
  double ptry = 2 * m.swap_rate; if (ptry > 1) ptry = 1;
  const double *bins = L.sbins + (size_t)rung * NP;
  const bool fullcov = m.like_kind == PTG_LIKE_GAUSS_FULLCOV;
  // all-uniform prior: this lane's box edges are loop-invariant (the PtgPrior1D structs sit 56 bytes apart in global memory)
  // staged once per CTA in the (otherwise unused) padding rows' neighbour: two rows of DP doubles behind rowB
  double *plo = L.rowB + (size_t)RP * DP, *phi = plo + DP;
  for (int c = threadIdx.x; c < DP; c += blockDim.x) {
    plo[c] = (c < D && m.all_uniform_prior) ? m.prior_w[c].a : -CUDART_INF;
    phi[c] = (c < D && m.all_uniform_prior) ? m.prior_w[c].b : CUDART_INF;
  }
  __syncthreads();

  for (int it = 0; it < n_steps; it++) {
    const uint64_t step = (uint64_t)(step0 + it);
    // publish, then warp 0 runs the ladder's swap phase on shuffles (falls back to the serial leader if a step has more
    // trials than lanes)
    if (maxswaps <= 32) {
#pragma unroll
      for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) myx[c] = ch.x[k]; }
      if (lane == 0) {
        L.sll[rung] = ch.llike; L.slpost[rung] = ch.lpost; L.slprior[rung] = ch.lprior; L.sbeta[rung] = ch.beta;
        L.n_lpost[rung] = ch.lpost; L.n_beta[rung] = ch.beta; L.perm[rung] = rung; L.napp[rung] = 0;
      }
      __syncthreads();
      if (R > 1 && rung == 0)
        xswap_warp(m, L, (uint64_t)(m.ladder_offset + ladder) * PTG_STREAM_STRIDE + PTG_STREAM_LADDER, step, lane, maxswaps, swap_thresh, ptry);
      __syncthreads();
    } else xpublish_and_swap<CPL, MODE>(m, L, ch, ls, step, myx, lane, rung, maxswaps, swap_thresh, ptry);
    const int na = L.napp[rung];
    ch.beta = L.n_beta[rung];
    const bool mh = (na == 0);
    if (!mh) xswapped_rung<CPL>(m, s, L, ch, na, lane, rung, DP);
    else if (m.evolve_rate > 0) ch.lpost = L.n_lpost[rung];

    // ---------------------------------------------------------------- phase A: proposal up to the rotation
    double newx[CPL], off[CPL];
    double prop_lh = 0;
    int type = 0, member = 0;
    bool need_t = false, is_gauss = false;
    uint32_t wB[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < CPL; k++) { newx[k] = ch.x[k]; off[k] = 0; }
    const double oldlprior = ch.lpost - ch.beta * ch.llike;
    if (mh) {
      rs.step = step;
      const int hsize = (int)(ch.nsize > m.hist_cap ? (long long)m.hist_cap : ch.nsize);
      uint32_t wA[4];
      rs.fetch(PTG_BLK_A, wA); rs.fetch(PTG_BLK_B, wB);
      if (m.wrap_in_set) {
        member = -1;
        const double x = (NP > 1) ? ptg_u32_to_unit(wA[0]) : 0.0;
        for (int i = 0; i < NP; i++) {
          const bool ready = (m.props[i].kind != PTG_PROP_DE) || hsize >= D * 10;
          if (member < 0 && ready && x < bins[i]) member = i;
        }
        if (member < 0) { rs.err = 2; member = 0; }
      }
      const PtgProp &p = m.props[member];
      if (p.kind == PTG_PROP_DE) {
        const double usnk = ptg_u32_to_unit(wA[1]), ug = ptg_u32_to_unit(wA[2]);
        if (!(p.snooker > usnk)) {
          double gamma = p.gamma_std;
          if (ug < p.g1frac) gamma = 1;
          int a1 = 0, a2 = 0;
          const int i1 = xde_index<CPL, MODE>(m, s, ch, p, rs, wA[3], 1, hsize, a1);
          const int i2 = xde_index<CPL, MODE>(m, s, ch, p, rs, wB[0], 2, hsize, a2);
          double a[CPL], b[CPL];
          xload_rec<CPL>(xhist<CPL>(m, s, ch, i1), a, lane, D);
          xload_rec<CPL>(xhist<CPL>(m, s, ch, i2), b, lane, D);
#pragma unroll
          for (int k = 0; k < CPL; k++) { const double t = ch.x[k] + a[k] * gamma; newx[k] = t + b[k] * (-gamma); }
        } else {
          const double gamma = (1.2 + ug) / p.reduce_gamma;
          double smznorm2 = 0, minusz[CPL], smz[CPL], t[CPL];
          int az = 0, isafe = 0;
          while (smznorm2 == 0) {
            const int iz = xde_index<CPL, MODE>(m, s, ch, p, rs, wB[1], 0, hsize, az);
            double zz[CPL];
            xload_rec<CPL>(xhist<CPL>(m, s, ch, iz), zz, lane, D);
#pragma unroll
            for (int k = 0; k < CPL; k++) { minusz[k] = zz[k] * (-1); smz[k] = ch.x[k] + minusz[k]; t[k] = smz[k] * smz[k]; }
            smznorm2 = xsum_tree<CPL>(t);
            if (++isafe > 1000) break;
          }
          int a1 = 0, a2 = 0;
          const int i1 = xde_index<CPL, MODE>(m, s, ch, p, rs, wA[3], 1, hsize, a1);
          const int i2 = xde_index<CPL, MODE>(m, s, ch, p, rs, wB[0], 2, hsize, a2);
          double a[CPL], b[CPL];
          xload_rec<CPL>(xhist<CPL>(m, s, ch, i1), a, lane, D);
          xload_rec<CPL>(xhist<CPL>(m, s, ch, i2), b, lane, D);
#pragma unroll
          for (int k = 0; k < CPL; k++) { const double ds12 = a[k] * gamma + b[k] * (-gamma); t[k] = ds12 * smz[k]; }
          const double fac = xsum_tree<CPL>(t) / smznorm2;
#pragma unroll
          for (int k = 0; k < CPL; k++) { newx[k] = ch.x[k] + smz[k] * fac; const double pmz = newx[k] + minusz[k]; t[k] = pmz * pmz; }
          prop_lh = (log(xsum_tree<CPL>(t)) - log(smznorm2)) * (D - 1) / 2.0;
          type = 1;
        }
      } else { // PTG_PROP_GAUSS
        is_gauss = true;
        xnormals<CPL, MODE>(m, rs, off, lane);
        const double *__restrict__ sig = m.prop_data + p.sigma_off;
#pragma unroll
        for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; off[k] = (c < D) ? off[k] * __ldg(sig + c) + 0.0 : 0.0; }
        if (p.one_d_frac > 0 && ptg_u32_to_unit(wA[1]) < p.one_d_frac) {
          const int ia = (int)(D * ptg_u32_to_unit(wA[2]));
#pragma unroll
          for (int k = 0; k < CPL; k++) if (CPL * lane + k != ia) off[k] = 0.0;
          type = 1;
        }
        if (p.has_transform) {
          if (p.trans_off == trans_off) need_t = true;                         // the batched DMMA rotation below
          else xtransform<CPL>(m, m.prop_data + p.trans_off, off, rowB, lane);   // a second, different matrix: exact per-warp path
        }
      }
      if (m.wrap_in_set) type = member + 10 * type;
    }
    // ---------------------------------------------------------------- batched rotation T = Off M^T on the tensor cores
    if (trans_off >= 0) {
      __syncwarp();
#pragma unroll
      for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) rowA[c] = need_t ? off[k] : 0.0; }
      if (__syncthreads_or(need_t ? 1 : 0)) {
        xcta_dmma(L.rowA, m.prop_data + trans_off, L.rowB, RP, D, DP, rung, R, lane);
        __syncthreads();
        if (need_t) {
#pragma unroll
          for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; off[k] = (c < D) ? rowB[c] : 0.0; }
        }
      }
    }
    // ---------------------------------------------------------------- phase B: enforce, prior, gate
    bool valid = m.zero_valid != 0, gate = false;
    double newlprior = -CUDART_INF;
    if (mh) {
      if (is_gauss) {
#pragma unroll
        for (int k = 0; k < CPL; k++) newx[k] = ch.x[k] + off[k];
      }
      if (valid) valid = xenforce<CPL>(m, newx, lane);
      if (m.all_uniform_prior) {
        bool in = valid;
#pragma unroll
        for (int k = 0; k < CPL; k++) in = in && !(newx[k] < plo[CPL * lane + k]) && !(newx[k] > phi[CPL * lane + k]);
        newlprior = __all_sync(0xffffffffu, in) ? m.uniform_lprior : -CUDART_INF;
      } else newlprior = xprior<CPL>(m, newx, valid, rowA, lane);
      gate = valid && ((newlprior > -1e200) || (newlprior - oldlprior > m.dprior_min));
    }
    // ---------------------------------------------------------------- batched quadratic form Y = X' Cinv^T on the tensor cores
    double newlike = -CUDART_INF, newlpost = -CUDART_INF;
    if (fullcov) {
      __syncwarp();
#pragma unroll
      for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) rowA[c] = (mh && gate) ? newx[k] : 0.0; }
      if (__syncthreads_or((mh && gate) ? 1 : 0)) {
        xcta_dmma(L.rowA, m.ldata, L.rowB, RP, D, DP, rung, R, lane);
        __syncthreads();
        if (mh && gate) {
          double t[CPL];
#pragma unroll
          for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; t[k] = (c < D) ? newx[k] * rowB[c] : 0.0; }
          newlike = __ldg(m.lparams) - 0.5 * xsum_tree<CPL>(t);
          if (!isfinite(newlike)) newlike = -CUDART_INF;
        }
      }
    } else if (mh && gate) {
      if (m.like_kind == PTG_LIKE_POLY_CHI2 || m.like_kind == PTG_LIKE_SINUSOID_CHI2) newlike = xlike_data_parallel(m, newx[0], lane);
      else newlike = xlike<CPL>(m, newx, rowA, rowB, lane);
    }
    // ---------------------------------------------------------------- phase C: Metropolis test, append
    double lhr = 0; int code = PTG_TRACE_SWAPPED;
    if (mh) {
      code = 0;
      bool accept = true;
      if (gate) newlpost = newlike * ch.beta + newlprior; else code |= PTG_TRACE_NOLIKE;
      lhr = prop_lh;
      if (isnan(lhr)) accept = false;
      lhr += newlpost - ch.lpost;
      if (!valid) { accept = false; code |= PTG_TRACE_INVALID; }
      if (accept && lhr < 0) accept = (log(ptg_u52_to_unit(wB[2], wB[3])) < lhr);
      ch.ntries++;
      if (accept) {
        ch.naccept++;
        ch.last_type = type;
#pragma unroll
        for (int k = 0; k < CPL; k++) ch.x[k] = newx[k];
        ch.llike = newlike; ch.lpost = newlpost; ch.lprior = newlprior;
        code |= PTG_TRACE_ACCEPT;
      }
      xappend<CPL>(m, s, ch, ch.x, ch.llike, ch.lpost, ch.beta, lane);
      code |= (type & PTG_TRACE_TYPE_MASK);
    }
    if (lane == 0 && (long long)step < m.trace_steps) {
      s.trace_lhr[step * m.n_chains + chain] = lhr;
      s.trace_code[step * m.n_chains + chain] = code;
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) s.cur_x[(long long)c * m.n_chains + chain] = ch.x[k]; }
  if (lane == 0) {
    s.lpost[chain] = ch.lpost; s.llike[chain] = ch.llike; s.lprior[chain] = ch.lprior; s.beta[chain] = ch.beta; s.map_lpost[chain] = ch.map_lpost;
    s.nhist[chain] = ch.nhist; s.nsize[chain] = ch.nsize; s.ntries[chain] = ch.ntries; s.naccept[chain] = ch.naccept; s.last_type[chain] = ch.last_type;
    stream_close<MODE>(s, rs, chain);
    if (rung == 0) stream_close<MODE>(s, ls, m.n_chains + ladder);
    s.directions[chain] = L.dir[rung]; s.ups[chain] = L.ups[rung]; s.downs[chain] = L.downs[rung]; s.instances[chain] = L.inst[rung];
    s.swap_count[chain] += L.scount[rung]; s.swap_accept[chain] += L.saccept[rung];
  }
}
