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
  const int mt_n = RP >> 3, nt_n = (D + 7) >> 3, KP = (D + 3) & ~3;
  const int g = lane >> 2, t = lane & 3;
  for (int tile = warp; tile < mt_n * nt_n; tile += nwarps) {
    const int mt = tile / nt_n, nt = tile - mt * nt_n;
    const double *arow = In + (size_t)(mt * 8 + g) * DP;
    const int n = nt * 8 + g;
    const double *__restrict__ brow = Cm + (size_t)n * D;
    const bool nok = n < D;
    double c0 = 0, c1 = 0;
    for (int k0 = 0; k0 < KP; k0 += 4) {
      const int k = k0 + t;
      const double a = arow[k];
      const double b = (nok && k < D) ? __ldg(brow + k) : 0.0;
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    }
    double *orow = Out + (size_t)(mt * 8 + g) * DP + nt * 8 + 2 * t;
    orow[0] = c0; orow[1] = c1;
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
  __syncthreads();

  for (int it = 0; it < n_steps; it++) {
    const uint64_t step = (uint64_t)(step0 + it);
    xpublish_and_swap<CPL, MODE>(m, L, ch, ls, step, myx, lane, rung, maxswaps, swap_thresh, ptry);
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
      newlprior = xprior<CPL>(m, newx, valid, rowA, lane);
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
    } else if (mh && gate) newlike = xlike<CPL>(m, newx, rowA, rowB, lane);
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
