// ptg_wide.cuh -- warp-per-chain kernels for 17 <= dim <= 128 (BASELINE config D: correlated Gaussian, d = 100).
//
// Layout: one CTA = one ladder, one WARP = one chain (rung); a chain's state is spread over the lanes, CPL components
// per lane (component c lives in lane c / CPL, slot c % CPL; CPL = 1, 2 or 4), so history records (dim+2 doubles,
// contiguous per sample) are read and written with fully coalesced accesses and every elementwise stage of a proposal
// is lane-parallel.  Control flow is uniform per warp (one chain takes one branch), so nothing diverges.
//
// Reductions follow the reference's summation ORDER (sequential over the dimension index; Eigen's GEMV blocking for the
// eigen-rotated Gaussian proposal): the vector is published to a shared-memory row and every lane adds it up serially,
// which keeps tape-mode runs bit-exact with the reference for d = 100 as well.  The two dense contractions of the path
// -- the proposal rotation M (z o sigma) and the quadratic form x^T Cinv x -- are row-per-lane here (exact order);
// the batched DMMA variant for Philox-mode production runs replaces them in ptg_wide_mma.cuh.
//
// Replica swaps: lane 0 of warp 0 replays the reference's serial schedule (chain.cc:1410-1538, pry_temps :1809-1846) on
// the published scalars, tracking only a PERMUTATION of the rungs' states; afterwards each warp fetches the state it ends
// up with (and the states of its up-to-two history appends) from the published rows.
#pragma once
#include "ptg_kernels.cuh"

template <int CPL>
struct XChain {
  double x[CPL];
  double lpost, llike, lprior, beta, map_lpost;
  long long nhist, nsize, ntries, naccept;
  int last_type, slot, since_save;
  long long chain;
};

struct XShared {
  double *sx, *rowA, *rowB;                       // [R][DP] each
  double *sll, *slpost, *slprior, *sbeta;         // published at step start [R]
  double *n_lpost, *n_beta, *app_lpost, *app_beta, *split, *udraw, *sbins; // [R],[R],[2R],[2R],[R],[3*SLOTS],[R*NP]
  long long *scount, *saccept;                    // [R]
  int *perm, *napp, *app_src, *dir, *ups, *downs, *inst, *iswap; // [R],[R],[2R],[R]x4,[SLOTS]
  __host__ __device__ void carve(unsigned char *base, int R, int DP, int NP) {
    double *d = reinterpret_cast<double *>(base);
    const int RP = (R + 7) & ~7; // the DMMA kernel treats rowA / rowB as [RP][DP] matrices (8-row tiles)
    sx = d; d += (size_t)RP * DP; rowA = d; d += (size_t)RP * DP; rowB = d; d += (size_t)(RP + 2) * DP; // +2 rows: prior box edges (DMMA kernel)
    sll = d; d += R; slpost = d; d += R; slprior = d; d += R; sbeta = d; d += R;
    n_lpost = d; d += R; n_beta = d; d += R; app_lpost = d; d += 2 * R; app_beta = d; d += 2 * R; split = d; d += R;
    udraw = d; d += 3 * PTG_SWAP_SLOTS; sbins = d; d += (size_t)R * NP;
    long long *l = reinterpret_cast<long long *>(d);
    scount = l; l += R; saccept = l; l += R;
    int *i = reinterpret_cast<int *>(l);
    perm = i; i += R; napp = i; i += R; app_src = i; i += 2 * R; dir = i; i += R; ups = i; i += R; downs = i; i += R; inst = i; i += R;
    iswap = i; i += PTG_SWAP_SLOTS;
  }
};
static __host__ __device__ inline size_t ptg_xshared_bytes(int R, int DP, int NP) {
  size_t b = sizeof(double) * ((size_t)(3 * ((R + 7) & ~7) + 2) * DP + (size_t)R * 11 + 3 * PTG_SWAP_SLOTS + (size_t)R * NP) + sizeof(long long) * 2 * (size_t)R +
             sizeof(int) * ((size_t)R * 8 + PTG_SWAP_SLOTS);
  return (b + 15) & ~(size_t)15;
}

// every lane adds row[0..D) up in index order: the reference's `for(i) sum += ...` loops
__device__ __forceinline__ double xsum_ordered(const double *row, int D) {
  double s = 0;
  for (int j = 0; j < D; j++) s += row[j];
  return s;
}

template <int CPL>
__device__ __forceinline__ void xpublish(double *row, const double v[CPL], int lane, int D) {
#pragma unroll
  for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) row[c] = v[k]; }
  __syncwarp();
}

// MH_chain::add_state (chain.cc:916-949): the warp appends x (spread over the lanes) with the given scalars
template <int CPL>
__device__ __forceinline__ void xappend(const PtgModel &m, const PtgState &s, XChain<CPL> &ch, const double x[CPL], double llike, double lpost, double beta,
                                        int lane) {
  const int D = m.dim;
  if (lpost > ch.map_lpost) {
    ch.map_lpost = lpost;
#pragma unroll
    for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) s.map_x[(long long)c * m.n_chains + ch.chain] = x[k]; }
  }
  if (ch.since_save == 0) {
    const long long rec = ch.chain * m.hist_cap + ch.slot;
    double *h = s.hist + rec * m.hx;
#pragma unroll
    for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) h[c] = x[k]; }
    if (lane == 0) {
      s.hist_lp[2 * rec] = lpost; s.hist_lp[2 * rec + 1] = llike;
      if (m.record_full) { s.hist_acc[rec] = ch.naccept / (double)ch.ntries; s.hist_beta[rec] = beta; s.hist_type[rec] = ch.last_type; }
    }
    ch.nsize++;
    ch.slot = (ch.slot + 1 == m.hist_cap) ? 0 : ch.slot + 1;
  }
  ch.since_save = (ch.since_save + 1 == m.save_every) ? 0 : ch.since_save + 1;
  ch.nhist++;
}

template <int CPL>
__device__ __forceinline__ const double *xhist(const PtgModel &m, const PtgState &s, const XChain<CPL> &ch, int index) {
  int p = index;
  if (ch.nsize > m.hist_cap) { p = ch.slot + index; if (p >= m.hist_cap) p -= m.hist_cap; }
  return s.hist + (ch.chain * m.hist_cap + p) * m.hx;
}
template <int CPL>
__device__ __forceinline__ double xhist_lpost(const PtgModel &m, const PtgState &s, const XChain<CPL> &ch, int index) {
  int p = index;
  if (ch.nsize > m.hist_cap) { p = ch.slot + index; if (p >= m.hist_cap) p -= m.hist_cap; }
  return s.hist_lp[2 * (ch.chain * m.hist_cap + p)];
}
template <int CPL>
__device__ __forceinline__ void xload_rec(const double *rec, double v[CPL], int lane, int D) {
#pragma unroll
  for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; v[k] = (c < D) ? rec[c] : 0.0; }
}

// ------------------------------------------------------------------------------------------------- functors (exact order)
// state::enforce (states.cc:11-58,86-102): per dimension; any failing dimension invalidates the state
template <int CPL>
__device__ __forceinline__ bool xenforce(const PtgModel &m, double x[CPL], int lane) {
  bool ok = true;
  if (m.any_bound) {
#pragma unroll
    for (int k = 0; k < CPL; k++) {
      const int c = CPL * lane + k;
      if (c < m.dim) {
        const int lt = m.lower_w[c], ut = m.upper_w[c];
        if (lt != PTG_BOUND_OPEN || ut != PTG_BOUND_OPEN) ok = bound_enforce(lt, ut, m.xmin_w[c], m.xmax_w[c], x[k]) && ok;
      }
    }
  }
  return __all_sync(0xffffffffu, ok);
}
// log(prod_i pdf_i) (probability_function.hh:59), product in index order
template <int CPL>
__device__ __forceinline__ double xprior(const PtgModel &m, const double x[CPL], bool valid, double *row, int lane) {
  if (!valid) return -CUDART_INF;
  const int D = m.dim;
  if (m.all_uniform_prior) {
    bool in = true;
#pragma unroll
    for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) in = in && !(x[k] < m.prior_w[c].a) && !(x[k] > m.prior_w[c].b); }
    return __all_sync(0xffffffffu, in) ? m.uniform_lprior : -CUDART_INF;
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) row[c] = prior_factor(m.prior_w[c], x[k], m.lower_w[c], m.upper_w[c], m.xmin_w[c], m.xmax_w[c]); }
  __syncwarp();
  double result = 1;
  for (int j = 0; j < D; j++) result *= row[j];
  return log(result);
}
// likelihood functors in the reference's operation order; rowA / rowB: this warp's scratch rows
template <int CPL>
__device__ __forceinline__ double xlike(const PtgModel &m, const double x[CPL], double *rowA, double *rowB, int lane) {
  const int D = m.dim;
  const double *__restrict__ P = m.lparams;
  double result = 0;
  if (m.like_kind == PTG_LIKE_FLAT) return 0;
  __syncwarp();
  if (m.like_kind == PTG_LIKE_GAUSS_ISO) { // example.cc:116-143
#pragma unroll
    for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) { const double dx = x[k] - __ldg(P + 2 + c); rowA[c] = dx * dx; } }
    __syncwarp();
    const double r2 = xsum_ordered(rowA, D);
    result = __ldg(P) - r2 / __ldg(P + 1);
  } else if (m.like_kind == PTG_LIKE_POLY_CHI2 || m.like_kind == PTG_LIKE_SINUSOID_CHI2) {
    // chi^2 over data (bayesian.hh:595-622) in the reference's summation order: the lanes evaluate 32 consecutive points' terms in
    // parallel (each term in the reference's own operation order), then every lane adds the 32 terms up in index order
#pragma unroll
    for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) rowA[c] = x[k]; }
    __syncwarp();
    const long long N = m.n_ldata / 3;
    const double *__restrict__ xs = m.ldata, *__restrict__ ys = m.ldata + N, *__restrict__ S = m.ldata + 2 * N;
    const bool poly = m.like_kind == PTG_LIKE_POLY_CHI2;
    double sum = 0;
    for (long long base = 0; base < N; base += 32) {
      const long long i = base + lane;
      double term = 0;
      if (i < N) {
        const double xi = __ldg(xs + i);
        double y = 0;
        if (poly) { double xn = 1; for (int j = 0; j < D; j++) { y += xn * rowA[j]; xn *= xi; } } // poly_example.cc:97-101
        else for (int k = 0; k + 2 < D; k += 3) y += rowA[k] * sin(2 * PTG_PI * rowA[k + 1] * xi + rowA[k + 2]);
        const double dd = y - __ldg(ys + i);
        term = dd * dd / __ldg(S + i);
      }
      rowB[lane] = term;
      __syncwarp();
      const int nq = (int)((N - base) < 32 ? (N - base) : 32);
      for (int q = 0; q < nq; q++) sum += rowB[q];
      __syncwarp();
    }
    sum += m.like_nsum;
    sum /= -2;
    result = sum - __ldg(P);
  } else { // PTG_LIKE_GAUSS_FULLCOV, cython/exampleGaussian.py:103-109: y_i = sum_j C_ij x_j (j in order), q = sum_i x_i y_i (i in order)
    const double *__restrict__ C = m.ldata;
#pragma unroll
    for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) rowA[c] = x[k]; }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < CPL; k++) {
      const int i = CPL * lane + k;
      if (i < D) {
        const double *__restrict__ Ci = C + (size_t)i * D;
        double y = 0;
        for (int j = 0; j < D; j++) y += __ldg(Ci + j) * rowA[j];
        rowB[i] = x[k] * y;
      }
    }
    __syncwarp();
    const double q = xsum_ordered(rowB, D);
    result = __ldg(P) - 0.5 * q;
  }
  if (!isfinite(result)) result = -CUDART_INF; // bayesian.hh:569-575
  return result;
}
// vec = diagTransform * vec in Eigen 3.3.7's GEMV summation order (see ptg_warp.cuh); off is replaced by M off
template <int CPL>
__device__ __forceinline__ void xtransform(const PtgModel &m, const double *__restrict__ M, double off[CPL], double *row, int lane) {
  const int D = m.dim, CB = (D / 4) * 4, EVEN_ROWS = D & ~1;
  __syncwarp();
#pragma unroll
  for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) row[c] = off[k]; }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < CPL; k++) {
    const int i = CPL * lane + k;
    if (i < D) {
      const double *__restrict__ a = M + (size_t)i * D;
      double acc = 0;
      for (int j = 0; j < CB; j += 4) {
        if (i < EVEN_ROWS) acc = acc + ((__ldg(a + j) * row[j] + __ldg(a + j + 1) * row[j + 1]) + (__ldg(a + j + 2) * row[j + 2] + __ldg(a + j + 3) * row[j + 3]));
        else { acc = __ldg(a + j) * row[j] + acc; acc = __ldg(a + j + 1) * row[j + 1] + acc; acc = __ldg(a + j + 2) * row[j + 2] + acc; acc = __ldg(a + j + 3) * row[j + 3] + acc; }
      }
      for (int j = CB; j < D; j++) acc += __ldg(a + j) * row[j];
      off[k] = acc;
    }
  }
  __syncwarp();
}

// d standard normals for this chain, component c in lane c / CPL (same addresses / tape positions as draw_normals<D>)
template <int CPL, int MODE>
__device__ __forceinline__ void xnormals(const PtgModel &m, Stream<MODE> &rs, double z[CPL], int lane) {
  const int D = m.dim;
  if constexpr (MODE == PTG_RNG_PHILOX) {
    if constexpr (CPL == 1) {
      double z0 = 0, z1 = 0;
      if (lane < D) { uint32_t w[4]; rs.fetch(PTG_BLK_NORMAL + lane / 2, w); box_muller(w, z0, z1); }
      z[0] = (lane & 1) ? z1 : z0;
    } else {
#pragma unroll
      for (int k = 0; k < CPL; k += 2) {
        const int c = CPL * lane + k;
        double z0 = 0, z1 = 0;
        if (c < D) { uint32_t w[4]; rs.fetch(PTG_BLK_NORMAL + c / 2, w); box_muller(w, z0, z1); }
        z[k] = z0; z[k + 1] = z1;
      }
    }
  } else {
    if (rs.zpos + D > rs.zend) { rs.err = 1; for (int k = 0; k < CPL; k++) z[k] = 0; return; }
#pragma unroll
    for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; z[k] = (c < D) ? rs.zt[rs.zpos + c] : 0.0; }
    rs.zpos += D;
  }
}

// prior drawSample (probability_function.cc:37-47,147-154,264-279): component c from block blk0 + c
template <int CPL, int MODE>
__device__ __forceinline__ bool xprior_draw(const PtgModel &m, Stream<MODE> &rs, uint32_t blk0, double x[CPL], int lane) {
  const int D = m.dim;
  if constexpr (MODE == PTG_RNG_PHILOX) {
#pragma unroll
    for (int k = 0; k < CPL; k++) {
      const int c = CPL * lane + k;
      x[k] = 0;
      if (c < D) {
        uint32_t w[4]; rs.fetch(blk0 + c, w);
        const PtgPrior1D p = m.prior_w[c];
        if (p.kind == PTG_PRIOR_GAUSSIAN || p.kind == PTG_PRIOR_GAUSSIAN_WRAPPED) { double z0, z1; box_muller(w, z0, z1); x[k] = z0 * p.b + p.a; }
        else x[k] = invcdf1d(p, ptg_u52_to_unit(w[0], w[1]));
      }
    }
  } else {
    // tape: dimension i consumes one normal (Gaussian factor) or one uniform (others), in dimension order
    int nu = 0, nz = 0;
#pragma unroll
    for (int kk = 0; kk < CPL; kk++) x[kk] = 0;
    for (int i = 0; i < D; i++) {
      const bool g = m.prior_w[i].kind == PTG_PRIOR_GAUSSIAN || m.prior_w[i].kind == PTG_PRIOR_GAUSSIAN_WRAPPED;
      const int owner = i / CPL, k = i - owner * CPL;
      if (owner == lane) {
        const PtgPrior1D p = m.prior_w[i];
        double v;
        if (g) v = (rs.zpos + nz < rs.zend) ? rs.zt[rs.zpos + nz] * p.b + p.a : 0.0;
        else v = invcdf1d(p, (rs.upos + nu < rs.uend) ? rs.ut[rs.upos + nu] : 0.5);
#pragma unroll
        for (int kk = 0; kk < CPL; kk++) if (kk == k) x[kk] = v;
      }
      if (g) nz++; else nu++;
    }
    if (rs.upos + nu > rs.uend || rs.zpos + nz > rs.zend) rs.err = 1;
    rs.upos += nu; rs.zpos += nz;
  }
  return xenforce<CPL>(m, x, lane);
}

// differential_evolution::draw_i_from_chain (proposal_distribution.cc:744-778); chain-uniform
template <int CPL, int MODE>
__device__ __forceinline__ int xde_index(const PtgModel &m, const PtgState &s, const XChain<CPL> &ch, const PtgProp &p, Stream<MODE> &rs, uint32_t w0, int which,
                                         int hsize, int &attempt) {
  const int D = m.dim;
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
      const double lpost = xhist_lpost<CPL>(m, s, ch, index);
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

// ------------------------------------------------------------------------------------------------- MH step (one warp = one chain)
template <int CPL, int MODE>
__device__ __forceinline__ MhOut xmh_step(const PtgModel &m, const PtgState &s, XChain<CPL> &ch, Stream<MODE> &rs, const double *bins, double *rowA,
                                          double *rowB, int lane) {
  const int D = m.dim;
  double newx[CPL];
  double prop_lh = 0;
  int type = 0;
  bool valid = m.zero_valid != 0;
  const double oldlprior = ch.lpost - ch.beta * ch.llike;
  const int hsize = (int)(ch.nsize > m.hist_cap ? (long long)m.hist_cap : ch.nsize);
  uint32_t wA[4] = {0u, 0u, 0u, 0u}, wB[4] = {0u, 0u, 0u, 0u};
  if constexpr (MODE == PTG_RNG_PHILOX) { rs.fetch(PTG_BLK_A, wA); rs.fetch(PTG_BLK_B, wB); }
  int member = 0;
  if (m.wrap_in_set) {
    member = -1;
    for (int count = 0; count <= 100 && member < 0; count++) {
      double x = 0.0;
      if (m.n_props > 1) { if constexpr (MODE == PTG_RNG_PHILOX) x = ptg_u32_to_unit(wA[0]); else x = rs.next_u(); }
      for (int i = 0; i < m.n_props; i++) {
        const bool ready = (m.props[i].kind != PTG_PROP_DE) || hsize >= D * 10;
        if (member < 0 && ready && x < bins[i]) member = i;
      }
      if constexpr (MODE == PTG_RNG_PHILOX) break;
    }
    if (member < 0) { rs.err = 2; member = 0; }
  }
  const PtgProp &p = m.props[member];
  if (p.kind == PTG_PROP_DE) {
    double usnk, ug;
    if constexpr (MODE == PTG_RNG_PHILOX) { usnk = ptg_u32_to_unit(wA[1]); ug = ptg_u32_to_unit(wA[2]); }
    else { usnk = rs.next_u(); ug = rs.next_u(); }
    if (!(p.snooker > usnk)) {
      double gamma = p.gamma_std;
      if (ug < p.g1frac) gamma = 1;
      int a1 = 0, a2 = 0;
      const int i1 = xde_index<CPL, MODE>(m, s, ch, p, rs, wA[3], 1, hsize, a1);
      const int i2 = xde_index<CPL, MODE>(m, s, ch, p, rs, wB[0], 2, hsize, a2);
      double a[CPL], b[CPL];
      xload_rec<CPL>(xhist<CPL>(m, s, ch, i1), a, lane, D);
      xload_rec<CPL>(xhist<CPL>(m, s, ch, i2), b, lane, D);
      if constexpr (MODE == PTG_RNG_TAPE) { // the d normals of the discarded jitter are still consumed (H8-1)
        if (rs.zpos + D > rs.zend) rs.err = 1;
        rs.zpos += D;
      }
#pragma unroll
      for (int k = 0; k < CPL; k++) {
        const double t = ch.x[k] + a[k] * gamma;
        newx[k] = t + b[k] * (-gamma);
      }
      type = 0;
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
        __syncwarp();
        xpublish<CPL>(rowA, t, lane, D);
        smznorm2 = xsum_ordered(rowA, D);
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
      __syncwarp();
      xpublish<CPL>(rowA, t, lane, D);
      const double dot = xsum_ordered(rowA, D);
      const double fac = dot / smznorm2;
#pragma unroll
      for (int k = 0; k < CPL; k++) {
        newx[k] = ch.x[k] + smz[k] * fac;
        const double pmz = newx[k] + minusz[k];
        t[k] = pmz * pmz;
      }
      __syncwarp();
      xpublish<CPL>(rowA, t, lane, D);
      const double pmz2 = xsum_ordered(rowA, D);
      prop_lh = (log(pmz2) - log(smznorm2)) * (D - 1) / 2.0;
      type = 1;
    }
  } else if (p.kind == PTG_PROP_PRIOR_DRAW) { // draw_from_dist::draw (proposal_distribution.hh:124-129)
    valid = xprior_draw<CPL, MODE>(m, rs, PTG_BLK_PRIOR, newx, lane);
    const double lp_old = xprior<CPL>(m, ch.x, true, rowA, lane);
    prop_lh = lp_old - xprior<CPL>(m, newx, valid, rowA, lane);
    type = 0;
  } else { // PTG_PROP_GAUSS (gaussian_prop::draw, proposal_distribution.hh:194-218)
    double off[CPL];
    xnormals<CPL, MODE>(m, rs, off, lane);
    const double *__restrict__ sig = m.prop_data + p.sigma_off;
#pragma unroll
    for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; off[k] = (c < D) ? off[k] * __ldg(sig + c) + 0.0 : 0.0; }
    double x1 = 1;
    if (p.one_d_frac > 0) { if constexpr (MODE == PTG_RNG_PHILOX) x1 = ptg_u32_to_unit(wA[1]); else x1 = rs.next_u(); }
    if (p.one_d_frac > 0 && x1 < p.one_d_frac) {
      double ua;
      if constexpr (MODE == PTG_RNG_PHILOX) ua = ptg_u32_to_unit(wA[2]); else ua = rs.next_u();
      const int ia = (int)(D * ua);
#pragma unroll
      for (int k = 0; k < CPL; k++) if (CPL * lane + k != ia) off[k] = 0.0;
      type = 1;
    } else type = 0;
    if (p.has_transform) xtransform<CPL>(m, m.prop_data + p.trans_off, off, rowA, lane);
#pragma unroll
    for (int k = 0; k < CPL; k++) newx[k] = ch.x[k] + off[k];
  }
  if (m.wrap_in_set) type = member + 10 * type;

  if (valid) valid = xenforce<CPL>(m, newx, lane);
  const double newlprior = xprior<CPL>(m, newx, valid, rowA, lane);
  double newlike, newlpost;
  int code = 0;
  bool accept = true;
  if (valid && ((newlprior > -1e200) || (newlprior - oldlprior > m.dprior_min))) {
    newlike = xlike<CPL>(m, newx, rowA, rowB, lane);
    newlpost = newlike * ch.beta + newlprior;
  } else { newlike = newlpost = -CUDART_INF; code |= PTG_TRACE_NOLIKE; }
  double lhr = prop_lh;
  if (isnan(lhr)) accept = false;
  lhr += newlpost - ch.lpost;
  if (!valid) { accept = false; code |= PTG_TRACE_INVALID; }
  if (accept && lhr < 0) {
    double u;
    if constexpr (MODE == PTG_RNG_PHILOX) u = ptg_u52_to_unit(wB[2], wB[3]); else u = rs.next_u();
    accept = (log(u) < lhr);
  }
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
  MhOut o; o.lhr = lhr; o.code = code;
  return o;
}

// ------------------------------------------------------------------------------------------------- swap phase leader
// lane 0 of warp 0: the reference's serial schedule on the published scalars; states are tracked as a permutation
template <int MODE>
__device__ __forceinline__ void xswap_trial(const PtgModel &m, XShared &L, Stream<MODE> &ls, int i, double u_philox) {
  const int R = m.n_rungs;
  if (i > 0) {
    if (L.dir[i] > 0) L.ups[i]++;
    if (L.dir[i] < 0) L.downs[i]++;
  }
  const int sa = L.perm[i], sb = L.perm[i + 1];
  double lla = L.sll[sa]; if (!(lla > -1e200)) lla = -1e200;
  double llb = L.sll[sb]; if (!(llb > -1e200)) llb = -1e200;
  const double lhr = -(L.n_beta[i + 1] - L.n_beta[i]) * (llb - lla);
  bool accept = true;
  if (lhr < 0) {
    double u;
    if constexpr (MODE == PTG_RNG_PHILOX) u = u_philox; else u = ls.next_u();
    accept = (log(u) < lhr);
  }
  if (accept) {
    L.perm[i] = sb; L.perm[i + 1] = sa;
    L.n_lpost[i + 1] = L.slprior[sa] + L.n_beta[i + 1] * L.sll[sa];
    L.n_lpost[i] = L.slprior[sb] + L.n_beta[i] * L.sll[sb];
  }
  for (int q = 0; q < 2; q++) { // appends of rungs i+1 and i with the state / posterior / beta they hold right now
    const int r = accept ? (q == 0 ? i + 1 : i) : (q == 0 ? i : i + 1);
    const int k = L.napp[r];
    if (k < 2) { L.app_src[2 * r + k] = L.perm[r]; L.app_lpost[2 * r + k] = L.n_lpost[r]; L.app_beta[2 * r + k] = L.n_beta[r]; }
    L.napp[r] = k + 1;
  }
  if (accept) {
    { const int t = L.dir[i]; L.dir[i] = L.dir[i + 1]; L.dir[i + 1] = t; }
    { const int t = L.inst[i]; L.inst[i] = L.inst[i + 1]; L.inst[i + 1] = t; }
    if (i == 0) L.dir[i] = 1;
    if (i + 1 == R - 1) L.dir[i + 1] = -1;
    L.saccept[i]++;
    if (m.evolve_rate > 0) { // pry_temps (chain.cc:1809-1846) + resetTemp (chain.cc:1088-1091)
      const double rate = m.evolve_rate;
      for (int k = 0; k < R - 1; k++) {
        double sp = L.n_beta[k] - L.n_beta[k + 1];
        if (m.evolve_lpost_cut >= 0 && L.n_lpost[k] - L.n_lpost[k + 1] > m.evolve_lpost_cut * L.n_beta[k]) sp *= (1.0 + rate);
        L.split[k] = sp;
      }
      L.split[i] *= 1.0 + rate;
      double sum = 0;
      for (int k = 0; k < R - 1; k++) sum += L.split[k];
      const double norm = sum / (1 - L.n_beta[R - 1]);
      double invtemp = 1;
      for (int k = 1; k < R - 1; k++) {
        invtemp -= L.split[k - 1] / norm;
        L.n_beta[k] = invtemp;
        L.n_lpost[k] = L.slprior[L.perm[k]] + invtemp * L.sll[L.perm[k]];
      }
    }
  }
  L.scount[i]++;
}

// steps 1-2 of a PT iteration, shared by the exact and the DMMA kernels: publish every chain's state, then the ladder's
// swap phase on lane 0 of warp 0.  Ends with a CTA barrier; afterwards L.napp / L.perm / L.n_lpost / L.n_beta describe the outcome.
template <int CPL, int MODE>
__device__ __forceinline__ void xpublish_and_swap(const PtgModel &m, XShared &L, const XChain<CPL> &ch, Stream<MODE> &ls, uint64_t step, double *myx, int lane,
                                                  int rung, int maxswaps, double swap_thresh, double ptry) {
  const int R = m.n_rungs, D = m.dim;
    // 1. publish
#pragma unroll
  for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) myx[c] = ch.x[k]; }
  if (lane == 0) {
    L.sll[rung] = ch.llike; L.slpost[rung] = ch.lpost; L.slprior[rung] = ch.lprior; L.sbeta[rung] = ch.beta;
    L.n_lpost[rung] = ch.lpost; L.n_beta[rung] = ch.beta; L.perm[rung] = rung; L.napp[rung] = 0;
  }
  if (R > 1 && rung == 0) {
    if constexpr (MODE == PTG_RNG_PHILOX) {
      ls.step = step;
      if (m.swap_mode == PTG_SWAP_REFERENCE) {
        for (int j = lane; j < maxswaps; j += 32) {
          uint32_t w[4]; ls.fetch((uint32_t)j, w);
          L.udraw[3 * j] = ptg_u32_to_unit(w[0]); L.udraw[3 * j + 1] = ptg_u32_to_unit(w[1]); L.udraw[3 * j + 2] = ptg_u52_to_unit(w[2], w[3]);
        }
      } else {
        const int parity = (int)(step & 1);
        if (lane + 1 < R && ((lane & 1) == parity)) {
          uint32_t w[4]; ls.fetch(PTG_BLK_SWAP_EVENODD + (uint32_t)lane, w);
          L.udraw[3 * (lane >> 1)] = ptg_u52_to_unit(w[0], w[1]); L.udraw[3 * (lane >> 1) + 2] = ptg_u52_to_unit(w[2], w[3]);
        }
      }
    }
  }
  __syncthreads();
  // 2. swap phase
  if (R > 1 && rung == 0 && lane == 0) {
    if (m.swap_mode == PTG_SWAP_REFERENCE) {
      for (int i = 0; i < maxswaps; i++) { // chain.cc:1410-1420
        int cand = -2;
        double x;
        if constexpr (MODE == PTG_RNG_PHILOX) x = L.udraw[3 * i]; else x = ls.next_u();
        if (x < swap_thresh) {
          if constexpr (MODE == PTG_RNG_PHILOX) x = L.udraw[3 * i + 1]; else x = ls.next_u();
          cand = (int)(x * (R - 1));
          for (int j = 0; j < i; j++) if (L.iswap[j] == cand || L.iswap[j] + 1 == cand) cand = -2;
        }
        L.iswap[i] = cand;
      }
      for (int j = 0; j < maxswaps; j++) if (L.iswap[j] >= 0) xswap_trial<MODE>(m, L, ls, L.iswap[j], L.udraw[3 * j + 2]);
    } else {
      const int parity = (int)(step & 1);
      int n = 0;
      for (int i = parity; i + 1 < R; i += 2) { // tries first (the oracle's order), then the trials
        double x;
        if constexpr (MODE == PTG_RNG_PHILOX) x = L.udraw[3 * (i >> 1)]; else x = ls.next_u();
        if (x < ptry) L.iswap[n++] = i;
      }
      for (int j = 0; j < n; j++) xswap_trial<MODE>(m, L, ls, L.iswap[j], L.udraw[3 * (L.iswap[j] >> 1) + 2]);
    }
  }
  __syncthreads();
}

// step 3a: a rung that took part in swap trials appends (up to twice) and takes over the state it ends up with
template <int CPL>
__device__ __forceinline__ void xswapped_rung(const PtgModel &m, const PtgState &s, XShared &L, XChain<CPL> &ch, int na, int lane, int rung, int DP) {
  const int D = m.dim;
  for (int k = 0; k < 2 && k < na; k++) {
    const int src = L.app_src[2 * rung + k];
    double xs[CPL];
    xload_rec<CPL>(L.sx + (size_t)src * DP, xs, lane, D);
    xappend<CPL>(m, s, ch, xs, L.sll[src], L.app_lpost[2 * rung + k], L.app_beta[2 * rung + k], lane);
  }
  const int src = L.perm[rung];
  xload_rec<CPL>(L.sx + (size_t)src * DP, ch.x, lane, D);
  ch.llike = L.sll[src]; ch.lprior = L.slprior[src]; ch.lpost = L.n_lpost[rung];
}

// ------------------------------------------------------------------------------------------------- step kernel
// MAXT: the CTA size the kernel is compiled for (32 * n_rungs rounded up to 512 / 768 / 1024): sets the register budget
template <int CPL, int MODE, int MAXT>
__global__ void __launch_bounds__(MAXT) ptg_xstep_kernel(const __grid_constant__ PtgModel m, PtgState s, long long step0, int n_steps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int R = m.n_rungs, D = m.dim, DP = 32 * CPL, NP = m.n_props;
  XShared L;
  L.carve(smem_raw, R, DP, NP);
  for (int i = threadIdx.x; i < R * NP; i += blockDim.x) L.sbins[i] = m.bins[i];
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
  __syncthreads();

  for (int it = 0; it < n_steps; it++) {
    const uint64_t step = (uint64_t)(step0 + it);
    if constexpr (MODE == PTG_RNG_TAPE) {
      if (s.u_mark && (long long)step < s.n_mark_steps) {
        const long long ns = m.n_chains + m.n_ladders, row = (long long)step * ns;
        rs.upos = s.u_mark[row + chain]; rs.zpos = s.z_mark[row + chain];
        ls.upos = s.u_mark[row + m.n_chains + ladder]; ls.zpos = s.z_mark[row + m.n_chains + ladder];
      }
    }
    xpublish_and_swap<CPL, MODE>(m, L, ch, ls, step, myx, lane, rung, maxswaps, swap_thresh, ptry);
    // 3. swap appends or MH step
    double lhr = 0; int code = PTG_TRACE_SWAPPED;
    const int na = L.napp[rung];
    ch.beta = L.n_beta[rung];
    if (na > 0) {
      xswapped_rung<CPL>(m, s, L, ch, na, lane, rung, DP);
    } else {
      if (m.evolve_rate > 0) ch.lpost = L.n_lpost[rung];
      rs.step = step;
      MhOut o = xmh_step<CPL, MODE>(m, s, ch, rs, bins, rowA, rowB, lane);
      lhr = o.lhr; code = o.code;
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

// ------------------------------------------------------------------------------------------------- init / eval
// MH_chain::initialize (chain.cc:846-876), one warp per chain; 4 warps per CTA, two scratch rows per warp
template <int CPL, int MODE>
__global__ void __launch_bounds__(128) ptg_xinit_kernel(const __grid_constant__ PtgModel m, PtgState s, const double *__restrict__ init_x) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int D = m.dim, DP = 32 * CPL, R = m.n_rungs;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long c = (long long)blockIdx.x * 4 + wib;
  if (c >= m.n_chains) return;
  double *rowA = reinterpret_cast<double *>(smem_raw) + (size_t)(2 * wib) * DP, *rowB = rowA + DP;
  const long long ladder = c / R; const int rung = (int)(c - ladder * R);
  XChain<CPL> ch;
  ch.chain = c;
  ch.beta = s.beta[c]; ch.map_lpost = s.map_lpost[c]; ch.nhist = 0; ch.nsize = s.nsize[c]; ch.ntries = s.ntries[c]; ch.naccept = s.naccept[c];
  ch.last_type = s.last_type[c]; ch.slot = (int)(ch.nsize % m.hist_cap); ch.since_save = 0; ch.lpost = ch.llike = ch.lprior = 0;
  Stream<MODE> rs;
  stream_open<MODE>(m, s, rs, c, (uint64_t)(m.ladder_offset + ladder) * PTG_STREAM_STRIDE + (uint64_t)rung, PTG_DOMAIN_INIT);
  for (int k = 0; k < m.n_init; k++) {
    double x[CPL];
    if (init_x) xload_rec<CPL>(init_x + ((long long)c * m.n_init + k) * D, x, lane, D);
    else {
      rs.step = (uint64_t)k;
      int icnt = 0;
      bool valid = xprior_draw<CPL, MODE>(m, rs, 0, x, lane);
      while (!valid || xlike<CPL>(m, x, rowA, rowB, lane) < -1e100) {
        icnt++;
        if (icnt >= 100000) { rs.err = 4; break; }
        valid = xprior_draw<CPL, MODE>(m, rs, (uint32_t)icnt * PTG_INIT_ATTEMPT_STRIDE, x, lane);
      }
    }
    const double ll = xlike<CPL>(m, x, rowA, rowB, lane);
    const double lp = xprior<CPL>(m, x, true, rowA, lane);
#pragma unroll
    for (int q = 0; q < CPL; q++) ch.x[q] = x[q];
    ch.llike = ll; ch.lprior = lp; ch.lpost = lp + ch.beta * ll;
    ch.nhist = 0; ch.since_save = 0;
    xappend<CPL>(m, s, ch, ch.x, ch.llike, ch.lpost, ch.beta, lane);
  }
#pragma unroll
  for (int q = 0; q < CPL; q++) { const int cc = CPL * lane + q; if (cc < D) s.cur_x[(long long)cc * m.n_chains + c] = ch.x[q]; }
  if (lane == 0) {
    s.lpost[c] = ch.lpost; s.llike[c] = ch.llike; s.lprior[c] = ch.lprior; s.map_lpost[c] = ch.map_lpost;
    s.nhist[c] = 0; s.nsize[c] = ch.nsize; s.last_type[c] = ch.last_type;
    stream_close<MODE>(s, rs, c);
  }
}

template <int CPL>
__global__ void __launch_bounds__(128) ptg_xeval_kernel(const __grid_constant__ PtgModel m, const double *__restrict__ x, long long n, double *__restrict__ out_ll,
                                                        double *__restrict__ out_lp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int D = m.dim, DP = 32 * CPL;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 4 + wib;
  if (i >= n) return;
  double *rowA = reinterpret_cast<double *>(smem_raw) + (size_t)(2 * wib) * DP, *rowB = rowA + DP;
  double v[CPL];
  xload_rec<CPL>(x + i * D, v, lane, D);
  if (out_ll) { const double ll = xlike<CPL>(m, v, rowA, rowB, lane); if (lane == 0) out_ll[i] = ll; }
  if (out_lp) {
    const bool valid = xenforce<CPL>(m, v, lane);
    const double lp = xprior<CPL>(m, v, valid, rowA, lane);
    if (lane == 0) out_lp[i] = lp;
  }
}
