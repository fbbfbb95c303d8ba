// ptg_wide_pipe.cuh -- the PIPELINED warp-per-chain production kernel for the full-covariance Gaussian workload (BASELINE config D:
// d = 100, eigen-rotated Gaussian proposal + differential evolution; Philox draws).  Same arithmetic as ptg_xmstep_kernel
// (ptg_wide_mma.cuh) -- chains are bit-identical between the two -- restructured around what the round-1 profile showed: the FP64
// tensor pipe ran 10 % of the time because every DMMA operand came through L1 / L2 behind a dependent load, seven CTA barriers
// separated the phases, and the ladder's swap phase ran on warp 0 while 23 warps waited.
//
//  * Both dim x dim matrices (rotation M and Cinv) are RESIDENT IN SHARED MEMORY for the whole launch, rows padded to a stride of
//    DS = 8 (mod 16) doubles so that the 128-bit fragment loads of the eight row groups of an MMA are bank-conflict free.
//  * One CTA = one ladder, one warp per chain.  The hottest rung's warp runs the ladder's swap phase (chain.cc:1410-1538) on shuffles
//    WHILE the other warps run the rotation contraction; nobody waits for it.
//  * Everything of a proposal that does not depend on the swap outcome (Philox blocks, member selection, the Gaussian offsets z o sigma,
//    the DE history indices -- their records are only PREFETCHED into L2 here) is done BEFORE the swap barrier; a rung that turns out to
//    have been in a swap trial discards it (as the reference skips its MH step, chain.cc:1553-1558).
//  * The rotation contracts only the rows of Gaussian-proposal chains (row indirection, 8-row tiles), the quadratic form only gated rows.
//  * Four barriers per PT iteration instead of seven, none of them behind a single busy warp; the current states that swapped rungs exchange travel through a small global
//    scratch row per chain (L2 resident) instead of a third shared-memory row buffer, which is what lets both matrices fit.
//
//  * Shared-memory addresses are constant-bank offsets (XPLayout, computed on the host); the two Philox blocks of a step are generated
//    once per warp (even / odd lanes) instead of by all 32 lanes.
//
//    S1  publish scalars + x (scratch); prep: draws, member, offsets z o sigma -> OFF rows, DE indices (records prefetched)
//    -- barrier --
//    S2  T = OFF M^T on DMMA (rows of the rotated proposals) on warps 0 .. R-2                       | warp R-1: the ladder's swap phase
//    -- barrier --
//    S3  swapped rungs append / take over states; MH rungs: proposal, enforce, prior, gate -> NEWX rows
//    -- barrier --
//    S4  Y = NEWX Cinv^T on DMMA (gated rows)
//    -- barrier --
//    S5  quadratic form, Metropolis test, append
#pragma once
#include "ptg_wide_mma.cuh"

// Warp-uniform scalars of one chain, parked in shared memory between the stages that use them: every lane of a chain's warp holds the same
// values, and keeping all of them in registers across the whole iteration is what pushed the 80-register kernel into 160 B of spills.
struct XPPark {
  double lprior, map_lpost;
  long long nhist, ntries, naccept;
  int last_type, since_save;
};
struct XPShared {
  double *Ms, *Cs;            // [D][DS] each, resident matrices (row n of the matrix = output component n)
  double *bufA, *bufB;        // [RP + 2][DS]: OFF / NEWX rows ; T / Y rows.  Row RP of bufA = zeros (padding rows of an MMA tile), row RP of bufB = dump
  double *plo, *phi;          // [32 CPL] prior box edges (all-uniform prior)
  int *kindflag;              // [R]: 1 = row of the rotation (set in S1), 2 = row of the quadratic form (set in S3)
  int *gsig, *gaxis;          // [R]: rotated Gaussian proposals of this iteration: offset of the member's sigmas in prop_data, axis of a 1-D step or -1
  XPPark *park;               // [R]
};

// Byte offsets of every shared-memory array from the start of dynamic shared memory, computed on the host and passed as a kernel parameter:
// the addresses are then constant-bank offsets instead of integer arithmetic the compiler re-derives inside the loop under register pressure.
struct XPLayout {
  int DS, RP;
  int Ms, Cs, bufA, bufB, plo, phi, kindflag, gsig, gaxis, park;
  int sll, slpost, slprior, sbeta, n_lpost, n_beta, app_lpost, app_beta, sbins, scount, saccept, perm, napp, app_src, dir, ups, downs, inst;
};
static inline int ptg_xp_stride(int D) { int ds = (D + 7) & ~7; while ((ds & 15) != 8) ds += 8; return ds; }
static inline size_t ptg_xp_shared_bytes(int R, int D, int NP, int CPL) {
  const int DS = ptg_xp_stride(D), RP = (R + 7) & ~7;
  size_t b = ptg_xshared_bytes(R, 0, NP);                                   // the ladder scalars / swap outcome block of XShared (no row buffers)
  b += sizeof(double) * ((size_t)2 * D * DS + (size_t)2 * (RP + 2) * DS + 2 * (size_t)32 * CPL);
  b += sizeof(int) * ((size_t)3 * R + 8);
  b = (b + 15) & ~(size_t)15;
  b += sizeof(XPPark) * (size_t)R;
  return (b + 15) & ~(size_t)15;
}

static inline XPLayout ptg_xp_layout(int R, int D, int NP, int CPL) {
  XPLayout y;
  y.DS = ptg_xp_stride(D); y.RP = (R + 7) & ~7;
  XShared L;
  L.carve(nullptr, R, 0, NP);
  auto off = [](const void *p) { return (int)(reinterpret_cast<const unsigned char *>(p) - reinterpret_cast<const unsigned char *>(0)); };
  y.sll = off(L.sll); y.slpost = off(L.slpost); y.slprior = off(L.slprior); y.sbeta = off(L.sbeta); y.n_lpost = off(L.n_lpost); y.n_beta = off(L.n_beta);
  y.app_lpost = off(L.app_lpost); y.app_beta = off(L.app_beta); y.sbins = off(L.sbins); y.scount = off(L.scount); y.saccept = off(L.saccept);
  y.perm = off(L.perm); y.napp = off(L.napp); y.app_src = off(L.app_src); y.dir = off(L.dir); y.ups = off(L.ups); y.downs = off(L.downs); y.inst = off(L.inst);
  int b = (int)ptg_xshared_bytes(R, 0, NP);
  y.Ms = b; b += 8 * D * y.DS; y.Cs = b; b += 8 * D * y.DS;
  y.bufA = b; b += 8 * (y.RP + 2) * y.DS; y.bufB = b; b += 8 * (y.RP + 2) * y.DS;
  y.plo = b; b += 8 * 32 * CPL; y.phi = b; b += 8 * 32 * CPL;
  y.kindflag = b; b += 4 * R; y.gsig = b; b += 4 * R; y.gaxis = b; b += 4 * R;
  b = (b + 15) & ~15; y.park = b;
  return y;
}

// Out[row][n] = sum_k In[row][k] * Bs[n][k] for the rows whose bit is set in `mask` (8-row MMA tiles; padding rows read the zero row
// `pad_row` of In and write the dump row `pad_row` of Out); Bs in shared memory with row stride DS.  k is consumed in the same permuted
// order as xcta_dmma (D % 4 == 0: lane t takes elements (2t, 2t+1) of each group of 8 for two consecutive MMAs), so every output row
// is bit-identical with the L1-streamed version.  Called by the `nwarps` chain warps.
__device__ __forceinline__ void xp_dmma(const double *In, const double *Bs, double *Out, unsigned mask, int D, int DS, int pad_row, int warp, int nwarps, int lane) {
  const int cnt = __popc(mask);
  const int mt_n = (cnt + 7) >> 3, nt_n = (D + 7) >> 3;
  const int g = lane >> 2, t = lane & 3;
  const int K8 = D & ~7;
  for (int tile = warp; tile < mt_n * nt_n; tile += nwarps) {
    const int mt = tile / nt_n, nt = tile - mt * nt_n;
    const int ridx = mt * 8 + g;
    const int row = (ridx < cnt) ? (int)__fns(mask, 0, ridx + 1) : pad_row;
    const double *arow = In + (size_t)row * DS;
    int n = nt * 8 + g;
    if (n >= D) n = D - 1;
    const double *brow = Bs + (size_t)n * DS;
    double c0 = 0, c1 = 0;
#pragma unroll 4
    for (int k0 = 0; k0 < K8; k0 += 8) { // (an explicit two-stage software pipeline of these loads measured slower: 2.20e8 vs 2.43e8 chain-steps/s)
      const double2 a2 = *reinterpret_cast<const double2 *>(arow + k0 + 2 * t);
      const double2 b2 = *reinterpret_cast<const double2 *>(brow + k0 + 2 * t);
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a2.x), "d"(b2.x));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a2.y), "d"(b2.y));
    }
    for (int k0 = K8; k0 < D; k0 += 4) { // tail: columns >= D of both operands are zero in shared memory
      const int k = k0 + t;
      const double a = arow[k], b = brow[k];
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    }
    double *orow = Out + (size_t)row * DS + nt * 8 + 2 * t; // accumulator fragment: lane g*4+t holds C[g][2t..2t+1]
    orow[0] = c0; orow[1] = c1;
  }
}

// POOL = true (experiment, PTG_XP_POOL=1): the Box-Muller normals of the rotated proposals pooled over the CTA after the first barrier (one
// pair per lane from the owner's stream address) and the rotation on all warps behind one more barrier.  Bit-identical, but MEASURED
// SLOWER on B200 (2.25e8 vs 2.38e8 chain-steps/s on config D): the extra barrier costs more than the balanced prep phase saves.
// Also measured and removed: one warp per 8-column strip holding the accumulators of all row tiles (up to four independent MMA chains per
// warp, each matrix fragment loaded once for all of them): bit-identical, 1.59e8 -- 13 busy warps and 280 B of spills lose far more than
// the covered accumulate latency gains.
template <int CPL, int MAXT, bool POOL>
__global__ void __launch_bounds__(MAXT) ptg_xpstep_kernel(const __grid_constant__ PtgModel m, PtgState s, long long step0, int n_steps, int trans_off, double *xscratch,
                                                          const __grid_constant__ XPLayout lay) {
  constexpr int MODE = PTG_RNG_PHILOX;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int R = m.n_rungs, D = m.dim, NP = m.n_props;
#define XP_SM(T, member) reinterpret_cast<T *>(smem_raw + lay.member)
  XShared L;
  L.sx = L.rowA = L.rowB = nullptr; L.split = L.udraw = nullptr; L.iswap = nullptr;
  L.sll = XP_SM(double, sll); L.slpost = XP_SM(double, slpost); L.slprior = XP_SM(double, slprior); L.sbeta = XP_SM(double, sbeta);
  L.n_lpost = XP_SM(double, n_lpost); L.n_beta = XP_SM(double, n_beta); L.app_lpost = XP_SM(double, app_lpost); L.app_beta = XP_SM(double, app_beta);
  L.sbins = XP_SM(double, sbins); L.scount = XP_SM(long long, scount); L.saccept = XP_SM(long long, saccept);
  L.perm = XP_SM(int, perm); L.napp = XP_SM(int, napp); L.app_src = XP_SM(int, app_src); L.dir = XP_SM(int, dir); L.ups = XP_SM(int, ups);
  L.downs = XP_SM(int, downs); L.inst = XP_SM(int, inst);
  XPShared P;
  P.Ms = XP_SM(double, Ms); P.Cs = XP_SM(double, Cs); P.bufA = XP_SM(double, bufA); P.bufB = XP_SM(double, bufB);
  P.plo = XP_SM(double, plo); P.phi = XP_SM(double, phi); P.kindflag = XP_SM(int, kindflag);
  P.gsig = XP_SM(int, gsig); P.gaxis = XP_SM(int, gaxis); P.park = XP_SM(XPPark, park);
#undef XP_SM
  const int DS = lay.DS, RP = lay.RP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rung = warp;
  const int swap_warp = R - 1;                    // this warp runs the ladder's swap phase while the others run the rotation
  const long long ladder = blockIdx.x;
  const long long chain = ladder * R + rung;
  // ---- one-time staging: matrices, zero / padding rows, box edges, bins
  {
    const double *__restrict__ Mg = m.prop_data + trans_off, *__restrict__ Cg = m.ldata;
    for (int i = threadIdx.x; i < D * DS; i += blockDim.x) {
      const int r = i / DS, c = i - r * DS;
      P.Ms[i] = (c < D) ? __ldg(Mg + (size_t)r * D + c) : 0.0;
      P.Cs[i] = (c < D) ? __ldg(Cg + (size_t)r * D + c) : 0.0;
    }
    for (int i = threadIdx.x; i < (RP + 2) * DS; i += blockDim.x) { P.bufA[i] = 0; P.bufB[i] = 0; }
    for (int c = threadIdx.x; c < 32 * CPL; c += blockDim.x) {
      P.plo[c] = (c < D && m.all_uniform_prior) ? m.prior_w[c].a : -CUDART_INF;
      P.phi[c] = (c < D && m.all_uniform_prior) ? m.prior_w[c].b : CUDART_INF;
    }
    for (int i = threadIdx.x; i < R * NP; i += blockDim.x) L.sbins[i] = m.bins[i];
  }
  double *rowA = P.bufA + (size_t)rung * DS, *rowB = P.bufB + (size_t)rung * DS;
  double *myscratch = xscratch + ((size_t)blockIdx.x * R + rung) * (size_t)(32 * CPL);
  const double *ladscratch = xscratch + (size_t)blockIdx.x * R * (size_t)(32 * CPL);

  XChain<CPL> ch;
  ch.chain = chain;
  Stream<MODE> rs;
  {
#pragma unroll
    for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; ch.x[k] = (c < D) ? s.cur_x[(long long)c * m.n_chains + chain] : 0.0; }
    ch.lpost = s.lpost[chain]; ch.llike = s.llike[chain]; ch.lprior = s.lprior[chain]; ch.beta = s.beta[chain]; ch.map_lpost = s.map_lpost[chain];
    ch.nhist = s.nhist[chain]; ch.nsize = s.nsize[chain]; ch.ntries = s.ntries[chain]; ch.naccept = s.naccept[chain]; ch.last_type = s.last_type[chain];
    ch.slot = (int)(ch.nsize % m.hist_cap); ch.since_save = (int)(ch.nhist % m.save_every);
    stream_open<MODE>(m, s, rs, chain, (uint64_t)(m.ladder_offset + ladder) * PTG_STREAM_STRIDE + (uint64_t)rung, PTG_DOMAIN_STEP);
    if (lane == 0) {
      L.dir[rung] = s.directions[chain]; L.ups[rung] = s.ups[chain]; L.downs[rung] = s.downs[chain]; L.inst[rung] = s.instances[chain];
      L.scount[rung] = 0; L.saccept[rung] = 0;
    }
  }
  XPPark *const PK = P.park + rung;
  // park: lane 0 writes the warp-uniform scalars the next stages do not need; unpark: every lane reads them back (broadcast)
#define XP_PARK() do { if (lane == 0) { PK->lprior = ch.lprior; PK->map_lpost = ch.map_lpost; PK->nhist = ch.nhist; PK->ntries = ch.ntries; PK->naccept = ch.naccept; \
                                        PK->last_type = ch.last_type; PK->since_save = ch.since_save; } __syncwarp(); } while (0)
#define XP_UNPARK() do { ch.lprior = PK->lprior; ch.map_lpost = PK->map_lpost; ch.nhist = PK->nhist; ch.ntries = PK->ntries; ch.naccept = PK->naccept; \
                         ch.last_type = PK->last_type; ch.since_save = PK->since_save; } while (0)
  XP_PARK();
  const int maxswaps = m.maxswaps;
  const double swap_thresh = (R - 1) * m.swap_rate / maxswaps;
  double ptry = 2 * m.swap_rate; if (ptry > 1) ptry = 1;
  const double *bins = L.sbins + (size_t)rung * NP;
  const uint64_t ladder_stream = (uint64_t)(m.ladder_offset + ladder) * PTG_STREAM_STRIDE + PTG_STREAM_LADDER;
  __syncthreads();

  for (int it = 0; it < n_steps; it++) {
    const uint64_t step = (uint64_t)(step0 + it);
    // ================================================================ S1: publish + everything that does not depend on the swap outcome
    int type = 0, member = 0, kind = 0, i1 = 0, i2 = 0, iz = 0, az = 0;
    bool snooker = false, need_t = false;
    double gamma = 0;
    uint32_t wB[4] = {0u, 0u, 0u, 0u};
    int hsize = 0;
    {
#pragma unroll
      for (int k = 0; k < CPL; k++) myscratch[CPL * lane + k] = ch.x[k];      // states of swapped rungs travel through this row
      if (lane == 0) {
        L.sll[rung] = ch.llike; L.slpost[rung] = ch.lpost; L.slprior[rung] = PK->lprior; L.sbeta[rung] = ch.beta;
        L.n_lpost[rung] = ch.lpost; L.n_beta[rung] = ch.beta; L.perm[rung] = rung; L.napp[rung] = 0;
      }
      rs.step = step;
      hsize = (int)(ch.nsize > m.hist_cap ? (long long)m.hist_cap : ch.nsize);
      uint32_t wA[4];
      { // the two blocks of the step: even lanes generate A, odd lanes B, then everybody takes both (instead of 32 lanes generating both)
        uint32_t q[4];
        rs.fetch((lane & 1) ? PTG_BLK_B : PTG_BLK_A, q);
#pragma unroll
        for (int i = 0; i < 4; i++) { wA[i] = __shfl_sync(0xffffffffu, q[i], 0); wB[i] = __shfl_sync(0xffffffffu, q[i], 1); }
      }
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
      kind = p.kind;
      if (kind == PTG_PROP_DE) {
        ch.map_lpost = PK->map_lpost;              // the unlikely_alpha test of the history draw reads it (proposal_distribution.cc:766)
        const double usnk = ptg_u32_to_unit(wA[1]), ug = ptg_u32_to_unit(wA[2]);
        snooker = p.snooker > usnk;
        int a1 = 0, a2 = 0;
        if (!snooker) { gamma = p.gamma_std; if (ug < p.g1frac) gamma = 1; }
        else {
          gamma = (1.2 + ug) / p.reduce_gamma;
          iz = xde_index<CPL, MODE>(m, s, ch, p, rs, wB[1], 0, hsize, az);
        }
        i1 = xde_index<CPL, MODE>(m, s, ch, p, rs, wA[3], 1, hsize, a1);
        i2 = xde_index<CPL, MODE>(m, s, ch, p, rs, wB[0], 2, hsize, a2);
        // pull the records towards the SM while the rotation runs: one 128-byte line per lane
        const int nline = (D * 8 + 127) >> 7;
        if (lane < nline) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(xhist<CPL>(m, s, ch, i1) + lane * 16));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(xhist<CPL>(m, s, ch, i2) + lane * 16));
          if (snooker) asm volatile("prefetch.global.L2 [%0];" ::"l"(xhist<CPL>(m, s, ch, iz) + lane * 16));
        }
      } else if (kind == PTG_PROP_GAUSS) {
        int ia = -1;
        if (p.one_d_frac > 0 && ptg_u32_to_unit(wA[1]) < p.one_d_frac) { ia = (int)(D * ptg_u32_to_unit(wA[2])); type = 1; }
        need_t = p.has_transform && p.trans_off == trans_off;
        if (need_t && POOL) {
          // the offsets z o sigma of the rotated proposals are generated after the barrier by ALL warps together (one Box-Muller pair per
          // lane, below) instead of 50 pairs by this warp while the differential-evolution warps wait
          if (lane == 0) { P.gsig[rung] = p.sigma_off; P.gaxis[rung] = ia; }
        } else {
          double off[CPL];
          xnormals<CPL, MODE>(m, rs, off, lane);
          const double *__restrict__ sig = m.prop_data + p.sigma_off;
#pragma unroll
          for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; off[k] = (c < D) ? off[k] * __ldg(sig + c) + 0.0 : 0.0; }
          if (ia >= 0) {
#pragma unroll
            for (int k = 0; k < CPL; k++) if (CPL * lane + k != ia) off[k] = 0.0;
          }
          if (need_t) {
#pragma unroll
            for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) rowA[c] = off[k]; }
          } else {
            if (p.has_transform) xtransform<CPL>(m, m.prop_data + p.trans_off, off, rowB, lane);   // a second, different matrix: exact per-warp path
            // an offset that needs no batched rotation waits in the chain's own T row (the rotation writes only rows of rotated proposals)
            __syncwarp();
#pragma unroll
            for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) rowB[c] = off[k]; }
          }
        }
      }
      if (lane == 0) P.kindflag[rung] = need_t ? 1 : 0;
    }
    __syncthreads();
    // ================================================================ S2a: pooled Gaussian offsets | swap phase on the hottest rung's warp
    {
      const unsigned mask = __ballot_sync(0xffffffffu, lane < R && P.kindflag[lane] == 1);   // rows of the rotation, the same list in every warp
      if (R > 1 && warp == swap_warp) xswap_warp(m, L, ladder_stream, step, lane, maxswaps, swap_thresh, ptry);
      else if (!POOL) {
        if (mask) xp_dmma(P.bufA, P.Ms, P.bufB, mask, D, DS, RP, warp, R > 1 ? R - 1 : 1, lane);
      }
      else if (mask) {
        // pooled offsets: work item (j-th rotated chain, Box-Muller pair q) -> components 2q, 2q+1 of that chain's OFF row, drawn from the
        // OWNER's stream address (PTG_BLK_NORMAL + q), so the values are those of the owner's own draw_normals
        const int npair = (D + 1) >> 1, total = __popc(mask) * npair;
        const int nthr = 32 * (R > 1 ? R - 1 : 1);
        for (int item = warp * 32 + lane; item < total; item += nthr) {
          const int j = item / npair, q = item - j * npair;
          const int owner = (int)__fns(mask, 0, j + 1);
          uint32_t w4[4];
          ptg_philox_draw(m.seed, (uint64_t)(m.ladder_offset + ladder) * PTG_STREAM_STRIDE + (uint64_t)owner, PTG_DOMAIN_STEP, step, PTG_BLK_NORMAL + (uint32_t)q, w4);
          double z0, z1;
          box_muller(w4, z0, z1);
          const double *__restrict__ sig = m.prop_data + P.gsig[owner];
          const int ia = P.gaxis[owner], c0i = 2 * q, c1i = 2 * q + 1;
          double v0 = z0 * __ldg(sig + c0i) + 0.0;
          if (ia >= 0 && c0i != ia) v0 = 0.0;
          P.bufA[(size_t)owner * DS + c0i] = v0;
          if (c1i < D) {
            double v1 = z1 * __ldg(sig + c1i) + 0.0;
            if (ia >= 0 && c1i != ia) v1 = 0.0;
            P.bufA[(size_t)owner * DS + c1i] = v1;
          }
        }
      }
      if (POOL) {
        __syncthreads();
        // ============================================================ S2b: rotation T = OFF M^T on the tensor cores, all warps
        if (mask) xp_dmma(P.bufA, P.Ms, P.bufB, mask, D, DS, RP, warp, R, lane);
      }
    }
    __syncthreads();
    // ================================================================ S3: swap outcome; proposal, enforce, prior, gate
    const int na = L.napp[rung];
    ch.beta = L.n_beta[rung];
    const bool mh = (na == 0);
    ch.llike = L.sll[rung]; ch.lpost = L.n_lpost[rung];   // an MH rung's own published values (pried if the ladder evolves); swapped rungs overwrite them below
    double newx[CPL];
    double prop_lh = 0;
#pragma unroll
    for (int k = 0; k < CPL; k++) newx[k] = ch.x[k];
    if (!mh) {
      // xswapped_rung with the published states in the global scratch rows
      XP_UNPARK();
      for (int k = 0; k < 2 && k < na; k++) {
        const int src = L.app_src[2 * rung + k];
        double xs[CPL];
#pragma unroll
        for (int q = 0; q < CPL; q++) xs[q] = ladscratch[(size_t)src * (32 * CPL) + CPL * lane + q];
        xappend<CPL>(m, s, ch, xs, L.sll[src], L.app_lpost[2 * rung + k], L.app_beta[2 * rung + k], lane);
      }
      const int src = L.perm[rung];
#pragma unroll
      for (int q = 0; q < CPL; q++) ch.x[q] = ladscratch[(size_t)src * (32 * CPL) + CPL * lane + q];
      ch.llike = L.sll[src]; ch.lprior = L.slprior[src]; ch.lpost = L.n_lpost[rung];
      XP_PARK();
    }
    const double oldlprior = ch.lpost - ch.beta * ch.llike;
    bool valid = m.zero_valid != 0, gate = false;
    double newlprior = -CUDART_INF;
    if (mh) {
      const PtgProp &p = m.props[member];
      if (kind == PTG_PROP_DE) {
        if (!snooker) {
          double a[CPL], b[CPL];
          xload_rec<CPL>(xhist<CPL>(m, s, ch, i1), a, lane, D);
          xload_rec<CPL>(xhist<CPL>(m, s, ch, i2), b, lane, D);
#pragma unroll
          for (int k = 0; k < CPL; k++) { const double t = ch.x[k] + a[k] * gamma; newx[k] = t + b[k] * (-gamma); }
        } else {
          double smznorm2 = 0, minusz[CPL], smz[CPL], t[CPL];
          int isafe = 0;
          while (true) {
            double zz[CPL];
            xload_rec<CPL>(xhist<CPL>(m, s, ch, iz), zz, lane, D);
#pragma unroll
            for (int k = 0; k < CPL; k++) { minusz[k] = zz[k] * (-1); smz[k] = ch.x[k] + minusz[k]; t[k] = smz[k] * smz[k]; }
            smznorm2 = xsum_tree<CPL>(t);
            if (++isafe > 1000 || smznorm2 != 0) break;
            ch.map_lpost = PK->map_lpost;
            iz = xde_index<CPL, MODE>(m, s, ch, p, rs, wB[1], 0, hsize, az);
          }
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
      } else if (kind == PTG_PROP_GAUSS) {
        double offs[CPL];
#pragma unroll
        for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; offs[k] = (c < D) ? rowB[c] : 0.0; }
#pragma unroll
        for (int k = 0; k < CPL; k++) newx[k] = ch.x[k] + offs[k];
      }
      if (m.wrap_in_set) type = member + 10 * type;
      if (valid) valid = xenforce<CPL>(m, newx, lane);
      if (m.all_uniform_prior) {
        bool in = valid;
#pragma unroll
        for (int k = 0; k < CPL; k++) in = in && !(newx[k] < P.plo[CPL * lane + k]) && !(newx[k] > P.phi[CPL * lane + k]);
        newlprior = __all_sync(0xffffffffu, in) ? m.uniform_lprior : -CUDART_INF;
      } else newlprior = xprior<CPL>(m, newx, valid, rowB, lane);
      gate = valid && ((newlprior > -1e200) || (newlprior - oldlprior > m.dprior_min));
    }
    __syncwarp();
    if (mh && gate) {
#pragma unroll
      for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) rowA[c] = newx[k]; }
    }
    if (lane == 0) P.kindflag[rung] = (mh && gate) ? 2 : 0;
    __syncthreads();
    // ================================================================ S4: quadratic form rows Y = NEWX Cinv^T on the tensor cores
    {
      const unsigned mask = __ballot_sync(0xffffffffu, lane < R && P.kindflag[lane] == 2);
      if (mask) xp_dmma(P.bufA, P.Cs, P.bufB, mask, D, DS, RP, warp, R, lane);
    }
    __syncthreads();
    // ================================================================ S5: quadratic form, Metropolis test, append
    double newlike = -CUDART_INF, newlpost = -CUDART_INF;
    if (mh && gate) {
      double t[CPL];
#pragma unroll
      for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; t[k] = (c < D) ? newx[k] * rowB[c] : 0.0; }
      newlike = __ldg(m.lparams) - 0.5 * xsum_tree<CPL>(t);
      if (!isfinite(newlike)) newlike = -CUDART_INF;
    }
    double lhr = 0; int code = PTG_TRACE_SWAPPED;
    if (mh) {
      code = 0;
      bool accept = true;
      XP_UNPARK();
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
      XP_PARK();
    }
    if (lane == 0 && (long long)step < m.trace_steps) {
      s.trace_lhr[step * m.n_chains + chain] = lhr;
      s.trace_code[step * m.n_chains + chain] = code;
    }
  }
  __syncthreads();
  XP_UNPARK();
#undef XP_PARK
#undef XP_UNPARK
#pragma unroll
  for (int k = 0; k < CPL; k++) { const int c = CPL * lane + k; if (c < D) s.cur_x[(long long)c * m.n_chains + chain] = ch.x[k]; }
  if (lane == 0) {
    s.lpost[chain] = ch.lpost; s.llike[chain] = ch.llike; s.lprior[chain] = ch.lprior; s.beta[chain] = ch.beta; s.map_lpost[chain] = ch.map_lpost;
    s.nhist[chain] = ch.nhist; s.nsize[chain] = ch.nsize; s.ntries[chain] = ch.ntries; s.naccept[chain] = ch.naccept; s.last_type[chain] = ch.last_type;
    stream_close<MODE>(s, rs, chain);
    s.directions[chain] = L.dir[rung]; s.ups[chain] = L.ups[rung]; s.downs[chain] = L.downs[rung]; s.instances[chain] = L.inst[rung];
    s.swap_count[chain] += L.scount[rung]; s.swap_accept[chain] += L.saccept[rung];
  }
}
