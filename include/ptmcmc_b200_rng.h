/* ptmcmc_b200_rng.h -- the published Philox draw layout of the ptg engine.
 *
 * The reference gives every chain its own sequential generator (chain.hh:45-70, MotherOfAll,
 * newran1.cxx:383-432) and consumes it in a data-dependent order.  The engine replaces
 * it by the counter-based Philox4x32-10 generator (Salmon et al., SC'11): every LOGICAL draw of the
 * hot path has a fixed address (stream, domain, step, block, word), so results do not depend on launch
 * partitioning, thread scheduling or the number of GPUs, and any third party can reproduce a stream
 * from this header alone.  Plain C, no device code: shared by the CUDA kernels and by the test oracle.
 *
 *   key      = (seed_lo, seed_hi)
 *   counter  = ( block,
 *                step & 0xffffffff,
 *                stream & 0xffffffff,
 *                ((step >> 32) & 0xfff) << 20 | domain << 16 | ((stream >> 32) & 0xffff) )
 *   stream   = global_ladder * 128 + rung          for a chain's own generator
 *            = global_ladder * 128 + 127           for the ladder's generator (swap scheduling/tests)
 *   domain   = 0 stepping (step = PT iteration), 1 initialisation (step = index of the prior draw), 2 cross-GPU boundary swaps
 *
 * Like MotherOfAll::Next() = (seed+0.5)/2^32 (newran1.cxx:432) every uniform lies in the OPEN interval
 * (0,1): u32 = (w+0.5)*2^-32, u52 = ((w0<<20 | w1>>12)+0.5)*2^-52.
 */
#ifndef PTMCMC_B200_RNG_H
#define PTMCMC_B200_RNG_H
#include <stdint.h>

#define PTG_STREAM_STRIDE 128
#define PTG_STREAM_LADDER 127
#define PTG_DOMAIN_STEP 0
#define PTG_DOMAIN_INIT 1
#define PTG_DOMAIN_BOUNDARY 2   /* rung-sharded ladders: swap trial across a GPU boundary; step = exchange index, block = boundary id,
                                   (w0,w1) u_swap (u52), drawn from the ladder's stream under the SHARED key */

/* blocks of a chain's stream within one MH step (domain 0): two blocks carry every draw of a common step */
#define PTG_BLK_A 0           /* w0 u_sel | w1 u_snooker (DE) / u_1d (GAUSS) | w2 u_gamma (DE) / u_axis (GAUSS) | w3 DE index s1  (u32) */
#define PTG_BLK_B 1           /* w0 DE index s2 | w1 DE index z (snooker), attempt 0                                    (u32)
                                 (w2,w3) u_acc, the Metropolis draw                                                      (u52) */
/* DE history index `which` (0 = z, 1 = s1, 2 = s2), attempt 0: block and word */
#define PTG_IDX_BLK(which) ((which) == 1 ? PTG_BLK_A : PTG_BLK_B)
#define PTG_IDX_WORD(which) ((which) == 1 ? 3 : ((which) == 2 ? 0 : 1))
#define PTG_BLK_NEST 2        /* w0 : selection draw of a NESTED proposal set (ptg_set_nested_set)                                   (u32) */
#define PTG_BLK_NORMAL 0x100  /* + j/2 : Box-Muller pair -> normals j, j+1 ; (w0,w1)=u_a (w2,w3)=u_b      (u52) */
#define PTG_BLK_RETRY 0x200   /* + which*0x100 + a : DE index attempt a>=1 (w0) and unlikely-alpha test (w1) (u32) */
#define PTG_BLK_PRIOR 0x600   /* + i : prior draw of dimension i; (w0,w1)=u (w2,w3)=u_b for Gaussian dims  (u52) */
#define PTG_BLK_MIX 0x10000   /* + which*0x1000 + k/4, word k%4 : k-th uniform of the temperature-mixing history draw `which`
                                 (draw_from_chain with support_mixing: 20 n_rungs + 2 uniforms, in the reference's order)   (u32) */
/* blocks of a ladder's stream within one PT step: trial j -> w0 u_try, w1 u_pair (u32), (w2,w3) u_swap (u52) */
#define PTG_BLK_SWAP_EVENODD 0x100 /* + lower rung : (w0,w1) u_try, (w2,w3) u_swap                           (u52) */
/* initialisation (domain 1): block = attempt*0x100 + i, same word use as PTG_BLK_PRIOR */
#define PTG_INIT_ATTEMPT_STRIDE 0x100

#define PTG_PHILOX_M0 0xD2511F53u
#define PTG_PHILOX_M1 0xCD9E8D57u
#define PTG_PHILOX_W0 0x9E3779B9u
#define PTG_PHILOX_W1 0xBB67AE85u

#if defined(__CUDACC__)
#define PTG_HD __host__ __device__ __forceinline__
#else
#define PTG_HD static inline
#endif

#define PTG_PRAGMA_(x) _Pragma(#x)
#define PTG_PRAGMA_UNROLL(n) PTG_PRAGMA_(unroll n)
PTG_HD void ptg_philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                              uint32_t out[4]) {
#if defined(__CUDA_ARCH__) && defined(PTG_PHILOX_UNROLL)
  PTG_PRAGMA_UNROLL(PTG_PHILOX_UNROLL)
#endif
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)PTG_PHILOX_M0 * c0;
    uint64_t p1 = (uint64_t)PTG_PHILOX_M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += PTG_PHILOX_W0; k1 += PTG_PHILOX_W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

PTG_HD void ptg_philox_draw(uint64_t seed, uint64_t stream, int domain, uint64_t step, uint32_t block, uint32_t out[4]) {
  uint32_t c3 = (uint32_t)(((step >> 32) & 0xfffu) << 20) | ((uint32_t)domain << 16) | (uint32_t)((stream >> 32) & 0xffffu);
  ptg_philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), block, (uint32_t)step, (uint32_t)stream, c3, out);
}

PTG_HD double ptg_u32_to_unit(uint32_t w) { return ((double)w + 0.5) * 2.3283064365386962890625e-10; /* 2^-32 */ }
PTG_HD double ptg_u52_to_unit(uint32_t w0, uint32_t w1) {
  uint64_t v = ((uint64_t)w0 << 20) | (uint64_t)(w1 >> 12);
  return ((double)v + 0.5) * 2.220446049250313080847263336181640625e-16; /* 2^-52 */
}
#endif
