// ptg_device.cuh -- device functions of the ptg engine: RNG streams, boundary enforcement, priors,
// likelihood functors.  sm_100a only.  This translation unit is compiled with -fmad=false: the reference's
// arithmetic is unfused IEEE fp64 in a fixed operation order and every expression below is
// written in that order, so states, proposals and decisions are bit-exact under injected draws; only the
// libm calls (log, exp, sin, cos) may differ from glibc in the last ulp.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "ptg_types.h"
#include "../../include/ptmcmc_b200_rng.h"

#define PTG_PI 3.14159265358979323846

// ------------------------------------------------------------------------------------------------- RNG
// One logical generator (a chain's, chain.hh:45, or a ladder's).  MODE = PTG_RNG_PHILOX: addressed
// counter-based draws (include/ptmcmc_b200_rng.h).  MODE = PTG_RNG_TAPE: injected draws consumed in the
// reference's sequential order (SURVEY.md 8c draw-order contract).
template <int MODE>
struct Stream {
  // philox
  uint64_t seed, id, step;
  int domain;
  // tape
  const double *ut, *zt;
  long long upos, uend, zpos, zend;
  int err;

  __device__ __forceinline__ void fetch(uint32_t blk, uint32_t w[4]) const {
    if constexpr (MODE == PTG_RNG_PHILOX) ptg_philox_draw(seed, id, domain, step, blk, w);
  }
  __device__ __forceinline__ double next_u() {
    if (upos >= uend) { err = 1; return 0.5; }
    return ut[upos++];
  }
  __device__ __forceinline__ double next_z() {
    if (zpos >= zend) { err = 1; return 0.0; }
    return zt[zpos++];
  }
  __device__ __forceinline__ double u32(const uint32_t w[4], int word) {
    if constexpr (MODE == PTG_RNG_PHILOX) return ptg_u32_to_unit(w[word]);
    else return next_u();
  }
  __device__ __forceinline__ double u52(const uint32_t w[4], int pair) {
    if constexpr (MODE == PTG_RNG_PHILOX) return ptg_u52_to_unit(w[2 * pair], w[2 * pair + 1]);
    else return next_u();
  }
};

// Box-Muller pair from one Philox block: (w0,w1)=u_a, (w2,w3)=u_b
__device__ __forceinline__ void box_muller(const uint32_t w[4], double &z0, double &z1) {
  double ua = ptg_u52_to_unit(w[0], w[1]), ub = ptg_u52_to_unit(w[2], w[3]);
  double r = sqrt(-2.0 * log(ua));
  double s, c;
  sincospi(2.0 * ub, &s, &c);
  z0 = r * c; z1 = r * s;
}

template <int D, int MODE>
__device__ __forceinline__ void draw_normals(Stream<MODE> &rs, double z[D]) {
  if constexpr (MODE == PTG_RNG_PHILOX) {
#pragma unroll
    for (int j = 0; j < D; j += 2) {
      uint32_t w[4]; rs.fetch(PTG_BLK_NORMAL + j / 2, w);
      double z0, z1; box_muller(w, z0, z1);
      z[j] = z0; if (j + 1 < D) z[j + 1] = z1;
    }
  } else {
#pragma unroll
    for (int j = 0; j < D; j++) z[j] = rs.next_z();
  }
}

// ------------------------------------------------------------------------------------------------- state space
// boundary::enforce, states.cc:11-58
__device__ __forceinline__ bool bound_enforce(int lt, int ut, double xmin, double xmax, double &x) {
  if ((lt == PTG_BOUND_WRAP) != (ut == PTG_BOUND_WRAP)) return false;
  else if (lt == PTG_BOUND_WRAP) {
    double width = xmax - xmin;
    if (width <= 0) return false;
    double xt = fmod(x - xmin, width);
    if (xt < 0) xt += width;
    x = xmin + xt;
    return true;
  }
  if (lt == PTG_BOUND_REFLECT && ut == PTG_BOUND_REFLECT) {
    double halfwidth = xmax - xmin;
    if (halfwidth <= 0) return false;
    double width = 2 * halfwidth;
    double xt = fmod(x - xmin, width);
    if (xt < 0) xt += width;
    if (xt >= halfwidth) xt = halfwidth - xt;
    x = xmin + xt;
    return true;
  }
  if (lt == PTG_BOUND_REFLECT && x < xmin) x = xmin + (xmin - x);
  else if (ut == PTG_BOUND_REFLECT && x > xmax) x = xmax - (x - xmax);
  if (lt == PTG_BOUND_LIMIT && x < xmin) return false;
  if (ut == PTG_BOUND_LIMIT && x > xmax) return false;
  return true;
}
// stateSpace::enforce, states.cc:86-102
template <int D>
__device__ __forceinline__ bool space_enforce(const PtgModel &m, double x[D]) {
  bool ok = true;
#pragma unroll
  for (int i = 0; i < D; i++) {
    if (ok) {
      int lt = m.lower[i], ut = m.upper[i];
      if (lt != PTG_BOUND_OPEN || ut != PTG_BOUND_OPEN) ok = bound_enforce(lt, ut, m.xmin[i], m.xmax[i], x[i]);
    }
  }
  return ok;
}

// ------------------------------------------------------------------------------------------------- prior
// ProbabilityDist.h:88-93,126-130,153-156,197-201,243-247
__device__ __forceinline__ double pdf1d(const PtgPrior1D &p, double x) {
  switch (p.kind) {
  case PTG_PRIOR_UNIFORM:
    if (x < p.a) return 0;
    if (x > p.b) return 0;
    return 1 / (p.b - p.a);
  case PTG_PRIOR_GAUSSIAN:
  case PTG_PRIOR_GAUSSIAN_WRAPPED: {
    double xnorm = (x - p.a) / p.b;
    return exp(-xnorm * xnorm / 2) / sqrt(2 * PTG_PI) / p.b;
  }
  case PTG_PRIOR_POLAR:
    if (x < p.a) return 0;
    if (x > p.b) return 0;
    return sin(x) / p.norm;
  case PTG_PRIOR_COPOLAR:
    if (x < p.a) return 0;
    if (x > p.b) return 0;
    return cos(x) / p.norm;
  case PTG_PRIOR_LOG:
    if (x < p.a) return 0;
    if (x > p.b) return 0;
    return 1 / (p.lb - p.la) / x;
  }
  return 0;
}
// one factor of the prior product: pdf1d, plus the images of a wrapped dimension for a wrapped Gaussian factor
// (gaussian_dist_product::evaluate with wrap_probability, probability_function.cc:57-78).  The image sum is a rare path kept out of line
// (one copy per translation unit instead of one per dimension of every unrolled prior product).
static __device__ __noinline__ double prior_wrap_images(double x0, double sigma, double x, double width, double resulti) {
  const double tol = 1e-12;
  double xplus = x, xminus = x, delta = 1;
  int count = 0;
  while (delta > tol && count < 100) {
    xplus += width;
    xminus -= width;
    const double xm = (xminus - x0) / sigma, xp = (xplus - x0) / sigma;
    delta = exp(-xm * xm / 2) / sqrt(2 * PTG_PI) / sigma + exp(-xp * xp / 2) / sqrt(2 * PTG_PI) / sigma;
    resulti += delta;
    count++;
  }
  return resulti;
}
__device__ __forceinline__ double prior_factor(const PtgPrior1D &p, double x, int lt, int ut, double xmin, double xmax) {
  double resulti = pdf1d(p, x);
  if (p.kind == PTG_PRIOR_GAUSSIAN_WRAPPED && lt == PTG_BOUND_WRAP && ut == PTG_BOUND_WRAP) resulti = prior_wrap_images(p.a, p.b, x, xmax - xmin, resulti);
  return resulti;
}
__device__ __forceinline__ double invcdf1d(const PtgPrior1D &p, double u) {
  switch (p.kind) {
  case PTG_PRIOR_UNIFORM: return (u * (p.b - p.a) + p.a);
  case PTG_PRIOR_POLAR: return acos(-p.norm * (u + p.cdfoff));
  case PTG_PRIOR_COPOLAR: return asin(p.norm * (u + p.cdfoff));
  case PTG_PRIOR_LOG: return exp(u * (p.lb - p.la) + p.la);
  }
  return CUDART_NAN;
}
// evaluate_log = log(prod_i pdf_i) (probability_function.hh:59); 0 probability for an invalid state
template <int D>
__device__ __forceinline__ double prior_eval_log(const PtgModel &m, const double x[D], bool valid) {
  if (!valid) return -CUDART_INF;
  if (m.all_uniform_prior) { // the product is a constant inside the box: its log is evaluated once at set-up
    bool in = true;
#pragma unroll
    for (int i = 0; i < D; i++) in = in && !(x[i] < m.prior[i].a) && !(x[i] > m.prior[i].b);
    return in ? m.uniform_lprior : -CUDART_INF;
  }
  double result = 1;
#pragma unroll
  for (int i = 0; i < D; i++) result *= prior_factor(m.prior[i], x[i], m.lower[i], m.upper[i], m.xmin[i], m.xmax[i]);
  return log(result);
}
// drawSample (probability_function.cc:37-47,147-154,264-279), returns validity after state(space,v) enforcement
template <int D, int MODE>
__device__ __forceinline__ bool prior_draw(const PtgModel &m, Stream<MODE> &rs, uint32_t blk0, double x[D]) {
#pragma unroll
  for (int i = 0; i < D; i++) {
    uint32_t w[4]; rs.fetch(blk0 + i, w);
    if (m.prior[i].kind == PTG_PRIOR_GAUSSIAN || m.prior[i].kind == PTG_PRIOR_GAUSSIAN_WRAPPED) {
      double z;
      if constexpr (MODE == PTG_RNG_PHILOX) { double z1; box_muller(w, z, z1); }
      else z = rs.next_z();
      x[i] = z * m.prior[i].b + m.prior[i].a;
    } else {
      double u = rs.u52(w, 0);
      x[i] = invcdf1d(m.prior[i], u);
    }
  }
  return space_enforce<D>(m, x);
}

// ------------------------------------------------------------------------------------------------- adaptive shares
// proposal_distribution_set::accept / reject (proposal_distribution.cc:132-166) on this chain's clone of the set: two accepts or two
// rejects in a row of `member` scale its share by 1 - adapt_rate / 4; adapt_count is never reset in the reference, so from the
// adapt_every-th (10 n) decision on reset_bins (:37-59) runs after every decision: shares normalised in place, bins rebuilt with the
// chain's CURRENT temperature when Tpow > 0.
// sh / bn: this chain's shares and bins of ONE set (n entries), bit0 of `last`: the set's first member
__device__ __forceinline__ void set_adapt_one(double *sh, double *bn, int n, int &last, int bit0, int32_t &count, double rate, int idx, bool accepted, double Tpow,
                                              const double *hot, double beta) {
  const int bit = bit0 + idx;
  if ((((last >> bit) & 1) != 0) == accepted) sh[idx] *= 1 - rate * 0.25;
  last = accepted ? (last | (1 << bit)) : (last & ~(1 << bit));
  count = count + 1;
  if (count >= 10 * n) {
    double Tfac = 0;
    if (Tpow > 0) Tfac = 1 - pow(beta, Tpow);
    double sum = 0;
    for (int i = 0; i < n; i++) sum += sh[i];
    double lastb = 0;
    for (int i = 0; i < n; i++) {
      const double v = sh[i] / sum;
      sh[i] = v;
      double b = lastb + v;
      if (Tpow > 0) b += (hot[i] - v) * Tfac;
      bn[i] = b;
      lastb = b;
    }
    const double back = bn[n - 1];
    for (int i = 0; i < n; i++) bn[i] /= back;
  }
}
// accept() / reject() of the top-level set, handed down to a nested set when one of its members drew (nested >= 0)
__device__ __forceinline__ void set_adapt(const PtgModel &m, const PtgState &s, long long chain, int slot, int nested, bool accepted, double beta) {
  double *sh = s.ad_shares + chain * m.n_bins, *bn = s.ad_bins + chain * m.n_bins;
  int last = s.ad_last[chain];
  if (m.adapt_rate != 0) { int32_t c = s.ad_count[chain]; set_adapt_one(sh, bn, m.n_slots, last, 0, c, m.adapt_rate, slot, accepted, m.Tpow, m.hot_norm, beta); s.ad_count[chain] = c; }
  if (nested >= 0 && m.nest_adapt != 0) {
    int32_t c = s.ad_count2[chain];
    set_adapt_one(sh + m.n_slots, bn + m.n_slots, m.nest_count, last, 16, c, m.nest_adapt, nested, accepted, 0.0, m.hot_norm, beta);
    s.ad_count2[chain] = c;
  }
  s.ad_last[chain] = last;
}

// ------------------------------------------------------------------------------------------------- likelihoods
// one functor per kind (compile-time KIND: the streamlined production kernels instantiate exactly one of them)
template <int D, int KIND>
__device__ __forceinline__ double like_eval_kind(const PtgModel &m, const double x[D]) {
  const double *__restrict__ P = m.lparams;
  double result = 0;
  if constexpr (KIND == PTG_LIKE_FLAT) return 0;
  else if constexpr (KIND == PTG_LIKE_GAUSS_ISO) { // example.cc:116-143
    double r2 = 0;
#pragma unroll
    for (int i = 0; i < D; i++) { double dx = x[i] - __ldg(P + 2 + i); r2 += dx * dx; }
    result = __ldg(P) - r2 / __ldg(P + 1);
  } else if constexpr (KIND == PTG_LIKE_SHELL2D) { // example.cc:195-206: Gaussian shell in (|p0|, p1)
    double dx = fabs(x[0]) - __ldg(P + 3);
    double r2 = dx * dx;
    dx = x[D >= 2 ? 1 : 0] - __ldg(P + 4);
    r2 += dx * dx;
    dx = sqrt(r2) - __ldg(P + 2);
    r2 = dx * dx;
    result = __ldg(P) - r2 / __ldg(P + 1);
  } else if constexpr (KIND == PTG_LIKE_SHELLS) { // example.cc:373-403: the better of the "plus" and the "minus" shell
    const double twosigmasq = __ldg(P + 1), r0 = __ldg(P + 2), x0 = __ldg(P + 3), spm = __ldg(P + 4), lnspm = __ldg(P + 5);
    double xx = x[0];
    if (__ldg(P + 6) != 0) {
      if (xx < 0) return -CUDART_INF;
      xx = log(xx);
    }
    double dx = xx - x0;
    double r2 = dx * dx;
#pragma unroll
    for (int i = 1; i < D; i++) { dx = x[i]; r2 += dx * dx; }
    dx = sqrt(r2) - r0;
    r2 = dx * dx;
    const double resultp = -r2 / (twosigmasq * spm) - 0.5 * lnspm;
    dx = xx + x0;
    r2 = dx * dx;
#pragma unroll
    for (int i = 1; i < D; i++) { dx = x[i]; r2 += dx * dx; }
    dx = sqrt(r2) - r0;
    r2 = dx * dx;
    const double resultm = -r2 / (twosigmasq / spm) + 0.5 * lnspm;
    result = __ldg(P);
    if (resultm > resultp) result += resultm; else result += resultp;
  } else if constexpr (KIND == PTG_LIKE_SINES) { // sines.hh:22-54
    const double height = __ldg(P), step_scale = __ldg(P + 1);
    double lprod = 0; int isum = 0;
    // rolled on purpose: ONE copy of sin() in the instruction stream and one set of temporaries (register budget and
    // instruction-cache footprint of the step kernels); x[i] is picked with selects, not by indexing (no local memory)
#ifndef PTG_SINES_UNROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int i = 0; i < D; i++) {
      double xi = x[0];
#pragma unroll
      for (int j = 1; j < D; j++) if (i == j) xi = x[j];
      const int k = (int)__ldg(P + 2 + i);
      const double mn = __ldg(P + 2 + D + i), mx = __ldg(P + 2 + 2 * D + i);
      const double xx = (xi - mn) / (mx - mn);
      double s = sin(k * PTG_PI * xx);
      s = s * s;
      lprod += (s * s - 1) * height;
      isum += (int)(xx * k);
    }
    return lprod + (-isum * step_scale);
  } else if constexpr (KIND == PTG_LIKE_POLY_CHI2) { // bayesian.hh:595-622 + poly_example.cc:85-106
    const long long N = m.n_ldata / 3;
    const double *__restrict__ xs = m.ldata, *__restrict__ ys = m.ldata + N, *__restrict__ S = m.ldata + 2 * N;
    double sum = 0;
    for (long long i = 0; i < N; i++) {
      double xi = __ldg(xs + i), y = 0, xn = 1;
#pragma unroll
      for (int j = 0; j < D; j++) { y += xn * x[j]; xn *= xi; }
      double dd = y - __ldg(ys + i);
      sum += dd * dd / __ldg(S + i);
    }
    sum += m.like_nsum;
    sum /= -2;
    result = sum - __ldg(P);
  } else if constexpr (KIND == PTG_LIKE_SINUSOID_CHI2) {
    const long long N = m.n_ldata / 3;
    const double *__restrict__ xs = m.ldata, *__restrict__ ys = m.ldata + N, *__restrict__ S = m.ldata + 2 * N;
    double sum = 0;
    for (long long i = 0; i < N; i++) {
      double ti = __ldg(xs + i), y = 0;
#pragma unroll
      for (int k = 0; k + 2 < D; k += 3) y += x[k] * sin(2 * PTG_PI * x[k + 1] * ti + x[k + 2]);
      double dd = y - __ldg(ys + i);
      sum += dd * dd / __ldg(S + i);
    }
    sum += m.like_nsum;
    sum /= -2;
    result = sum - __ldg(P);
  } else if constexpr (KIND == PTG_LIKE_GAUSS_FULLCOV) { // cython/exampleGaussian.py:103-109
    const double *__restrict__ C = m.ldata;
    double q = 0;
#pragma unroll
    for (int i = 0; i < D; i++) {
      double y = 0;
#pragma unroll
      for (int j = 0; j < D; j++) y += __ldg(C + i * D + j) * x[j];
      q += x[i] * y;
    }
    result = __ldg(P) - 0.5 * q;
  }
  if (!isfinite(result)) result = -CUDART_INF; // bayesian.hh:569-575
  return result;
}
template <int D>
__device__ __forceinline__ double like_eval(const PtgModel &m, const double x[D]) {
  switch (m.like_kind) {
  case PTG_LIKE_FLAT: return 0;
  case PTG_LIKE_GAUSS_ISO: return like_eval_kind<D, PTG_LIKE_GAUSS_ISO>(m, x);
  case PTG_LIKE_SHELL2D: if constexpr (D >= 2) return like_eval_kind<D, PTG_LIKE_SHELL2D>(m, x); else return 0;
  case PTG_LIKE_SHELLS: return like_eval_kind<D, PTG_LIKE_SHELLS>(m, x);
  case PTG_LIKE_SINES: return like_eval_kind<D, PTG_LIKE_SINES>(m, x);
  case PTG_LIKE_POLY_CHI2: return like_eval_kind<D, PTG_LIKE_POLY_CHI2>(m, x);
  case PTG_LIKE_SINUSOID_CHI2: return like_eval_kind<D, PTG_LIKE_SINUSOID_CHI2>(m, x);
  case PTG_LIKE_GAUSS_FULLCOV: return like_eval_kind<D, PTG_LIKE_GAUSS_FULLCOV>(m, x);
  }
  return 0;
}
