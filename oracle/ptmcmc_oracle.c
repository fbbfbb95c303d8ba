/* ptmcmc_oracle.c -- TEST INFRASTRUCTURE: CPU restatement of the reference's chain-stepping hot path.
 * See ptmcmc_oracle.h for the rules (checker only; never imported by the product) and the pin status.
 *
 * Every function cites the reference file:line (relative to JohnGBaker/ptmcmc) that it restates.
 * Arithmetic is written in the reference's operation order and must be compiled with
 * -ffp-contract=off so that it stays unfused IEEE fp64 like the reference build.
 */
#include "ptmcmc_oracle.h"
#include "newran_port.h"
#include "../include/ptmcmc_b200_rng.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>

static char g_err[512] = "";
const char *pto_last_error(void) { return g_err; }
static int fail(int code, const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
  return code;
}

/* ------------------------------------------------------------------------------------------------ types */
typedef struct {
  int kind;
  double a, b;          /* as passed: (xmin,xmax) or (x0,sigma) */
  double norm, cdfoff;  /* polar / copolar (ProbabilityDist.h:187-196, 228-236) */
  double la, lb;        /* log: log_xmin, log_xmax (ProbabilityDist.h:118-119) */
} prior1d_t;

typedef struct {
  int kind;
  double share, hot_share;
  double snooker, g1frac, b_small, ignore_frac, unlikely_alpha, reduce_gamma;
  double one_d_frac;
  double *sigmas, *transform;
} prop_t;

typedef struct {
  int mode;
  mother_t mom;
  const double *ut, *zt;
  int64_t upos, uend, zpos, zend;
  double *urec, *zrec;
  int64_t nu, nz, cu, cz;
  uint64_t id;
} stream_t;

typedef struct { /* one proposal_distribution_set as a rung's clone holds it (proposal_distribution.hh:308-321) */
  int n;
  double shares[PTG_MAX_PROPOSALS], bins[PTG_MAX_PROPOSALS];
  int last_accepted[PTG_MAX_PROPOSALS];
  int adapt_count;
} setstate_t;

typedef struct {
  double *x;
  double lpost, llike, beta;
  int64_t nhist, nsize, ntries, naccept;
  int last_type;
  double map_lpost;
  double *map_x;
  double *hx, *hlpost, *hllike, *hacc, *hbeta;
  int32_t *htype;
  int64_t hcap;
  stream_t rng;
  setstate_t top, nest; /* this rung's clone of the proposal set (and of its nested set) owns shares, bins and the adaptation state */
  int64_t frozen;   /* history_freeze (chain.cc:1552, chain.hh:86): the size other chains see during the MH phase of a PT step */
} chain_t;

struct pto_handle {
  ptg_config cfg;
  int d, R, L;
  int64_t nchains;
  int have_space, have_prior, have_like, have_props, inited;
  int lower[PTG_MAX_DIM], upper[PTG_MAX_DIM];
  double xmin[PTG_MAX_DIM], xmax[PTG_MAX_DIM];
  int zero_valid;
  prior1d_t prior[PTG_MAX_DIM];
  int like_kind; double *lparams; int n_lparams; double *ldata; int64_t n_ldata; double like_nsum;
  int nprops; prop_t props[PTG_MAX_PROPOSALS]; double Tpow; int wrap_in_set;
  double adapt_rate; int de_mixing; double de_Tmix; double hot_norm[PTG_MAX_PROPOSALS];
  int nest_first, nest_count; double nest_share, nest_hot, nest_adapt;
  double *betas0; /* optional explicit */
  chain_t *chains;
  stream_t *lstreams;
  int maxswaps;
  int64_t *swap_count, *swap_accept;
  int32_t *directions, *ups, *downs, *instances;
  int64_t istep, total_steps;
  double *trace_lhr; int32_t *trace_code;
  int record, tape_err;
  double *tape_u, *tape_z; /* owned copies */
  int64_t *mark_u, *mark_z, n_marks, cap_marks; /* per PT step: draws each stream had consumed when the step began */
};

/* ---------------------------------------------------------------------------------------------- streams */
static void rec_push(double **buf, int64_t *n, int64_t *cap, double v) {
  if (*n == *cap) { *cap = *cap ? *cap * 2 : 1024; *buf = (double *)realloc(*buf, (size_t)*cap * sizeof(double)); }
  (*buf)[(*n)++] = v;
}

/* next uniform of a SEQUENTIAL stream: chain::get_uniform (chain.hh:48-54) / rng.Next() sites */
static double seq_uniform(pto_handle *h, stream_t *s) {
  double u;
  if (s->mode == PTO_RNG_NEWRAN) u = mother_next(&s->mom);
  else { /* tape */
    if (s->upos >= s->uend) { h->tape_err = 1; u = 0.5; }
    else u = s->ut[s->upos++];
  }
  if (h->record) rec_push(&s->urec, &s->nu, &s->cu, u);
  return u;
}
/* next standard normal: `Normal normal; normal.Next()` (ProbabilityDist.cxx:78-79) */
static double seq_normal(pto_handle *h, stream_t *s) {
  double z;
  if (s->mode == PTO_RNG_NEWRAN) z = newran_normal(&s->mom);
  else {
    if (s->zpos >= s->zend) { h->tape_err = 1; z = 0.0; }
    else z = s->zt[s->zpos++];
  }
  if (h->record) rec_push(&s->zrec, &s->nz, &s->cz, z);
  return z;
}
static int is_philox(const stream_t *s) { return s->mode == PTG_RNG_PHILOX; }

/* addressed draws (Philox layout of include/ptmcmc_b200_rng.h); sequential modes ignore the address */
static double draw_u32(pto_handle *h, stream_t *s, int domain, uint64_t step, uint32_t blk, int word) {
  if (!is_philox(s)) return seq_uniform(h, s);
  uint32_t w[4]; ptg_philox_draw(h->cfg.seed, s->id, domain, step, blk, w);
  double u = ptg_u32_to_unit(w[word]);
  if (h->record) rec_push(&s->urec, &s->nu, &s->cu, u);
  return u;
}
static double draw_u52(pto_handle *h, stream_t *s, int domain, uint64_t step, uint32_t blk, int pair) {
  if (!is_philox(s)) return seq_uniform(h, s);
  uint32_t w[4]; ptg_philox_draw(h->cfg.seed, s->id, domain, step, blk, w);
  double u = ptg_u52_to_unit(w[2 * pair], w[2 * pair + 1]);
  if (h->record) rec_push(&s->urec, &s->nu, &s->cu, u);
  return u;
}
/* Box-Muller pair from one Philox block */
static void philox_normal_pair(pto_handle *h, stream_t *s, int domain, uint64_t step, uint32_t blk, double *z0, double *z1) {
  uint32_t w[4]; ptg_philox_draw(h->cfg.seed, s->id, domain, step, blk, w);
  double ua = ptg_u52_to_unit(w[0], w[1]), ub = ptg_u52_to_unit(w[2], w[3]);
  double r = sqrt(-2.0 * log(ua)), th = 2.0 * M_PI * ub;
  *z0 = r * cos(th); *z1 = r * sin(th);
}
/* d standard normals */
static void draw_normals(pto_handle *h, stream_t *s, uint64_t step, int d, double *z) {
  if (!is_philox(s)) { for (int j = 0; j < d; j++) z[j] = seq_normal(h, s); return; }
  for (int j = 0; j < d; j += 2) {
    double z0, z1; philox_normal_pair(h, s, PTG_DOMAIN_STEP, step, PTG_BLK_NORMAL + j / 2, &z0, &z1);
    z[j] = z0; if (j + 1 < d) z[j + 1] = z1;
    if (h->record) { rec_push(&s->zrec, &s->nz, &s->cz, z0); if (j + 1 < d) rec_push(&s->zrec, &s->nz, &s->cz, z1); }
  }
}

/* ---------------------------------------------------------------------------------------------- state space */
/* boundary::enforce, states.cc:11-58 */
static int bound_enforce(int lt, int ut, double xmin, double xmax, double *px) {
  double x = *px;
  if ((lt == PTG_BOUND_WRAP) != (ut == PTG_BOUND_WRAP)) return 0;
  else if (lt == PTG_BOUND_WRAP) {
    double width = xmax - xmin;
    if (width <= 0) return 0;
    double xt = fmod(x - xmin, width);
    if (xt < 0) xt += width;
    *px = xmin + xt;
    return 1;
  }
  if (lt == PTG_BOUND_REFLECT && ut == PTG_BOUND_REFLECT) {
    double halfwidth = xmax - xmin;
    if (halfwidth <= 0) return 0;
    double width = 2 * halfwidth;
    double xt = fmod(x - xmin, width);
    if (xt < 0) xt += width;
    if (xt >= halfwidth) xt = halfwidth - xt;
    *px = xmin + xt;
    return 1;
  }
  if (lt == PTG_BOUND_REFLECT && x < xmin) x = xmin + (xmin - x);
  else if (ut == PTG_BOUND_REFLECT && x > xmax) x = xmax - (x - xmax);
  *px = x;
  if (lt == PTG_BOUND_LIMIT && x < xmin) return 0;
  if (ut == PTG_BOUND_LIMIT && x > xmax) return 0;
  return 1;
}
/* stateSpace::enforce, states.cc:86-102 (stops at the first failing dimension) */
static int space_enforce(const pto_handle *h, double *x) {
  for (int i = 0; i < h->d; i++)
    if (!bound_enforce(h->lower[i], h->upper[i], h->xmin[i], h->xmax[i], &x[i])) return 0;
  return 1;
}

/* ---------------------------------------------------------------------------------------------- prior */
/* 1-D pdfs: ProbabilityDist.h:88-93 (uniform), :126-130 (log), :153-156 (gaussian), :197-201 (polar), :243-247 (copolar) */
static double pdf1d(const prior1d_t *p, double x) {
  switch (p->kind) {
  case PTG_PRIOR_UNIFORM:
    if (x < p->a) return 0;
    if (x > p->b) return 0;
    return 1 / (p->b - p->a);
  case PTG_PRIOR_GAUSSIAN:
  case PTG_PRIOR_GAUSSIAN_WRAPPED: {
    double xnorm = (x - p->a) / p->b;
    return exp(-xnorm * xnorm / 2) / sqrt(2 * M_PI) / p->b;
  }
  case PTG_PRIOR_POLAR:
    if (x < p->a) return 0;
    if (x > p->b) return 0;
    return sin(x) / p->norm;
  case PTG_PRIOR_COPOLAR:
    if (x < p->a) return 0;
    if (x > p->b) return 0;
    return cos(x) / p->norm;
  case PTG_PRIOR_LOG:
    if (x < p->a) return 0;
    if (x > p->b) return 0;
    return 1 / (p->lb - p->la) / x;
  }
  return 0;
}
/* invcdf: ProbabilityDist.h:94-97,131-134,202-205,248-251 */
static double invcdf1d(const prior1d_t *p, double u) {
  switch (p->kind) {
  case PTG_PRIOR_UNIFORM: return (u * (p->b - p->a) + p->a);
  case PTG_PRIOR_POLAR: return acos(-p->norm * (u + p->cdfoff));
  case PTG_PRIOR_COPOLAR: return asin(p->norm * (u + p->cdfoff));
  case PTG_PRIOR_LOG: return exp(u * (p->lb - p->la) + p->la);
  }
  return NAN;
}
/* sampleable_probability_function::evaluate_log = log(evaluate(s)) (probability_function.hh:59) with
 * evaluate = prod_i pdf_i, 0 for an invalid state (probability_function.cc:49-81,156-166,281-304) */
static double prior_eval_log(const pto_handle *h, const double *x, int valid) {
  if (!valid) return log(0.0);
  double result = 1;
  for (int i = 0; i < h->d; i++) {
    double resulti = pdf1d(&h->prior[i], x[i]);
    if (h->prior[i].kind == PTG_PRIOR_GAUSSIAN_WRAPPED && h->lower[i] == PTG_BOUND_WRAP && h->upper[i] == PTG_BOUND_WRAP) {
      /* gaussian_dist_product::evaluate with wrap_probability (probability_function.cc:57-78): add the images of a wrapped dimension */
      const double tol = 1e-12;
      double xplus = x[i], xminus = x[i], delta = 1;
      double width = h->xmax[i] - h->xmin[i];
      int count = 0;
      while (delta > tol && count < 100) {
        xplus += width;
        xminus -= width;
        delta = pdf1d(&h->prior[i], xminus) + pdf1d(&h->prior[i], xplus);
        resulti += delta;
        count++;
      }
    }
    result *= resulti;
  }
  return log(result);
}
/* drawSample (probability_function.cc:37-47,147-154,264-279) + ProbabilityDist::draw (ProbabilityDist.cxx:33-59)
 * + GaussianDist::draw (:64-89), then state(space,v) which enforces (states.cc:178-182).  Returns validity. */
static int prior_draw(pto_handle *h, stream_t *s, int domain, uint64_t step, uint32_t blk0, double *x) {
  for (int i = 0; i < h->d; i++) {
    const prior1d_t *p = &h->prior[i];
    if (p->kind == PTG_PRIOR_GAUSSIAN || p->kind == PTG_PRIOR_GAUSSIAN_WRAPPED) {
      double z;
      if (is_philox(s)) {
        double z1; philox_normal_pair(h, s, domain, step, blk0 + i, &z, &z1);
        if (h->record) rec_push(&s->zrec, &s->nz, &s->cz, z);
      } else z = seq_normal(h, s);
      x[i] = z * p->b + p->a;
    } else {
      double u = draw_u52(h, s, domain, step, blk0 + i, 0);
      x[i] = invcdf1d(p, u);
    }
  }
  return space_enforce(h, x);
}

/* ---------------------------------------------------------------------------------------------- likelihoods */
static double like_eval(const pto_handle *h, const double *x) {
  const int d = h->d;
  const double *P = h->lparams;
  double result = 0;
  switch (h->like_kind) {
  case PTG_LIKE_FLAT: return 0;
  case PTG_LIKE_GAUSS_ISO: { /* example.cc:116-143 */
    double r2 = 0;
    for (int i = 0; i < d; i++) { double dx = x[i] - P[2 + i]; r2 += dx * dx; }
    result = P[0] - r2 / P[1];
    break;
  }
  case PTG_LIKE_SHELL2D: { /* example.cc:195-206: Gaussian shell in (|p0|, p1); P = lnnormfac, twosigmasq, r0, x0[2] */
    double r2 = 0;
    double dx = fabs(x[0]) - P[3];
    r2 += dx * dx;
    dx = x[1] - P[4];
    r2 += dx * dx;
    dx = sqrt(r2) - P[2];
    r2 = dx * dx;
    result = P[0] - r2 / P[1];
    break;
  }
  case PTG_LIKE_SHELLS: { /* example.cc:373-403; P = lnnormfac, twosigmasq, r0, x0, sigmapoverm, lnsigmapoverm, logx */
    const double twosigmasq = P[1], r0 = P[2], x0 = P[3], spm = P[4], lnspm = P[5];
    double xx = x[0];
    if (P[6] != 0) {
      if (xx < 0) return -INFINITY;
      xx = log(xx);
    }
    double dx = xx - x0;
    double r2 = dx * dx;
    for (int i = 1; i < d; i++) { dx = x[i]; r2 += dx * dx; }
    dx = sqrt(r2) - r0;
    r2 = dx * dx;
    double resultp = -r2 / (twosigmasq * spm) - 0.5 * lnspm;
    dx = xx + x0;
    r2 = dx * dx;
    for (int i = 1; i < d; i++) { dx = x[i]; r2 += dx * dx; }
    dx = sqrt(r2) - r0;
    r2 = dx * dx;
    double resultm = -r2 / (twosigmasq / spm) + 0.5 * lnspm;
    result = P[0];
    if (resultm > resultp) result += resultm; else result += resultp;
    break;
  }
  case PTG_LIKE_SINES: { /* sines.hh:22-54 */
    const double height = P[0], step_scale = P[1];
    const double *ks = P + 2, *mins = P + 2 + d, *maxs = P + 2 + 2 * d;
    double lprod = 0; int isum = 0;
    for (int i = 0; i < d; i++) {
      int k = (int)ks[i];
      double xx = (x[i] - mins[i]) / (maxs[i] - mins[i]);
      double s = sin(k * M_PI * xx);
      s = s * s;
      lprod += (s * s - 1) * height;
    }
    for (int j = 0; j < d; j++) {
      int k = (int)ks[j];
      double xx = (x[j] - mins[j]) / (maxs[j] - mins[j]);
      isum += (int)(xx * k);
    }
    return lprod + (-isum * step_scale);
  }
  case PTG_LIKE_POLY_CHI2:      /* bayesian.hh:595-622 + poly_example.cc:85-106 */
  case PTG_LIKE_SINUSOID_CHI2: {
    const int64_t N = h->n_ldata / 3;
    const double *xs = h->ldata, *ys = h->ldata + N, *S = h->ldata + 2 * N;
    double sum = 0;
    for (int64_t i = 0; i < N; i++) {
      double y = 0;
      if (h->like_kind == PTG_LIKE_POLY_CHI2) {
        double xn = 1;
        for (int j = 0; j < d; j++) { y += xn * x[j]; xn *= xs[i]; }
      } else {
        for (int k = 0; k + 2 < d; k += 3) y += x[k] * sin(2 * M_PI * x[k + 1] * xs[i] + x[k + 2]);
      }
      double dd = y - ys[i];
      sum += dd * dd / S[i];
    }
    sum += h->like_nsum; /* nsum = sum_i log(S_i) does not depend on the state */
    sum /= -2;
    result = sum - P[0];
    break;
  }
  case PTG_LIKE_GAUSS_FULLCOV: { /* cython/exampleGaussian.py:103-109 */
    const double *C = h->ldata;
    double q = 0;
    for (int i = 0; i < d; i++) {
      double y = 0;
      for (int j = 0; j < d; j++) y += C[(size_t)i * d + j] * x[j];
      q += x[i] * y;
    }
    result = P[0] - 0.5 * q;
    break;
  }
  }
  if (!isfinite(result)) result = -INFINITY; /* bayesian.hh:569-575 ; example.cc:64-68,134-138 */
  return result;
}

/* ---------------------------------------------------------------------------------------------- history */
static void hist_reserve(pto_handle *h, chain_t *c, int64_t n) {
  if (n <= c->hcap) return;
  int64_t cap = c->hcap ? c->hcap : 256;
  while (cap < n) cap *= 2;
  c->hx = (double *)realloc(c->hx, (size_t)cap * h->d * sizeof(double));
  c->hlpost = (double *)realloc(c->hlpost, (size_t)cap * sizeof(double));
  c->hllike = (double *)realloc(c->hllike, (size_t)cap * sizeof(double));
  c->hacc = (double *)realloc(c->hacc, (size_t)cap * sizeof(double));
  c->hbeta = (double *)realloc(c->hbeta, (size_t)cap * sizeof(double));
  c->htype = (int32_t *)realloc(c->htype, (size_t)cap * sizeof(int32_t));
  c->hcap = cap;
}
/* MH_chain::add_state with explicit log_like, log_post (chain.cc:916-949) */
static void add_state(pto_handle *h, chain_t *c, const double *x, double llike, double lpost) {
  c->llike = llike;
  c->lpost = lpost;
  if (x != c->x) memcpy(c->x, x, (size_t)h->d * sizeof(double));
  if (c->lpost > c->map_lpost) { c->map_lpost = c->lpost; memcpy(c->map_x, c->x, (size_t)h->d * sizeof(double)); }
  if (c->nhist % h->cfg.save_every == 0) {
    hist_reserve(h, c, c->nsize + 1);
    memcpy(c->hx + (size_t)c->nsize * h->d, c->x, (size_t)h->d * sizeof(double));
    c->hlpost[c->nsize] = c->lpost;
    c->hllike[c->nsize] = c->llike;
    c->hacc[c->nsize] = c->naccept / (double)c->ntries;
    c->hbeta[c->nsize] = c->beta;
    c->htype[c->nsize] = c->last_type;
    c->nsize++;
  }
  c->nhist++;
  h->total_steps++;
}

/* ---------------------------------------------------------------------------------------------- proposals */
typedef struct { int type; double log_hastings; int valid; int member, slot, nested; } draw_result_t;

/* differential_evolution::draw_i_from_chain (proposal_distribution.cc:744-778); `size` is the frozen chain
 * size (chain.hh:86); with a ring of capacity C only the newest min(size,C) samples are eligible. */
static int64_t de_draw_index(pto_handle *h, chain_t *c, const prop_t *p, uint64_t step, int which, int *attempt) {
  int64_t nsz = c->nsize, W = nsz, base = 0;
  if (h->cfg.hist_capacity > 0 && W > h->cfg.hist_capacity) { W = h->cfg.hist_capacity; base = nsz - W; }
  int size = (int)W, start = 0, mins = h->d * 10, minc = h->d * 100;
  if ((size - minc) * (1 - p->ignore_frac) > mins) start = (int)((size - minc) * p->ignore_frac);
  double lpost0 = c->map_lpost - h->d;
  double alpha = p->unlikely_alpha;
  while (1) {
    int a = *attempt;
    double xrnd = (a == 0) ? draw_u32(h, &c->rng, PTG_DOMAIN_STEP, step, PTG_IDX_BLK(which), PTG_IDX_WORD(which))
                           : draw_u32(h, &c->rng, PTG_DOMAIN_STEP, step, PTG_BLK_RETRY + which * 0x100 + (a & 0xff), 0);
    (*attempt)++;
    int index = (int)(start + (size - start) * xrnd);
    double lpost = c->hlpost[base + index];
    if (alpha > 0 && lpost0 > lpost) {
      double pr = exp(alpha * (lpost - lpost0));
      xrnd = draw_u32(h, &c->rng, PTG_DOMAIN_STEP, step, PTG_BLK_RETRY + which * 0x100 + (a & 0xff), 1);
      if (xrnd < pr) return base + index;
      alpha *= 0.9;
    } else return base + index;
  }
}


/* ---- temperature mixing: differential_evolution::draw_from_chain with support_mixing (proposal_distribution.cc:594-741) ----
 * Reached when a BARE differential_evolution with support_mixing(true) is handed to parallel_tempering_chains::set_proposal
 * (chain.cc:1375-1379: set_chain(this), so ch = the ladder and multiplicity() = Ntemps).  unlikely_alpha = 0 is required, so
 * draw_i_from_chain consumes exactly one uniform.  Sizes are the FROZEN sizes (history_freeze, chain.cc:1552). */
static double mix_u(pto_handle *h, chain_t *caller, uint64_t step, int which, int *k) {
  const int kk = (*k)++;
  return draw_u32(h, &caller->rng, PTG_DOMAIN_STEP, step, PTG_BLK_MIX + (uint32_t)which * 0x1000u + (uint32_t)(kk >> 2), kk & 3);
}
/* draw_i_from_chain(caller, c) (proposal_distribution.cc:744-778): window index into c's frozen history */
static int mix_draw_i(pto_handle *h, chain_t *caller, const chain_t *c, const prop_t *p, uint64_t step, int which, int *k) {
  int64_t W = c->frozen;
  if (h->cfg.hist_capacity > 0 && W > h->cfg.hist_capacity) W = h->cfg.hist_capacity;
  int size = (int)W, start = 0, mins = h->d * 10, minc = h->d * 100;
  if ((size - minc) * (1 - p->ignore_frac) > mins) start = (int)((size - minc) * p->ignore_frac);
  double xrnd = mix_u(h, caller, step, which, k);
  return (int)(start + (size - start) * xrnd);
}
static int64_t win_base(const pto_handle *h, int64_t nsz) {
  return (h->cfg.hist_capacity > 0 && nsz > h->cfg.hist_capacity) ? nsz - h->cfg.hist_capacity : 0;
}
/* MH_chain::getLogLike(index, true) (chain.cc:1078-1085): beyond the chain's own size it is the current value */
static double mix_llike_at(const pto_handle *h, const chain_t *c, int64_t nsz, int index) {
  int64_t base = win_base(h, nsz);
  if (index < 0 || index >= nsz - base) return c->llike;
  return c->hllike[base + index];
}
static const double *de_draw_state_mixed(pto_handle *h, chain_t *caller, const prop_t *p, uint64_t step, int which, int *kcount) {
  const int R = h->R;
  chain_t *lad = h->chains + ((size_t)(caller - h->chains) / R) * R;
  enum { Nmean = 10, Nmedian = 10 };
  const double pmix = h->de_Tmix;
  double k[PTG_MAX_RUNGS + 1], l0[Nmean], l[Nmedian];
  int kd = *kcount, index, guard = 0; /* the uniform counter runs on across the snooker's redraws of z */
  double beta = caller->beta;
  k[0] = 0;
  int ithis = 0;
  for (int i = 0; i < R; i++) {
    chain_t *ci = &lad[i];
    double l0max = -1e100, l0min = 1e100;
    for (int j = 0; j < Nmean; j++) {
      index = mix_draw_i(h, caller, ci, p, step, which, &kd);
      double dl = mix_llike_at(h, caller, caller->nsize, index);
      if (isfinite(dl)) {
        if (dl > l0max) l0max = dl;
        if (dl < l0min) l0min = dl;
        l0[j] = dl;
      } else { j--; if (++guard > 100000) { h->tape_err = 1; *kcount = kd; return caller->x; } }
    }
    double alpha = ci->beta, amb = alpha - beta;
    amb = -amb;
    if (amb == 0) ithis = i;
    double sum = 0;
    double l0scale = amb < 0 ? l0min : l0max;
    for (int ii = 0; ii < Nmean; ii++) sum += exp((l0[ii] - l0scale) * amb);
    double ll0 = log(sum / Nmean) + l0scale * amb;
    for (int j = 0; j < Nmedian; j++) {
      index = mix_draw_i(h, caller, ci, p, step, which, &kd);
      l[j] = mix_llike_at(h, ci, ci->frozen, index);
    }
    for (int a = 1; a < Nmedian; a++) { /* sort(l.begin(), l.end()) */
      double v = l[a]; int b = a - 1;
      while (b >= 0 && l[b] > v) { l[b + 1] = l[b]; b--; }
      l[b + 1] = v;
    }
    double ll = l[(int)Nmedian / 2];
    double lk = ll0 - ll * amb;
    lk = -lk;
    lk /= pmix;
    if (lk > 0) lk = 0;
    k[i + 1] = k[i] + exp(lk);
  }
  int ipick = ithis;
  double xrnd = mix_u(h, caller, step, which, &kd) * k[R];
  for (int i = 0; i < R; i++)
    if (xrnd <= k[i + 1]) { ipick = i; break; }
  chain_t *ci = &lad[ipick];
  index = mix_draw_i(h, caller, ci, p, step, which, &kd);
  *kcount = kd;
  return ci->hx + (size_t)(win_base(h, ci->frozen) + index) * h->d;
}
static int de_mixing_on(const pto_handle *h) { return h->de_mixing && !h->wrap_in_set && h->R > 1; }
/* draw_from_chain: the state a DE proposal reads for `which` (0 = z, 1 = s1, 2 = s2) */
static const double *de_draw_state(pto_handle *h, chain_t *c, const prop_t *p, uint64_t step, int which, int *attempt) {
  if (de_mixing_on(h)) return de_draw_state_mixed(h, c, p, step, which, attempt);
  return c->hx + (size_t)de_draw_index(h, c, p, step, which, attempt) * h->d;
}

/* differential_evolution::draw_standard (proposal_distribution.cc:489-535) */
static void de_draw_standard(pto_handle *h, chain_t *c, const prop_t *p, uint64_t step, double *prop, draw_result_t *r) {
  const int d = h->d;
  double gamma = 1.68 / sqrt(d) / p->reduce_gamma;
  double xgamma = draw_u32(h, &c->rng, PTG_DOMAIN_STEP, step, PTG_BLK_A, 2);
  if (xgamma < p->g1frac) gamma = 1;
  int a1 = 0, a2 = 0;
  const double *s1 = de_draw_state(h, c, p, step, 1, &a1);
  const double *s2 = de_draw_state(h, c, p, step, 2, &a2);
  if (!is_philox(&c->rng)) { /* edist.drawSample: d normals drawn and discarded (H8-1, :518-526) */
    double zz[PTG_MAX_DIM]; draw_normals(h, &c->rng, step, d, zz);
  }
  for (int i = 0; i < d; i++) {
    double t = c->x[i] + s1[i] * gamma;   /* prop=prop.add(s1.scalar_mult(gamma))  */
    prop[i] = t + s2[i] * (-gamma);       /* prop=prop.add(s2.scalar_mult(-gamma)) */
  }
  r->log_hastings = 0; r->type = 0;
  r->valid = h->zero_valid; /* state::add builds its result from state(space,n) (H8-3, states.cc:194-204) */
}

/* differential_evolution::draw_snooker (proposal_distribution.cc:538-591) */
static void de_draw_snooker(pto_handle *h, chain_t *c, const prop_t *p, uint64_t step, double *prop, draw_result_t *r) {
  const int d = h->d;
  double xgamma = draw_u32(h, &c->rng, PTG_DOMAIN_STEP, step, PTG_BLK_A, 2);
  double gamma = (1.2 + xgamma) / p->reduce_gamma;
  double smznorm2 = 0, minusz[PTG_MAX_DIM], smz[PTG_MAX_DIM];
  int az = 0, isafe = 0;
  while (smznorm2 == 0) {
    const double *z = de_draw_state(h, c, p, step, 0, &az);
    for (int i = 0; i < d; i++) { minusz[i] = z[i] * (-1); smz[i] = c->x[i] + minusz[i]; }
    smznorm2 = 0;
    for (int i = 0; i < d; i++) smznorm2 += smz[i] * smz[i];
    isafe++;
    if (isafe > 1000) break;
  }
  int a1 = 0, a2 = 0;
  const double *s1 = de_draw_state(h, c, p, step, 1, &a1);
  const double *s2 = de_draw_state(h, c, p, step, 2, &a2);
  double dot = 0;
  for (int i = 0; i < d; i++) {
    double ds12 = s1[i] * gamma + s2[i] * (-gamma);
    dot += ds12 * smz[i];
  }
  double fac = dot / smznorm2, pmz2 = 0;
  for (int i = 0; i < d; i++) {
    prop[i] = c->x[i] + smz[i] * fac;
    double pmz = prop[i] + minusz[i];
    pmz2 += pmz * pmz;
  }
  r->log_hastings = (log(pmz2) - log(smznorm2)) * (d - 1) / 2.0;
  r->type = 1;
  r->valid = h->zero_valid;
}

/* `vec = diagTransform*vec` (proposal_distribution.hh:212) is Eigen 3.3.7's column-major dense GEMV
 * (Eigen/src/Core/products/GeneralMatrixVector.h): the result starts at zero, columns are consumed four at a time as
 * res += (a0 v0 + a1 v1) + (a2 v2 + a3 v3) for packet-aligned rows, and the leftover columns one at a time.  With
 * SSE2 packets of two doubles and 16-byte-aligned heap storage every row of an even-dimensional problem takes the
 * packet path; an odd trailing row takes the scalar path res += a0 v0; res += a1 v1; ... (same four-column blocks). */
static void eigen_gemv(const double *M, const double *v, double *t, int d) {
  const int cb = (d / 4) * 4, even_rows = d & ~1;
  for (int i = 0; i < d; i++) {
    const double *a = M + (size_t)i * d;
    double acc = 0;
    for (int j = 0; j < cb; j += 4) {
      if (i < even_rows) acc = acc + ((a[j] * v[j] + a[j + 1] * v[j + 1]) + (a[j + 2] * v[j + 2] + a[j + 3] * v[j + 3]));
      else { acc = a[j] * v[j] + acc; acc = a[j + 1] * v[j + 1] + acc; acc = a[j + 2] * v[j + 2] + acc; acc = a[j + 3] * v[j + 3] + acc; }
    }
    for (int j = cb; j < d; j++) acc += a[j] * v[j];
    t[i] = acc;
  }
}

/* gaussian_prop::draw (proposal_distribution.hh:194-218) */
static void gauss_draw(pto_handle *h, chain_t *c, const prop_t *p, uint64_t step, double *prop, draw_result_t *r) {
  const int d = h->d;
  double z[PTG_MAX_DIM], off[PTG_MAX_DIM];
  draw_normals(h, &c->rng, step, d, z);
  for (int i = 0; i < d; i++) off[i] = z[i] * p->sigmas[i] + 0.0; /* GaussianDist::draw: normal*sigma+x0 */
  double x = 1;
  if (p->one_d_frac > 0) x = draw_u32(h, &c->rng, PTG_DOMAIN_STEP, step, PTG_BLK_A, 1);
  if (p->one_d_frac > 0 && x < p->one_d_frac) {
    int i = (int)(d * draw_u32(h, &c->rng, PTG_DOMAIN_STEP, step, PTG_BLK_A, 2));
    for (int j = 0; j < d; j++) if (j != i) off[j] = 0.0;
    r->type = 1;
  } else r->type = 0;
  if (p->transform) { /* vec = diagTransform*vec */
    double t[PTG_MAX_DIM];
    eigen_gemv(p->transform, off, t, d);
    memcpy(off, t, (size_t)d * sizeof(double));
  }
  for (int i = 0; i < d; i++) prop[i] = c->x[i] + off[i];
  r->log_hastings = 0;
  r->valid = h->zero_valid;
}

/* draw_from_dist::draw (proposal_distribution.hh:124-129) */
static void prior_prop_draw(pto_handle *h, chain_t *c, uint64_t step, double *prop, draw_result_t *r) {
  r->valid = prior_draw(h, &c->rng, PTG_DOMAIN_STEP, step, PTG_BLK_PRIOR, prop);
  r->log_hastings = prior_eval_log(h, c->x, 1) - prior_eval_log(h, prop, r->valid);
  r->type = 0;
}

static int prop_ready(const pto_handle *h, const chain_t *c, const prop_t *p) {
  if (p->kind == PTG_PROP_DE) { /* differential_evolution::is_ready (proposal_distribution.hh:407) */
    int64_t W = c->nsize;
    if (h->cfg.hist_capacity > 0 && W > h->cfg.hist_capacity) W = h->cfg.hist_capacity;
    return W >= h->d * 10;
  }
  return 1;
}

static void member_draw(pto_handle *h, chain_t *c, const prop_t *p, uint64_t step, double *prop, draw_result_t *r) {
  switch (p->kind) {
  case PTG_PROP_DE: { /* differential_evolution::draw (proposal_distribution.cc:790-801) */
    double x = draw_u32(h, &c->rng, PTG_DOMAIN_STEP, step, PTG_BLK_A, 1);
    if (p->snooker > x) de_draw_snooker(h, c, p, step, prop, r);
    else de_draw_standard(h, c, p, step, prop, r);
    break;
  }
  case PTG_PROP_GAUSS: gauss_draw(h, c, p, step, prop, r); break;
  case PTG_PROP_PRIOR_DRAW: prior_prop_draw(h, c, step, prop, r); break;
  }
}

/* ---- proposal sets.  The engine's member list is flat; optionally members [nest_first, nest_first + nest_count) form ONE nested
 * proposal_distribution_set that occupies a single slot of the top-level set (what ptmcmc_sampler builds for prop_adapt_rate > 0,
 * ptmcmc.cc:123-131).  slot -> member: slots before the nested one map to themselves, the nested slot to -1, later slots skip the
 * nested members. */
static int n_top(const pto_handle *h) { return h->nest_count ? h->nprops - h->nest_count + 1 : h->nprops; }
static int slot_member(const pto_handle *h, int slot) {
  if (!h->nest_count || slot < h->nest_first) return slot;
  if (slot == h->nest_first) return -1;
  return slot + h->nest_count - 1;
}
/* proposal_distribution_set::reset_bins (proposal_distribution.cc:37-59): normalises the shares IN PLACE */
static void set_reset_bins(setstate_t *s, double Tpow, const double *hot, double beta, int have_chain) {
  int n = s->n;
  double Tfac = 0;
  if (Tpow > 0 && have_chain) Tfac = 1 - pow(beta, Tpow);
  double sum = 0;
  for (int i = 0; i < n; i++) sum += s->shares[i];
  double last = 0;
  for (int i = 0; i < n; i++) {
    s->shares[i] /= sum;
    s->bins[i] = last + s->shares[i];
    if (Tpow > 0) s->bins[i] += (hot[i] - s->shares[i]) * Tfac;
    last = s->bins[i];
  }
  double back = s->bins[n - 1];
  for (int i = 0; i < n; i++) s->bins[i] /= back;
}
/* constructor (proposal_distribution.cc:61-93: hot shares normalised, reset_bins without a chain) followed by set_chain on the rung's clone
 * (proposal_distribution.hh:336: reset_bins with the rung's temperature) */
static void set_init(setstate_t *s, int n, const double *shares, double *hot, double Tpow, double beta) {
  s->n = n; s->adapt_count = 0;
  for (int i = 0; i < n; i++) { s->shares[i] = shares[i]; s->last_accepted[i] = 1; } /* last_accepted.resize(Nsize,true) (:88) */
  if (Tpow > 0) {
    double sum = 0;
    for (int i = 0; i < n; i++) sum += hot[i];
    if (sum <= 0) for (int i = 0; i < n; i++) hot[i] = shares[i];
    else for (int i = 0; i < n; i++) hot[i] /= sum;
  }
  set_reset_bins(s, Tpow, hot, beta, 0);
  set_reset_bins(s, Tpow, hot, beta, 1);
}
/* proposal_distribution_set::accept / reject (proposal_distribution.cc:132-166).  adapt_count is never reset, so from the
 * adapt_every-th (10 Nsize) decision on the bins are rebuilt after every decision. */
static void set_adapt(setstate_t *s, double rate, int idx, int accepted, double Tpow, const double *hot, double beta) {
  if (rate == 0) return;
  if (s->last_accepted[idx] == accepted) s->shares[idx] *= 1 - rate * 0.25;
  s->last_accepted[idx] = accepted;
  s->adapt_count++;
  if (s->adapt_count >= 10 * s->n) set_reset_bins(s, Tpow, hot, beta, 1);
}
static void chain_sets_init(pto_handle *h, chain_t *c) {
  double sh[PTG_MAX_PROPOSALS], hot[PTG_MAX_PROPOSALS];
  const int nt = n_top(h);
  for (int sl = 0; sl < nt; sl++) {
    int m = slot_member(h, sl);
    sh[sl] = m < 0 ? h->nest_share : h->props[m].share;
    hot[sl] = m < 0 ? h->nest_hot : h->props[m].hot_share;
  }
  set_init(&c->top, nt, sh, hot, h->Tpow, c->beta);
  for (int sl = 0; sl < nt; sl++) h->hot_norm[sl] = hot[sl];
  c->nest.n = 0;
  if (h->nest_count) {
    for (int j = 0; j < h->nest_count; j++) { sh[j] = h->props[h->nest_first + j].share; hot[j] = 0; }
    set_init(&c->nest, h->nest_count, sh, hot, 0.0, c->beta); /* the nested set has no thermal weighting (ptmcmc.cc:130) */
  }
}

/* proposal_distribution_set::draw (proposal_distribution.cc:99-129), for the top-level set and a nested one */
static int set_draw(pto_handle *h, chain_t *c, uint64_t step, double *prop, draw_result_t *r) {
  if (!h->wrap_in_set) { member_draw(h, c, &h->props[0], step, prop, r); return 0; }
  int count = 0;
  const int nt = n_top(h);
  while (1) {
    double x;
    if (nt > 1) x = draw_u32(h, &c->rng, PTG_DOMAIN_STEP, step, PTG_BLK_A, 0);
    else x = 0;
    for (int sl = 0; sl < nt; sl++) {
      const int m = slot_member(h, sl);
      if ((m < 0 || prop_ready(h, c, &h->props[m])) && x < c->top.bins[sl]) { /* a set is always ready (proposal_distribution.hh:70) */
        r->slot = sl; r->nested = -1;
        if (m >= 0) {
          member_draw(h, c, &h->props[m], step, prop, r);
          r->member = m;
        } else { /* the nested set's own draw */
          int ncount = 0, done = 0;
          while (!done) {
            double x2;
            if (h->nest_count > 1) x2 = draw_u32(h, &c->rng, PTG_DOMAIN_STEP, step, PTG_BLK_NEST, 0);
            else x2 = 0;
            for (int j = 0; j < h->nest_count && !done; j++) {
              if (prop_ready(h, c, &h->props[h->nest_first + j]) && x2 < c->nest.bins[j]) {
                member_draw(h, c, &h->props[h->nest_first + j], step, prop, r);
                r->type = j + 10 * r->type;
                r->member = h->nest_first + j; r->nested = j;
                done = 1;
              }
            }
            if (done) break;
            ncount++;
            if (ncount > 100 || is_philox(&c->rng)) return -1;
          }
        }
        r->type = sl + 10 * r->type;
        return 0;
      }
    }
    count++;
    if (count > 100) return -1;
    if (is_philox(&c->rng)) return -1; /* addressed draws cannot be redrawn */
  }
}
/* accept() / reject() of the top-level set, handed down to the member that drew (proposal_distribution.cc:132-166) */
static void sets_adapt(pto_handle *h, chain_t *c, const draw_result_t *r, int accepted) {
  if (!h->wrap_in_set) return;
  set_adapt(&c->top, h->adapt_rate, r->slot, accepted, h->Tpow, h->hot_norm, c->beta);
  if (r->nested >= 0) set_adapt(&c->nest, h->nest_adapt, r->nested, accepted, 0.0, h->hot_norm, c->beta);
}

/* ---------------------------------------------------------------------------------------------- MH step */
/* MH_chain::step(prop) (chain.cc:966-1022) */
static int mh_step(pto_handle *h, chain_t *c, uint64_t step, double *lhr_out, int *code_out) {
  const int d = h->d;
  double newx[PTG_MAX_DIM];
  draw_result_t r; r.type = 0; r.log_hastings = 0; r.valid = 1; r.member = 0; r.slot = 0; r.nested = -1;
  double oldlprior = c->lpost - c->beta * c->llike;
  if (set_draw(h, c, step, newx, &r) != 0) return fail(PTG_EINVAL, "proposal set: no member ready");
  int valid = r.valid;
  if (valid) valid = space_enforce(h, newx);           /* newstate.enforce() */
  double newlprior = prior_eval_log(h, newx, valid);
  double newlike, newlpost;
  int accept = 1, code = 0;
  if (valid && ((newlprior > -1e200) || (newlprior - oldlprior > h->cfg.dprior_min))) { /* H8-4 */
    newlike = like_eval(h, newx);
    newlpost = newlike * c->beta + newlprior;
  } else {
    newlike = newlpost = -INFINITY;
    code |= PTG_TRACE_NOLIKE;
  }
  double lhr = r.log_hastings;
  if (isnan(lhr)) accept = 0;
  lhr += newlpost - c->lpost;
  if (!valid) { accept = 0; code |= PTG_TRACE_INVALID; }
  if (accept && lhr < 0) {
    double x = draw_u52(h, &c->rng, PTG_DOMAIN_STEP, step, PTG_BLK_B, 1);
    accept = (log(x) < lhr);
  }
  c->ntries++;
  if (accept) {
    c->naccept++;
    c->last_type = r.type;
    sets_adapt(h, c, &r, 1);
    add_state(h, c, newx, newlike, newlpost);
    code |= PTG_TRACE_ACCEPT;
  } else {
    sets_adapt(h, c, &r, 0);
    add_state(h, c, c->x, c->llike, c->lpost);
  }
  code |= (r.type & PTG_TRACE_TYPE_MASK);
  *lhr_out = lhr; *code_out = code;
  return 0;
}

/* ---------------------------------------------------------------------------------------------- PT step */
/* parallel_tempering_chains::pry_temps, vector version with one pried gap (chain.cc:1809-1846),
 * + MH_chain::resetTemp (chain.cc:1088-1091) */
static void pry_temps(pto_handle *h, chain_t *ch, int ipry, double rate, double *invtemps) {
  const int R = h->R;
  double splits[PTG_MAX_RUNGS];
  for (int i = 0; i < R - 1; i++) {
    splits[i] = invtemps[i] - invtemps[i + 1];
    if (h->cfg.evolve_lpost_cut >= 0 && ch[i].lpost - ch[i + 1].lpost > h->cfg.evolve_lpost_cut * invtemps[i])
      splits[i] *= pow(1.0 + rate, 1);
  }
  splits[ipry] *= 1.0 + rate;
  double sum = 0;
  for (int i = 0; i < R - 1; i++) sum += splits[i];
  double norm = sum / (1 - invtemps[R - 1]);
  double invtemp = 1;
  for (int i = 1; i < R - 1; i++) {
    invtemp -= splits[i - 1] / norm;
    invtemps[i] = invtemp;
    ch[i].beta = invtemp;
    ch[i].lpost = prior_eval_log(h, ch[i].x, 1) + invtemp * ch[i].llike;
  }
}

/* parallel_tempering_chains::step for one ladder (chain.cc:1393-1570) */
static int pt_step_ladder(pto_handle *h, int l) {
  const int R = h->R, d = h->d;
  chain_t *ch = h->chains + (size_t)l * R;
  stream_t *ls = &h->lstreams[l];
  const uint64_t step = (uint64_t)h->istep;
  int iswaps[64];
  const int maxswaps = h->maxswaps;
  if (h->cfg.swap_mode == PTG_SWAP_REFERENCE) {
    /* choose swap candidates (chain.cc:1410-1420) */
    for (int i = 0; i < maxswaps; i++) {
      iswaps[i] = -2;
      double x = draw_u32(h, ls, PTG_DOMAIN_STEP, step, (uint32_t)i, 0);
      if (R > 1 && x < (R - 1) * h->cfg.swap_rate / maxswaps) {
        x = draw_u32(h, ls, PTG_DOMAIN_STEP, step, (uint32_t)i, 1);
        iswaps[i] = (int)(x * (R - 1));
        for (int j = 0; j < i; j++)
          if (iswaps[j] == iswaps[i] || iswaps[j] + 1 == iswaps[i]) iswaps[i] = -2;
      }
    }
  } else {
    /* even/odd performance mode: pairs (i,i+1), i = parity, parity+2, ... each tried with probability
     * min(1, 2*swap_rate); statistically validated, not a reference mode */
    int parity = (int)(step & 1), n = 0;
    double ptry = 2 * h->cfg.swap_rate; if (ptry > 1) ptry = 1;
    for (int i = parity; i + 1 < R; i += 2) {
      double x = draw_u52(h, ls, PTG_DOMAIN_STEP, step, PTG_BLK_SWAP_EVENODD + (uint32_t)i, 0);
      if (x < ptry) iswaps[n++] = i;
    }
    for (int i = n; i < 64; i++) iswaps[i] = -2;
  }
  const int ntrial = (h->cfg.swap_mode == PTG_SWAP_REFERENCE) ? maxswaps : 64;
  /* gather_llikes / gather_invtemps (chain.cc:1433-1435): states and loglikes always mirror the chains here */
  double invtemps[PTG_MAX_RUNGS];
  for (int i = 0; i < R; i++) invtemps[i] = ch[i].beta;
  for (int j = 0; j < ntrial; j++) {
    if (iswaps[j] < 0) continue;
    int accept = 1;
    int i = iswaps[j];
    int32_t *dir = h->directions + (size_t)l * R, *ups = h->ups + (size_t)l * R, *downs = h->downs + (size_t)l * R,
            *inst = h->instances + (size_t)l * R;
    if (i > 0) {
      if (dir[i] > 0) ups[i]++;
      if (dir[i] < 0) downs[i]++;
    }
    double lla = ch[i].llike; if (!(lla > -1e200)) lla = -1e200;
    double llb = ch[i + 1].llike; if (!(llb > -1e200)) llb = -1e200;
    double lhr = -(invtemps[i + 1] - invtemps[i]) * (llb - lla);
    if (lhr < 0) {
      double x = (h->cfg.swap_mode == PTG_SWAP_REFERENCE)
                     ? draw_u52(h, ls, PTG_DOMAIN_STEP, step, (uint32_t)j, 1)
                     : draw_u52(h, ls, PTG_DOMAIN_STEP, step, PTG_BLK_SWAP_EVENODD + (uint32_t)i, 1);
      accept = (log(x) < lhr);
    }
    if (accept) {
      double xa[PTG_MAX_DIM], xb[PTG_MAX_DIM];
      memcpy(xa, ch[i].x, (size_t)d * sizeof(double)); memcpy(xb, ch[i + 1].x, (size_t)d * sizeof(double));
      double la = ch[i].llike, lb = ch[i + 1].llike;
      /* add_state(state, llike) with log_post=999 => lpost recomputed (chain.cc:925-928) */
      add_state(h, &ch[i + 1], xa, la, prior_eval_log(h, xa, 1) + ch[i + 1].beta * la);
      add_state(h, &ch[i], xb, lb, prior_eval_log(h, xb, 1) + ch[i].beta * lb);
      { int t = dir[i]; dir[i] = dir[i + 1]; dir[i + 1] = t; }
      { int t = inst[i]; inst[i] = inst[i + 1]; inst[i + 1] = t; }
      if (i == 0) dir[i] = 1;
      if (i + 1 == R - 1) dir[i + 1] = -1;
      h->swap_accept[(size_t)l * (R - 1) + i]++;
      if (h->cfg.evolve_rate > 0) pry_temps(h, ch, i, h->cfg.evolve_rate, invtemps);
    } else {
      add_state(h, &ch[i], ch[i].x, ch[i].llike, ch[i].lpost);
      add_state(h, &ch[i + 1], ch[i + 1].x, ch[i + 1].llike, ch[i + 1].lpost);
    }
    h->swap_count[(size_t)l * (R - 1) + i]++;
  }
  /* standard step for the rungs not touched by a swap trial (chain.cc:1544-1559) */
  for (int i = 0; i < R; i++) ch[i].frozen = ch[i].nsize; /* history_freeze: nothing below changes another chain's visible size */
  for (int i = 0; i < R; i++) {
    int skip = 0;
    for (int j = 0; j < ntrial; j++) if (i == iswaps[j] || i == iswaps[j] + 1) skip = 1;
    double lhr = 0; int code = PTG_TRACE_SWAPPED;
    if (!skip) { int rc = mh_step(h, &ch[i], step, &lhr, &code); if (rc) return rc; }
    if (h->istep < h->cfg.trace_steps) {
      size_t k = (size_t)h->istep * h->nchains + (size_t)l * R + i;
      h->trace_lhr[k] = lhr; h->trace_code[k] = code;
    }
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------- API */
int pto_create(const ptg_config *cfg, pto_handle **out) {
  if (!cfg || !out) return fail(PTG_EINVAL, "null argument");
  if (cfg->abi_version != PTG_ABI_VERSION) return fail(PTG_EINVAL, "abi version mismatch");
  if (cfg->dim < 1 || cfg->dim > PTG_MAX_DIM || cfg->n_rungs < 1 || cfg->n_rungs > PTG_MAX_RUNGS || cfg->n_ladders < 1 ||
      cfg->save_every < 1 || cfg->n_init < 1)
    return fail(PTG_EINVAL, "bad config");
  pto_handle *h = (pto_handle *)calloc(1, sizeof(*h));
  h->cfg = *cfg; h->d = cfg->dim; h->R = cfg->n_rungs; h->L = cfg->n_ladders;
  h->nchains = (int64_t)h->R * h->L;
  h->zero_valid = 1;
  for (int i = 0; i < h->d; i++) { h->xmin[i] = -INFINITY; h->xmax[i] = INFINITY; }
  h->chains = (chain_t *)calloc((size_t)h->nchains, sizeof(chain_t));
  h->lstreams = (stream_t *)calloc((size_t)h->L, sizeof(stream_t));
  h->maxswaps = (int)(1 + 2 * cfg->swap_rate * h->R); /* chain.cc:1192 */
  if (h->maxswaps > 64) h->maxswaps = 64;
  h->swap_count = (int64_t *)calloc((size_t)h->L * h->R, sizeof(int64_t));
  h->swap_accept = (int64_t *)calloc((size_t)h->L * h->R, sizeof(int64_t));
  h->directions = (int32_t *)calloc((size_t)h->nchains, sizeof(int32_t));
  h->ups = (int32_t *)calloc((size_t)h->nchains, sizeof(int32_t));
  h->downs = (int32_t *)calloc((size_t)h->nchains, sizeof(int32_t));
  h->instances = (int32_t *)calloc((size_t)h->nchains, sizeof(int32_t));
  if (cfg->trace_steps > 0) {
    h->trace_lhr = (double *)calloc((size_t)cfg->trace_steps * h->nchains, sizeof(double));
    h->trace_code = (int32_t *)calloc((size_t)cfg->trace_steps * h->nchains, sizeof(int32_t));
  }
  double tratio = (h->R > 1) ? exp(log(cfg->Tmax) / (h->R - 1)) : 1.0; /* chain.cc:1181-1183 */
  for (int l = 0; l < h->L; l++) {
    double temp = 1;
    for (int r = 0; r < h->R; r++) {
      chain_t *c = &h->chains[(size_t)l * h->R + r];
      if (r > 0) temp = temp * tratio;
      c->beta = 1 / temp; /* chain.cc:1339 */
      c->x = (double *)calloc((size_t)h->d, sizeof(double));
      c->map_x = (double *)calloc((size_t)h->d, sizeof(double));
      c->ntries = 1; c->naccept = 1; c->last_type = -1; c->map_lpost = -1e200; /* chain.cc:649, chain.hh:69 */
      c->rng.mode = cfg->rng_mode;
      c->rng.id = (uint64_t)(cfg->ladder_offset + l) * PTG_STREAM_STRIDE + (uint64_t)r;
    }
    h->lstreams[l].mode = cfg->rng_mode;
    h->lstreams[l].id = (uint64_t)(cfg->ladder_offset + l) * PTG_STREAM_STRIDE + PTG_STREAM_LADDER;
  }
  *out = h;
  return 0;
}

int pto_destroy(pto_handle *h) {
  if (!h) return 0;
  for (int64_t i = 0; i < h->nchains; i++) {
    chain_t *c = &h->chains[i];
    free(c->x); free(c->map_x); free(c->hx); free(c->hlpost); free(c->hllike); free(c->hacc); free(c->hbeta); free(c->htype);
    free(c->rng.urec); free(c->rng.zrec);
  }
  for (int l = 0; l < h->L; l++) { free(h->lstreams[l].urec); free(h->lstreams[l].zrec); }
  for (int i = 0; i < h->nprops; i++) { free(h->props[i].sigmas); free(h->props[i].transform); }
  free(h->chains); free(h->lstreams); free(h->lparams); free(h->ldata); free(h->betas0);
  free(h->swap_count); free(h->swap_accept); free(h->directions); free(h->ups); free(h->downs); free(h->instances);
  free(h->trace_lhr); free(h->trace_code); free(h->tape_u); free(h->tape_z); free(h->mark_u); free(h->mark_z);
  free(h);
  return 0;
}

int pto_set_space(pto_handle *h, const int32_t *lt, const int32_t *ut, const double *xmin, const double *xmax) {
  for (int i = 0; i < h->d; i++) { h->lower[i] = lt[i]; h->upper[i] = ut[i]; h->xmin[i] = xmin[i]; h->xmax[i] = xmax[i]; }
  double zero[PTG_MAX_DIM]; memset(zero, 0, sizeof zero);
  h->zero_valid = space_enforce(h, zero); /* state(space,n) enforces the zero vector (states.cc:168-176) */
  h->have_space = 1;
  return 0;
}

int pto_set_prior(pto_handle *h, const int32_t *type, const double *a, const double *b) {
  for (int i = 0; i < h->d; i++) {
    prior1d_t *p = &h->prior[i];
    p->kind = type[i]; p->a = a[i]; p->b = b[i];
    if (p->kind == PTG_PRIOR_POLAR) { /* ProbabilityDist.h:187-192: the clamp acts on the ctor arguments only */
      double lo = a[i], hi = b[i];
      if (lo < 0) lo = 0;
      if (hi > M_PI) hi = M_PI;
      p->norm = -cos(hi) + cos(lo); p->cdfoff = -cos(lo) / p->norm;
    } else if (p->kind == PTG_PRIOR_COPOLAR) {
      double lo = a[i], hi = b[i];
      if (lo < -M_PI / 2) lo = -M_PI / 2;
      if (hi > M_PI / 2) hi = M_PI / 2;
      p->norm = sin(hi) - sin(lo); p->cdfoff = sin(lo) / p->norm;
    } else if (p->kind == PTG_PRIOR_LOG) {
      if (a[i] <= 0 || b[i] <= a[i]) return fail(PTG_EINVAL, "log prior needs 0<xmin<xmax");
      p->la = log(a[i]); p->lb = log(b[i]);
    } else if (p->kind != PTG_PRIOR_UNIFORM && p->kind != PTG_PRIOR_GAUSSIAN && p->kind != PTG_PRIOR_GAUSSIAN_WRAPPED) return fail(PTG_EINVAL, "bad prior type");
  }
  h->have_prior = 1;
  return 0;
}

int pto_set_likelihood(pto_handle *h, int32_t kind, const double *params, int32_t n_params, const double *data, int64_t n_data) {
  h->like_kind = kind;
  free(h->lparams); free(h->ldata); h->lparams = NULL; h->ldata = NULL;
  h->n_lparams = n_params; h->n_ldata = n_data;
  h->lparams = (double *)calloc((size_t)(n_params > 0 ? n_params : 1), sizeof(double));
  if (n_params) memcpy(h->lparams, params, (size_t)n_params * sizeof(double));
  if (n_data) { h->ldata = (double *)malloc((size_t)n_data * sizeof(double)); memcpy(h->ldata, data, (size_t)n_data * sizeof(double)); }
  h->like_nsum = 0;
  if (kind == PTG_LIKE_POLY_CHI2 || kind == PTG_LIKE_SINUSOID_CHI2) { /* bayesian.hh:613: nsum+=log(S[i]) */
    int64_t N = n_data / 3; double nsum = 0;
    for (int64_t i = 0; i < N; i++) nsum += log(h->ldata[2 * N + i]);
    h->like_nsum = nsum;
  }
  h->have_like = 1;
  return 0;
}

int pto_set_proposals(pto_handle *h, int32_t n, const ptg_proposal *props, double Tpow, int32_t wrap_in_set) {
  if (n < 1 || n > PTG_MAX_PROPOSALS) return fail(PTG_EINVAL, "bad proposal count");
  if (!wrap_in_set && n != 1) return fail(PTG_EINVAL, "a bare proposal must be single");
  h->nprops = n; h->Tpow = Tpow; h->wrap_in_set = wrap_in_set;
  h->adapt_rate = 0; h->de_mixing = 0; h->de_Tmix = 1; h->nest_first = h->nest_count = 0; h->nest_adapt = 0;
  for (int i = 0; i < n; i++) {
    prop_t *p = &h->props[i]; const ptg_proposal *q = &props[i];
    p->kind = q->kind; p->share = q->share; p->hot_share = q->hot_share;
    p->snooker = q->snooker; p->g1frac = q->gamma_one_frac; p->b_small = q->b_small; p->ignore_frac = q->ignore_frac;
    p->unlikely_alpha = q->unlikely_alpha; p->reduce_gamma = q->reduce_gamma; p->one_d_frac = q->one_d_frac;
    p->sigmas = NULL; p->transform = NULL;
    if (q->kind == PTG_PROP_GAUSS) {
      p->sigmas = (double *)malloc((size_t)h->d * sizeof(double)); memcpy(p->sigmas, q->sigmas, (size_t)h->d * sizeof(double));
      if (q->transform) { p->transform = (double *)malloc((size_t)h->d * h->d * sizeof(double)); memcpy(p->transform, q->transform, (size_t)h->d * h->d * sizeof(double)); }
    }
  }
  h->have_props = 1;
  return 0;
}

int pto_set_proposal_options(pto_handle *h, double adapt_rate, int32_t de_mixing, double de_Tmix) {
  if (!h->have_props) return fail(PTG_EINVAL, "set the proposals first");
  if (adapt_rate != 0 && !h->wrap_in_set) return fail(PTG_EINVAL, "adaptive shares need a proposal set");
  if (de_mixing) {
    if (h->wrap_in_set || h->props[0].kind != PTG_PROP_DE) return fail(PTG_EINVAL, "temperature mixing needs a bare differential-evolution proposal");
    if (h->props[0].unlikely_alpha != 0) return fail(PTG_EINVAL, "temperature mixing needs unlikely_alpha = 0");
  }
  h->adapt_rate = adapt_rate; h->de_mixing = de_mixing; h->de_Tmix = de_Tmix;
  return 0;
}
int pto_set_nested_set(pto_handle *h, int32_t first, int32_t count, double share, double hot_share, double adapt_rate) {
  if (!h->have_props || !h->wrap_in_set) return fail(PTG_EINVAL, "a nested set needs a proposal set");
  if (first < 0 || count < 1 || first + count > h->nprops) return fail(PTG_EINVAL, "nested set out of range");
  h->nest_first = first; h->nest_count = count; h->nest_share = share; h->nest_hot = hot_share; h->nest_adapt = adapt_rate;
  return 0;
}
int pto_get_proposal_shares(pto_handle *h, double *shares) {
  const int nt = n_top(h), nb = nt + h->nest_count;
  for (int64_t i = 0; i < h->nchains; i++) {
    for (int k = 0; k < nt; k++) shares[i * nb + k] = h->chains[i].top.shares[k];
    for (int k = 0; k < h->nest_count; k++) shares[i * nb + nt + k] = h->chains[i].nest.shares[k];
  }
  return 0;
}
int pto_set_betas(pto_handle *h, const double *betas) {
  if (!betas) return 0;
  for (int64_t i = 0; i < h->nchains; i++) h->chains[i].beta = betas[i];
  return 0;
}

int pto_seed(pto_handle *h, uint64_t seed) { h->cfg.seed = seed; return 0; }

int pto_seed_newran(pto_handle *h, double seed) {
  mother_t master; mother_init(&master, seed);
  for (int l = 0; l < h->L; l++) {
    h->lstreams[l].mode = PTO_RNG_NEWRAN;
    mother_init(&h->lstreams[l].mom, mother_next(&master));
    for (int r = 0; r < h->R; r++) {
      chain_t *c = &h->chains[(size_t)l * h->R + r];
      c->rng.mode = PTO_RNG_NEWRAN;
      mother_init(&c->rng.mom, mother_next(&master));
    }
  }
  return 0;
}

int pto_inject_tapes(pto_handle *h, const double *u, const int64_t *u_off, const double *z, const int64_t *z_off) {
  int64_t ns = h->nchains + h->L;
  free(h->tape_u); free(h->tape_z);
  h->tape_u = (double *)malloc((size_t)(u_off[ns] > 0 ? u_off[ns] : 1) * sizeof(double));
  h->tape_z = (double *)malloc((size_t)(z_off[ns] > 0 ? z_off[ns] : 1) * sizeof(double));
  memcpy(h->tape_u, u, (size_t)u_off[ns] * sizeof(double));
  memcpy(h->tape_z, z, (size_t)z_off[ns] * sizeof(double));
  for (int64_t s = 0; s < ns; s++) {
    stream_t *st = s < h->nchains ? &h->chains[s].rng : &h->lstreams[s - h->nchains];
    st->mode = PTG_RNG_TAPE;
    st->ut = h->tape_u; st->zt = h->tape_z;
    st->upos = u_off[s]; st->uend = u_off[s + 1]; st->zpos = z_off[s]; st->zend = z_off[s + 1];
  }
  return 0;
}

int pto_record_tapes(pto_handle *h, int on) { h->record = on; return 0; }
int pto_get_tape_sizes(pto_handle *h, int64_t *uc, int64_t *zc) {
  for (int64_t s = 0; s < h->nchains + h->L; s++) {
    stream_t *st = s < h->nchains ? &h->chains[s].rng : &h->lstreams[s - h->nchains];
    uc[s] = st->nu; zc[s] = st->nz;
  }
  return 0;
}
int pto_get_tapes(pto_handle *h, double *u, double *z) {
  int64_t pu = 0, pz = 0;
  for (int64_t s = 0; s < h->nchains + h->L; s++) {
    stream_t *st = s < h->nchains ? &h->chains[s].rng : &h->lstreams[s - h->nchains];
    if (st->nu) memcpy(u + pu, st->urec, (size_t)st->nu * sizeof(double));
    if (st->nz) memcpy(z + pz, st->zrec, (size_t)st->nz * sizeof(double));
    pu += st->nu; pz += st->nz;
  }
  return 0;
}

/* per-stream draw counts at the start of each recorded PT step, [n_marks][n_streams]; pass NULLs to query n_marks */
int pto_get_tape_marks(pto_handle *h, int64_t *n_marks, int64_t *u_mark, int64_t *z_mark) {
  const int64_t ns = h->nchains + h->L;
  if (n_marks) *n_marks = h->n_marks;
  if (u_mark && h->n_marks) memcpy(u_mark, h->mark_u, (size_t)h->n_marks * ns * sizeof(int64_t));
  if (z_mark && h->n_marks) memcpy(z_mark, h->mark_z, (size_t)h->n_marks * ns * sizeof(int64_t));
  return 0;
}

static int finish_init(pto_handle *h) {
  for (int64_t i = 0; i < h->nchains; i++) {
    chain_t *c = &h->chains[i];
    c->nhist = 0;
    chain_sets_init(h, c); /* set_proposal after initialize (ptmcmc.cc:514-522) */
  }
  for (int l = 0; l < h->L; l++)
    for (int r = 0; r < h->R; r++) { /* chain.cc:1345-1358 */
      size_t k = (size_t)l * h->R + r;
      h->instances[k] = r;
      h->directions[k] = (r == 0) ? -1 : (r == h->R - 1 ? 1 : 0);
      h->ups[k] = h->downs[k] = 0;
    }
  h->total_steps = 0; h->istep = 0; h->inited = 1;
  return h->tape_err ? fail(PTG_ETAPE, "tape exhausted during init") : 0;
}

/* MH_chain::initialize(n) (chain.cc:846-876) for every chain */
int pto_init_from_prior(pto_handle *h) {
  if (!h->have_prior || !h->have_like || !h->have_props) return fail(PTG_EINVAL, "set prior, likelihood and proposals first");
  double x[PTG_MAX_DIM];
  for (int64_t ic = 0; ic < h->nchains; ic++) {
    chain_t *c = &h->chains[ic];
    for (int k = 0; k < h->cfg.n_init; k++) {
      int icnt = 0;
      int valid = prior_draw(h, &c->rng, PTG_DOMAIN_INIT, (uint64_t)k, 0, x);
      double slike = 0;
      while (!valid || (slike = like_eval(h, x)) < -1e100) {
        icnt++;
        if (icnt >= 100000) return fail(PTG_ESTUCK, "init: cannot draw a valid state");
        valid = prior_draw(h, &c->rng, PTG_DOMAIN_INIT, (uint64_t)k, (uint32_t)icnt * PTG_INIT_ATTEMPT_STRIDE, x);
      }
      c->nhist = 0;
      /* add_state(s): likelihood and posterior re-evaluated (log_like=log_post=999, chain.cc:925-928) */
      double ll = like_eval(h, x);
      add_state(h, c, x, ll, prior_eval_log(h, x, 1) + c->beta * ll);
    }
  }
  return finish_init(h);
}

int pto_init_states(pto_handle *h, const double *x) {
  if (!h->have_prior || !h->have_like || !h->have_props) return fail(PTG_EINVAL, "set prior, likelihood and proposals first");
  for (int64_t ic = 0; ic < h->nchains; ic++) {
    chain_t *c = &h->chains[ic];
    for (int k = 0; k < h->cfg.n_init; k++) {
      const double *xs = x + ((size_t)ic * h->cfg.n_init + k) * h->d;
      c->nhist = 0;
      double ll = like_eval(h, xs);
      add_state(h, c, xs, ll, prior_eval_log(h, xs, 1) + c->beta * ll);
    }
  }
  return finish_init(h);
}

int pto_step(pto_handle *h, int64_t n_steps) {
  if (!h->inited) return fail(PTG_EINVAL, "not initialised");
  for (int64_t s = 0; s < n_steps; s++) {
    if (h->record) {
      const int64_t ns = h->nchains + h->L;
      if (h->n_marks == h->cap_marks) {
        h->cap_marks = h->cap_marks ? 2 * h->cap_marks : 256;
        h->mark_u = (int64_t *)realloc(h->mark_u, (size_t)h->cap_marks * ns * sizeof(int64_t));
        h->mark_z = (int64_t *)realloc(h->mark_z, (size_t)h->cap_marks * ns * sizeof(int64_t));
      }
      for (int64_t k = 0; k < ns; k++) {
        stream_t *st = k < h->nchains ? &h->chains[k].rng : &h->lstreams[k - h->nchains];
        h->mark_u[h->n_marks * ns + k] = st->nu; h->mark_z[h->n_marks * ns + k] = st->nz;
      }
      h->n_marks++;
    }
    for (int l = 0; l < h->L; l++) { int rc = pt_step_ladder(h, l); if (rc) return rc; }
    h->istep++;
  }
  return h->tape_err ? fail(PTG_ETAPE, "tape exhausted") : 0;
}

int pto_get_current(pto_handle *h, double *x, double *lpost, double *llike, double *beta) {
  for (int64_t i = 0; i < h->nchains; i++) {
    chain_t *c = &h->chains[i];
    if (x) memcpy(x + (size_t)i * h->d, c->x, (size_t)h->d * sizeof(double));
    if (lpost) lpost[i] = c->lpost;
    if (llike) llike[i] = c->llike;
    if (beta) beta[i] = c->beta;
  }
  return 0;
}
int pto_get_counters(pto_handle *h, int64_t *nhist, int64_t *nsize, int64_t *ntries, int64_t *naccept, int32_t *last_type, double *map_lpost) {
  for (int64_t i = 0; i < h->nchains; i++) {
    chain_t *c = &h->chains[i];
    if (nhist) nhist[i] = c->nhist;
    if (nsize) nsize[i] = c->nsize;
    if (ntries) ntries[i] = c->ntries;
    if (naccept) naccept[i] = c->naccept;
    if (last_type) last_type[i] = c->last_type;
    if (map_lpost) map_lpost[i] = c->map_lpost;
  }
  return 0;
}
int pto_get_history(pto_handle *h, int32_t ladder, int32_t rung, int64_t first, int64_t count,
                    double *x, double *lpost, double *llike, double *acc, double *beta, int32_t *type) {
  if (ladder < 0 || ladder >= h->L || rung < 0 || rung >= h->R) return fail(PTG_EINVAL, "bad chain");
  chain_t *c = &h->chains[(size_t)ladder * h->R + rung];
  if (first < 0 || first + count > c->nsize) return fail(PTG_EINVAL, "history range");
  if (x) memcpy(x, c->hx + (size_t)first * h->d, (size_t)count * h->d * sizeof(double));
  if (lpost) memcpy(lpost, c->hlpost + first, (size_t)count * sizeof(double));
  if (llike) memcpy(llike, c->hllike + first, (size_t)count * sizeof(double));
  if (acc) memcpy(acc, c->hacc + first, (size_t)count * sizeof(double));
  if (beta) memcpy(beta, c->hbeta + first, (size_t)count * sizeof(double));
  if (type) memcpy(type, c->htype + first, (size_t)count * sizeof(int32_t));
  return 0;
}
int pto_get_swap_stats(pto_handle *h, int64_t *sc, int64_t *sa, int32_t *dir, int32_t *ups, int32_t *downs, int32_t *inst) {
  for (int l = 0; l < h->L; l++)
    for (int r = 0; r < h->R - 1; r++) {
      if (sc) sc[(size_t)l * (h->R - 1) + r] = h->swap_count[(size_t)l * (h->R - 1) + r];
      if (sa) sa[(size_t)l * (h->R - 1) + r] = h->swap_accept[(size_t)l * (h->R - 1) + r];
    }
  if (dir) memcpy(dir, h->directions, (size_t)h->nchains * sizeof(int32_t));
  if (ups) memcpy(ups, h->ups, (size_t)h->nchains * sizeof(int32_t));
  if (downs) memcpy(downs, h->downs, (size_t)h->nchains * sizeof(int32_t));
  if (inst) memcpy(inst, h->instances, (size_t)h->nchains * sizeof(int32_t));
  return 0;
}
/* rung-sharded ladders: restatement of ptg_boundary_pack / ptg_boundary_swap (host memory) */
int pto_boundary_pack(pto_handle *h, int32_t rung, void *out) {
  double *o = (double *)out;
  for (int l = 0; l < h->L; l++) {
    chain_t *c = &h->chains[(size_t)l * h->R + rung];
    double *rec = o + (size_t)l * (h->d + 3);
    memcpy(rec, c->x, (size_t)h->d * sizeof(double));
    rec[h->d] = c->llike; rec[h->d + 1] = prior_eval_log(h, c->x, 1); rec[h->d + 2] = c->beta;
  }
  return 0;
}
int pto_boundary_swap(pto_handle *h, int32_t my_rung, const void *nb, int32_t i_am_lower, uint64_t shared_seed, int64_t boundary_id, int64_t exchange_index) {
  const double *pk = (const double *)nb;
  for (int l = 0; l < h->L; l++) {
    chain_t *c = &h->chains[(size_t)l * h->R + my_rung];
    const double *rec = pk + (size_t)l * (h->d + 3);
    const double nb_ll = rec[h->d], nb_lprior = rec[h->d + 1], nb_beta = rec[h->d + 2];
    double lla = i_am_lower ? c->llike : nb_ll; if (!(lla > -1e200)) lla = -1e200;
    double llb = i_am_lower ? nb_ll : c->llike; if (!(llb > -1e200)) llb = -1e200;
    const double ba = i_am_lower ? c->beta : nb_beta, bb = i_am_lower ? nb_beta : c->beta;
    const double lhr = -(bb - ba) * (llb - lla);
    int accept = 1;
    if (lhr < 0) {
      uint32_t w[4];
      ptg_philox_draw(shared_seed, (uint64_t)(h->cfg.ladder_offset + l) * PTG_STREAM_STRIDE + PTG_STREAM_LADDER, PTG_DOMAIN_BOUNDARY, (uint64_t)exchange_index,
                      (uint32_t)boundary_id, w);
      accept = (log(ptg_u52_to_unit(w[0], w[1])) < lhr);
    }
    if (accept) add_state(h, c, rec, nb_ll, nb_lprior + c->beta * nb_ll);
    else add_state(h, c, c->x, c->llike, c->lpost);
    if (i_am_lower) { h->swap_count[(size_t)l * (h->R > 1 ? h->R - 1 : 1) + (my_rung < h->R - 1 ? my_rung : 0)] += 0; }
  }
  return 0;
}
int pto_get_trace(pto_handle *h, int64_t first, int64_t count, double *lhr, int32_t *code) {
  if (first < 0 || first + count > h->cfg.trace_steps || first + count > h->istep) return fail(PTG_EINVAL, "trace range");
  if (lhr) memcpy(lhr, h->trace_lhr + (size_t)first * h->nchains, (size_t)count * h->nchains * sizeof(double));
  if (code) memcpy(code, h->trace_code + (size_t)first * h->nchains, (size_t)count * h->nchains * sizeof(int32_t));
  return 0;
}
int pto_get_total_steps(pto_handle *h, int64_t *total) { *total = h->total_steps; return 0; }
int pto_eval_loglike(pto_handle *h, const double *x, int64_t n, double *out) {
  for (int64_t i = 0; i < n; i++) out[i] = like_eval(h, x + (size_t)i * h->d);
  return 0;
}
int pto_eval_logprior(pto_handle *h, const double *x, int64_t n, double *out) {
  for (int64_t i = 0; i < n; i++) {
    double t[PTG_MAX_DIM]; memcpy(t, x + (size_t)i * h->d, (size_t)h->d * sizeof(double));
    int valid = space_enforce(h, t);
    out[i] = prior_eval_log(h, t, valid);
  }
  return 0;
}
