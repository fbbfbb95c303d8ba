/* ptmcmc_b200.h -- C ABI of the B200-native chain-stepping engine ("ptg").
 *
 * This is the drop-in boundary for the hot path of JohnGBaker/ptmcmc:
 *     parallel_tempering_chains::step      chain.cc:1393-1761
 *       -> swap scheduling + swap tests    chain.cc:1410-1538   (+ pry_temps chain.cc:1809-1846)
 *       -> MH_chain::step(prop)            chain.cc:966-1022
 *            -> proposal_distribution_set::draw / gaussian_prop::draw / differential_evolution::draw
 *                                           proposal_distribution.cc:99-129,489-591,744-801 ; .hh:119-129,194-218
 *            -> state::enforce             states.cc:11-58,86-102,161-166
 *            -> prior evaluate_log         probability_function.hh:59 ; .cc:49-81,156-166,281-304
 *            -> likelihood evaluate_log    (device functor; bayesian.hh:553-581 is the host hook it replaces)
 *            -> MH_chain::add_state        chain.cc:916-949
 * batched over n_ladders independent temperature ladders resident on ONE GPU.
 *
 * The reference has no C ABI: its seams are C++ virtuals (chain.hh:34-142, proposal_distribution.hh:38-88,
 * probability_function.hh:31-83).  Every entry point below cites the reference interface it stands in for;
 * INTEGRATION.md shows the C++ facade (`gpu_parallel_tempering_chains : chain`) a maintainer adds on the
 * reference side to bind these.
 *
 * Conventions: plain pointers and sizes only; all arrays are HOST memory unless a name ends in `_dev`;
 * every call returns 0 on success or a negative PTG_E* code, with a message in ptg_last_error();
 * no C++ exception crosses the boundary; one handle = one GPU = one host thread at a time.
 * Chains are indexed  chain = ladder * n_rungs + rung  (rung 0 = coldest, beta = 1).
 * There is NO CPU fallback: ptg_create fails if no CUDA device is usable.
 */
#ifndef PTMCMC_B200_H
#define PTMCMC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PTG_ABI_VERSION 1
#define PTG_MAX_DIM 128        /* thread-per-chain kernels: dim 1-10, 12, 16; warp-per-chain kernels: dim 17-128 (n_rungs <= 32;
                                  likelihoods flat / gaussian / full-covariance gaussian, no prior-draw member) */
#define PTG_MAX_PROPOSALS 16
#define PTG_MAX_RUNGS 64

/* error codes */
#define PTG_OK 0
#define PTG_EINVAL (-1)   /* bad argument / call order            */
#define PTG_ECUDA (-2)    /* CUDA runtime error                   */
#define PTG_ETAPE (-3)    /* injected tape exhausted              */
#define PTG_ESTUCK (-4)   /* init could not draw a valid state (chain.cc:854-866) */
#define PTG_ENOMEM (-5)
#define PTG_EXCHANGE (-6) /* a cross-GPU boundary exchange was aborted (neighbour missing / watchdog): the handle refuses further steps */

/* boundary types: states.hh:35-38 */
enum { PTG_BOUND_OPEN = 0, PTG_BOUND_LIMIT = 1, PTG_BOUND_REFLECT = 2, PTG_BOUND_WRAP = 3 };
/* 1-D prior factor types: mixed_dist_product, probability_function.hh:151-155 */
enum { PTG_PRIOR_UNIFORM = 1, PTG_PRIOR_GAUSSIAN = 2, PTG_PRIOR_POLAR = 3, PTG_PRIOR_COPOLAR = 4, PTG_PRIOR_LOG = 5,
       /* a Gaussian factor of gaussian_dist_product(..., wrap_probability = true) (probability_function.cc:57-78): on a dimension whose
        * boundary wraps, the pdf sums the images x +- k (xmax - xmin) until a pair of images adds less than 1e-12 (at most 100 pairs) */
       PTG_PRIOR_GAUSSIAN_WRAPPED = 6 };
/* device likelihood functors */
enum {
  PTG_LIKE_FLAT = 0,          /* log L = 0                                  (example.cc:22-72 "constant") */
  PTG_LIKE_GAUSS_ISO = 1,     /* lnnormfac - r^2/twosigmasq                  (example.cc:116-143)          */
  PTG_LIKE_SINES = 2,         /* sin^4 multimodal surface                    (sines.hh:22-54)              */
  PTG_LIKE_POLY_CHI2 = 3,     /* chi^2 of polynomial model over data         (bayesian.hh:595-622, poly_example.cc:85-106) */
  PTG_LIKE_SINUSOID_CHI2 = 4, /* chi^2 of sum of sinusoids over data         (same chi^2) */
  PTG_LIKE_GAUSS_FULLCOV = 5, /* like0 - x^T Cinv x / 2                      (cython/exampleGaussian.py:103-109) */
  PTG_LIKE_HOST_CALLBACK = 6, /* the caller's own likelihood, evaluated on the host for all chains at once (ptg_register_evaluate_log) */
  PTG_LIKE_SHELL2D = 7,       /* 2-D double Gaussian shell, reflected in p0  (example.cc:147-222)  params: lnnormfac, twosigmasq, r0, x0[2] */
  PTG_LIKE_SHELLS = 8         /* d-dim Gaussian shell pair (one or two, optionally in ln p0) (example.cc:226-421)
                                 params: lnnormfac, twosigmasq, r0, x0, sigmapoverm, lnsigmapoverm, logx (0/1) */
};
/* proposal kinds (members of a proposal_distribution_set, proposal_distribution.cc:99-129) */
enum { PTG_PROP_DE = 1, PTG_PROP_GAUSS = 2, PTG_PROP_PRIOR_DRAW = 3 };
enum { PTG_SWAP_REFERENCE = 0, PTG_SWAP_EVEN_ODD = 1 };
enum { PTG_RNG_PHILOX = 0, PTG_RNG_TAPE = 1 };
/* history record level: 0 = x, lpost, llike per stored sample; 1 = + acceptance_ratio, invtemp, type
 * (the full record of MH_chain::add_state, chain.cc:936-943) */
enum { PTG_RECORD_BASIC = 0, PTG_RECORD_FULL = 1 };

typedef struct ptg_handle ptg_handle;

/* Stands in for the constructor arguments of parallel_tempering_chains (chain.cc:1163) and
 * MH_chain (chain.cc:647), plus evolve_temps (chain.hh:302) and the RNG seeding of chain::chain() (chain.hh:58). */
typedef struct ptg_config {
  int32_t abi_version;      /* must be PTG_ABI_VERSION */
  int32_t device;           /* CUDA device ordinal */
  int32_t n_ladders;        /* independent ladders on this GPU */
  int32_t n_rungs;          /* Ntemps; 1 = plain MH_chain */
  int32_t dim;
  int32_t save_every;       /* add_every_N */
  int32_t hist_capacity;    /* history slots per chain (ring once exceeded) */
  int32_t n_init;           /* Ninit: prior draws that seed each chain's history (ptmcmc.cc:86) */
  int32_t swap_mode;        /* PTG_SWAP_* */
  int32_t rng_mode;         /* PTG_RNG_* */
  int32_t record_level;     /* PTG_RECORD_* */
  int32_t trace_steps;      /* >0: keep a decision trace for this many PT steps (parity runs) */
  double swap_rate;         /* pt_swap_rate */
  double Tmax;              /* pt_Tmax: geometric ladder 1..Tmax (chain.cc:1181-1183) */
  double dprior_min;        /* chain_dprior_min (minPrior, chain.cc:980) */
  double evolve_rate;       /* pt_evolve_rate; 0 = fixed temperatures */
  double evolve_lpost_cut;  /* pt_evolve_lpost_cut; <0 = off */
  uint64_t seed;            /* Philox key */
  int64_t ladder_offset;    /* global id of local ladder 0: makes results invariant to the GPU count */
} ptg_config;

/* one member of the proposal set; fields of differential_evolution (proposal_distribution.hh:361-414)
 * and gaussian_prop (proposal_distribution.hh:145-227) */
typedef struct ptg_proposal {
  int32_t kind;             /* PTG_PROP_* */
  int32_t reserved;
  double share;             /* static share (normalised by the engine like reset_bins, proposal_distribution.cc:37-59) */
  double hot_share;         /* share at beta -> 0 when Tpow > 0 */
  /* PTG_PROP_DE */
  double snooker, gamma_one_frac, b_small, ignore_frac, unlikely_alpha, reduce_gamma;
  /* PTG_PROP_GAUSS */
  double one_d_frac;
  const double *sigmas;     /* [dim] */
  const double *transform;  /* [dim*dim] row-major eigenvector matrix M (offset = M (z o sigma)), or NULL = identity */
} ptg_proposal;

const char *ptg_last_error(void);
int ptg_abi_version(void);

/* lifecycle ------------------------------------------------------------------------------------ */
int ptg_create(const ptg_config *cfg, ptg_handle **out);
int ptg_destroy(ptg_handle *h);

/* model set-up (all before ptg_init_*) --------------------------------------------------------- */
/* stateSpace::set_bound (states.hh) per dimension */
int ptg_set_space(ptg_handle *h, const int32_t *lower_type, const int32_t *upper_type,
                  const double *xmin, const double *xmax);
/* uniform/gaussian/mixed_dist_product: per dimension (type, a, b): uniform/polar/copolar/log (xmin,xmax); gaussian (x0,sigma) */
int ptg_set_prior(ptg_handle *h, const int32_t *type, const double *a, const double *b);
/* device likelihood functor; replaces bayes_likelihood::register_evaluate_log (bayesian.hh:544-552).
 *   GAUSS_ISO      params = [lnnormfac, twosigmasq, x0[dim]]
 *   SINES          params = [height, step_scale, k[dim], min[dim], max[dim]]
 *   POLY_CHI2      params = [like0]; data = [x[N], y[N], var[N]]  (n_data = 3N doubles)
 *   SINUSOID_CHI2  params = [like0]; data = [t[N], y[N], var[N]]  (n_data = 3N); state = (A,f,phi) triples
 *   GAUSS_FULLCOV  params = [like0]; data = Cinv[dim*dim] row-major (n_data = dim*dim) */
int ptg_set_likelihood(ptg_handle *h, int32_t kind, const double *params, int32_t n_params,
                       const double *data, int64_t n_data);
/* Host-callback likelihood: the batched form of bayes_likelihood::register_evaluate_log / register_reference_object
 * (bayesian.hh:544-552) for likelihoods that exist only as host code (C++ classes, Python through Cython).  `fn(user, x, n, out)`
 * receives n states x[n][dim] (row-major) and writes their log-likelihoods; non-finite values count as -inf (bayesian.hh:569-575).
 * Everything else of the step stays on the GPU: per PT iteration one kernel runs the swap phase and generates every chain's
 * proposal, the proposals that pass the prior gate (chain.cc:980) are handed to `fn` in ONE call, a second kernel does the
 * Metropolis test and the history append.  Replaces ptg_set_likelihood.  Philox draws, dim <= 16. */
typedef void (*ptg_batch_loglike_fn)(void *user, const double *x, int64_t n, double *loglike);
int ptg_register_evaluate_log(ptg_handle *h, ptg_batch_loglike_fn fn, void *user);
/* proposal_distribution_set constructor (proposal_distribution.cc:61-93).  wrap_in_set=0 (n must be 1) uses the
 * single proposal bare, as ptmcmc drivers do when they pass e.g. a differential_evolution directly to
 * set_proposal (testMH.cpp:91-160): no selection draw, type() not multiplied by 10. */
int ptg_set_proposals(ptg_handle *h, int32_t n, const ptg_proposal *props, double Tpow, int32_t wrap_in_set);
/* Options of the proposal objects beyond their constructors (call after ptg_set_proposals, before initialising):
 *   adapt_rate  != 0: the adaptive shares of proposal_distribution_set (proposal_distribution.cc:132-166; ptmcmc's prop_adapt_rate with
 *               prop_adapt_more): every chain owns its copy of the shares; two accepts or two rejects in a row of member i scale
 *               shares[i] by 1 - adapt_rate / 4, and from the (10 n)-th decision on reset_bins renormalises after every decision.
 *               Needs a set (wrap_in_set = 1).
 *   de_mixing   != 0: differential_evolution::support_mixing(true) on a BARE differential-evolution proposal (wrap_in_set = 0) of a
 *               ladder: every history draw first weighs the rungs by 10 + 10 log-likelihood samples of each (draw_from_chain,
 *               proposal_distribution.cc:594-741) and then draws from the chosen rung's history.  Inside a set the reference never
 *               mixes (proposal_distribution_set does not override support_mixing(), chain.cc:1375), so the flag is refused there.
 *               Needs unlikely_alpha = 0, dim <= 16, n_rungs <= 32.
 *   de_Tmix     temperature_mixing_factor (mix_temperatures_more, proposal_distribution.hh:403; ptmcmc's de_Tmix).
 * Both run in the tape-capable warp kernel (PTG_KERNEL_WARP) in either RNG mode. */
int ptg_set_proposal_options(ptg_handle *h, double adapt_rate, int32_t de_mixing, double de_Tmix);
/* Members [first, first + count) of the list given to ptg_set_proposals form ONE NESTED proposal_distribution_set that takes a single slot of
 * the top-level set (at position `first`) with top-level share `share` (`hot_share` when Tpow > 0); the members' own `share` fields are
 * their shares INSIDE the nested set; adapt_rate is the nested set's own (ptmcmc_sampler::select_proposal builds exactly this for
 * prop_adapt_rate > 0: the six Gaussian scales in an adaptive sub-set, the top level adaptive only with prop_adapt_more; ptmcmc.cc:70-72,
 * 123-143).  A step that selects the nested slot draws a second selection uniform (Philox block PTG_BLK_NEST) and reports the type
 * slot + 10 (j + 10 member_type).  Call after ptg_set_proposals / ptg_set_proposal_options, before initialising.  Same kernel and limits as
 * the adaptive shares. */
int ptg_set_nested_set(ptg_handle *h, int32_t first, int32_t count, double share, double hot_share, double adapt_rate);
/* the current shares of every chain's proposal set, shares[n_chains][n_slots + count]: the top-level slots (n_slots = n_props, or
 * n_props - count + 1 with a nested set), then the nested set's members (proposal_distribution_set::report(1)) */
int ptg_get_proposal_shares(ptg_handle *h, double *shares);
/* explicit inverse temperatures instead of the geometric ladder; [n_ladders*n_rungs] or NULL */
int ptg_set_betas(ptg_handle *h, const double *betas);

/* random numbers --------------------------------------------------------------------------------- */
int ptg_seed(ptg_handle *h, uint64_t seed);
/* PTG_RNG_TAPE: injected draws.  Streams 0..n_chains-1 are the chains' own generators (chain.hh:45),
 * streams n_chains..n_chains+n_ladders-1 the ladders' (parallel_tempering_chains is itself a chain).
 * u_off/z_off have n_streams+1 entries; the tapes are copied to the device. */
int ptg_inject_tapes(ptg_handle *h, const double *u, const int64_t *u_off, const double *z, const int64_t *z_off);
/* optional: absolute tape cursors of every stream at the start of PT step s, [n_steps][n_streams] each.  The engine
 * re-synchronises to them at every step boundary, so a last-ulp libm difference that changes how many draws one
 * step CONSUMES (e.g. lhr = -1e-21 vs 0: same decision, one acceptance draw fewer; chain.cc:998-1001) stays local
 * to that step instead of shifting the rest of the tape. */
int ptg_inject_tape_marks(ptg_handle *h, int64_t n_steps, const int64_t *u_mark, const int64_t *z_mark);

/* initialisation: MH_chain::initialize (chain.cc:846-876) ------------------------------------------- */
int ptg_init_from_prior(ptg_handle *h);
/* seed each chain's history from caller-provided states x[n_chains][n_init][dim] (chain_init_file path) */
int ptg_init_states(ptg_handle *h, const double *x);

/* the hot path: n_steps iterations of parallel_tempering_chains::step for every ladder ------------------ */
int ptg_step(ptg_handle *h, int64_t n_steps);
int ptg_synchronize(ptg_handle *h);
/* Kernel selection is automatic: ladders of <= 32 rungs run register-resident warp-shuffle kernels (PTG_KERNEL_FAST for
 * Philox draws, PTG_KERNEL_WARP for tape replay), longer ladders the shared-memory kernel (PTG_KERNEL_SHARED).
 * ptg_select_kernel pins PTG_KERNEL_WARP or PTG_KERNEL_SHARED where they apply (tests / profiling: in Philox mode all
 * three produce bit-identical chains); ptg_get_launch_count reports how many step kernels this handle has launched. */
enum { PTG_KERNEL_AUTO = 0, PTG_KERNEL_SHARED = 1, PTG_KERNEL_WARP = 2, PTG_KERNEL_FAST = 3,
       PTG_KERNEL_FAST_GENERAL = 4 /* the production kernel's general instantiation even where a streamlined one applies (tests: bit-identical) */ };
int ptg_select_kernel(ptg_handle *h, int32_t kernel);
int ptg_get_launch_count(ptg_handle *h, int64_t *n);
/* same, end-to-end with host buffers: steps, then copies the cold chains' newest `n_out` stored samples
 * of every ladder to host memory x_out[n_ladders][n_out][dim], lpost_out/llike_out[n_ladders][n_out] */
int ptg_step_host(ptg_handle *h, int64_t n_steps, int32_t n_out, double *x_out, double *lpost_out, double *llike_out);
/* the same in two halves: _begin enqueues the steps, the gather and the device-to-host copies (on a copy stream of the handle, double-
 * buffered staging) and returns; _wait blocks until the samples of every begun block have landed.  With pinned host buffers the copy of
 * block k overlaps the kernel of block k+1: the shape of the reference's run loop, which dumps every cold sample of a block while the
 * next block is already stepping would (ptmcmc.cc:601-616, chain.cc:1112-1135). */
int ptg_step_host_begin(ptg_handle *h, int64_t n_steps, int32_t n_out, double *x_out, double *lpost_out, double *llike_out);
int ptg_step_host_wait(ptg_handle *h);
/* effective capacity of the history ring (hist_capacity = 0 at creation means n_init + 1024) */
int ptg_get_hist_capacity(ptg_handle *h, int32_t *capacity);

/* the device functors applied to caller-provided states x[n][dim] (host): log-likelihood (the batched form of
 * bayes_likelihood::evaluate_log, bayesian.hh:553-581) and log-prior after boundary enforcement
 * (probability_function.hh:59, states.cc:161-166).  Either output may be NULL.  Needs prior+likelihood+proposals set. */
int ptg_eval(ptg_handle *h, const double *x, int64_t n, double *loglike, double *logprior);

/* read-back: chain::getState/getLogPost/getLogLike/invTemp (chain.hh:86-124) ---------------------------- */
int ptg_get_current(ptg_handle *h, double *x, double *lpost, double *llike, double *beta);
int ptg_get_lprior(ptg_handle *h, double *lprior);
int ptg_get_counters(ptg_handle *h, int64_t *nhist, int64_t *nsize, int64_t *ntries, int64_t *naccept,
                     int32_t *last_type, double *map_lpost);
/* raw-indexed history of one chain, elements [first, first+count); entries older than the ring are an error.
 * acc/beta/type may be NULL (and must be unless record_level = FULL) */
int ptg_get_history(ptg_handle *h, int32_t ladder, int32_t rung, int64_t first, int64_t count,
                    double *x, double *lpost, double *llike, double *acc, double *beta, int32_t *type);
/* swap_count / swap_accept_count (chain.hh:244-245) [n_ladders][n_rungs-1]; directions/ups/downs/instances
 * (chain.hh:238-242) [n_ladders][n_rungs]; any pointer may be NULL */
int ptg_get_swap_stats(ptg_handle *h, int64_t *swap_count, int64_t *swap_accept, int32_t *directions,
                       int32_t *ups, int32_t *downs, int32_t *instances);
/* decision trace (parity harness): for PT step s in [first, first+count) and every chain,
 * lhr[s][chain] = log Hastings ratio tested, code[s][chain] = see PTG_TRACE_* */
int ptg_get_trace(ptg_handle *h, int64_t first, int64_t count, double *lhr, int32_t *code);
#define PTG_TRACE_TYPE_MASK 0xff     /* proposal type() as recorded by the set (i + 10*member type) */
#define PTG_TRACE_ACCEPT 0x100
#define PTG_TRACE_INVALID 0x200      /* proposed state failed boundary enforcement */
#define PTG_TRACE_SWAPPED 0x400      /* rung took part in a swap trial this step (no MH update) */
#define PTG_TRACE_NOLIKE 0x800       /* prior gate skipped the likelihood (chain.cc:980-987) */
/* total history appends (= tempered chain-steps, the BASELINE metric) since init */
int ptg_get_total_steps(ptg_handle *h, int64_t *total);
/* device pointers + stream for zero-copy consumers (torch / NCCL gather of cold samples): history records
 * hist_dev[chain][hist_capacity][hx] = x[dim] padded to hx doubles (hx = dim rounded up to a multiple of 4 for dim <= 16, whole
 * 32-byte sectors per record; hx = dim for dim > 16), *hist_stride = hist_capacity * hx = doubles per chain; (lpost, llike) of the
 * same slot live in a separate ring that ptg_get_history reads;
 * cur_x_dev[dim][n_chains] */
int ptg_get_device_views(ptg_handle *h, void **hist_dev, void **cur_x_dev, void **stream, int64_t *hist_stride);
/* run all further work of this handle on a caller-owned cudaStream_t (e.g. torch's current stream) */
int ptg_set_stream(ptg_handle *h, void *cuda_stream);
/* overwrite the current state of every chain from host arrays x[n_chains][dim], lpost, llike, lprior[n_chains]
 * (the restart path of MH_chain, chain.cc:691-731; also what a host facade that keeps `state` objects
 * authoritative does before each block of steps).  Asynchronous on the handle's stream. */
int ptg_set_current(ptg_handle *h, const double *x, const double *lpost, const double *llike, const double *lprior);

/* Thermodynamic-integration evidence, computed on the device from the history ring:
 *   ptg_get_mean_loglike: mean log-likelihood of every chain over its newest `n_last` stored samples -> mean_ll[n_chains];
 *   ptg_get_log_evidence: per ladder, the reference's trapezoid over the rungs plus its tail term below the hottest rung
 *                         (parallel_tempering_chains::log_evidence_ratio and "Total log-evidence", chain.cc:1582-1600,1984-2012)
 *                         -> log_evidence[n_ladders]. */
int ptg_get_mean_loglike(ptg_handle *h, int32_t n_last, double *mean_ll);
/* Device-side ESS input: integrated autocorrelation time (Sokal window, c = 5) of every parameter of rung `rung` of every ladder over
 * its newest n_last stored samples, lags < max_lag -> tau[n_ladders][dim]; ESS of a chain = n_last / max_j tau_j (in stored samples). */
int ptg_get_act(ptg_handle *h, int32_t rung, int32_t n_last, int32_t max_lag, double *tau);
int ptg_get_log_evidence(ptg_handle *h, int32_t n_last, double *log_evidence);
/* The reference's own effective-sample-size recipe, device part: the windowed lag statistics of chain::compute_autocovar_windows
 * (chain.cc:126-289) of rung `rung` of every ladder, summed in the reference's order.  Windows are counted in stored records:
 * window k (0 = oldest) of ladder l covers records [end_l - (n_win - k) * swidth, + swidth), end_l = end_rec[l] (NULL: every record
 * stored so far); lag_rec[n_lag] in records, lag_rec[0] = 0; the first n_feat parameters are the features.
 *   means[l][f][k][j] = sum_i (f_i + f_{i - lag_j}) / swidth / 2,   covar[l][f][k][j] = sum_i f_i f_{i - lag_j} / swidth - means^2.
 * chain::compute_effective_samples / report_effective_samples (chain.cc:292-545) combine them on the host
 * (ptmcmc_b200/analysis.py).  Fails if a window or its lagged copy has left the history ring. */
int ptg_get_autocovar_windows(ptg_handle *h, int32_t rung, int32_t swidth, int32_t n_win, int32_t n_lag, const int32_t *lag_rec,
                              const int64_t *end_rec, int32_t n_feat, double *means, double *covar);

/* Rung-sharded ladders (the reference's MPI layout, chain.cc:1290-1311,1433-1435, with block rung assignment): every GPU holds
 * a contiguous block of rungs of EVERY ladder (explicit betas via ptg_set_betas); replica swaps inside a block run in the step
 * kernels, swaps ACROSS a block boundary are a separate exchange step between launches:
 *   ptg_boundary_pack  writes rung `rung` of every ladder as rec[ladder] = (x[dim], llike, lprior, beta) into DEVICE memory
 *                      `out_dev` ([n_ladders][dim+3] doubles) -- the payload that crosses NVLink (NCCL or peer copy);
 *   ptg_boundary_swap  performs the swap trial (chain.cc:1459-1490) between this engine's rung `my_rung` and the neighbour's packed
 *                      rung for every ladder and appends the outcome to this chain's history (both chains of a trial append,
 *                      chain.cc:1487-1534).  Both sides call it with each other's packs; the acceptance draw comes from the ladder's
 *                      Philox stream under `shared_seed` (domain PTG_DOMAIN_BOUNDARY, step = exchange_index, block = boundary_id),
 *                      so they reach the same decision without communicating it -- as the reference replicates swap decisions
 *                      from the shared seed (ptmcmc.cc:263-268).  Engines of different ranks must use different `seed`s for
 *                      their chains' own streams. */
int ptg_boundary_pack(ptg_handle *h, int32_t rung, void *out_dev);
int ptg_boundary_swap(ptg_handle *h, int32_t my_rung, const void *neighbour_pack_dev, int32_t i_am_lower, uint64_t shared_seed,
                      int64_t boundary_id, int64_t exchange_index);

/* The same exchange FUSED into the production step kernel over peer memory (NVLink loads, no collective and no extra launch):
 * every engine owns an exchange area in its own HBM; the epilogue of a ptg_step_exchange launch publishes the block's two edge
 * rungs there and raises a per-ladder flag, and the prologue of the NEXT launch waits (per ladder, bounded) for the neighbour's
 * flag, reads the neighbour's record through a peer pointer and runs the trial of ptg_boundary_swap with the same Philox address
 * (exchange_index = number of publishes so far, boundary ids as given), so fused and unfused runs produce identical chains.
 *   ptg_xchg_export   allocates the area; returns its CUDA IPC handle (64 bytes, for other processes) and/or its device pointer
 *                     (for engines of the same process);
 *   ptg_xchg_connect  takes the colder / hotter neighbour's handle or pointer (NULL at the ends of the ladder) and the two
 *                     boundary ids (= rank of the pair's lower block);
 *   ptg_step_exchange n_steps (<= 16384, may be 0) PT iterations in one launch; apply_pending runs the pending boundary trials
 *                     first, publish publishes the edges at the end.  every > 0 (n_steps a multiple of it): the exchange ALSO runs
 *                     inside the launch after every `every`-th iteration -- each ladder's warp publishes, waits for the neighbour
 *                     GPU's warp of the same ladder and swaps, while the other warps keep stepping; needs the whole grid resident
 *                     (one wave: refused otherwise) and the neighbours to run the same schedule.  every = 0: launch boundaries only.
 * Engines that wait on each other must run on DIFFERENT GPUs (or, with every = 0, be launched strictly one after the other). */
int ptg_xchg_export(ptg_handle *h, void *ipc_handle_64_bytes, void **local_ptr);
int ptg_xchg_connect(ptg_handle *h, const void *colder, const void *hotter, int32_t handles_are_ipc, uint64_t shared_seed,
                     int64_t colder_boundary_id, int64_t hotter_boundary_id);
int ptg_step_exchange(ptg_handle *h, int64_t n_steps, int32_t every, int32_t apply_pending, int32_t publish);
/* Watchdog of the fused exchange.  A boundary wait inside a launch is unbounded by itself; the host bounds it: when a launch
 * overstays the caller's deadline (a neighbour rank died or never launched), ptg_xchg_abort -- callable from another host thread
 * while ptg_synchronize blocks -- makes every wait give up.  The launch then ends, ptg_synchronize returns PTG_EXCHANGE and the handle
 * refuses further steps: the skipped trial left the boundary pair inconsistent, continuing would duplicate or lose a state. */
int ptg_xchg_abort(ptg_handle *h);

/* FP64 peak microbenchmarks on `device` (the FP64 roofline denominators): out[0] = DFMA TFLOP/s,
 * out[1] = DMUL+DADD pairs (the engine's unfused arithmetic) TFLOP/s, out[2] = DMMA (mma.sync.m8n8k4.f64) TFLOP/s, out[3] = SM count */
int ptg_measure_fp64_peaks(int32_t device, double *out);

/* checkpoint / restore of the complete engine state (restart.hh semantics; format in DESIGN.md) */
int ptg_checkpoint(ptg_handle *h, const char *path);
int ptg_restore(ptg_handle *h, const char *path);

#ifdef __cplusplus
}
#endif
#endif /* PTMCMC_B200_H */
