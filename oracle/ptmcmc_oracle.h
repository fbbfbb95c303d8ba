/* ptmcmc_oracle.h -- TEST INFRASTRUCTURE.  CPU restatement ("oracle") of the reference's chain-stepping
 * hot path (JohnGBaker/ptmcmc, chain.cc / proposal_distribution.cc / probability_function.cc / states.cc).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; it is the
 * CHECKER, never the thing measured or shipped.  The product (ptmcmc_b200/) does not import, link or
 * execute anything under oracle/.
 *
 * Parity status: PINNED.  The oracle, driven by its restatement of the reference's own RNG (newran_port.c),
 * reproduces bit-for-bit the per-rung histories that the unmodified reference produces through
 * oracle/_ref/ref_trace (tests/test_oracle_ref.py, fixtures in tests/golden/).
 *
 * The API mirrors include/ptmcmc_b200.h one to one (pto_* instead of ptg_*), so that parity tests run the
 * same call sequence on both.  Extra entry points: the NEWRAN RNG mode and tape recording.
 */
#ifndef PTMCMC_ORACLE_H
#define PTMCMC_ORACLE_H
#include "../include/ptmcmc_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
#define PTO_RNG_NEWRAN 2
typedef struct pto_handle pto_handle;
const char *pto_last_error(void);
int pto_create(const ptg_config *cfg, pto_handle **out);
int pto_destroy(pto_handle *h);
int pto_set_space(pto_handle *h, const int32_t *lower_type, const int32_t *upper_type, const double *xmin, const double *xmax);
int pto_set_prior(pto_handle *h, const int32_t *type, const double *a, const double *b);
int pto_set_likelihood(pto_handle *h, int32_t kind, const double *params, int32_t n_params, const double *data, int64_t n_data);
int pto_set_proposals(pto_handle *h, int32_t n, const ptg_proposal *props, double Tpow, int32_t wrap_in_set);
int pto_set_proposal_options(pto_handle *h, double adapt_rate, int32_t de_mixing, double de_Tmix);
int pto_set_nested_set(pto_handle *h, int32_t first, int32_t count, double share, double hot_share, double adapt_rate);
int pto_get_proposal_shares(pto_handle *h, double *shares);
int pto_set_betas(pto_handle *h, const double *betas);
int pto_seed(pto_handle *h, uint64_t seed);
/* reference RNG: master MotherOfAll(seed); per ladder one draw for the ladder's generator, then one per rung
 * in order (chain.hh:58-59, chain.cc:1323-1326) */
int pto_seed_newran(pto_handle *h, double seed);
int pto_inject_tapes(pto_handle *h, const double *u, const int64_t *u_off, const double *z, const int64_t *z_off);
/* record every uniform / standard normal each stream consumes (any RNG mode) */
int pto_record_tapes(pto_handle *h, int on);
int pto_get_tape_sizes(pto_handle *h, int64_t *u_count, int64_t *z_count); /* [n_streams] each */
int pto_get_tapes(pto_handle *h, double *u, double *z);                    /* concatenated in stream order */
/* draws each stream had consumed when each recorded PT step began, [n_marks][n_streams]; NULL arrays = query n_marks */
int pto_get_tape_marks(pto_handle *h, int64_t *n_marks, int64_t *u_mark, int64_t *z_mark);
int pto_init_from_prior(pto_handle *h);
int pto_init_states(pto_handle *h, const double *x);
int pto_step(pto_handle *h, int64_t n_steps);
int pto_get_current(pto_handle *h, double *x, double *lpost, double *llike, double *beta);
int pto_get_counters(pto_handle *h, int64_t *nhist, int64_t *nsize, int64_t *ntries, int64_t *naccept, int32_t *last_type, double *map_lpost);
int pto_get_history(pto_handle *h, int32_t ladder, int32_t rung, int64_t first, int64_t count,
                    double *x, double *lpost, double *llike, double *acc, double *beta, int32_t *type);
int pto_get_swap_stats(pto_handle *h, int64_t *swap_count, int64_t *swap_accept, int32_t *directions, int32_t *ups, int32_t *downs, int32_t *instances);
int pto_get_trace(pto_handle *h, int64_t first, int64_t count, double *lhr, int32_t *code);
int pto_get_total_steps(pto_handle *h, int64_t *total);
int pto_boundary_pack(pto_handle *h, int32_t rung, void *out);
int pto_boundary_swap(pto_handle *h, int32_t my_rung, const void *nb, int32_t i_am_lower, uint64_t shared_seed, int64_t boundary_id, int64_t exchange_index);
/* stand-alone evaluation helpers for unit tests (n states, row-major x[n][dim]) */
int pto_eval_loglike(pto_handle *h, const double *x, int64_t n, double *out);
int pto_eval_logprior(pto_handle *h, const double *x, int64_t n, double *out);
#ifdef __cplusplus
}
#endif
#endif
