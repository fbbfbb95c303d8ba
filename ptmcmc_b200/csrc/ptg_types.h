// ptg_types.h -- host/device shared plain structs of the ptg engine (no CUDA-only constructs).
#ifndef PTG_TYPES_H
#define PTG_TYPES_H
#include <stdint.h>
#include "../../include/ptmcmc_b200.h"

#define PTG_TPC_MAX_DIM 16   // thread-per-chain kernels
// doubles per record of the history x-ring: the thread-per-chain kernels (dim <= 16) pad a record to a multiple of 4 doubles so that
// a record is a whole number of 32-byte sectors (a DE gather at dim <= 4 is exactly one sector, one 256-bit load); the warp-per-chain
// kernels read whole records coalesced and keep the natural stride
#define PTG_HX(dim) ((dim) <= PTG_TPC_MAX_DIM ? (((dim) + 3) & ~3) : (dim))
#define PTG_SWAP_SLOTS 64    // swap trials per PT step: maxswapsperstep = 1+2*swap_rate*Ntemps (chain.cc:1192), capped (the oracle caps at 64 too)

// 1-D prior factor (ProbabilityDist.h:76-257)
struct PtgPrior1D {
  int32_t kind, pad;
  double a, b;         // (xmin,xmax) or (x0,sigma)
  double norm, cdfoff; // polar/copolar
  double la, lb;       // log
};

struct PtgProp {
  int32_t kind, has_transform;
  double snooker, g1frac, ignore_frac, unlikely_alpha, reduce_gamma, one_d_frac;
  double gamma_std;           // 1.68/sqrt(d)/reduce_gamma, proposal_distribution.cc:495
  int32_t sigma_off, trans_off; // offsets (in doubles) into Model::prop_data
};

// Everything the kernels need to know about the model; passed by value (__grid_constant__).
struct PtgModel {
  int32_t dim, n_rungs, n_ladders, n_props;
  int32_t save_every, hist_cap, n_init, maxswaps;
  int32_t swap_mode, record_full, wrap_in_set, zero_valid;
  int32_t like_kind, n_lparams, trace_steps, all_uniform_prior;
  int32_t any_bound, hx;      // any_bound: some dimension has a non-open boundary (state::enforce does work); hx: doubles per history x-record (ptg_hx)
  int64_t n_ldata;
  int64_t n_chains;
  double swap_rate, dprior_min, evolve_rate, evolve_lpost_cut;
  double like_nsum;
  double like_t0, like_dt;    // data chi^2 likelihoods: abscissae are the uniform grid t0 + i dt (like_uniform_t), e.g. config C2's time samples
  int32_t like_uniform_t, de_mixing; // de_mixing: temperature mixing of a bare DE proposal's history draws (proposal_distribution.cc:594-741)
  double adapt_rate, de_Tmix, Tpow;  // adaptive shares of the set (proposal_distribution.cc:132-166); reset_bins' thermal exponent
  double hot_norm[PTG_MAX_PROPOSALS]; // normalised hot shares of the top-level slots (constructor, proposal_distribution.cc:72-79)
  // one nested proposal set (ptg_set_nested_set): members [nest_first, nest_first + nest_count) share top-level slot nest_first;
  // n_slots = top-level slots, n_bins = n_slots + nest_count = entries per rung / per chain of the bins and shares tables
  int32_t nest_first, nest_count, n_slots, n_bins;
  double nest_adapt;
  double uniform_lprior;      // log(prod 1/(b-a)) when every factor is uniform (evaluated once on the device)
  uint64_t seed;
  int64_t ladder_offset;
  int32_t lower[PTG_TPC_MAX_DIM], upper[PTG_TPC_MAX_DIM];
  double xmin[PTG_TPC_MAX_DIM], xmax[PTG_TPC_MAX_DIM];
  PtgPrior1D prior[PTG_TPC_MAX_DIM];
  PtgProp props[PTG_MAX_PROPOSALS];
  // per-dimension tables in device memory for every dim (the wide, warp-per-chain kernels index these; dim <= 16 kernels
  // use the by-value copies above)
  const int32_t *lower_w, *upper_w;
  const double *xmin_w, *xmax_w;
  const PtgPrior1D *prior_w;
  const double *lparams;      // device
  const double *ldata;        // device
  const double *prop_data;    // device: sigmas / transforms
  const double *bins;         // device [n_rungs][n_bins] cumulative shares (reset_bins): top-level slots, then the nested set's members
};

// Device-resident chain state (SoA over chains) + history + ladder statistics.
struct PtgState {
  double *cur_x;      // [dim][n_chains]
  double *lpost, *llike, *lprior, *beta; // [n_chains]
  double *map_lpost;  // [n_chains]
  double *map_x;      // [dim][n_chains]
  long long *nhist, *nsize, *ntries, *naccept; // [n_chains]
  int32_t *last_type; // [n_chains]
  double *hist;       // [n_chains][hist_cap][hx]  x-ring: record = x[dim], padded to whole 32-byte sectors for dim <= 16 (ptg_hx)
  double *hist_lp;    // [n_chains][hist_cap][2]   (lpost, llike) of the same slot: read only by dumps, evidence and unlikely_alpha
  double *hist_acc, *hist_beta; // [n_chains][hist_cap] (record_full)
  int32_t *hist_type; // [n_chains][hist_cap]        (record_full)
  long long *swap_count, *swap_accept; // [n_ladders][n_rungs]
  int32_t *directions, *ups, *downs, *instances; // [n_ladders][n_rungs]
  // tapes (PTG_RNG_TAPE)
  const double *tape_u, *tape_z;
  long long *u_pos, *z_pos;            // [n_streams] cursors
  const long long *u_end, *z_end;      // [n_streams]
  const long long *u_mark, *z_mark;    // optional [n_mark_steps][n_streams]: cursor of every stream at the start of PT step s
  long long n_mark_steps;
  // trace
  double *trace_lhr;  // [trace_steps][n_chains]
  int32_t *trace_code;
  // host-callback likelihood mode (ptg_register_evaluate_log): the parked proposal of every chain
  double *pend_x;                       // [n_chains][dim]  (row-major: handed to the caller's batched likelihood as is)
  double *pend_lprior, *pend_lh, *pend_like; // [n_chains]
  int32_t *pend_type, *pend_flags;      // flags: bit0 valid, bit1 gate (likelihood wanted), bit2 MH step pending
  uint32_t *pend_w;                     // [n_chains][2] acceptance-draw words
  // adaptive shares: every chain's clone of the proposal set owns shares, bins, last_accepted (bit per member) and adapt_count
  double *ad_shares, *ad_bins;          // [n_chains][n_props]
  int32_t *ad_last, *ad_count;          // [n_chains]: last_accepted bits (top-level slots in bits 0-15, nested members in bits 16-31), top adapt_count
  int32_t *ad_count2;                   // [n_chains]: the nested set's adapt_count
  int32_t *err;       // device error flag (PTG_ETAPE, PTG_ESTUCK)
};
// Rung-sharded ladders, exchange fused into the production step kernel over NVLink peer memory (ptg_fast.cuh).
// Every rank owns one exchange area:  edges [2 parity][2 edge: 0 = coldest rung, 1 = hottest rung][n_ladders][dim + 3] doubles
// (x, llike, lprior, beta) followed by flags [2 edge][n_ladders] int32 = number of publishes completed for that ladder's edge.
// The epilogue of launch p publishes into parity p & 1 and raises the flags to p + 1; the prologue of the NEXT launch waits for the
// neighbour's flag to reach p + 1, reads the neighbour's record through its peer pointer and runs the boundary swap trial.
#define PTG_ERR_XCHG_TIMEOUT 6      // device error flag: a boundary-exchange wait was aborted
struct PtgXchg {
  int on, swap_in, publish_out, has_lo, has_hi;
  int every;                         // > 0: also exchange INSIDE the launch after every `every`-th iteration
  long long index;                   // publish index of this launch's epilogue; its prologue consumes index - 1
  double *my_edges; int *my_flags;
  const double *lo_edges; const int *lo_flags; // colder neighbour (its hottest rung pairs with my coldest)
  const double *hi_edges; const int *hi_flags; // hotter neighbour (its coldest rung pairs with my hottest)
  unsigned long long shared_seed;
  const int *abort;                  // watchdog word in mapped host memory: non-zero = every wait gives up (ptg_xchg_abort)
  long long lo_boundary, hi_boundary; // boundary ids for the Philox address (= rank of the pair's lower block)
};

#endif
