// gpu_parallel_tempering_chains.hh -- the reference-side binding of the ptg engine, compiled against the
// REAL reference headers (chain.hh, proposal_distribution.hh, probability_function.hh of JohnGBaker/ptmcmc).
//
// The class IS a parallel_tempering_chains (chain.hh:206-327): ptmcmc_sampler::run (ptmcmc.cc:563-661) dynamic_casts its chain to
// that type for dumpChain / bestEvidenceErr, and every host-side service of the reference -- dumpChain (chain.cc:1112-1135), status,
// report_effective_samples (chain.cc:549-643), checkpoint in the MHchain.cp / PTchain.cp layout (chain.cc:656-731, 1213-1239) -- keeps
// running UNMODIFIED on the base class's own MH_chain objects.  What changes is who produces the samples:
//   * initialize()   : the reference draws the Ninit start-up samples of every rung (its RNG, the host likelihood), then the engine is
//                      created and takes those samples as its initial histories (ptg_init_states);
//   * step()         : one PT iteration of the engine (ptg_step) instead of chain.cc:1393-1761;
//   * sync()         : every `sync_every` steps the new records of every rung's device history are appended to the base class's
//                      MH_chain vectors (states, lposts, llikes, acceptance_ratio, invtemps, types) and its counters / current state /
//                      temperatures / swap statistics are refreshed -- the lazily synchronised host mirror of SURVEY.md H6.
// The mirror writes private members of MH_chain / parallel_tempering_chains: in-tree this class is a friend of both; the scratch build
// of this repository's tests compiles the translation unit with -fno-access-control instead of touching the reference headers.
// A likelihood becomes GPU-capable by deriving from ptg_device_likelihood and naming its device functor; everything else (priors,
// state space, proposals) is translated from the reference objects themselves.
#pragma once
#include <vector>
#include <string>
#include <iostream>
#include <cstdlib>
#include "chain.hh"
#include "proposal_distribution.hh"
#include "probability_function.hh"
extern "C" {
#include "ptmcmc_b200.h"
}

/// the reference-side hook a likelihood implements to run on the device (next to register_evaluate_log, bayesian.hh:544-552)
struct ptg_device_likelihood {
  virtual ~ptg_device_likelihood() {}
  /// kind = PTG_LIKE_*; params / data as ptg_set_likelihood takes them.  Return false for a host-only likelihood.
  virtual bool describe_device_likelihood(int &kind, std::vector<double> &params, std::vector<double> &data) const = 0;
};

class gpu_parallel_tempering_chains : public parallel_tempering_chains {
  ptg_handle *h;
  ptg_config cfg;
  int nt, ninit, sync_every_;
  long long nsteps;
  bool engine_inited;
  std::vector<long long> mirrored; // records of each rung already appended to the base class's MH_chain
  const stateSpace *space;
  probability_function *the_llike;
  const sampleable_probability_function *the_prior;

  static void check(int rc, const char *what) {
    if (rc) { std::cout << "gpu_parallel_tempering_chains::" << what << ": " << ptg_last_error() << std::endl; exit(1); } // the reference's print-and-exit
  }
  void push_space() {
    const int d = space->size();
    std::vector<int32_t> lt(d), ut(d); std::vector<double> lo(d), hi(d);
    for (int i = 0; i < d; i++) { const boundary b = space->get_bound(i); lt[i] = b.lowertype; ut[i] = b.uppertype; lo[i] = b.xmin; hi[i] = b.xmax; }
    check(ptg_set_space(h, lt.data(), ut.data(), lo.data(), hi.data()), "set_space"); // boundary::open/limit/reflect/wrap = PTG_BOUND_* (states.hh:35-38)
  }
  void push_prior() {
    const int d = space->size();
    std::vector<int32_t> ty(d); std::vector<double> a(d), b(d);
    if (const uniform_dist_product *u = dynamic_cast<const uniform_dist_product *>(the_prior)) {
      for (int i = 0; i < d; i++) { ty[i] = PTG_PRIOR_UNIFORM; a[i] = u->min[i]; b[i] = u->max[i]; }
    } else if (const gaussian_dist_product *g = dynamic_cast<const gaussian_dist_product *>(the_prior)) {
      // wrap_probability (probability_function.cc:57-78): images of wrapped dimensions are added to the pdf
      for (int i = 0; i < d; i++) { ty[i] = g->wrap_probability ? PTG_PRIOR_GAUSSIAN_WRAPPED : PTG_PRIOR_GAUSSIAN; a[i] = g->x0s[i]; b[i] = g->sigmas[i]; }
    } else if (const mixed_dist_product *m = dynamic_cast<const mixed_dist_product *>(the_prior)) {
      for (int i = 0; i < d; i++) {
        ty[i] = m->types[i]; // uniform 1, gaussian 2, polar 3, copolar 4, log 5 = PTG_PRIOR_* (probability_function.hh:151-155)
        const double c = m->centers[i], w = m->halfwidths[i];
        if (ty[i] == mixed_dist_product::gaussian) { a[i] = c; b[i] = w; }
        else if (ty[i] == mixed_dist_product::log) { a[i] = c / w; b[i] = c * w; } // probability_function.cc:243-249
        else { a[i] = c - w; b[i] = c + w; }
      }
    } else { std::cout << "gpu_parallel_tempering_chains: prior type has no device form" << std::endl; exit(1); }
    check(ptg_set_prior(h, ty.data(), a.data(), b.data()), "set_prior");
  }
  void push_likelihood() {
    const ptg_device_likelihood *dl = dynamic_cast<const ptg_device_likelihood *>(the_llike);
    int kind = 0; std::vector<double> params, data;
    if (!dl || !dl->describe_device_likelihood(kind, params, data)) {
      std::cout << "gpu_parallel_tempering_chains: the likelihood names no device functor (ptg_device_likelihood)" << std::endl; exit(1);
    }
    check(ptg_set_likelihood(h, kind, params.data(), (int32_t)params.size(), data.empty() ? nullptr : data.data(), (int64_t)data.size()), "set_likelihood");
  }
  // one member of a proposal set, or a bare proposal
  bool describe_member(proposal_distribution *p, ptg_proposal &q, std::vector<std::vector<double> > &keep) {
    q = ptg_proposal();
    if (differential_evolution *de = dynamic_cast<differential_evolution *>(p)) {
      q.kind = PTG_PROP_DE; q.snooker = de->snooker; q.gamma_one_frac = de->gamma_one_frac; q.b_small = de->b_small;
      q.ignore_frac = de->ignore_frac; q.unlikely_alpha = de->unlikely_alpha; q.reduce_gamma = de->reduce_gamma_fac;
      return true;
    }
    if (gaussian_prop *g = dynamic_cast<gaussian_prop *>(p)) {
      q.kind = PTG_PROP_GAUSS; q.one_d_frac = g->oneDfrac;
      const int d = (int)g->sigmas.size();
      keep.push_back(std::vector<double>(d));
      for (int i = 0; i < d; i++) keep.back()[i] = g->sigmas[i];
      q.sigmas = keep.back().data();
      if (!g->identity_trans) {
        keep.push_back(std::vector<double>((size_t)d * d));
        for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) keep.back()[(size_t)i * d + j] = g->diagTransform(i, j);
        q.transform = keep.back().data();
      }
      return true;
    }
    if (dynamic_cast<draw_from_dist *>(p)) { q.kind = PTG_PROP_PRIOR_DRAW; return true; }
    return false;
  }
  void push_proposal(proposal_distribution &prop) {
    std::vector<ptg_proposal> props; std::vector<std::vector<double> > keep; keep.reserve(64);
    double Tpow = 0, adapt_rate = 0, de_Tmix = 1, nest_share = 0, nest_hot = 0, nest_adapt = 0; int wrap = 0, de_mixing = 0, nest_first = 0, nest_count = 0;
    if (proposal_distribution_set *set = dynamic_cast<proposal_distribution_set *>(&prop)) {
      wrap = 1; Tpow = set->Tpow;
      adapt_rate = set->adapt_rate; // the adaptive shares of the set (proposal_distribution.cc:132-166) run on the device, per rung
      for (size_t i = 0; i < set->proposals.size(); i++) {
        ptg_proposal q;
        // ONE nested set (ptmcmc.cc:123-131: the Gaussian scales of prop_adapt_rate > 0) is flattened into the member list and declared with
        // ptg_set_nested_set; its members' shares are their shares inside it, the slot's own share is the nested set's
        if (proposal_distribution_set *sub = dynamic_cast<proposal_distribution_set *>(set->proposals[i])) {
          if (nest_count || sub->Tpow > 0) { std::cout << "gpu_parallel_tempering_chains: only one nested proposal set without thermal weighting has a device form" << std::endl; exit(1); }
          nest_first = (int)props.size(); nest_count = (int)sub->proposals.size(); nest_share = set->shares[i]; nest_hot = Tpow > 0 ? set->hot_shares[i] : 0;
          nest_adapt = sub->adapt_rate;
          for (size_t j = 0; j < sub->proposals.size(); j++) {
            if (!describe_member(sub->proposals[j], q, keep)) { std::cout << "gpu_parallel_tempering_chains: nested proposal member " << j << " has no device form" << std::endl; exit(1); }
            q.share = sub->shares[j]; q.hot_share = 0;
            props.push_back(q);
          }
          continue;
        }
        if (!describe_member(set->proposals[i], q, keep)) { std::cout << "gpu_parallel_tempering_chains: proposal member " << i << " has no device form" << std::endl; exit(1); }
        q.share = set->shares[i]; q.hot_share = Tpow > 0 ? set->hot_shares[i] : 0;
        props.push_back(q);
      }
    } else {
      ptg_proposal q;
      if (!describe_member(&prop, q, keep)) { std::cout << "gpu_parallel_tempering_chains: proposal has no device form" << std::endl; exit(1); }
      q.share = 1; props.push_back(q);
      // temperature mixing acts only on a BARE differential_evolution: parallel_tempering_chains::set_proposal asks the top-level proposal
      // (chain.cc:1375), and a proposal_distribution_set answers support_mixing() = false whatever its members say
      if (differential_evolution *de = dynamic_cast<differential_evolution *>(&prop))
        if (de->do_support_mixing && nt > 1) { de_mixing = 1; de_Tmix = de->temperature_mixing_factor; }
    }
    check(ptg_set_proposals(h, (int32_t)props.size(), props.data(), Tpow, wrap), "set_proposals");
    if (adapt_rate != 0 || de_mixing) check(ptg_set_proposal_options(h, adapt_rate, de_mixing, de_Tmix), "set_proposal_options");
    if (nest_count) check(ptg_set_nested_set(h, nest_first, nest_count, nest_share, nest_hot, nest_adapt), "set_nested_set");
  }

  /// reboot of stuck hot replicas (do_reboot, chain.hh:308; chain.cc:1689-1759) swaps whole MH_chain objects -- histories, generators and
  /// counters -- between rungs; the engine's per-rung rings have no device form of that, so a run that asks for it stops here
  void refuse_reboot() const {
    if (max_reboot_rate > 0) { std::cout << "gpu_parallel_tempering_chains: reboot of hot replicas (pt_reboot_rate > 0) has no device form" << std::endl; exit(1); }
  }

public:
  /// same arguments as parallel_tempering_chains (chain.cc:1163-1211) + the engine's ring capacity per rung and the mirror cadence
  gpu_parallel_tempering_chains(int Ntemps, double Tmax, double swap_rate = 0.01, int add_every_N = 1, bool do_evid = false, bool verbose_evid = true,
                                double dpriormin = -30, int hist_capacity = 0, int sync_every = 1)
      : parallel_tempering_chains(Ntemps, Tmax, swap_rate, add_every_N, do_evid, verbose_evid, dpriormin), h(nullptr), nt(Ntemps), ninit(0),
        sync_every_(sync_every < 1 ? 1 : sync_every), nsteps(0), engine_inited(false), space(nullptr), the_llike(nullptr), the_prior(nullptr) {
    cfg = ptg_config();
    cfg.abi_version = PTG_ABI_VERSION; cfg.device = 0; cfg.n_ladders = 1; cfg.n_rungs = Ntemps; cfg.save_every = add_every_N;
    cfg.hist_capacity = hist_capacity; cfg.swap_mode = PTG_SWAP_REFERENCE; cfg.rng_mode = PTG_RNG_PHILOX; cfg.record_level = PTG_RECORD_FULL;
    cfg.swap_rate = swap_rate; cfg.Tmax = Tmax; cfg.dprior_min = dpriormin; cfg.evolve_rate = 0; cfg.evolve_lpost_cut = -1;
    // the engine's Philox key comes from the reference's master generator, where a chain takes its seed (chain.hh:58-59)
    cfg.seed = (uint64_t)(ProbabilityDist::getPRNG()->Next() * 18446744073709551615.0);
  }
  ~gpu_parallel_tempering_chains() { if (h) ptg_destroy(h); }
  bool evolve_temps(double rate = 0.01, double lpost_cut = -1) {
    cfg.evolve_rate = rate; cfg.evolve_lpost_cut = lpost_cut;
    return parallel_tempering_chains::evolve_temps(rate, lpost_cut);
  }
  void initialize(probability_function *log_likelihood, const sampleable_probability_function *log_prior, int n = 1, std::string initialization_file = "") {
    // the reference builds its rungs and draws their start-up samples (chain.cc:1281-1365, 846-876)
    parallel_tempering_chains::initialize(log_likelihood, log_prior, n, initialization_file);
    the_llike = log_likelihood; the_prior = log_prior; space = log_prior->get_space(); ninit = n;
    cfg.dim = space->size(); cfg.n_init = n;
    if (cfg.hist_capacity > 0 && cfg.hist_capacity < n) cfg.hist_capacity = n;
    check(ptg_create(&cfg, &h), "create");
    push_space(); push_prior(); push_likelihood();
  }
  void set_proposal(proposal_distribution &proposal) override {
    parallel_tempering_chains::set_proposal(proposal); // per-rung clones: report_prop / show keep working
    push_proposal(proposal);
    // the engine starts from the samples the reference drew: x[rung][k][dim]
    const int d = cfg.dim;
    std::vector<double> x((size_t)nt * ninit * d);
    for (int r = 0; r < nt; r++)
      for (int k = 0; k < ninit; k++)
        for (int j = 0; j < d; j++) x[((size_t)r * ninit + k) * d + j] = chains[r].states[k].get_param(j);
    check(ptg_init_states(h, x.data()), "init_states");
    mirrored.assign(nt, ninit);
    engine_inited = true;
  }
  /// one PT iteration on the device (replaces chain.cc:1393-1761)
  void step() override {
    refuse_reboot();
    check(ptg_step(h, 1), "step");
    nsteps++;
    // the run loop dumps after step k * Nevery + 1 (ptmcmc.cc:599-601: `cc->step(); if (0 == istep % Nevery) dump`): mirror on that cadence
    if ((nsteps - 1) % sync_every_ == 0) sync();
  }
  /// services that read the base class's chains at other times mirror first
  void checkpoint(std::string path) override { sync(); parallel_tempering_chains::checkpoint(path); }
  std::string status() override { sync(); return parallel_tempering_chains::status(); }
  /// n iterations in one launch (what a run loop aware of the engine calls between dumps)
  void step_many(long long n) {
    refuse_reboot();
    check(ptg_step(h, n), "step");
    nsteps += n;
    sync();
  }
  /// append the device histories' new records to the base class's chains and refresh its counters: after this every reference routine
  /// (dumpChain, status, report_effective_samples, checkpoint) sees exactly the state a CPU run would have built
  void sync() {
    const int d = cfg.dim;
    std::vector<int64_t> nhist(nt), nsize(nt), ntries(nt), naccept(nt); std::vector<int32_t> last_type(nt); std::vector<double> maplp(nt);
    check(ptg_get_counters(h, nhist.data(), nsize.data(), ntries.data(), naccept.data(), last_type.data(), maplp.data()), "get_counters");
    std::vector<double> cx((size_t)nt * d), clp(nt), cll(nt), cb(nt);
    check(ptg_get_current(h, cx.data(), clp.data(), cll.data(), cb.data()), "get_current");
    for (int r = 0; r < nt; r++) {
      MH_chain &c = chains[r];
      const long long first = mirrored[r], count = nsize[r] - first;
      if (count > 0) {
        std::vector<double> x((size_t)count * d), lp(count), ll(count), acc(count), beta(count); std::vector<int32_t> type(count);
        check(ptg_get_history(h, 0, r, first, count, x.data(), lp.data(), ll.data(), acc.data(), beta.data(), type.data()), "get_history");
        for (long long k = 0; k < count; k++) {
          std::valarray<double> v(d);
          for (int j = 0; j < d; j++) v[j] = x[(size_t)k * d + j];
          c.states.push_back(state(space, v));
          c.lposts.push_back(lp[k]); c.llikes.push_back(ll[k]); c.acceptance_ratio.push_back(acc[k]); c.invtemps.push_back(beta[k]); c.types.push_back(type[k]);
        }
        mirrored[r] = nsize[r];
      }
      c.Nsize = (int)nsize[r]; c.Nhist = (int)nhist[r]; c.Ntries = (int)ntries[r]; c.Naccept = (int)naccept[r]; c.last_type = last_type[r];
      std::valarray<double> v(d);
      for (int j = 0; j < d; j++) v[j] = cx[(size_t)r * d + j];
      c.current_state = state(space, v); c.current_lpost = clp[r]; c.current_llike = cll[r]; c.invtemp = cb[r];
      if (maplp[r] > c.MAPlpost) c.MAPlpost = maplp[r];
      temps[r] = 1 / cb[r];
    }
    std::vector<int64_t> sc(nt > 1 ? nt - 1 : 1), sa(nt > 1 ? nt - 1 : 1); std::vector<int32_t> di(nt), up(nt), dn(nt), in(nt);
    check(ptg_get_swap_stats(h, sc.data(), sa.data(), di.data(), up.data(), dn.data(), in.data()), "get_swap_stats");
    for (int r = 0; r < nt; r++) { directions[r] = di[r]; ups[r] = up[r]; downs[r] = dn[r]; instances[r] = in[r]; }
    for (int r = 0; r + 1 < nt; r++) { swap_count[r] = (int)sc[r]; swap_accept_count[r] = (int)sa[r]; }
    Nsize = chains[0].Nsize;
  }
  ptg_handle *engine() { return h; }
};
