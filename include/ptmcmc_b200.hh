// ptmcmc_b200.hh -- C++ host facade over the C ABI (ptmcmc_b200.h), mirroring the reference's class interface for the
// chain-stepping path: same class names, constructor arguments and accessor semantics as JohnGBaker/ptmcmc, so that a
// driver written against the reference (example.cc, testMH.cpp) reads the same against this header.  Header-only C++11,
// links against libptmcmc_b200.so only; it does NOT include or need the reference's headers (namespace ptg keeps the
// names apart when both are present -- INTEGRATION.md shows the in-tree variant deriving from the reference's `chain`).
//
//   boundary / stateSpace / state                     states.hh:29-234
//   uniform_ / gaussian_ / mixed_dist_product         probability_function.hh:88-180
//   gaussian_prop / differential_evolution / draw_from_dist / proposal_distribution_set   proposal_distribution.hh:119-414
//   device likelihoods (sines, gaussian, chi-squared fits, full-covariance Gaussian)      bayesian.hh:544-622 hook
//   gpu_parallel_tempering_chains                      chain.hh:34-330 (chain, MH_chain, parallel_tempering_chains)
//
// Error behaviour follows the reference: a message on stdout and exit(1) (e.g. chain.cc:967-971); install another
// handler with ptg::set_error_handler for embedding / tests.
#ifndef PTMCMC_B200_HH
#define PTMCMC_B200_HH
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include <utility>
#include <valarray>
#include <vector>
extern "C" {
#include "ptmcmc_b200.h"
}

namespace ptg {

typedef void (*error_handler_t)(const std::string &);
inline void default_error_handler(const std::string &msg) { std::cout << msg << std::endl; exit(1); }
inline error_handler_t &error_handler() { static error_handler_t h = default_error_handler; return h; }
inline void set_error_handler(error_handler_t h) { error_handler() = h; }
inline void check(int rc, const char *where) {
  if (rc != 0) error_handler()(std::string(where) + ": " + ptg_last_error());
}

// ------------------------------------------------------------------------------------------------ states.hh
class boundary {
public:
  enum { open = PTG_BOUND_OPEN, limit = PTG_BOUND_LIMIT, reflect = PTG_BOUND_REFLECT, wrap = PTG_BOUND_WRAP };
  int lowertype, uppertype;
  double xmin, xmax;
  boundary(int lowertype = open, int uppertype = open, double min = -INFINITY, double max = INFINITY)
      : lowertype(lowertype), uppertype(uppertype), xmin(min), xmax(max) {}
};

class stateSpace {
  int dim;
  std::vector<boundary> bounds;
  std::vector<std::string> names;
public:
  explicit stateSpace(int dim = 0) : dim(dim), bounds(dim), names(dim) {
    for (int i = 0; i < dim; i++) { std::ostringstream s; s << "param(" << i << ")"; names[i] = s.str(); }
  }
  int size() const { return dim; }
  void set_bound(int i, const boundary &b) { if (i < 0 || i >= dim) error_handler()("stateSpace::set_bound: Index out of range."); else bounds[i] = b; }
  boundary get_bound(int i) const { return bounds[i]; }
  void set_names(const std::vector<std::string> &n) { names = n; names.resize(dim); }
  std::string get_name(int i) const { return names[i]; }
  int get_index(const std::string &name) const { for (int i = 0; i < dim; i++) if (names[i] == name) return i; return -1; }
};

class state {
  const stateSpace *space;
  std::valarray<double> params;
  bool valid;
public:
  state(const stateSpace *space = nullptr, const std::valarray<double> &p = std::valarray<double>()) : space(space), params(p), valid(space != nullptr) {}
  int size() const { return (int)params.size(); }
  const stateSpace *getSpace() const { return space; }
  double get_param(int i) const { return params[i]; }
  std::vector<double> get_params_vector() const { return std::vector<double>(std::begin(params), std::end(params)); }
  std::valarray<double> get_params() const { return params; }
  bool invalid() const { return !valid; }
  std::string get_string() const { std::ostringstream s; s.precision(15); for (int i = 0; i < size(); i++) s << (i ? ", " : "") << params[i]; return s.str(); }
};

// ------------------------------------------------------------------------------------------------ probability_function.hh
class sampleable_probability_function {
protected:
  const stateSpace *space;
  std::vector<int32_t> types;
  std::vector<double> a, b; // uniform-like: (min,max); gaussian: (x0,sigma)
public:
  explicit sampleable_probability_function(const stateSpace *space) : space(space) {}
  virtual ~sampleable_probability_function() {}
  const stateSpace *get_space() const { return space; }
  int getDim() const { return (int)types.size(); }
  // uniform half-width or Gaussian sigma (probability_function.hh:110,132-135,170-179)
  void getScales(std::valarray<double> &scales) const {
    scales.resize(types.size());
    for (size_t i = 0; i < types.size(); i++) scales[i] = types[i] == PTG_PRIOR_GAUSSIAN ? b[i] : (b[i] - a[i]) / 2.0;
  }
  void push(ptg_handle *h) const { check(ptg_set_prior(h, types.data(), a.data(), b.data()), "ptg_set_prior"); }
};
class uniform_dist_product : public sampleable_probability_function {
public:
  uniform_dist_product(const stateSpace *space, const std::valarray<double> &min, const std::valarray<double> &max) : sampleable_probability_function(space) {
    for (size_t i = 0; i < min.size(); i++) { types.push_back(PTG_PRIOR_UNIFORM); a.push_back(min[i]); b.push_back(max[i]); }
  }
};
class gaussian_dist_product : public sampleable_probability_function {
public:
  gaussian_dist_product(const stateSpace *space, const std::valarray<double> &x0s, const std::valarray<double> &sigmas) : sampleable_probability_function(space) {
    for (size_t i = 0; i < x0s.size(); i++) { types.push_back(PTG_PRIOR_GAUSSIAN); a.push_back(x0s[i]); b.push_back(sigmas[i]); }
  }
};
class mixed_dist_product : public sampleable_probability_function {
public:
  enum { uniform = PTG_PRIOR_UNIFORM, gaussian = PTG_PRIOR_GAUSSIAN, polar = PTG_PRIOR_POLAR, copolar = PTG_PRIOR_COPOLAR, log = PTG_PRIOR_LOG };
  // centers / halfwidths as in the reference (probability_function.cc:235-249): gaussian (x0, sigma), others centre +- halfwidth
  mixed_dist_product(const stateSpace *space, const std::valarray<int> &t, const std::valarray<double> &centers, const std::valarray<double> &halfwidths)
      : sampleable_probability_function(space) {
    for (size_t i = 0; i < t.size(); i++) {
      types.push_back(t[i]);
      if (t[i] == gaussian) { a.push_back(centers[i]); b.push_back(halfwidths[i]); }
      else { a.push_back(centers[i] - halfwidths[i]); b.push_back(centers[i] + halfwidths[i]); }
    }
  }
};

// ------------------------------------------------------------------------------------------------ device likelihoods
// What bayes_likelihood::register_evaluate_log (bayesian.hh:544-552) is to a host likelihood: states which device functor
// evaluates log L and with which parameters / data.
class probability_function {
protected:
  const stateSpace *space;
  int32_t kind;
  std::vector<double> params, data;
public:
  probability_function(const stateSpace *space, int32_t kind) : space(space), kind(kind) {}
  virtual ~probability_function() {}
  virtual void push(ptg_handle *h) {
    check(ptg_set_likelihood(h, kind, params.data(), (int32_t)params.size(), data.empty() ? nullptr : data.data(), (int64_t)data.size()), "ptg_set_likelihood");
  }
};
// A likelihood that exists only as host code: derive and override evaluate_log(state&) exactly as with the reference's
// probability_function / bayes_likelihood (probability_function.hh:31-45, bayesian.hh:553-581).  The engine calls it for all
// chains' proposals of a PT iteration in one batch (ptg_register_evaluate_log); everything else of the step stays on the GPU.
class host_probability_function : public probability_function {
  static void trampoline(void *user, const double *x, int64_t n, double *out) {
    host_probability_function *self = static_cast<host_probability_function *>(user);
    const int d = self->space->size();
    std::valarray<double> p(d);
    for (int64_t i = 0; i < n; i++) {
      for (int j = 0; j < d; j++) p[j] = x[i * d + j];
      state s(self->space, p);
      out[i] = self->evaluate_log(s);
    }
  }
public:
  explicit host_probability_function(const stateSpace *space) : probability_function(space, PTG_LIKE_HOST_CALLBACK) {}
  virtual double evaluate_log(state &s) = 0;
  void push(ptg_handle *h) { check(ptg_register_evaluate_log(h, trampoline, this), "ptg_register_evaluate_log"); }
};
// sines.hh:12-61
class sines : public probability_function {
public:
  sines(const stateSpace *sp, double height, const std::valarray<int> &ks, const std::valarray<double> &mins, const std::valarray<double> &maxs, double step_scale = 0)
      : probability_function(sp, PTG_LIKE_SINES) {
    params.push_back(height); params.push_back(step_scale);
    for (size_t i = 0; i < ks.size(); i++) params.push_back(ks[i]);
    for (size_t i = 0; i < mins.size(); i++) params.push_back(mins[i]);
    for (size_t i = 0; i < maxs.size(); i++) params.push_back(maxs[i]);
  }
};
// example.cc:74-143: isotropic Gaussian lnnormfac - r^2/(2 sigma^2)
class gaussian_likelihood : public probability_function {
public:
  gaussian_likelihood(const stateSpace *sp, const std::valarray<double> &x0, double sigma) : probability_function(sp, PTG_LIKE_GAUSS_ISO) {
    const double twosigmasq = 2 * sigma * sigma;
    params.push_back(-0.5 * (double)x0.size() * std::log(M_PI * twosigmasq)); params.push_back(twosigmasq);
    for (size_t i = 0; i < x0.size(); i++) params.push_back(x0[i]);
  }
};
// bayes_likelihood::log_chi_squared (bayesian.hh:595-622) with the polynomial model of poly_example.cc:85-106 (poly=true)
// or a sum of sinusoids A sin(2 pi f t + phi) with state (A,f,phi) triples (poly=false)
class chi_squared_likelihood : public probability_function {
public:
  chi_squared_likelihood(const stateSpace *sp, bool poly, const std::vector<double> &x, const std::vector<double> &y, const std::vector<double> &dy, double like0 = 0)
      : probability_function(sp, poly ? PTG_LIKE_POLY_CHI2 : PTG_LIKE_SINUSOID_CHI2) {
    params.push_back(like0);
    data = x; data.insert(data.end(), y.begin(), y.end());
    for (size_t i = 0; i < dy.size(); i++) data.push_back(dy[i] * dy[i]);
  }
};
// cython/exampleGaussian.py:103-109: like0 - x^T Cinv x / 2
class fullcov_gaussian_likelihood : public probability_function {
public:
  fullcov_gaussian_likelihood(const stateSpace *sp, const std::vector<double> &cinv_rowmajor, double like0) : probability_function(sp, PTG_LIKE_GAUSS_FULLCOV) {
    params.push_back(like0); data = cinv_rowmajor;
  }
};

// ------------------------------------------------------------------------------------------------ proposal_distribution.hh
class proposal_distribution {
public:
  virtual ~proposal_distribution() {}
  virtual void collect(std::vector<ptg_proposal> &out, double share, double hot_share) const = 0;
  virtual bool is_set() const { return false; }
  virtual double Tpow() const { return 0; }
};
class gaussian_prop : public proposal_distribution {
  std::vector<double> sigmas, transform;
  double oneDfrac;
public:
  gaussian_prop(const std::valarray<double> &sig, double oneDfrac = 0.0, bool scaleWithTemp = false) : sigmas(std::begin(sig), std::end(sig)), oneDfrac(oneDfrac) { (void)scaleWithTemp; }
  // covariance constructor (proposal_distribution.hh:164-187): eigen-decomposition on the host (cyclic Jacobi), ascending
  // eigenvalues like Eigen::SelfAdjointEigenSolver; sigmas = sqrt(eigenvalues), offset = V (z o sigma)
  gaussian_prop(const std::vector<double> &covar_rowmajor, int n, double oneDfrac = 0.0, bool scaleWithTemp = false) : oneDfrac(oneDfrac) {
    (void)scaleWithTemp;
    std::vector<double> A(covar_rowmajor), V((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) V[(size_t)i * n + i] = 1;
    for (int sweep = 0; sweep < 100; sweep++) {
      double off = 0;
      for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += A[(size_t)p * n + q] * A[(size_t)p * n + q];
      if (off < 1e-300) break;
      for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
        const double apq = A[(size_t)p * n + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1)), c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < n; k++) { const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q]; A[(size_t)k * n + p] = c * akp - s * akq; A[(size_t)k * n + q] = s * akp + c * akq; }
        for (int k = 0; k < n; k++) { const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k]; A[(size_t)p * n + k] = c * apk - s * aqk; A[(size_t)q * n + k] = s * apk + c * aqk; }
        for (int k = 0; k < n; k++) { const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q]; V[(size_t)k * n + p] = c * vkp - s * vkq; V[(size_t)k * n + q] = s * vkp + c * vkq; }
      }
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) if (A[(size_t)order[j] * n + order[j]] < A[(size_t)order[i] * n + order[i]]) std::swap(order[i], order[j]);
    sigmas.resize(n); transform.resize((size_t)n * n);
    for (int j = 0; j < n; j++) {
      sigmas[j] = std::sqrt(A[(size_t)order[j] * n + order[j]]);
      for (int i = 0; i < n; i++) transform[(size_t)i * n + j] = V[(size_t)i * n + order[j]];
    }
  }
  void collect(std::vector<ptg_proposal> &out, double share, double hot_share) const {
    ptg_proposal p = ptg_proposal();
    p.kind = PTG_PROP_GAUSS; p.share = share; p.hot_share = hot_share; p.one_d_frac = oneDfrac;
    p.sigmas = sigmas.data(); p.transform = transform.empty() ? nullptr : transform.data();
    out.push_back(p);
  }
};
class differential_evolution : public proposal_distribution {
  double snooker, gamma_one_frac, b_small, ignore_frac, unlikely_alpha, reduce_gamma_fac;
public:
  differential_evolution(double snooker = 0.0, double gamma_one_frac = 0.1, double b_small = 0.0001, double ignore_frac = 0.3, double unlikely_alpha = 0)
      : snooker(snooker), gamma_one_frac(gamma_one_frac), b_small(b_small), ignore_frac(ignore_frac), unlikely_alpha(unlikely_alpha), reduce_gamma_fac(1) {}
  void reduce_gamma(double factor) { reduce_gamma_fac = factor; }
  void collect(std::vector<ptg_proposal> &out, double share, double hot_share) const {
    ptg_proposal p = ptg_proposal();
    p.kind = PTG_PROP_DE; p.share = share; p.hot_share = hot_share; p.snooker = snooker; p.gamma_one_frac = gamma_one_frac; p.b_small = b_small;
    p.ignore_frac = ignore_frac; p.unlikely_alpha = unlikely_alpha; p.reduce_gamma = reduce_gamma_fac;
    out.push_back(p);
  }
};
class draw_from_dist : public proposal_distribution {
public:
  explicit draw_from_dist(const sampleable_probability_function &) {}
  void collect(std::vector<ptg_proposal> &out, double share, double hot_share) const {
    ptg_proposal p = ptg_proposal(); p.kind = PTG_PROP_PRIOR_DRAW; p.share = share; p.hot_share = hot_share; out.push_back(p);
  }
};
class proposal_distribution_set : public proposal_distribution {
  std::vector<proposal_distribution *> members;
  std::vector<double> shares, hot_shares;
  double tpow;
public:
  proposal_distribution_set(const std::vector<proposal_distribution *> &props, const std::vector<double> &shares, double adapt_rate = 0, double Tpow = 0,
                            const std::vector<double> &hot_shares = std::vector<double>())
      : members(props), shares(shares), hot_shares(hot_shares), tpow(Tpow) {
    if (adapt_rate != 0) error_handler()("proposal_distribution_set: adaptive shares are not supported by the GPU engine (host-side feature, SURVEY.md section 2)");
    this->hot_shares.resize(members.size(), 0.0);
  }
  bool is_set() const { return true; }
  double Tpow() const { return tpow; }
  void collect(std::vector<ptg_proposal> &out, double, double) const {
    for (size_t i = 0; i < members.size(); i++) {
      if (members[i]->is_set()) error_handler()("proposal_distribution_set: nested sets are not supported by the GPU engine");
      members[i]->collect(out, shares[i], hot_shares[i]);
    }
  }
};

// ------------------------------------------------------------------------------------------------ chain.hh
// parallel_tempering_chains (chain.cc:1163-1761) on the GPU.  n_ladders independent ladders step together; ladder 0 serves
// the single-chain accessors of the reference interface (getState, dumpChain, ...), the others are read with *_of().
class gpu_parallel_tempering_chains {
  ptg_handle *h;
  ptg_config cfg;
  const stateSpace *space;
  int Ntemps, dim, Ninit;
  bool have_model, inited;
  std::vector<double> hx, hlp, hll, hacc, hbeta; std::vector<int32_t> htype; // host mirror of one chain's history (lazy)
  int mirror_ladder, mirror_rung; long long mirror_step; long long nstep;
  std::vector<int64_t> nhist, nsize, ntries, naccept; std::vector<int32_t> last_type; std::vector<double> map_lpost;
  std::vector<double> cur_x, cur_lpost, cur_llike, cur_beta;
  long long counters_step;

  void sync_counters() {
    if (counters_step == nstep) return;
    const size_t n = (size_t)cfg.n_ladders * Ntemps;
    nhist.resize(n); nsize.resize(n); ntries.resize(n); naccept.resize(n); last_type.resize(n); map_lpost.resize(n);
    cur_x.resize(n * dim); cur_lpost.resize(n); cur_llike.resize(n); cur_beta.resize(n);
    check(ptg_get_counters(h, nhist.data(), nsize.data(), ntries.data(), naccept.data(), last_type.data(), map_lpost.data()), "ptg_get_counters");
    check(ptg_get_current(h, cur_x.data(), cur_lpost.data(), cur_llike.data(), cur_beta.data()), "ptg_get_current");
    counters_step = nstep;
  }
  // one D2H of a chain's whole (ring-resident) history per dump, not per call
  void sync_mirror(int ladder, int rung) {
    sync_counters();
    if (mirror_step == nstep && mirror_ladder == ladder && mirror_rung == rung) return;
    const size_t c = (size_t)ladder * Ntemps + rung;
    const int64_t n = nsize[c], first = n > cfg.hist_capacity ? n - cfg.hist_capacity : 0, cnt = n - first;
    hx.assign((size_t)n * dim, NAN); hlp.assign(n, NAN); hll.assign(n, NAN); hacc.assign(n, NAN); hbeta.assign(n, NAN); htype.assign(n, -1);
    check(ptg_get_history(h, ladder, rung, first, cnt, hx.data() + first * dim, hlp.data() + first, hll.data() + first, hacc.data() + first,
                          hbeta.data() + first, htype.data() + first), "ptg_get_history");
    mirror_step = nstep; mirror_ladder = ladder; mirror_rung = rung;
  }
  // MH_chain::get_state_idx (chain.cc:1041-1051)
  int state_idx(int i, size_t c) const {
    if (i < 0 || i >= nhist[c]) i = (int)nhist[c] - 1;
    return Ninit + i / cfg.save_every;
  }
public:
  gpu_parallel_tempering_chains(int Ntemps, double Tmax, double swap_rate = 0.01, int add_every_N = 1, bool do_evid = false, bool verbose_evid = false,
                                double dpriormin = -30)
      : h(nullptr), space(nullptr), Ntemps(Ntemps), dim(0), Ninit(0), have_model(false), inited(false), mirror_ladder(-1), mirror_rung(-1),
        mirror_step(-1), nstep(0), counters_step(-1) {
    (void)do_evid; (void)verbose_evid;
    cfg = ptg_config();
    cfg.abi_version = PTG_ABI_VERSION; cfg.device = 0; cfg.n_ladders = 1; cfg.n_rungs = Ntemps; cfg.save_every = add_every_N; cfg.hist_capacity = 0;
    cfg.swap_mode = PTG_SWAP_REFERENCE; cfg.rng_mode = PTG_RNG_PHILOX; cfg.record_level = PTG_RECORD_FULL; cfg.swap_rate = swap_rate; cfg.Tmax = Tmax;
    cfg.dprior_min = dpriormin; cfg.evolve_rate = 0; cfg.evolve_lpost_cut = -1; cfg.seed = 0xB2000003ull; cfg.ladder_offset = 0;
  }
  ~gpu_parallel_tempering_chains() { if (h) ptg_destroy(h); }
  // engine-only knobs (no reference counterpart): batch size, device, history ring, Philox key, swap schedule
  void set_ladders(int n_ladders, long long ladder_offset = 0) { cfg.n_ladders = n_ladders; cfg.ladder_offset = ladder_offset; }
  void set_device(int device) { cfg.device = device; }
  void set_history_capacity(int slots) { cfg.hist_capacity = slots; }
  void setRNGseed(uint64_t key) { cfg.seed = key; }
  void set_swap_mode(int mode) { cfg.swap_mode = mode; }
  void evolve_temps(double rate = 0.01, double lpost_cut = -1) { cfg.evolve_rate = rate; cfg.evolve_lpost_cut = lpost_cut; } // chain.hh:302

  // parallel_tempering_chains::initialize(llike, lprior, n) (chain.cc:1281-1365); the prior draws happen in set_proposal,
  // once the proposal is known (the reference's call order: initialize, then set_proposal, ptmcmc.cc:509-522)
  void initialize(probability_function *log_likelihood, const sampleable_probability_function *log_prior, int n = 1) {
    if (have_model) { error_handler()("gpu_parallel_tempering_chains::initialize: Cannot re-initialize."); return; }
    space = log_prior->get_space(); dim = space->size(); Ninit = n;
    cfg.dim = dim; cfg.n_init = n;
    if (cfg.hist_capacity <= 0) cfg.hist_capacity = n + 4096;
    check(ptg_create(&cfg, &h), "ptg_create");
    std::vector<int32_t> lo(dim), up(dim); std::vector<double> xmin(dim), xmax(dim);
    for (int i = 0; i < dim; i++) { const boundary b = space->get_bound(i); lo[i] = b.lowertype; up[i] = b.uppertype; xmin[i] = b.xmin; xmax[i] = b.xmax; }
    check(ptg_set_space(h, lo.data(), up.data(), xmin.data(), xmax.data()), "ptg_set_space");
    log_prior->push(h);
    log_likelihood->push(h);
    have_model = true;
  }
  // parallel_tempering_chains::set_proposal (chain.cc:1367-1386): one clone of the proposal per rung
  void set_proposal(proposal_distribution &proposal) {
    if (!have_model) { error_handler()("gpu_parallel_tempering_chains::set_proposal: initialize first."); return; }
    std::vector<ptg_proposal> props;
    proposal.collect(props, 1.0, 0.0);
    check(ptg_set_proposals(h, (int32_t)props.size(), props.data(), proposal.Tpow(), proposal.is_set() ? 1 : 0), "ptg_set_proposals");
    check(ptg_init_from_prior(h), "ptg_init_from_prior");
    inited = true;
  }
  void step() { step(1); } // chain.cc:1393
  void step(long long n) {
    if (!h) { error_handler()("MH_chain:step: Can't step before initializing chain (chain.cc:967-971)"); return; }
    check(ptg_step(h, n), "ptg_step"); nstep += n;
  }
  int multiplicity() const { return Ntemps; }
  int n_ladders() const { return cfg.n_ladders; }
  // ---- accessors of the cold chain of ladder 0 (chain.hh:86-124); *_of(ladder, rung, ...) for any chain
  int size() { sync_counters(); return (int)nsize[0]; }
  int getStep() { sync_counters(); return (int)nhist[0]; }
  double invTemp(int rung = 0, int ladder = 0) { sync_counters(); return cur_beta[(size_t)ladder * Ntemps + rung]; }
  double getMAPlpost(int rung = 0, int ladder = 0) { sync_counters(); return map_lpost[(size_t)ladder * Ntemps + rung]; }
  state getState(int elem = -1, bool raw_indexing = false) { return getState_of(0, 0, elem, raw_indexing); }
  double getLogPost(int elem = -1, bool raw_indexing = false) { return getLogPost_of(0, 0, elem, raw_indexing); }
  double getLogLike(int elem = -1, bool raw_indexing = false) { return getLogLike_of(0, 0, elem, raw_indexing); }
  state getState_of(int ladder, int rung, int elem = -1, bool raw_indexing = false) {
    sync_counters();
    const size_t c = (size_t)ladder * Ntemps + rung;
    std::valarray<double> p(dim);
    if (elem < 0 || (raw_indexing && elem >= nsize[c]) || (!raw_indexing && elem >= nhist[c])) { // current state by default (chain.cc:1057-1061)
      for (int i = 0; i < dim; i++) p[i] = cur_x[c * dim + i];
      return state(space, p);
    }
    sync_mirror(ladder, rung);
    const int idx = raw_indexing ? elem : state_idx(elem, c);
    for (int i = 0; i < dim; i++) p[i] = hx[(size_t)idx * dim + i];
    return state(space, p);
  }
  double getLogPost_of(int ladder, int rung, int elem = -1, bool raw_indexing = false) {
    sync_counters();
    const size_t c = (size_t)ladder * Ntemps + rung;
    if (elem < 0 || (raw_indexing && elem >= nsize[c]) || (!raw_indexing && elem >= nhist[c])) return cur_lpost[c];
    sync_mirror(ladder, rung);
    return hlp[raw_indexing ? elem : state_idx(elem, c)];
  }
  double getLogLike_of(int ladder, int rung, int elem = -1, bool raw_indexing = false) {
    sync_counters();
    const size_t c = (size_t)ladder * Ntemps + rung;
    if (elem < 0 || (raw_indexing && elem >= nsize[c]) || (!raw_indexing && elem >= nhist[c])) return cur_llike[c];
    sync_mirror(ladder, rung);
    return hll[raw_indexing ? elem : state_idx(elem, c)];
  }
  // parallel_tempering_chains::dumpChain(ichain, os, Nburn, ievery) -> MH_chain::dumpChain (chain.cc:1112-1135): same text format
  void dumpChain(int ichain, std::ostream &os, int Nburn = 0, int ievery = 1, int ladder = 0) {
    sync_mirror(ladder, ichain);
    const size_t c = (size_t)ladder * Ntemps + ichain;
    if (nsize[c] == 0) return;
    os << "#Ninit=" << Ninit << ", Nburn=" << Nburn << "\n";
    os << "#eval: log(posterior) log(likelihood) acceptance_ratio prop_type: ";
    for (int i = 0; i < dim; i++) os << space->get_name(i) << " ";
    os << std::endl;
    if (Nburn + Ninit < 0) Nburn = -Ninit;
    for (int i = Nburn; i < nhist[c]; i += ievery) {
      int idx = Ninit + i;
      if (i >= 0) idx = state_idx(i, c);
      os << i << " " << hlp[idx] << " " << hll[idx] << " " << hacc[idx] << " " << htype[idx] << ": ";
      for (int j = 0; j < dim - 1; j++) os << hx[(size_t)idx * dim + j] << " ";
      os << hx[(size_t)idx * dim + dim - 1];
      os << " " << cur_beta[c];
      os << std::endl;
    }
  }
  // swap statistics of one ladder (status text inputs, chain.cc:2053-2094)
  void swap_stats(int ladder, std::vector<int64_t> &swap_count, std::vector<int64_t> &swap_accept) {
    std::vector<int64_t> sc((size_t)cfg.n_ladders * (Ntemps > 1 ? Ntemps - 1 : 1)), sa(sc.size());
    check(ptg_get_swap_stats(h, sc.data(), sa.data(), nullptr, nullptr, nullptr, nullptr), "ptg_get_swap_stats");
    swap_count.assign(sc.begin() + (size_t)ladder * (Ntemps - 1), sc.begin() + (size_t)(ladder + 1) * (Ntemps - 1));
    swap_accept.assign(sa.begin() + (size_t)ladder * (Ntemps - 1), sa.begin() + (size_t)(ladder + 1) * (Ntemps - 1));
  }
  // chain::report_effective_samples(imax, width, every, esslimit < 0) (chain.cc:457-481, 549-643) for one chain -> (ess, useful length):
  // the windowed lag statistics come from the device in the reference's summation order (ptg_get_autocovar_windows), the combination
  // below is chain::compute_effective_samples (chain.cc:340-416).  The run loop calls it as (-1, save_every*1000, save_every), ptmcmc.cc:645.
  // The whole run must still be in the history ring.
  std::pair<double, int> report_effective_samples(int imax = -1, int width = 40000, int every = 100, int ladder = 0, int rung = 0) {
    sync_counters();
    const size_t ci = (size_t)ladder * Ntemps + rung;
    const long long istep = nhist[ci];
    const int se = cfg.save_every;
    while (width < istep * 0.05) width *= 2;
    if (imax < 0 || imax > dim) imax = dim;
    if (imax > 20) imax = 20;
    if (every < 1) every = 1;
    if (every != se) error_handler()("gpu_parallel_tempering_chains::report_effective_samples: `every` must equal add_every_N (as in the run loop)");
    const int minburn = 2, maxbins = 20;
    while ((long long)width * (maxbins + minburn) < istep) width *= 2;
    const int swidth = width / every;
    width = swidth * every;
    const int Nwin = (int)(istep / width) - minburn;
    if (Nwin < 1 || imax < 1) return std::make_pair(0.0, 0);
    std::vector<int> lags(1, 0);                                      // logarithmic lag grid, dlag = 1.1 (chain.cc:207-222)
    { double fac = 1; int idx = 1; while (idx < minburn * swidth) { lags.push_back(every * idx); const int last = idx; while (last == idx) { fac *= 1.1; idx = (int)fac; } } }
    const int Nlag = (int)lags.size();
    std::vector<int32_t> lag_rec(Nlag);
    for (int j = 0; j < Nlag; j++) lag_rec[j] = lags[j] / se;
    std::vector<int64_t> end_rec((size_t)cfg.n_ladders);
    for (int l = 0; l < cfg.n_ladders; l++) end_rec[l] = Ninit + nhist[(size_t)l * Ntemps + rung] / se;
    const size_t per = (size_t)imax * Nwin * Nlag;
    std::vector<double> means((size_t)cfg.n_ladders * per), covar(means.size());
    check(ptg_get_autocovar_windows(h, rung, swidth, Nwin, Nlag, lag_rec.data(), end_rec.data(), imax, means.data(), covar.data()), "ptg_get_autocovar_windows");
    const double *M = means.data() + (size_t)ladder * per, *C = covar.data() + (size_t)ladder * per;
    auto at = [&](const double *a, int f, int k, int j) { return a[((size_t)f * Nwin + k) * Nlag + j]; };
    double ess_max = 0; int nwin_max = 0;
    for (int nwin = 1; nwin <= Nwin; nwin++) {
      double ess = 1e100;
      for (int f = 0; f < imax; f++) {
        double sum = 0;
        for (int i = Nwin - nwin; i < Nwin; i++) sum += at(M, f, i, 0);
        const double mean = sum / nwin;
        int last_lag = 0; double ac_len = 1.0, lastcorr = 1, dacl = 0;
        for (int il = 1; il < Nlag; il++) {
          double num = 0, denom = 0;
          for (int iw = Nwin - nwin; iw < Nwin; iw++) {
            const double dmean = mean - at(M, f, iw, il), dmean0 = mean - at(M, f, iw, 0);
            num += (at(C, f, iw, il) + dmean * dmean) * swidth;
            denom += (at(C, f, iw, 0) + dmean0 * dmean0) * swidth;
          }
          const double corr = num / denom;
          if (lastcorr < 0 && corr < 0) { ac_len -= dacl; break; }   // initially-positive-sequence cut
          lastcorr = corr;
          dacl = 2.0 * (lags[il] - last_lag) * corr;
          ac_len += dacl;
          last_lag = lags[il];
        }
        double essi = nwin * (double)width / ac_len;
        if (ac_len < every) essi = nwin * (double)width / 3.0 / every;
        if (essi < ess) ess = essi;
      }
      if (ess > ess_max) { ess_max = ess; nwin_max = nwin; }
    }
    return std::make_pair(ess_max, width * nwin_max);
  }
  long long total_steps() { int64_t t = 0; check(ptg_get_total_steps(h, &t), "ptg_get_total_steps"); return t; }
  void checkpoint(const std::string &path) { check(ptg_checkpoint(h, path.c_str()), "ptg_checkpoint"); }
  void restart(const std::string &path) { check(ptg_restore(h, path.c_str()), "ptg_restore"); mirror_step = counters_step = -1; }
  ptg_handle *handle() { return h; }
};

// the default proposal mix of ptmcmc_sampler::select_proposal with all flags at their defaults (ptmcmc.cc:67-139):
// 80 % differential evolution (snooker 0.1, gamma_one_frac 0.3, reduce_gamma 4) + 20 % six-scale Gaussians (1-D fraction 0.5)
struct default_proposal_mix {
  std::vector<proposal_distribution *> members;
  std::vector<double> shares;
  proposal_distribution_set *set;
  explicit default_proposal_mix(const sampleable_probability_function &prior, double gauss_draw_frac = 0.2, double gauss_1d_frac = 0.5) : set(nullptr) {
    std::valarray<double> scales; prior.getScales(scales);
    differential_evolution *de = new differential_evolution(0.1, 0.3, 1e-4, 0.0, 0);
    de->reduce_gamma(4);
    members.push_back(de); shares.push_back(1 - gauss_draw_frac);
    const int Ng = 6;
    const double sum = std::pow(2.0, Ng + 1) - 2, stepfac = 2;
    double fac = std::pow(2.0 / stepfac, 4.0), sharefac = 1;
    for (int i = 0; i < Ng; i++) {
      fac *= stepfac;
      std::valarray<double> sig = scales / 100.0 / fac;
      members.push_back(new gaussian_prop(sig, gauss_1d_frac, false));
      sharefac *= 2; shares.push_back(sharefac / sum * gauss_draw_frac);
    }
    set = new proposal_distribution_set(members, shares, 0, 0, std::vector<double>(members.size(), 0.0));
  }
  ~default_proposal_mix() { delete set; for (size_t i = 0; i < members.size(); i++) delete members[i]; }
};

} // namespace ptg
#endif
