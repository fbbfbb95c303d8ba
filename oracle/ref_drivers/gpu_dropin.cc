// gpu_dropin: the drop-in claim compiled and run for real (TEST INFRASTRUCTURE, built by oracle/build_ref.sh against the untouched
// reference sources under /root/reference; nothing from the reference is copied).
//
// A driver in the style of example.cc: a bayes_likelihood (the 3-D Gaussian of example.cc:76-143 restated against the reference API),
// ptmcmc_sampler with its own option parsing, select_proposal(), initialize(), and the UNMODIFIED run loop ptmcmc_sampler::run
// (ptmcmc.cc:563-661: step, dumpChain, status, report_prop, report_effective_samples, checkpoint).  The only addition is the branch
// INTEGRATION.md describes: gpu_ptmcmc_sampler::initialize creates a gpu_parallel_tempering_chains (include/gpu_parallel_tempering_chains.hh,
// a parallel_tempering_chains whose step() is the ptg engine) instead of a parallel_tempering_chains when PTMCMC_GPU is set.
// In-tree that branch lives in ptmcmc_sampler::initialize itself (ptmcmc.cc:505-522); here it is an override so that the reference
// translation units stay byte-for-byte the reference's.  Compiled with -fno-access-control (the in-tree form is a friend declaration).
//
//   gpu_dropin --seed=0.5 --pt=8 --nsteps=20000 --nevery=5000 --outname=run      (reference on the CPU)
//   PTMCMC_GPU=1 gpu_dropin ... same flags ...                                      (same run loop, chains stepped by the engine)
#include <cstdlib>
#include <cmath>
#include <iostream>
#include <sstream>
#include <string>
#include <valarray>
#include "omp.h"
#include "options.hh"
#include "bayesian.hh"
#include "proposal_distribution.hh"
#include "ptmcmc.hh"
#include "gpu_parallel_tempering_chains.hh"

using namespace std;
shared_ptr<Random> globalRNG; // declared extern at probability_function.hh:22

/// 3-D Gaussian likelihood with a uniform box prior, against the reference API (the model of example.cc:76-143); it names its device functor
class dropin_gaussian_likelihood : public bayes_likelihood, public ptg_device_likelihood {
  int idx[3];
  double x0[3];
  double lnnormfac, twosigmasq;
public:
  double Ztheor;
  dropin_gaussian_likelihood() : bayes_likelihood(nullptr, nullptr, nullptr) {}
  virtual void setup() {
    haveSetup();
    const int npar = 3;
    stateSpace space(npar);
    string names[] = {"p0", "p1", "p2"};
    space.set_names(names);
    nativeSpace = space;
    defWorkingStateSpace(nativeSpace);
    best = state(&space, space.size());
    const int uni = mixed_dist_product::uniform;
    valarray<double> centers((initializer_list<double>){2.0, -3.0, 5.0});
    valarray<double> halfwidths((initializer_list<double>){2.0, 3.0, 5.0});
    valarray<int> types((initializer_list<int>){uni, uni, uni});
    setPrior(new mixed_dist_product(&nativeSpace, types, centers, halfwidths));
    for (int i = 0; i < 3; i++) x0[i] = centers[i];
    const double sigma = 0.5;
    twosigmasq = 2 * sigma * sigma;
    lnnormfac = -1.5 * std::log(M_PI * twosigmasq);
    Ztheor = -std::log(8 * halfwidths[0] * halfwidths[1] * halfwidths[2]);
  }
  void defWorkingStateSpace(const stateSpace &sp) {
    checkSetup();
    idx[0] = sp.requireIndex("p0"); idx[1] = sp.requireIndex("p1"); idx[2] = sp.requireIndex("p2");
    haveWorkingStateSpace();
  }
  int size() const { return 0; }
  double evaluate_log(state &s) {
    valarray<double> params = s.get_params();
    double r2 = 0;
    for (int i = 0; i < 3; i++) { double dx = params[idx[i]] - x0[i]; r2 += dx * dx; }
    double result = lnnormfac - r2 / twosigmasq;
    double post = result + nativePrior->evaluate_log(s);
#pragma omp critical
    {
      if (post > best_post) { best_post = post; best = state(s); }
      if (!isfinite(result)) result = -INFINITY;
    }
    return result;
  }
  // PTG_LIKE_GAUSS_ISO: params = lnnormfac, twosigmasq, x0[dim]
  bool describe_device_likelihood(int &kind, vector<double> &params, vector<double> &data) const {
    kind = PTG_LIKE_GAUSS_ISO;
    params.assign({lnnormfac, twosigmasq, x0[0], x0[1], x0[2]});
    data.clear();
    return true;
  }
};

/// ptmcmc_sampler with the one extra branch of INTEGRATION.md in initialize(); run(), options, checkpointing: the reference's own
class gpu_ptmcmc_sampler : public ptmcmc_sampler {
public:
  bool use_gpu; int gpu_hist_capacity;
  gpu_ptmcmc_sampler() : ptmcmc_sampler(), use_gpu(false), gpu_hist_capacity(0) {}
  bayes_sampler *clone() { // ptmcmc_sampler::clone (ptmcmc.hh:66-84) for this type
    if (have_cc) { cout << "gpu_ptmcmc_sampler::clone(): Cannot clone after instantiating chain/prop." << endl; exit(1); }
    gpu_ptmcmc_sampler *s = new gpu_ptmcmc_sampler();
    s->use_gpu = use_gpu; s->gpu_hist_capacity = gpu_hist_capacity;
    s->copyOptioned(*this);
    if (have_setup) s->setup(*chain_llike, *chain_prior, output_precision);
    if (have_cprop) { s->cprop = cprop->clone(); s->have_cprop = true; s->chain_Ninit = chain_Ninit; }
    return s;
  }
  int initialize() {
    if (!use_gpu || !parallel_tempering) return ptmcmc_sampler::initialize();
    // ptmcmc.cc:497-528 with the parallel-tempering branch creating the engine-backed chain set
    if (!have_setup or !have_cprop) { cout << "ptmcmc_sampler::initialize.  Must call setup() and set proposal before initialization!" << endl; exit(1); }
    int Ninit = chain_Ninit;
    if (restarting or Nstep <= 0) Ninit = 0;
    const int cap = gpu_hist_capacity > 0 ? gpu_hist_capacity : Ninit + 2 * (Nstep / save_every) + 1024;
    gpu_parallel_tempering_chains *ptc =
        new gpu_parallel_tempering_chains(Nptc, Tmax, swap_rate, save_every, pt_stop_evid_err > 0, pt_stop_evid_err > 0, dpriormin, cap, /*sync_every=*/Nevery);
    cc = ptc;
    have_cc = true;
    if (pt_evolve_rate > 0) ptc->evolve_temps(pt_evolve_rate, pt_evolve_lpost_cut);
    ptc->initialize(chain_llike, chain_prior, Ninit, initialization_file);
    cc->set_proposal(*cprop);
    cprop->set_chain(cc);
    return 0;
  }
};

int main(int argc, char *argv[]) {
  ptmcmc_sampler::Init();
  Options opt(true);
  gpu_ptmcmc_sampler mcmc;
  mcmc.use_gpu = getenv("PTMCMC_GPU") != nullptr;
  bayes_sampler *s0 = &mcmc;
  bayes_likelihood *like = new dropin_gaussian_likelihood();
  s0->addOptions(opt);
  like->addOptions(opt);
  opt.add(Option("seed", "Pseudo random number generator seed in [0,1). (Default=-1, use clock to seed.)", "-1"));
  opt.add(Option("precision", "Set output precision digits. (Default 13).", "13"));
  opt.add(Option("outname", "Base name for output files (Default 'mcmc_output').", "mcmc_output"));
  if (opt.parse(argc, argv)) { cout << "Usage:\n gpu_dropin [-options=vals]\n" << opt.print_usage() << endl; return 1; }
  like->setup();
  double seed; int output_precision; string outname;
  istringstream(opt.value("seed")) >> seed;
  if (seed < 0) seed = 0.5;
  istringstream(opt.value("precision")) >> output_precision;
  istringstream(opt.value("outname")) >> outname;
  cout.precision(output_precision);
  ProbabilityDist::setSeed(seed);
  globalRNG.reset(ProbabilityDist::getPRNG());
  shared_ptr<const sampleable_probability_function> prior = like->getObjectPrior();
  mcmc.setup(*like, *prior, output_precision);
  mcmc.select_proposal();
  cout << "gpu_dropin: chains stepped by " << (mcmc.use_gpu ? "the ptg engine (gpu_parallel_tempering_chains)" : "the reference (parallel_tempering_chains)") << endl;
  bayes_sampler *s = s0->clone();
  s->initialize();
  s->run(outname, 0);
  delete s;
  cout << "best_post " << like->bestPost() << ", state=" << like->bestState().get_string() << endl;
  delete like;
  return 0;
}
