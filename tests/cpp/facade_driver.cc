// facade_driver.cc -- a driver written against include/ptmcmc_b200.hh the way testMH.cpp / example.cc are written against
// the reference: state space, prior, likelihood, proposal mix, parallel_tempering_chains, step loop, dumpChain.
//   facade_driver <model: sines|gauss|hostgauss> <dim> <Ntemps> <nsteps> <n_ladders> <outfile> [evolve_rate]
// Writes the cold chain of ladder 0 (and of the last ladder, to <outfile>.last) in the reference's chain-file format.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include "ptmcmc_b200.hh"
using namespace ptg;
using namespace std;

// a likelihood written the reference's way: a class with evaluate_log(state&) on the host (example.cc:74-143)
class host_gaussian : public host_probability_function {
  valarray<double> x0; double lnnormfac, twosigmasq;
public:
  host_gaussian(const stateSpace *sp, const valarray<double> &x0, double sigma) : host_probability_function(sp), x0(x0) {
    twosigmasq = 2 * sigma * sigma;
    lnnormfac = -0.5 * (double)x0.size() * std::log(M_PI * twosigmasq);
  }
  double evaluate_log(state &s) {
    valarray<double> p = s.get_params();
    double r2 = 0;
    for (size_t i = 0; i < x0.size(); i++) { double dx = p[i] - x0[i]; r2 += dx * dx; }
    return lnnormfac - r2 / twosigmasq;
  }
};

int main(int argc, char **argv) {
  if (argc < 7) { fprintf(stderr, "usage: %s model dim Ntemps nsteps n_ladders outfile [evolve_rate]\n", argv[0]); return 2; }
  const string model = argv[1];
  const int dim = atoi(argv[2]), Ntemps = atoi(argv[3]), nsteps = atoi(argv[4]), nladders = atoi(argv[5]);
  const string out = argv[6];
  const double evolve = argc > 7 ? atof(argv[7]) : 0.0;

  stateSpace space(dim);
  vector<string> names(dim);
  for (int i = 0; i < dim; i++) { names[i] = "p" + to_string(i); }
  space.set_names(names);
  valarray<double> lo(dim), hi(dim), centers(dim);
  probability_function *like = nullptr;
  if (model == "sines") { // testMH.cpp:15-21,65,74-85 in d dimensions
    lo = 0.0; hi = 1.0;
    valarray<int> ks(2, dim);
    like = new sines(&space, 64.0, ks, lo, hi, log(2.0));
  } else { // example.cc:88-104: uniform prior box, isotropic Gaussian likelihood, bounds left open
    for (int i = 0; i < dim; i++) { centers[i] = 2.0 - 5.0 * (i % 2); lo[i] = centers[i] - 2 - i; hi[i] = centers[i] + 2 + i; }
    if (model == "hostgauss") like = new host_gaussian(&space, centers, 0.5);   // same arithmetic, evaluated by the caller on the host
    else like = new gaussian_likelihood(&space, centers, 0.5);
  }
  uniform_dist_product prior(&space, lo, hi);
  default_proposal_mix mix(prior);

  gpu_parallel_tempering_chains ptc(Ntemps, 1e9, 0.1, 1, false, false, -30);
  ptc.set_ladders(nladders);
  ptc.set_history_capacity(50 * dim + 2 * nsteps + 8);
  if (evolve > 0) ptc.evolve_temps(evolve);
  ptc.initialize(like, &prior, 50 * dim); // Ninit = de_ni * Npar (ptmcmc.cc:86)
  ptc.set_proposal(*mix.set);
  for (int i = 0; i < nsteps / 2; i++) ptc.step();   // the reference's outer loop (ptmcmc.cc:563-599) ...
  ptc.step(nsteps - nsteps / 2);                     // ... and the batched form
  {
    ofstream os(out.c_str()); os.precision(17);
    ptc.dumpChain(0, os, -50 * dim, 1, 0);
  }
  {
    ofstream os((out + ".last").c_str()); os.precision(17);
    ptc.dumpChain(0, os, 0, 1, nladders - 1);
  }
  state s = ptc.getState();
  printf("facade_driver: size=%d step=%d total=%lld lpost=%.17g llike=%.17g invtemp_hot=%.17g x0=%.17g first=%.17g\n", ptc.size(), ptc.getStep(),
         ptc.total_steps(), ptc.getLogPost(), ptc.getLogLike(), ptc.invTemp(Ntemps - 1), s.get_param(0), ptc.getState(0).get_param(0));
  if (nsteps >= 6000) { // the run loop's effective-sample report (ptmcmc.cc:645)
    std::pair<double, int> el = ptc.report_effective_samples(-1, 1000, 1, 0, 0);
    printf("facade_ess: ess=%.17g length=%d\n", el.first, el.second);
  }
  delete like;
  return 0;
}
