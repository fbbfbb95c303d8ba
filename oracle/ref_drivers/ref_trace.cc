// ref_trace: drive the UNMODIFIED reference (JohnGBaker/ptmcmc) through its own public
// API -- parallel_tempering_chains::initialize / set_proposal / step (chain.cc:1281,1367,1393)
// -- for the hot-path models of SURVEY.md section 8(d), and dump every rung's raw history
// (states, lposts, llikes, acceptance_ratio, invtemps, types) in binary.
//
// This file is OUR code (test infrastructure).  It is compiled by oracle/build_ref.sh
// against the reference sources where they lie under /root/reference; nothing from the
// reference is copied.  The likelihood classes below are small probability_function
// subclasses written against the reference API, exactly as BASELINE.md section 3 prescribes
// for the configs that have no (compiling) reference driver.
//
// It is compiled with g++ -fno-access-control ONLY so that this trace driver can read the private
// history vectors (MH_chain::types, acceptance_ratio, invtemps) and swap counters at full
// precision; it does not change any reference behaviour.
//
// usage: ref_trace key=value ...   (see parse below);  writes <out> (binary) and prints a summary.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <map>
#include <vector>
#include <valarray>
#include <iostream>
#include <sstream>
#include <fstream>
#include <memory>
#include <new>
#include <chrono>
#include "chain.hh"
#include "proposal_distribution.hh"
#include "probability_function.hh"
#include "sines.hh"

using namespace std;
shared_ptr<Random> globalRNG; // declared extern at probability_function.hh:22

// ---- likelihoods written against the reference API ------------------------------------
// d-dim isotropic Gaussian: arithmetic of example.cc:116-143 (lnnormfac - r2/twosigmasq).
class iso_gauss_like : public probability_function {
public:
  vector<double> x0; double lnnormfac, twosigmasq;
  iso_gauss_like(const stateSpace *sp, const vector<double> &x0, double sigma)
      : probability_function(sp), x0(x0) {
    twosigmasq = 2 * sigma * sigma;
    lnnormfac = -0.5 * (double)x0.size() * std::log(M_PI * twosigmasq);
  }
  double evaluate_log(state &s) {
    valarray<double> p = s.get_params();
    double r2 = 0;
    for (size_t i = 0; i < x0.size(); i++) { double dx = p[i] - x0[i]; r2 += dx * dx; }
    double result = lnnormfac - r2 / twosigmasq;
    if (!isfinite(result)) result = -INFINITY;
    return result;
  }
};

// chi-squared over data: arithmetic of bayes_likelihood::log_chi_squared (bayesian.hh:595-622)
// with the polynomial model of poly_signal::get_model_signal (poly_example.cc:85-106),
// or a sum of sinusoids  y(t)=sum_k A_k sin(2 pi f_k t + phi_k)  (SURVEY.md 8(d) config C2,
// params ordered A_0,f_0,phi_0,A_1,...).
class chi2_like : public probability_function {
public:
  int kind; // 0 poly, 1 sinusoids
  vector<double> xs, ys, S; double like0;
  chi2_like(const stateSpace *sp, int kind, const vector<double> &xs, const vector<double> &ys,
            const vector<double> &dys) : probability_function(sp), kind(kind), xs(xs), ys(ys), like0(0) {
    S.resize(dys.size());
    for (size_t i = 0; i < dys.size(); i++) S[i] = dys[i] * dys[i];
  }
  double evaluate_log(state &s) {
    valarray<double> p = s.get_params();
    int d = p.size();
    double sum = 0, nsum = 0;
    for (size_t i = 0; i < xs.size(); i++) {
      double y = 0;
      if (kind == 0) {
        double xn = 1;
        for (int j = 0; j < d; j++) { y += xn * p[j]; xn *= xs[i]; }
      } else {
        for (int k = 0; k + 2 < d; k += 3) y += p[k] * sin(2 * M_PI * p[k + 1] * xs[i] + p[k + 2]);
      }
      double dd = y - ys[i];
      sum += dd * dd / S[i];
      nsum += log(S[i]);
    }
    sum += nsum;
    sum /= -2;
    double result = sum - like0;
    if (!isfinite(result)) result = -INFINITY; // bayesian.hh:569-575
    return result;
  }
};

// full-covariance Gaussian, arithmetic of cython/exampleGaussian.py:103-109:
//   like0 - 0.5 * x^T Cinv x, evaluated as y_i = sum_j Cinv[i][j] x_j ; q = sum_i x_i y_i.
class fullcov_like : public probability_function {
public:
  int d; vector<double> cinv; double like0;
  fullcov_like(const stateSpace *sp, int d, const vector<double> &cinv, double like0)
      : probability_function(sp), d(d), cinv(cinv), like0(like0) {}
  double evaluate_log(state &s) {
    valarray<double> p = s.get_params();
    double q = 0;
    for (int i = 0; i < d; i++) {
      double y = 0;
      for (int j = 0; j < d; j++) y += cinv[(size_t)i * d + j] * p[j];
      q += p[i] * y;
    }
    double result = like0 - 0.5 * q;
    if (!isfinite(result)) result = -INFINITY;
    return result;
  }
};

// 2-D double Gaussian shell, reflected in p0: arithmetic of gaussian_shell_2D_likelihood::evaluate_log (example.cc:195-206)
class shell2d_like : public probability_function {
public:
  double x0[2], r0, lnnormfac, twosigmasq;
  shell2d_like(const stateSpace *sp, double x00, double x01, double r0, double sigma) : probability_function(sp), r0(r0) {
    x0[0] = x00; x0[1] = x01;
    twosigmasq = 2 * sigma * sigma;
    lnnormfac = -0.5 * std::log(M_PI * twosigmasq); // example.cc:179
  }
  double evaluate_log(state &s) {
    valarray<double> params = s.get_params();
    double r2 = 0;
    double dx = abs(params[0]) - x0[0];
    r2 += dx * dx;
    dx = params[1] - x0[1];
    r2 += dx * dx;
    dx = sqrt(r2) - r0;
    r2 = dx * dx;
    double result = lnnormfac - r2 / twosigmasq;
    if (!isfinite(result)) result = -INFINITY;
    return result;
  }
};
// d-dim Gaussian shell pair: arithmetic of gaussian_shell_likelihood::evaluate_log (example.cc:373-403)
class shells_like : public probability_function {
public:
  int dim; bool logx; double x0, r0, lnnormfac, twosigmasq, sigmapoverm, lnsigmapoverm;
  shells_like(const stateSpace *sp, int dim, double x0, double r0, double sigma, double sigmapoverm, bool logx)
      : probability_function(sp), dim(dim), logx(logx), x0(x0), r0(r0), sigmapoverm(sigmapoverm) {
    twosigmasq = 2 * sigma * sigma;
    lnnormfac = -0.5 * std::log(M_PI * twosigmasq); // example.cc:341-342
    lnsigmapoverm = std::log(sigmapoverm);
  }
  double evaluate_log(state &s) {
    valarray<double> params = s.get_params();
    double x = params[0];
    if (logx) {
      if (x < 0) return -INFINITY;
      x = std::log(x);
    }
    double dx = x - x0;
    double r2 = dx * dx;
    for (int i = 1; i < dim; i++) { dx = params[i]; r2 += dx * dx; }
    dx = sqrt(r2) - r0;
    r2 = dx * dx;
    double resultp = -r2 / (twosigmasq * sigmapoverm) - 0.5 * lnsigmapoverm;
    dx = x + x0;
    r2 = dx * dx;
    for (int i = 1; i < dim; i++) { dx = params[i]; r2 += dx * dx; }
    dx = sqrt(r2) - r0;
    r2 = dx * dx;
    double resultm = -r2 / (twosigmasq / sigmapoverm) + 0.5 * lnsigmapoverm;
    double result = lnnormfac;
    if (resultm > resultp) result += resultm;
    else result += resultp;
    if (!isfinite(result)) result = -INFINITY;
    return result;
  }
};

static map<string, string> kv;
static string S(const string &k, const string &def) { return kv.count(k) ? kv[k] : def; }
static double D(const string &k, double def) { return kv.count(k) ? atof(kv[k].c_str()) : def; }
static int I(const string &k, int def) { return kv.count(k) ? atoi(kv[k].c_str()) : def; }
static vector<double> readvec(const string &path) {
  vector<double> v; ifstream in(path.c_str(), ios::binary);
  if (!in) { cerr << "cannot read " << path << endl; exit(2); }
  in.seekg(0, ios::end); size_t n = in.tellg() / sizeof(double); in.seekg(0);
  v.resize(n); in.read((char *)v.data(), n * sizeof(double)); return v;
}
static void wi(FILE *f, long long v) { fwrite(&v, sizeof(v), 1, f); }
static void wd(FILE *f, double v) { fwrite(&v, sizeof(v), 1, f); }

int main(int argc, char **argv) {
  for (int i = 1; i < argc; i++) {
    string a(argv[i]); size_t p = a.find('=');
    if (p == string::npos) { cerr << "bad arg " << a << endl; return 2; }
    kv[a.substr(0, p)] = a.substr(p + 1);
  }
  string model = S("model", "gauss");
  int d = I("dim", 2), nrungs = I("rungs", 8), nsteps = I("steps", 1000), save_every = I("save_every", 1);
  double seed = D("seed", 0.224), Tmax = D("Tmax", 1e9), swap_rate = D("swap_rate", 0.1);
  double evolve_rate = D("evolve_rate", 0.0), evolve_cut = D("evolve_lpost_cut", -1.0);
  double dpriormin = D("dprior_min", -30);
  int de_ni = I("de_ni", 50);
  string prop = S("prop", "default"), prior_kind = S("prior", "uniform"), out = S("out", "ref_trace.bin");
  string bound = S("bound", "open");

  ProbabilityDist::setSeed(seed);               // example.cc:493 ; newran1.cxx:341-346
  globalRNG.reset(ProbabilityDist::getPRNG());

  // ---- state space, prior ---------------------------------------------------------------
  valarray<double> centers(d), halfw(d);
  vector<double> pc = kv.count("centers") ? readvec(kv["centers"]) : vector<double>();
  vector<double> ph = kv.count("halfwidths") ? readvec(kv["halfwidths"]) : vector<double>();
  for (int i = 0; i < d; i++) {
    centers[i] = pc.size() ? pc[i] : 0.5;
    halfw[i] = ph.size() ? ph[i] : 0.5;
  }
  stateSpace space(d);
  {
    vector<string> names(d);
    for (int i = 0; i < d; i++) { ostringstream ss; ss << "p" << i; names[i] = ss.str(); }
    space.set_names(names);
    // bound=open: the prior alone cuts support (example.cc:88-100); bound=limit/wrap/reflect use the
    // prior box as the domain (states.cc:11-58).  bound may also be a per-dimension string e.g. "olwr".
    for (int i = 0; i < d; i++) {
      char c = bound.size() == (size_t)d ? bound[i] : bound[0];
      int t = c == 'l' ? boundary::limit : c == 'w' ? boundary::wrap : c == 'r' ? boundary::reflect : boundary::open;
      if (t != boundary::open) space.set_bound(i, boundary(t, t, centers[i] - halfw[i], centers[i] + halfw[i]));
    }
  }
  sampleable_probability_function *prior;
  valarray<double> scales;
  if (prior_kind == "uniform") {
    valarray<double> lo = centers - halfw, hi = centers + halfw;
    prior = new uniform_dist_product(&space, lo, hi);
  } else if (prior_kind == "gaussian") {
    prior = new gaussian_dist_product(&space, centers, halfw);
  } else if (prior_kind == "gaussian_wrap") { // wrap_probability: images of wrapped dimensions are summed (probability_function.cc:57-78)
    prior = new gaussian_dist_product(&space, centers, halfw, true);
  } else { // mixed: types file of doubles (1 uniform, 2 gaussian)
    valarray<int> types(d);
    vector<double> pt = kv.count("types") ? readvec(kv["types"]) : vector<double>(d, 1.0);
    for (int i = 0; i < d; i++) types[i] = (int)pt[i];
    prior = new mixed_dist_product(&space, types, centers, halfw);
  }
  prior->getScales(scales);

  // ---- likelihood ---------------------------------------------------------------------------
  probability_function *like = 0;
  if (model == "gauss") {
    vector<double> x0(d);
    vector<double> px = kv.count("x0") ? readvec(kv["x0"]) : vector<double>();
    for (int i = 0; i < d; i++) x0[i] = px.size() ? px[i] : centers[i];
    like = new iso_gauss_like(&space, x0, D("sigma", 0.5));
  } else if (model == "sines") {
    valarray<int> ks(I("k", 2), d);
    valarray<double> mins = centers - halfw, maxs = centers + halfw;
    like = new sines(&space, D("height", 64), ks, mins, maxs, D("step_scale", log(2.0)));
  } else if (model == "poly" || model == "sinusoid") {
    vector<double> xs = readvec(kv["data_x"]), ys = readvec(kv["data_y"]), dys = readvec(kv["data_dy"]);
    like = new chi2_like(&space, model == "poly" ? 0 : 1, xs, ys, dys);
  } else if (model == "shell2d") {
    like = new shell2d_like(&space, D("shell_x0", 3.0), D("shell_x1", 0.0), D("shell_r0", 2.0), D("shell_sigma", 0.1));
  } else if (model == "shells") {
    like = new shells_like(&space, d, D("shell_x0", 3.0), D("shell_r0", 2.0), D("shell_sigma", 0.1), D("shell_spm", 1.0), I("shell_logx", 0) != 0);
  } else if (model == "fullcov") {
    vector<double> cinv = readvec(kv["cinv"]);
    like = new fullcov_like(&space, d, cinv, D("like0", 0.0));
  } else { cerr << "unknown model " << model << endl; return 2; }

  // ---- proposal -------------------------------------------------------------------------------
  proposal_distribution *cprop = 0;
  int Ninit = de_ni * d;
  if (prop == "default") {
    // the default mix of ptmcmc_sampler::select_proposal (ptmcmc.cc:67-139) with all flags at default
    int Ng = 6; double gshare = D("gauss_draw_frac", 0.2), g1d = D("gauss_1d_frac", 0.5);
    vector<proposal_distribution *> set(1 + Ng); vector<double> shares(1 + Ng), hot(1 + Ng);
    differential_evolution *de = new differential_evolution(0.1, D("de_g1_frac", 0.3), D("de_eps", 1e-4), 0.0, D("de_unlikely_alpha", 0));
    de->reduce_gamma(D("de_reduce_gamma", 4)); de->mix_temperatures_more(D("de_Tmix", 300));
    if (I("de_mixing", 0)) de->support_mixing(true); // as ptmcmc.cc:84 does; inert inside a set (chain.cc:1375 asks the SET, which says no)
    set[0] = de; shares[0] = 1 - gshare;
    double sum = (pow(2, Ng + 1) - 2), stepfac = 2, fac = pow(2.0 / stepfac, 4.0), sharefac = 1;
    double prop_adapt_rate = D("prop_adapt_rate", 0);
    if (prop_adapt_rate > 0) {
      // hierarchical: the Gaussian scales in a nested adaptive set, the top level adaptive only with prop_adapt_more (ptmcmc.cc:70-72,123-143)
      vector<proposal_distribution *> gset(Ng); vector<double> gshares(Ng);
      for (int i = 0; i < Ng; i++) {
        fac *= stepfac;
        gset[i] = new gaussian_prop(scales / 100.0 / fac, g1d, false);
        sharefac *= 2; gshares[i] = sharefac / sum;
      }
      set.resize(2); shares.resize(2); hot.resize(2);
      set[1] = new proposal_distribution_set(gset, gshares, prop_adapt_rate);
      shares[1] = gshare;
      cprop = new proposal_distribution_set(set, shares, I("prop_adapt_more", 0) ? prop_adapt_rate : 0, 0, hot);
    } else {
    for (int i = 1; i < 1 + Ng; i++) {
      fac *= stepfac;
      set[i] = new gaussian_prop(scales / 100.0 / fac, g1d, false);
      sharefac *= 2; shares[i] = sharefac / sum * gshare;
    }
    cprop = new proposal_distribution_set(set, shares, D("adapt_rate", 0), 0, hot);
    }
  } else if (prop == "de") {
    differential_evolution *de = new differential_evolution(D("de_snooker", 0.1), D("de_g1_frac", 0.3), D("de_eps", 1e-4),
                                                            D("de_ignore_frac", 0.0), D("de_unlikely_alpha", 0));
    de->reduce_gamma(D("de_reduce_gamma", 4));
    if (I("de_mixing", 0)) { de->support_mixing(true); de->mix_temperatures_more(D("de_Tmix", 1)); }
    cprop = de;
  } else if (prop == "gauss") {
    cprop = new gaussian_prop(scales / D("gauss_div", 10.0), D("gauss_1d_frac", 0.5), false);
  } else if (prop == "cov" || prop == "covde") {
    vector<double> c = readvec(kv["prop_cov"]);
    Eigen::MatrixXd covar(d, d);
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) covar(i, j) = c[(size_t)i * d + j];
    streambuf *old = cout.rdbuf(); ostringstream sink; cout.rdbuf(sink.rdbuf()); // constructor is chatty
    // The reference's covariance constructor self-initialises a member (`:sigmas(sigmas)`, proposal_distribution.hh:164),
    // i.e. copy-constructs a valarray from its own uninitialised storage: undefined behaviour that throws bad_alloc whenever
    // the heap garbage looks like a huge size (seen for d >= 28).  Constructing the object in ZEROED storage makes that
    // self-copy see an empty valarray, which is what the author evidently relied on; nothing else changes.
    void *gmem = calloc(1, sizeof(gaussian_prop));
    gaussian_prop *g = new (gmem) gaussian_prop(covar, D("gauss_1d_frac", 0.0), false);
    cout.rdbuf(old);
    { // dump the eigen-decomposition the reference actually uses (proposal_distribution.hh:173-176)
      FILE *f = fopen((out + ".eig").c_str(), "wb");
      for (int i = 0; i < d; i++) wd(f, g->sigmas[i]);
      for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) wd(f, g->diagTransform(i, j));
      fclose(f);
    }
    if (prop == "cov") cprop = g;
    else {
      differential_evolution *de = new differential_evolution(0.1, 0.3, 1e-4, 0.0, 0);
      de->reduce_gamma(4);
      vector<proposal_distribution *> set(2); vector<double> shares(2, 0.5), hot(2);
      set[0] = g; set[1] = de;
      cprop = new proposal_distribution_set(set, shares, 0, 0, hot);
    }
  } else if (prop == "prior") { // DE + prior draws (draw_from_dist, proposal_distribution.hh:119-132)
    differential_evolution *de = new differential_evolution(0.1, 0.3, 1e-4, 0.0, 0);
    de->reduce_gamma(4);
    vector<proposal_distribution *> set(2); vector<double> shares(2), hot(2);
    set[0] = de; shares[0] = 1 - D("prior_draw_frac", 0.3);
    set[1] = new draw_from_dist(*prior); shares[1] = D("prior_draw_frac", 0.3);
    // thermal weighting of the prior draws as ptmcmc_sampler::select_proposal sets it up (ptmcmc.cc:95-101): hot share 1 for the prior draw
    double Tpow = D("Tpow", 0);
    if (Tpow > 0) { hot[0] = D("hot_de", 0.0); hot[1] = D("hot_prior", 1.0); }
    cprop = new proposal_distribution_set(set, shares, D("adapt_rate", 0), Tpow, hot);
  } else { cerr << "unknown prop " << prop << endl; return 2; }

  // ---- chain: exactly ptmcmc_sampler::initialize (ptmcmc.cc:508-522) ------------------------------
  streambuf *old = cout.rdbuf(); ostringstream sink;
  if (!I("verbose", 0)) cout.rdbuf(sink.rdbuf());
  parallel_tempering_chains *ptc = new parallel_tempering_chains(nrungs, Tmax, swap_rate, save_every, false, false, dpriormin);
  if (evolve_rate > 0) ptc->evolve_temps(evolve_rate, evolve_cut);
  ptc->initialize(like, prior, Ninit, "");
  ptc->set_proposal(*cprop);
  cprop->set_chain(ptc);
  chrono::steady_clock::time_point t_begin = chrono::steady_clock::now();
  for (int s = 0; s < nsteps; s++) ptc->step();
  double step_seconds = chrono::duration<double>(chrono::steady_clock::now() - t_begin).count();
  cout.rdbuf(old);

  // ---- dump ---------------------------------------------------------------------------------------
  FILE *f = fopen(out.c_str(), "wb");
  wi(f, 0x7074726566LL); wi(f, nrungs); wi(f, d); wi(f, nsteps); wi(f, Ninit); wi(f, save_every);
  for (int r = 0; r < nrungs; r++) {
    MH_chain &c = ptc->chains[r];
    wi(f, c.Nsize); wi(f, c.Nhist); wi(f, c.Ntries); wi(f, c.Naccept); wi(f, c.last_type);
    wd(f, c.invtemp); wd(f, c.current_lpost); wd(f, c.current_llike); wd(f, c.MAPlpost);
    for (int k = 0; k < c.Nsize; k++) {
      for (int j = 0; j < d; j++) wd(f, c.states[k].get_param(j));
      wd(f, c.lposts[k]); wd(f, c.llikes[k]); wd(f, c.acceptance_ratio[k]); wd(f, c.invtemps[k]);
      wd(f, (double)c.types[k]);
    }
  }
  for (int r = 0; r < nrungs - 1; r++) { wi(f, ptc->swap_count[r]); wi(f, ptc->swap_accept_count[r]); }
  for (int r = 0; r < nrungs; r++) { wi(f, ptc->directions[r]); wi(f, ptc->ups[r]); wi(f, ptc->downs[r]); wi(f, ptc->instances[r]); }
  fclose(f);
  if (I("ess", 0)) {
    // the run loop's effective-sample report (ptmcmc.cc:645): report_effective_samples(-1, save_every*1000, save_every, esslimit)
    cout.rdbuf(sink.rdbuf());
    ptc->reporting = false;
    pair<double, int> el = ptc->report_effective_samples(-1, I("ess_width", save_every * 1000), I("ess_every", save_every), D("esslimit", -1));
    cout.rdbuf(old);
    printf("ref_ess: ess=%.17g length=%d\n", el.first, el.second);
  }
  long long total = 0;
  for (int r = 0; r < nrungs; r++) total += ptc->chains[r].Nhist;
  printf("ref_trace: model=%s d=%d rungs=%d steps=%d Ninit=%d total_Nhist=%lld cold_Nsize=%d step_seconds=%.6f\n",
         model.c_str(), d, nrungs, nsteps, Ninit, total, ptc->chains[0].Nsize, step_seconds);
  return 0;
}
