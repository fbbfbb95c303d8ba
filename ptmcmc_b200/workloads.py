"""Workload specifications: one `Spec` describes a hot-path configuration (model, prior, state space, proposal mix,
ladder) and configures any engine behind the C ABI; it also produces the argument list of oracle/_ref/ref_trace, the driver that
runs the unmodified reference on the same configuration (used by the parity tests and by bench.py's CPU arm).  The data
generators reproduce the synthetic inputs of BASELINE configs B, C2 and D.
"""
import os
import numpy as np
from ptmcmc_b200 import _capi as K


def default_mix(scales, gauss_draw_frac=0.2, g1d=0.5, unlikely_alpha=0.0):
    """the 7-member default set of ptmcmc_sampler::select_proposal (ptmcmc.cc:67-139)"""
    Ng = 6
    props = [dict(kind=K.PROP_DE, share=1 - gauss_draw_frac, snooker=0.1, gamma_one_frac=0.3, b_small=1e-4,
                  ignore_frac=0.0, unlikely_alpha=unlikely_alpha, reduce_gamma=4.0)]
    total = 2.0 ** (Ng + 1) - 2
    stepfac = 2.0
    fac = (2.0 / stepfac) ** 4.0
    sharefac = 1.0
    for _ in range(Ng):
        fac *= stepfac
        sharefac *= 2
        props.append(dict(kind=K.PROP_GAUSS, share=sharefac / total * gauss_draw_frac, sigmas=scales / 100.0 / fac,
                          one_d_frac=g1d))
    return props


class Spec:
    def __init__(self, model, dim, rungs, *, prop="default", prior="uniform", centers=None, halfwidths=None,
                 bound="open", Tmax=1e9, swap_rate=0.1, evolve_rate=0.0, evolve_lpost_cut=-1.0, save_every=1,
                 de_ni=50, seed=0.224, extra=None, prior_types=None):
        self.model, self.dim, self.rungs, self.prop, self.prior = model, dim, rungs, prop, prior
        self.centers = np.full(dim, 0.5) if centers is None else np.asarray(centers, dtype=float)
        self.halfwidths = np.full(dim, 0.5) if halfwidths is None else np.asarray(halfwidths, dtype=float)
        self.bound, self.Tmax, self.swap_rate = bound, Tmax, swap_rate
        self.evolve_rate, self.evolve_lpost_cut, self.save_every = evolve_rate, evolve_lpost_cut, save_every
        self.de_ni, self.seed = de_ni, seed
        self.extra = extra or {}
        self.prior_types = prior_types
        self.eig = None  # (sigmas, transform) for prop=cov*, filled from the reference dump or by the caller

    # ---------------------------------------------------------------- engine side
    def config(self, n_ladders=1, **kw):
        kw.setdefault("n_init", self.de_ni * self.dim)
        kw.setdefault("save_every", self.save_every)
        kw.setdefault("swap_rate", self.swap_rate)
        kw.setdefault("Tmax", self.Tmax)
        kw.setdefault("evolve_rate", self.evolve_rate)
        kw.setdefault("evolve_lpost_cut", self.evolve_lpost_cut)
        return K.make_config(n_ladders, self.rungs, self.dim, **kw)

    def scales(self):
        """getScales: uniform half-width or Gaussian sigma; log-type dimensions of a mixed prior: center * min(1, (h - 1/h) / 2)
        (probability_function.hh:172-181)"""
        sc = self.halfwidths.copy()
        if self.prior == "mixed":
            for i, t in enumerate(self.prior_types):
                if int(t) == K.PRIOR_LOG:
                    fac = (self.halfwidths[i] - 1 / self.halfwidths[i]) / 2
                    if fac > 1:
                        fac = 1
                    sc[i] = self.centers[i] * fac
        return sc

    def prior_ab(self):
        """(a, b) of every 1-D factor as ptg_set_prior takes them: (min, max), (x0, sigma) for Gaussian factors, and for log-type factors
        the multiplicative range (center / halfwidth, center * halfwidth) of mixed_dist_product (probability_function.cc:243-249)"""
        types = np.asarray(self.prior_types, dtype=np.int32)
        a = np.where(types == K.PRIOR_GAUSSIAN, self.centers, self.centers - self.halfwidths)
        b = np.where(types == K.PRIOR_GAUSSIAN, self.halfwidths, self.centers + self.halfwidths)
        lg = types == K.PRIOR_LOG
        a = np.where(lg, self.centers / self.halfwidths, a)
        b = np.where(lg, self.centers * self.halfwidths, b)
        return types, a, b

    def bounds(self):
        code = {"o": K.BOUND_OPEN, "l": K.BOUND_LIMIT, "w": K.BOUND_WRAP, "r": K.BOUND_REFLECT}
        b = self.bound if len(self.bound) == self.dim and self.dim > 1 else self.bound[0] * self.dim
        if self.bound == "open":
            b = "o" * self.dim
        t = np.array([code[c] for c in b], dtype=np.int32)
        lo = np.where(t == K.BOUND_OPEN, -np.inf, self.centers - self.halfwidths)
        hi = np.where(t == K.BOUND_OPEN, np.inf, self.centers + self.halfwidths)
        return t, lo, hi

    def setup(self, api):
        d = self.dim
        t, lo, hi = self.bounds()
        api.set_space(t, t, lo, hi)
        if self.prior == "uniform":
            api.set_prior([K.PRIOR_UNIFORM] * d, self.centers - self.halfwidths, self.centers + self.halfwidths)
        elif self.prior == "gaussian":
            api.set_prior([K.PRIOR_GAUSSIAN] * d, self.centers, self.halfwidths)
        elif self.prior == "gaussian_wrap":   # gaussian_dist_product(..., wrap_probability = true)
            api.set_prior([K.PRIOR_GAUSSIAN_WRAPPED] * d, self.centers, self.halfwidths)
        else:
            api.set_prior(*self.prior_ab())
        e = self.extra
        if self.model == "gauss":
            sigma = e.get("sigma", 0.5)
            tw = 2 * sigma * sigma
            x0 = np.asarray(e.get("x0", self.centers), dtype=float)
            api.set_likelihood(K.LIKE_GAUSS_ISO, np.concatenate([[-0.5 * d * np.log(np.pi * tw), tw], x0]))
        elif self.model == "shell2d":   # gaussian_shell_2D_likelihood (example.cc:147-222)
            sigma = e.get("shell_sigma", 0.1)
            tw = 2 * sigma * sigma
            api.set_likelihood(K.LIKE_SHELL2D, [-0.5 * np.log(np.pi * tw), tw, e.get("shell_r0", 2.0), e.get("shell_x0", 3.0), e.get("shell_x1", 0.0)])
        elif self.model == "shells":    # gaussian_shell_likelihood (example.cc:226-421)
            sigma, spm = e.get("shell_sigma", 0.1), e.get("shell_spm", 1.0)
            tw = 2 * sigma * sigma
            api.set_likelihood(K.LIKE_SHELLS, [-0.5 * np.log(np.pi * tw), tw, e.get("shell_r0", 2.0), e.get("shell_x0", 3.0), spm, np.log(spm),
                                               float(e.get("shell_logx", 0))])
        elif self.model == "sines":
            k = e.get("k", 2)
            api.set_likelihood(K.LIKE_SINES, np.concatenate([[e.get("height", 64.0), e.get("step_scale", np.log(2.0))],
                                                             np.full(d, float(k)), self.centers - self.halfwidths,
                                                             self.centers + self.halfwidths]))
        elif self.model in ("poly", "sinusoid"):
            kind = K.LIKE_POLY_CHI2 if self.model == "poly" else K.LIKE_SINUSOID_CHI2
            dy = np.asarray(e["data_dy"], dtype=float)
            api.set_likelihood(kind, [0.0], np.concatenate([e["data_x"], e["data_y"], dy * dy]))
        elif self.model == "fullcov":
            api.set_likelihood(K.LIKE_GAUSS_FULLCOV, [e.get("like0", 0.0)], np.asarray(e["cinv"], dtype=float).ravel())
        elif self.model == "flat":
            api.set_likelihood(K.LIKE_FLAT, [0.0])
        else:
            raise ValueError(self.model)
        sc = self.scales()
        de = dict(kind=K.PROP_DE, share=1.0, snooker=e.get("de_snooker", 0.1), gamma_one_frac=0.3, b_small=1e-4,
                  ignore_frac=e.get("de_ignore_frac", 0.0), unlikely_alpha=e.get("de_unlikely_alpha", 0.0), reduce_gamma=4.0)
        if self.prop == "default" and e.get("prop_adapt_rate", 0) > 0:
            # ptmcmc_sampler::select_proposal with prop_adapt_rate > 0 (ptmcmc.cc:70-72,123-143): the six Gaussian scales form a nested,
            # adaptive set; the top level [DE, nested set] adapts only with prop_adapt_more
            mix = default_mix(sc, unlikely_alpha=e.get("de_unlikely_alpha", 0.0))
            for i, p in enumerate(mix[1:]):
                p["share"] = 2.0 ** (i + 1) / (2.0 ** len(mix) - 2)   # gshares[i] = sharefac / sum (:128)
            api.set_proposals(mix)
            api.set_nested_set(1, len(mix) - 1, share=0.2, adapt_rate=e["prop_adapt_rate"])
            if e.get("prop_adapt_more", 0):
                api.set_proposal_options(adapt_rate=e["prop_adapt_rate"])
        elif self.prop == "default":
            api.set_proposals(default_mix(sc, unlikely_alpha=e.get("de_unlikely_alpha", 0.0)))
        elif self.prop == "de":
            api.set_proposals([de], wrap_in_set=False)
        elif self.prop == "gauss":
            api.set_proposals([dict(kind=K.PROP_GAUSS, sigmas=sc / e.get("gauss_div", 10.0), one_d_frac=e.get("gauss_1d_frac", 0.5))],
                              wrap_in_set=False)
        elif self.prop in ("cov", "covde"):
            sig, M = self.eig
            g = dict(kind=K.PROP_GAUSS, sigmas=sig, transform=M, one_d_frac=e.get("gauss_1d_frac", 0.0), share=0.5)
            if self.prop == "cov":
                api.set_proposals([g], wrap_in_set=False)
            else:
                de2 = dict(de); de2["share"] = 0.5
                api.set_proposals([g, de2])
        elif self.prop == "prior":
            f = e.get("prior_draw_frac", 0.3)
            de2 = dict(de); de2["share"] = 1 - f
            # thermal weighting of the prior draws as ptmcmc_sampler::select_proposal sets it up (ptmcmc.cc:95-101, prior_draw_Tpow)
            Tpow = e.get("Tpow", 0.0)
            de2["hot_share"] = e.get("hot_de", 0.0) if Tpow > 0 else 0.0
            api.set_proposals([de2, dict(kind=K.PROP_PRIOR_DRAW, share=f, hot_share=e.get("hot_prior", 1.0) if Tpow > 0 else 0.0)], Tpow=Tpow)
        else:
            raise ValueError(self.prop)
        if e.get("adapt_rate", 0) != 0 or e.get("de_mixing", 0):
            bare = self.prop in ("de", "gauss", "cov")
            # inside a set the reference never mixes (chain.cc:1375 asks the set, which does not support mixing): only a bare DE does
            api.set_proposal_options(adapt_rate=e.get("adapt_rate", 0.0), de_mixing=bool(e.get("de_mixing", 0)) and bare,
                                     de_Tmix=e.get("de_Tmix", 1.0 if bare else 300.0))

    # ---------------------------------------------------------------- reference side
    def ref_args(self, tmpdir, steps, out):
        def dump(name, arr):
            p = os.path.join(tmpdir, name + ".f64")
            np.asarray(arr, dtype=np.float64).tofile(p)
            return p
        a = ["model=%s" % self.model, "dim=%d" % self.dim, "rungs=%d" % self.rungs, "steps=%d" % steps,
             "save_every=%d" % self.save_every, "seed=%.17g" % self.seed, "Tmax=%.17g" % self.Tmax,
             "swap_rate=%.17g" % self.swap_rate, "evolve_rate=%.17g" % self.evolve_rate,
             "evolve_lpost_cut=%.17g" % self.evolve_lpost_cut, "de_ni=%d" % self.de_ni, "prop=%s" % self.prop,
             "prior=%s" % self.prior, "bound=%s" % self.bound, "out=%s" % out,
             "centers=" + dump("centers", self.centers), "halfwidths=" + dump("halfwidths", self.halfwidths)]
        if self.prior == "mixed":
            a.append("types=" + dump("types", np.asarray(self.prior_types, dtype=float)))
        for k, v in self.extra.items():
            if isinstance(v, np.ndarray):
                a.append("%s=%s" % (k, dump(k, v)))
            else:
                a.append("%s=%.17g" % (k, v))
        return a



# ---------------------------------------------------------------------------------------------------- named cases
def poly_data(n=1000, d=5, seed=5):
    """config B: x_k = -10 + 0.02 (k + 1/2), truth c ~ U(-10,10)^d, unit noise"""
    rng = np.random.default_rng(seed)
    xs = -10 + 0.02 * (np.arange(n) + 0.5)
    truth = rng.uniform(-10, 10, d)
    ys = sum(truth[j] * xs ** j for j in range(d)) + rng.normal(size=n)
    return dict(data_x=xs, data_y=ys, data_dy=np.ones(n))


def sinusoid_data(n=10000, dt=1e-3, seed=7):
    """config C2: y(t) = sum_k A_k sin(2 pi f_k t + phi_k) + N(0,1)"""
    rng = np.random.default_rng(seed)
    t = np.arange(n) * dt
    A, f, ph = [1, 0.7, 0.4], [1.3, 3.1, 7.7], [0.3, 1.1, 2.0]
    y = sum(A[k] * np.sin(2 * np.pi * f[k] * t + ph[k]) for k in range(3)) + rng.normal(size=n)
    return dict(data_x=t, data_y=y, data_dy=np.ones(n))


def sinusoid_spec(rungs, n=10000, dt=1e-3, **kw):
    c = np.array([1, 5, np.pi] * 3, dtype=float)
    return Spec("sinusoid", 9, rungs, centers=c, halfwidths=c.copy(), bound="oowoowoow", extra=sinusoid_data(n, dt), **kw)


def fullcov_spec(d, rungs, seed=100, prop="covde", prior_scale=None, **kw):
    """config D: C = Wishart(nu = d+5, I) sample (cython/exampleGaussian.py:181-182); prior +-prior_scale sqrt(C_ii);
    Gaussian proposal with covariance 2.38^2/d C, eigen-rotated (exampleGaussian.py:88,95).

    prior_scale: the example's priorscale is 100 (exampleGaussian.py:70).  The reference evaluates a prior as log(prod_i pdf_i)
    (probability_function.hh:59, .cc:156-166), and prod_i 1 / (200 sqrt(C_ii)) underflows to 0 for d = 100 (~1e-330): the log prior is then
    -inf EVERYWHERE, the likelihood gate (chain.cc:980) never opens, every proposal is accepted through the NaN Hastings ratio
    (chain.cc:989-1001) and the run is a likelihood-free random walk -- in the reference as much as here (the oracle and the engine
    reproduce it bit for bit).  A d = 100 workload that exercises the likelihood therefore needs a box whose density survives the product:
    the default is 100 up to d = 64, 10 up to d = 110 (prod ~ 1e-231) and 5 above."""
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(d + 5, d))
    Cm = A.T @ A
    cinv = np.linalg.inv(Cm)
    like0 = -0.5 * (d * np.log(2 * np.pi) + np.linalg.slogdet(Cm)[1])
    if prior_scale is None:
        prior_scale = 100 if d <= 64 else (10 if d <= 110 else 5)
    assert np.sum(np.log(2 * prior_scale * np.sqrt(np.diag(Cm)))) < 700, "the prior density product underflows: lower prior_scale"
    sp = Spec("fullcov", d, rungs, centers=np.zeros(d), halfwidths=prior_scale * np.sqrt(np.diag(Cm)), prop=prop,
              extra=dict(cinv=cinv.ravel(), like0=like0), **kw)
    w, V = np.linalg.eigh(2.38 ** 2 / d * Cm)
    sp.eig = (np.sqrt(w), V)
    sp.extra["prop_cov"] = (2.38 ** 2 / d * Cm).ravel()
    return sp


