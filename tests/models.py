"""Model / proposal specifications shared by the parity tests.

One `Spec` describes a hot-path configuration and can (a) configure any engine behind the C ABI
(the CUDA engine or the CPU oracle) and (b) produce the command line of oracle/_ref/ref_trace, the driver that runs
the unmodified reference on the same configuration.
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
        return self.halfwidths.copy()  # getScales: uniform half-width or Gaussian sigma

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
        else:
            types = np.asarray(self.prior_types, dtype=np.int32)
            a = np.where(types == K.PRIOR_GAUSSIAN, self.centers, self.centers - self.halfwidths)
            b = np.where(types == K.PRIOR_GAUSSIAN, self.halfwidths, self.centers + self.halfwidths)
            api.set_prior(types, a, b)
        e = self.extra
        if self.model == "gauss":
            sigma = e.get("sigma", 0.5)
            tw = 2 * sigma * sigma
            x0 = np.asarray(e.get("x0", self.centers), dtype=float)
            api.set_likelihood(K.LIKE_GAUSS_ISO, np.concatenate([[-0.5 * d * np.log(np.pi * tw), tw], x0]))
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
        if self.prop == "default":
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
            api.set_proposals([de2, dict(kind=K.PROP_PRIOR_DRAW, share=f)])
        else:
            raise ValueError(self.prop)

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


def read_ref_trace(path):
    """binary dump written by oracle/ref_drivers/ref_trace.cc"""
    raw = np.fromfile(path, dtype=np.int64)
    f64 = raw.view(np.float64)
    assert raw[0] == 0x7074726566
    R, d, steps, ninit, save_every = (int(v) for v in raw[1:6])
    p = 6
    rungs = []
    for _ in range(R):
        nsize, nhist, ntries, naccept, last_type = (int(v) for v in raw[p:p + 5]); p += 5
        invtemp, lpost, llike, maplpost = (float(v) for v in f64[p:p + 4]); p += 4
        rec = f64[p:p + nsize * (d + 5)].reshape(nsize, d + 5); p += nsize * (d + 5)
        rungs.append(dict(nsize=nsize, nhist=nhist, ntries=ntries, naccept=naccept, last_type=last_type, beta=invtemp,
                          lpost=lpost, llike=llike, map_lpost=maplpost, x=rec[:, :d].copy(), hlpost=rec[:, d].copy(),
                          hllike=rec[:, d + 1].copy(), hacc=rec[:, d + 2].copy(), hbeta=rec[:, d + 3].copy(),
                          htype=rec[:, d + 4].astype(np.int32)))
    sw = raw[p:p + 2 * (R - 1)].reshape(R - 1, 2); p += 2 * (R - 1)
    misc = raw[p:p + 4 * R].reshape(R, 4)
    return dict(R=R, dim=d, steps=steps, ninit=ninit, save_every=save_every, rungs=rungs, swap_count=sw[:, 0].copy(),
                swap_accept=sw[:, 1].copy(), directions=misc[:, 0].copy(), ups=misc[:, 1].copy(), downs=misc[:, 2].copy(),
                instances=misc[:, 3].copy())


def engine_dump(api, ladder=0):
    """same structure as read_ref_trace, from an engine behind the C ABI"""
    R, d = api.cfg.n_rungs, api.cfg.dim
    cnt = api.get_counters(); cur = api.get_current(); sw = api.get_swap_stats()
    rungs = []
    for r in range(R):
        i = ladder * R + r
        n = int(cnt["nsize"][i])
        hst = api.get_history(ladder, r, 0, n)
        rungs.append(dict(nsize=n, nhist=int(cnt["nhist"][i]), ntries=int(cnt["ntries"][i]), naccept=int(cnt["naccept"][i]),
                          last_type=int(cnt["last_type"][i]), beta=float(cur["beta"][i]), lpost=float(cur["lpost"][i]),
                          llike=float(cur["llike"][i]), map_lpost=float(cnt["map_lpost"][i]), x=hst["x"], hlpost=hst["lpost"],
                          hllike=hst["llike"], hacc=hst["acc"], hbeta=hst["beta"], htype=hst["type"]))
    return dict(R=R, dim=d, rungs=rungs, swap_count=sw["swap_count"][ladder], swap_accept=sw["swap_accept"][ladder],
                directions=sw["directions"][ladder], ups=sw["ups"][ladder], downs=sw["downs"][ladder],
                instances=sw["instances"][ladder])


def compare_dumps(a, b, rtol=0.0, what="", max_report=5, exact_x=True):
    """bit-exact (rtol=0) or relative comparison of two dumps; returns list of mismatch strings.
    exact_x: positions and acceptance ratios must be bit-identical even when rtol > 0 (tape parity)"""
    bad = []

    def chk(name, u, v, exact=False):
        u = np.asarray(u); v = np.asarray(v)
        if u.shape != v.shape:
            bad.append("%s %s: shape %s vs %s" % (what, name, u.shape, v.shape)); return
        if rtol == 0.0 or exact or u.dtype.kind in "iu":
            ok = (u == v) | ((u != u) & (v != v))
        else:
            with np.errstate(invalid="ignore"):
                ok = (np.abs(u - v) <= rtol * np.maximum(np.abs(u), np.abs(v))) | (u == v) | ((u != u) & (v != v))
        if not np.all(ok):
            idx = np.argwhere(~ok)[0]
            bad.append("%s %s: %d mismatches, first at %s: %r vs %r" % (what, name, int((~ok).sum()), tuple(idx), u[tuple(idx)], v[tuple(idx)]))

    for r in range(a["R"]):
        ra, rb = a["rungs"][r], b["rungs"][r]
        for k in ("nsize", "nhist", "ntries", "naccept", "last_type"):
            if ra[k] != rb[k]:
                bad.append("%s rung %d %s: %r vs %r" % (what, r, k, ra[k], rb[k]))
        if ra["nsize"] != rb["nsize"]:
            continue
        chk("rung %d x" % r, ra["x"], rb["x"], exact=exact_x)
        chk("rung %d htype" % r, ra["htype"], rb["htype"])
        chk("rung %d hacc" % r, ra["hacc"], rb["hacc"], exact=exact_x)
        chk("rung %d hbeta" % r, ra["hbeta"], rb["hbeta"])
        chk("rung %d hllike" % r, ra["hllike"], rb["hllike"])
        chk("rung %d hlpost" % r, ra["hlpost"], rb["hlpost"])
        chk("rung %d beta" % r, ra["beta"], rb["beta"])
        chk("rung %d map" % r, ra["map_lpost"], rb["map_lpost"])
    for k in ("swap_count", "swap_accept", "directions", "ups", "downs", "instances"):
        chk(k, a[k], b[k])
    return bad[:max_report] if max_report else bad


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


def fullcov_spec(d, rungs, seed=100, prop="covde", **kw):
    """config D: C = Wishart(nu = d+5, I) sample (cython/exampleGaussian.py:181-182); prior +-100 sqrt(C_ii);
    Gaussian proposal with covariance 2.38^2/d C, eigen-rotated (exampleGaussian.py:88,95)"""
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(d + 5, d))
    Cm = A.T @ A
    cinv = np.linalg.inv(Cm)
    like0 = -0.5 * (d * np.log(2 * np.pi) + np.linalg.slogdet(Cm)[1])
    sp = Spec("fullcov", d, rungs, centers=np.zeros(d), halfwidths=100 * np.sqrt(np.diag(Cm)), prop=prop,
              extra=dict(cinv=cinv.ravel(), like0=like0), **kw)
    w, V = np.linalg.eigh(2.38 ** 2 / d * Cm)
    sp.eig = (np.sqrt(w), V)
    sp.extra["prop_cov"] = (2.38 ** 2 / d * Cm).ravel()
    return sp


def parity_cases():
    """(name, spec, steps, n_ladders): small versions of BASELINE configs A-D plus the reference's edge cases"""
    return [
        ("A_gauss2d_default", Spec("gauss", 2, 8, centers=[2, -3], halfwidths=[2, 3]), 1000, 3),
        ("C1_sines_d3_R32", Spec("sines", 3, 32, seed=0.1234), 400, 2),
        ("sines_evolve", Spec("sines", 2, 8, seed=0.012556, evolve_rate=0.01), 1500, 1),
        ("sines_evolve_cut", Spec("sines", 2, 8, seed=0.012556, evolve_rate=0.01, evolve_lpost_cut=0.5), 1500, 2),
        ("gauss3_de_save3", Spec("gauss", 3, 6, centers=[2, -3, 5], halfwidths=[2, 3, 5], prop="de", save_every=3), 1500, 1),
        ("wrap", Spec("gauss", 2, 5, centers=[2, -3], halfwidths=[2, 3], bound="w", extra=dict(sigma=3.0)), 1000, 1),
        ("reflect", Spec("gauss", 2, 5, centers=[2, -3], halfwidths=[2, 3], bound="r", extra=dict(sigma=3.0)), 1000, 1),
        ("limit_gauss", Spec("gauss", 3, 6, centers=[2, -3, 5], halfwidths=[2, 3, 5], prop="gauss", bound="l"), 1000, 1),
        ("prior_draw_mixed", Spec("gauss", 2, 5, centers=[2, -3], halfwidths=[2, 3], prop="prior", prior="mixed", prior_types=[1, 2]), 1000, 1),
        ("gaussian_prior", Spec("gauss", 4, 4, centers=[0, 1, 2, 3], halfwidths=[1, 2, 3, 4], prior="gaussian", Tmax=100, extra=dict(sigma=0.7)), 800, 1),
        ("unlikely_alpha", Spec("gauss", 2, 5, centers=[2, -3], halfwidths=[2, 3], extra=dict(de_unlikely_alpha=0.5)), 1000, 1),
        ("single_MH_chain", Spec("sines", 2, 1, prop="de", seed=0.1234), 2000, 2),
        ("B_poly", Spec("poly", 5, 4, centers=np.zeros(5), halfwidths=np.full(5, 10.0), prop="de", Tmax=1e6, extra=poly_data()), 200, 1),
        ("C2_sinusoid", sinusoid_spec(4, n=500, dt=0.02), 150, 1),
        ("D_fullcov_d6", fullcov_spec(6, 4, Tmax=100), 800, 1),
        # ladder widths of the warp kernel: 24 rungs in a 32-lane group (ghost lanes), 40 rungs (shared-memory kernel)
        ("R24_evolve", Spec("sines", 2, 24, seed=0.31, evolve_rate=0.01, swap_rate=0.2), 500, 2),
        ("R40_two_warps", Spec("gauss", 2, 40, centers=[2, -3], halfwidths=[2, 3], seed=0.77), 300, 1),
        # history shorter than 10 d at the start: differential evolution is not ready and the set falls through to the next
        # ready member whose bin covers the draw (proposal_distribution.cc:105-112, proposal_distribution.hh:399-407)
        ("de_not_ready", Spec("gauss", 2, 4, centers=[2, -3], halfwidths=[2, 3], seed=0.83, de_ni=3, Tmax=100), 300, 3),
        ("R3_high_swap_rate", Spec("gauss", 2, 3, centers=[2, -3], halfwidths=[2, 3], seed=0.41, swap_rate=0.9, Tmax=50), 800, 5),
    ] + wide_cases()


def wide_cases():
    """dim > 16: the warp-per-chain kernels (BASELINE config D in small: full-covariance Gaussian, eigen-rotated proposal + DE)"""
    d1 = fullcov_spec(100, 24, Tmax=1e4, de_ni=11, prop="covde", swap_rate=0.2)
    d1.extra["gauss_1d_frac"] = 0.3
    return [
        ("W_fullcov_d20_R4", fullcov_spec(20, 4, Tmax=100), 300, 2),
        ("W_fullcov_d40_R6_evolve", fullcov_spec(40, 6, Tmax=100, evolve_rate=0.01), 200, 1),
        ("W_D_fullcov_d100_R8", fullcov_spec(100, 8, Tmax=1e4, de_ni=12), 120, 1),
        ("W_gauss_d33_default_wrap", Spec("gauss", 33, 5, centers=np.linspace(-1, 1, 33), halfwidths=np.full(33, 3.0), bound="w",
                                          extra=dict(sigma=1.5), de_ni=12), 200, 2),
        ("W_gauss_d24_gaussprior_de", Spec("gauss", 24, 4, centers=np.zeros(24), halfwidths=np.full(24, 2.0), prior="gaussian", prop="de",
                                           Tmax=100, extra=dict(sigma=1.0, de_unlikely_alpha=0.3), de_ni=12), 200, 1),
        ("W_D_fullcov_d100_R24_cov1d", d1, 60, 1),
    ]
