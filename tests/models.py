"""Test-side helpers: reference-trace reader, engine dumps and their comparison, and the named parity cases.
The workload specifications themselves live in ptmcmc_b200/workloads.py (bench.py uses them too)."""
import os
import numpy as np
from ptmcmc_b200 import _capi as K
from ptmcmc_b200.workloads import Spec, default_mix, poly_data, sinusoid_data, sinusoid_spec, fullcov_spec  # noqa: F401


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

    def chk(name, u, v, exact=False, scale_to_column=False):
        u = np.asarray(u); v = np.asarray(v)
        if u.shape != v.shape:
            bad.append("%s %s: shape %s vs %s" % (what, name, u.shape, v.shape)); return
        if rtol == 0.0 or exact or u.dtype.kind in "iu":
            ok = (u == v) | ((u != u) & (v != v))
        else:
            with np.errstate(invalid="ignore"):
                ref = np.maximum(np.abs(u), np.abs(v))
                if scale_to_column and u.ndim == 2 and len(u):
                    # positions: a proposal x + gamma (a - b) cancels, so an ulp of the parameter's scale is the honest unit near zero
                    ref = np.maximum(ref, np.nanmax(np.where(np.isfinite(u), np.abs(u), 0.0), axis=0, keepdims=True))
                if not exact_x:
                    ref = np.maximum(ref, 1.0)   # libm-dependent positions: log-posteriors that cancel to ~0 are compared on the scale of their terms
                ok = (np.abs(u - v) <= rtol * ref) | (u == v) | ((u != u) & (v != v))
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
        chk("rung %d x" % r, ra["x"], rb["x"], exact=exact_x, scale_to_column=True)
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
        # the remaining mixed_dist_product factor types (probability_function.cc:235-249; ProbabilityDist.h:107-139,171-232): polar angle,
        # co-latitude, logarithmic; with the default proposal mix (scales of log-type dimensions, probability_function.hh:172-181) and with
        # prior draws, whose Hastings ratio evaluates every pdf and whose draws go through every invcdf
        ("mixed_polar_copolar_log", Spec("gauss", 4, 5, centers=[np.pi / 2, 0.0, 2.0, 1.0], halfwidths=[np.pi / 2, np.pi / 2, 3.0, 2.0], prior="mixed",
                                         prior_types=[3, 4, 5, 1], seed=0.37, Tmax=100, extra=dict(sigma=0.6, x0=np.array([1.2, 0.3, 1.5, 0.5]))), 800, 2),
        ("mixed_types_prior_draw", Spec("gauss", 4, 4, centers=[1.6, 0.1, 3.0, 0.0], halfwidths=[1.2, 1.0, 2.5, 1.5], prior="mixed", prop="prior",
                                        prior_types=[3, 4, 5, 2], seed=0.29, Tmax=30, extra=dict(sigma=0.8, x0=np.array([1.5, 0.2, 2.0, 0.3]))), 800, 1),
        # temperature-dependent shares (proposal_distribution.cc:40-54,72-79): hot rungs draw from the prior, as prior_draw_Tpow sets it up
        ("Tpow_hot_prior_draws", Spec("gauss", 2, 6, centers=[2, -3], halfwidths=[2, 3], prop="prior", seed=0.53, Tmax=1e3,
                                      extra=dict(Tpow=0.5, prior_draw_frac=0.1)), 1000, 2),
        ("Tpow2_custom_hot_shares", Spec("sines", 2, 5, prop="prior", seed=0.61, extra=dict(Tpow=2.0, prior_draw_frac=0.2, hot_de=0.3, hot_prior=0.7)), 800, 1),
        # shell likelihoods of example.cc (2-D reflected shell :147-222; d-dim shell pair :226-421, also in ln p0 with a log-type prior factor)
        ("shell2d", Spec("shell2d", 2, 8, centers=[0, 0], halfwidths=[6, 6], seed=0.47, Tmax=100), 800, 2),
        ("shells_d3_pair", Spec("shells", 3, 6, centers=[0, 0, 0], halfwidths=[6, 6, 6], seed=0.43, Tmax=100, extra=dict(shell_spm=1.5)), 600, 1),
        ("shells_d2_logx", Spec("shells", 2, 6, centers=[np.exp(0.001) * np.sqrt(np.exp(5.998)), 0.0], halfwidths=[np.sqrt(np.exp(5.998)), 6.0], prior="mixed",
                                prior_types=[5, 1], seed=0.59, Tmax=100, extra=dict(shell_logx=1)), 600, 1),
        # adaptive shares of the set (proposal_distribution.cc:132-166): every rung's clone adapts its own shares; with Tpow > 0 the
        # rebuilt bins follow the rung's CURRENT temperature (evolving here)
        ("adaptive_shares_default", Spec("gauss", 2, 6, centers=[2, -3], halfwidths=[2, 3], seed=0.67, Tmax=100, extra=dict(adapt_rate=0.05)), 1200, 2),
        ("adaptive_shares_Tpow_evolve", Spec("sines", 2, 6, prop="prior", seed=0.73, evolve_rate=0.01,
                                             extra=dict(Tpow=1.5, prior_draw_frac=0.2, hot_de=0.4, hot_prior=0.6, adapt_rate=0.2)), 1000, 1),
        # what ptmcmc_sampler builds for --prop_adapt_rate (ptmcmc.cc:70-72,123-143; the reference's own exampleLISA test runs with it): the
        # Gaussian scales in a NESTED adaptive set, and with --prop_adapt_more the top level adaptive as well
        ("nested_adaptive_gaussians", Spec("gauss", 3, 6, centers=[2, -3, 1], halfwidths=[2, 3, 1], seed=0.101, Tmax=100, extra=dict(prop_adapt_rate=0.05)), 1000, 2),
        ("nested_adaptive_more_evolve", Spec("sines", 2, 8, seed=0.103, evolve_rate=0.01, extra=dict(prop_adapt_rate=0.01, prop_adapt_more=1)), 1200, 1),
        # the wrapped Gaussian prior (probability_function.cc:57-78): images of the wrapped dimensions are added to the pdf
        ("gaussian_wrap_prior", Spec("gauss", 3, 5, centers=[0.5, -1.0, 2.0], halfwidths=[0.8, 1.5, 0.6], prior="gaussian_wrap", bound="wow",
                                     seed=0.79, Tmax=50, extra=dict(sigma=2.5)), 800, 2),
        # temperature mixing of the history draws (proposal_distribution.cc:594-741): a bare differential_evolution with
        # support_mixing(true) on the ladder; and the same flag inside the default set, where the reference never mixes (chain.cc:1375)
        ("de_mixing_bare", Spec("gauss", 2, 5, centers=[2, -3], halfwidths=[2, 3], prop="de", seed=0.89, Tmax=100, de_ni=20,
                                extra=dict(de_mixing=1, de_Tmix=3.0)), 400, 2),
        ("de_mixing_sines_R8", Spec("sines", 2, 8, prop="de", seed=0.97, de_ni=15, swap_rate=0.3, extra=dict(de_mixing=1, de_Tmix=300.0, de_snooker=0.3)), 300, 1),
        ("de_mixing_inert_in_set", Spec("gauss", 2, 4, centers=[2, -3], halfwidths=[2, 3], seed=0.93, Tmax=100, extra=dict(de_mixing=1, de_Tmix=300.0)), 400, 1),
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
        # the data chi^2 likelihoods and the prior-draw member above 16 dimensions (exact-summation-order kernel)
        ("W_poly_d18", Spec("poly", 18, 4, centers=np.zeros(18), halfwidths=np.full(18, 5.0), prop="de", Tmax=1e4, de_ni=12, seed=0.35,
                            extra=wide_poly_data()), 120, 1),
        ("W_sinusoid_d18", wide_sinusoid_spec(4), 100, 1),
        ("W_prior_draw_d20_mixed", Spec("gauss", 20, 4, centers=np.linspace(-1, 1, 20), halfwidths=np.full(20, 2.0), prior="mixed", prop="prior",
                                        prior_types=[1, 2] * 10, Tmax=100, de_ni=12, seed=0.71, extra=dict(sigma=1.2)), 200, 1),
    ]


def wide_poly_data(n=203, d=18, seed=11):
    """a degree-17 polynomial on [-1, 1] (203 points: a ragged last group of the 32-point rounds)"""
    rng = np.random.default_rng(seed)
    xs = np.linspace(-1, 1, n)
    truth = rng.uniform(-3, 3, d)
    ys = sum(truth[j] * xs ** j for j in range(d)) + rng.normal(size=n)
    return dict(data_x=xs, data_y=ys, data_dy=np.full(n, 0.8))


def wide_sinusoid_spec(rungs, n=150, dt=0.02):
    """six sinusoids (d = 18) on a short series"""
    rng = np.random.default_rng(13)
    t = np.arange(n) * dt
    A, f, ph = rng.uniform(0.3, 1.5, 6), rng.uniform(0.5, 8, 6), rng.uniform(0, 2 * np.pi, 6)
    y = sum(A[k] * np.sin(2 * np.pi * f[k] * t + ph[k]) for k in range(6)) + rng.normal(size=n)
    c = np.array([1, 5, np.pi] * 6, dtype=float)
    return Spec("sinusoid", 18, rungs, centers=c, halfwidths=c.copy(), bound="oow" * 6, de_ni=12, seed=0.19,
                extra=dict(data_x=t, data_y=y, data_dy=np.ones(n)))
