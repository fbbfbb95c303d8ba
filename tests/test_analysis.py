"""Closed-form checks of the host analysis helpers (cython/testKL.py:32-34 does the same for the reference's estimators)."""
import numpy as np
from ptmcmc_b200.analysis import knn_kl, integrated_act, ess_per_sample


def test_knn_kl_matches_closed_form_for_two_gaussians():
    rng = np.random.default_rng(0)
    d, n = 2, 40000
    p = rng.normal(size=(n, d)); q = rng.normal(size=(n, d)) * 1.5 + 0.5
    exact = 0.5 * (d / 2.25 + d * 0.25 / 2.25 - d + d * np.log(2.25))   # KL(N(0,I) || N(0.5, 1.5^2 I))
    assert abs(knn_kl(p, q) - exact) < 0.03
    assert abs(knn_kl(p, rng.normal(size=(n, d)))) < 0.01                 # same distribution: ~0


def test_integrated_act_of_ar1():
    rng = np.random.default_rng(1)
    phi, n, nch = 0.9, 20000, 16
    x = np.zeros((nch, n))
    e = rng.normal(size=(nch, n))
    for t in range(1, n):
        x[:, t] = phi * x[:, t - 1] + e[:, t]
    tau = integrated_act(x)
    assert abs(tau - (1 + phi) / (1 - phi)) / 19.0 < 0.1
    eps, taus = ess_per_sample(np.stack([x, x], axis=2))
    assert abs(eps - 1 / tau) < 1e-12 and len(taus) == 2


def test_reference_ess_recipe_matches_golden():
    """analysis.report_effective_samples restates chain::report_effective_samples (chain.cc:457-643); the fixtures are the
    (ess, length) pairs the reference build itself returned for these cold-chain histories (tests/golden/make_golden.py ess)."""
    import os
    from ptmcmc_b200.analysis import report_effective_samples
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_ess.npz"))
    n = len([k for k in g.files if k.startswith("x")])
    assert n >= 4
    for i in range(n):
        nhist, ninit, se, esslimit, ess_ref, len_ref = g["meta%d" % i]
        ess, length = report_effective_samples(g["x%d" % i], int(nhist), n_init=int(ninit), add_every=int(se), width=int(se) * 1000,
                                               every=int(se), esslimit=esslimit)
        assert length == int(len_ref)
        assert abs(ess - ess_ref) <= 1e-10 * ess_ref, (i, ess, ess_ref)


def test_reference_ess_recipe_on_ar1():
    """AR(1) with coefficient a has autocorrelation time (1+a)/(1-a): the recipe's ESS/length lands on it."""
    from ptmcmc_b200.analysis import report_effective_samples
    rng = np.random.default_rng(5)
    a, n = 0.9, 400000
    e = rng.normal(size=n)
    x = np.empty(n); x[0] = e[0]
    for i in range(1, n):
        x[i] = a * x[i - 1] + e[i]
    ess, length = report_effective_samples(x, n, width=1000, every=1)
    assert length > n // 2
    assert abs(length / ess / ((1 + a) / (1 - a)) - 1) < 0.15


def test_batched_recipe_equals_per_chain_recipe():
    """effective_samples_from_windows (the combination used with the device lag statistics) against report_effective_samples chain by chain"""
    from ptmcmc_b200.analysis import report_effective_samples, autocovar_windows, effective_samples_from_windows, recipe_geometry
    rng = np.random.default_rng(11)
    N, se = 24000, 2
    chains = []
    for a in (0.3, 0.9, 0.98, 0.0):
        e = rng.normal(size=(N // se + 50, 2))
        x = np.empty_like(e); x[0] = e[0]
        for i in range(1, len(x)):
            x[i] = a * x[i - 1] + e[i]
        x[:, 1] = np.abs(x[:, 1])
        chains.append(x)
    width, swidth, n_win, lags = recipe_geometry(N, se * 1000, se)
    assert n_win >= 3 and lags[0] == 0 and lags[1] == se
    cm = [autocovar_windows(x, N, 50, se, width, se, 2, 0, 1.1) for x in chains]
    ess, nwin = effective_samples_from_windows(np.stack([c[0] for c in cm]), np.stack([c[1] for c in cm]), lags, width, se, swidth)
    for i, x in enumerate(chains):
        e1, l1 = report_effective_samples(x, N, n_init=50, add_every=se, width=se * 1000, every=se)
        assert l1 == nwin[i] * width and abs(e1 - ess[i]) <= 1e-12 * e1
