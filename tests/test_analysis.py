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
