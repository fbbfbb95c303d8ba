"""Known-answer values that pin the oracle's likelihood / prior / boundary arithmetic."""
import numpy as np
import pytest
from ptmcmc_b200 import _capi as K
from tests.models import Spec, poly_data, fullcov_spec
from tests.oracle_binding import Oracle


def mk(spec):
    o = Oracle(spec.config(n_ladders=1))
    spec.setup(o)
    return o


def test_sines_peak_values():
    """sines.hh:26-37,55-60: at peak centres x_i = min_i + (idx_i + 1/2)/k_i  log L = -step_scale * sum idx exactly;
    at x = min log L = -d * height"""
    d, k = 3, 2
    o = mk(Spec("sines", d, 4))
    pts, want = [], []
    for idx in np.ndindex(*(k,) * d):
        pts.append([(i + 0.5) / k for i in idx]); want.append(-np.log(2.0) * sum(idx))
    got = o.eval_loglike(np.array(pts))
    assert np.allclose(got, want, rtol=0, atol=1e-13)
    assert o.eval_loglike(np.zeros((1, d)))[0] == -d * 64.0
    # peak weights P ~ exp(-sum idx ln 2): {4/9,2/9,2/9,1/9} for a 2x2 surface (testMH.cpp:20,186,195)
    w = np.exp(-np.log(2.0) * np.array([0, 1, 1, 2])); w /= w.sum()
    assert np.allclose(w, [4 / 9, 2 / 9, 2 / 9, 1 / 9])


def test_gauss3d_max_and_prior_norm():
    """example.cc:102-105: max log L = -1.5 ln(pi * 0.5) for sigma = 0.5; uniform prior log density = -ln(8*2*3*5)"""
    sp = Spec("gauss", 3, 4, centers=[2, -3, 5], halfwidths=[2, 3, 5])
    o = mk(sp)
    assert o.eval_loglike(np.array([[2.0, -3.0, 5.0]]))[0] == pytest.approx(-1.5 * np.log(np.pi * 0.5), rel=1e-15)
    assert o.eval_logprior(np.array([[2.0, -3.0, 5.0]]))[0] == pytest.approx(-np.log(8 * 2 * 3 * 5), rel=1e-15)
    # closed at both ends (ProbabilityDist.h:88-93), zero outside
    assert np.isfinite(o.eval_logprior(np.array([[0.0, -6.0, 0.0], [4.0, 0.0, 10.0]]))).all()
    assert o.eval_logprior(np.array([[4.0000001, 0.0, 0.0]]))[0] == -np.inf


def test_gaussian_prior_density():
    sp = Spec("flat", 2, 1, centers=[1.0, -2.0], halfwidths=[0.5, 3.0], prior="gaussian", prop="gauss")
    o = mk(sp)
    x = np.array([[1.3, -4.0]])
    want = sum(-0.5 * ((x[0, i] - c) / s) ** 2 - np.log(s * np.sqrt(2 * np.pi)) for i, (c, s) in enumerate([(1.0, 0.5), (-2.0, 3.0)]))
    assert o.eval_logprior(x)[0] == pytest.approx(want, rel=1e-14)


def test_poly_chi2_closed_form():
    """bayesian.hh:595-622 with poly_example.cc:85-106: -1/2 sum[(m-y)^2/S + log S] - like0"""
    e = poly_data()
    sp = Spec("poly", 5, 2, centers=np.zeros(5), halfwidths=np.full(5, 10.0), prop="de", extra=e)
    o = mk(sp)
    c = np.array([[0.3, -1.0, 0.25, 0.01, -0.002]])
    m = sum(c[0, j] * e["data_x"] ** j for j in range(5))
    want = -0.5 * np.sum((m - e["data_y"]) ** 2)
    assert o.eval_loglike(c)[0] == pytest.approx(want, rel=1e-12)


def test_fullcov_quadratic_form():
    sp = fullcov_spec(6, 2)
    o = mk(sp)
    x = np.random.default_rng(1).normal(size=(5, 6))
    cinv = sp.extra["cinv"].reshape(6, 6)
    want = sp.extra["like0"] - 0.5 * np.einsum("ni,ij,nj->n", x, cinv, x)
    assert np.allclose(o.eval_loglike(x), want, rtol=1e-13)


@pytest.mark.parametrize("bound,x,want", [
    ("w", 4.5, 0.5), ("w", -0.25, 3.75), ("w", 4.0, 0.0),   # wrap is half-open [xmin,xmax) via fmod (states.cc:24-27)
    ("r", 3.0, 3.0),
    # reference quirk kept: double reflection folds with `xt=halfwidth-xt` (states.cc:40), which leaves the box
    ("r", 4.5, -0.5), ("r", -0.5, -3.5),
])
def test_boundary_enforce(bound, x, want):
    """boundary::enforce (states.cc:11-58) on the box [0,4], seen through a narrow Gaussian prior centred at `want`"""
    t = {"w": K.BOUND_WRAP, "r": K.BOUND_REFLECT}[bound]
    sp = Spec("flat", 1, 1, prop="gauss")
    o = Oracle(sp.config(n_ladders=1))
    o.set_space([t], [t], [0.0], [4.0])
    o.set_prior([K.PRIOR_GAUSSIAN], [want], [1e-3])
    peak = -np.log(1e-3 * np.sqrt(2 * np.pi))
    assert o.eval_logprior(np.array([[x]]))[0] == pytest.approx(peak, abs=1e-9)


def test_limit_bound_rejects():
    sp = Spec("flat", 1, 1, centers=[2.0], halfwidths=[2.0], bound="l", prop="gauss")
    o = Oracle(sp.config(n_ladders=1)); sp.setup(o)
    assert o.eval_logprior(np.array([[4.5]]))[0] == -np.inf   # invalid state -> evaluate = 0
    assert np.isfinite(o.eval_logprior(np.array([[3.5]]))[0])


def test_geometric_ladder():
    """chain.cc:1181-1183,1339: temps[i] = temps[i-1]*Tmax^(1/(N-1)), beta = 1/temp"""
    sp = Spec("flat", 1, 8, Tmax=1e6, prop="gauss")
    o = Oracle(sp.config(n_ladders=2)); sp.setup(o); o.init_from_prior()
    b = o.get_current()["beta"].reshape(2, 8)
    assert b[0, 0] == 1.0 and b[0, -1] == pytest.approx(1e-6, rel=1e-12)
    assert np.allclose(b[0, :-1] / b[0, 1:], 1e6 ** (1 / 7), rtol=1e-12)
    assert (b[0] == b[1]).all()
