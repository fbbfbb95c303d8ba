"""Adaptive proposal shares (proposal_distribution.cc:132-166), temperature mixing of the DE history draws (:594-741) and the wrapped
Gaussian prior (probability_function.cc:57-78) on the engine.  The chains themselves are compared with the oracle by the named cases of
tests.models.parity_cases() (`adaptive_shares_*`, `de_mixing_*`, `gaussian_wrap_prior`) in test_gpu_parity.py; here: the adapted shares,
the functor, checkpoint / restore of the adapted state, and the error behaviour of ptg_set_proposal_options."""
import os
import numpy as np
import pytest
from ptmcmc_b200 import _capi as K
from tests.models import Spec, parity_cases
from tests.parity import record_reference_run, RTOL

pytestmark = pytest.mark.gpu
CASES = {c[0]: c for c in parity_cases()}


@pytest.mark.parametrize("name", ["adaptive_shares_default", "adaptive_shares_Tpow_evolve", "nested_adaptive_gaussians", "nested_adaptive_more_evolve"])
def test_adapted_shares_equal_the_reference_algorithm(name, oracle_cls, engine_cls):
    """after a tape replay every chain's shares are the ones the reference's accept / reject bookkeeping produced (bit-exact without
    Tpow; with Tpow the rebuilt bins go through pow(), the shares themselves do not)"""
    _, spec, steps, L = CASES[name]
    cap = spec.de_ni * spec.dim + 2 * steps + 8   # the ring never wraps: the reference history is unbounded
    o, tapes, marks = record_reference_run(oracle_cls, spec, steps, L, hist_capacity=cap)
    g = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_TAPE, trace_steps=steps, hist_capacity=cap))
    spec.setup(g); g.inject_tapes(*tapes); g.inject_tape_marks(*marks); g.init_from_prior(); g.step(steps); g.synchronize()
    so, sg = o.get_proposal_shares(), g.get_proposal_shares()
    assert so.shape == sg.shape
    assert so.tobytes() == sg.tobytes()
    assert np.ptp(so, axis=0).max() > 1e-3          # the rungs really adapted differently
    nn = int(getattr(g, "n_nested", 0))                # every set's shares are normalised: the top level, and the nested set behind it
    nt = so.shape[1] - nn
    assert np.allclose(so[:, :nt].sum(axis=1), 1.0, atol=1e-12)
    if nn:
        assert np.allclose(so[:, nt:].sum(axis=1), 1.0, atol=1e-12)


def test_adaptive_run_survives_checkpoint_restore(engine_cls, tmp_path):
    spec = CASES["adaptive_shares_default"][1]
    mk = lambda: engine_cls(spec.config(n_ladders=8, rng_mode=K.RNG_PHILOX, hist_capacity=600))
    a = mk(); spec.setup(a); a.init_from_prior(); a.step(300); a.synchronize()
    b = mk(); spec.setup(b); b.init_from_prior(); b.step(120)
    path = os.path.join(str(tmp_path), "adaptive.ckpt")
    b.checkpoint(path)
    c = mk(); spec.setup(c); c.restore(path); c.step(180); c.synchronize()
    assert a.get_current()["x"].tobytes() == c.get_current()["x"].tobytes()
    assert a.get_proposal_shares().tobytes() == c.get_proposal_shares().tobytes()


def test_wrapped_gaussian_prior_functor(oracle_cls, engine_cls):
    """log prior = log prod_i (pdf_i + images) on wrapped dimensions, plain Gaussian elsewhere; both wide and narrow kernels"""
    rng = np.random.default_rng(5)
    for d, bound in ((3, "wow"), (20, "wo" * 10)):
        c = rng.uniform(-1, 1, d); hw = rng.uniform(0.5, 2.0, d)
        sp = Spec("flat", d, 2, centers=c, halfwidths=hw, prior="gaussian_wrap", prop="gauss", bound=bound)
        plain = Spec("flat", d, 2, centers=c, halfwidths=hw, prior="gaussian", prop="gauss", bound=bound)
        x = c + rng.normal(size=(1500, d)) * hw * 1.5
        o = oracle_cls(sp.config(n_ladders=1)); sp.setup(o)
        g = engine_cls(sp.config(n_ladders=1)); sp.setup(g)
        p = engine_cls(plain.config(n_ladders=1)); plain.setup(p)
        a, b, q = o.eval_logprior(x), g.eval_logprior(x), p.eval_logprior(x)
        assert np.isfinite(a).all()
        assert np.allclose(a, b, rtol=RTOL, atol=1e-13)
        assert (b > q + 1e-6).all()                 # the images add probability everywhere on a wrapped dimension


def test_mixing_draws_from_other_rungs(engine_cls):
    """with temperature mixing a cold chain's proposals use hotter rungs' histories: on a well-separated bimodal target the cold chain
    of a SHORT run visits the second mode more often than without mixing (the purpose of the feature); both runs sample the same target"""
    spec_m = Spec("sines", 2, 8, prop="de", seed=0.5, de_ni=15, extra=dict(de_mixing=1, de_Tmix=300.0))
    spec_p = Spec("sines", 2, 8, prop="de", seed=0.5, de_ni=15)
    res = []
    for sp in (spec_m, spec_p):
        e = engine_cls(sp.config(n_ladders=64, rng_mode=K.RNG_PHILOX, hist_capacity=2500, record_level=K.RECORD_BASIC))
        sp.setup(e); e.init_from_prior(); e.step(2000); e.synchronize()
        cnt = e.get_counters()
        acc = (cnt["naccept"] / cnt["ntries"]).reshape(64, 8).mean(axis=0)
        res.append((e.get_current()["x"].copy(), acc))
        e.close()
    assert not np.array_equal(res[0][0], res[1][0])
    # hot-rung acceptance is unaffected in order of magnitude; every chain keeps moving
    assert (res[0][1] > 0.01).all() and (res[1][1] > 0.01).all()


def test_nested_set_errors_and_static_shares(engine_cls):
    sp = Spec("gauss", 2, 4, centers=[2, -3], halfwidths=[2, 3])
    e = engine_cls(sp.config(n_ladders=2)); sp.setup(e)
    with pytest.raises(K.CApiError, match="outside"):
        e.set_nested_set(3, 6, share=0.2)
    e.set_nested_set(1, 6, share=0.2)                # a static nested set: same selection probabilities, two selection draws
    e.init_from_prior(); e.step(50); e.synchronize()
    sh = e.get_proposal_shares()
    assert sh.shape == (8, 8) and np.allclose(sh[:, :2], [0.8, 0.2]) and np.allclose(sh[:, 2:].sum(axis=1), 1.0)
    e.close()
    sp = Spec("gauss", 2, 4, centers=[2, -3], halfwidths=[2, 3], prop="de")
    e = engine_cls(sp.config(n_ladders=1)); sp.setup(e)
    with pytest.raises(K.CApiError, match="needs a proposal set"):
        e.set_nested_set(0, 1, share=1.0)
    e.close()


def test_proposal_option_errors(engine_cls):
    sp = Spec("gauss", 2, 4, centers=[2, -3], halfwidths=[2, 3])
    e = engine_cls(sp.config(n_ladders=1)); sp.setup(e)
    with pytest.raises(K.CApiError, match="bare differential-evolution"):
        e.set_proposal_options(de_mixing=True)                         # inside a set the reference never mixes
    e.set_proposal_options(adapt_rate=0.1)
    e.close()
    sp = Spec("gauss", 2, 4, centers=[2, -3], halfwidths=[2, 3], prop="de")
    e = engine_cls(sp.config(n_ladders=1)); sp.setup(e)
    with pytest.raises(K.CApiError, match="need a proposal set"):
        e.set_proposal_options(adapt_rate=0.1)
    e.set_proposal_options(de_mixing=True, de_Tmix=10.0)
    e.close()
    sp = Spec("gauss", 2, 4, centers=[2, -3], halfwidths=[2, 3], prop="de", extra=dict(de_unlikely_alpha=0.5))
    e = engine_cls(sp.config(n_ladders=1)); sp.setup(e)
    with pytest.raises(K.CApiError, match="unlikely_alpha"):
        e.set_proposal_options(de_mixing=True)
    e.close()
    sp = Spec("gauss", 20, 4, centers=np.zeros(20), halfwidths=np.ones(20))
    e = engine_cls(sp.config(n_ladders=1)); sp.setup(e)
    with pytest.raises(K.CApiError, match="dim <= 16"):
        e.set_proposal_options(adapt_rate=0.1)
    e.close()
