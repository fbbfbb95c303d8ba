"""BASELINE.json correctness bar (2): posteriors of the CUDA engine (Philox draws) match the reference run statistically --
KL divergence between cold-chain samples < 0.01 nats (k-NN estimator, the method of the reference's testKL /
test_proposal::KL_divergence) and ESS per sample within 10 % -- against the reference algorithm driven by the reference's
own RNG (the oracle, which reproduces the unmodified reference bit for bit, tests/test_oracle_ref.py)."""
import numpy as np
import pytest
from ptmcmc_b200 import _capi as K
from ptmcmc_b200.analysis import knn_kl, ess_per_sample
from tests.models import Spec

pytestmark = pytest.mark.gpu

BURN, STEPS, THIN = 1000, 4000, 20
CASES = [
    ("A_gauss2d", Spec("gauss", 2, 8, centers=[2, -3], halfwidths=[2, 3], seed=0.224)),
    ("sines_2x2", Spec("sines", 2, 8, seed=0.1234)),
]


def cold_chains(api, spec, L, steps):
    """[L, steps, d] cold-chain states at PT iterations BURN..BURN+steps (one per iteration: the newest append of each step)"""
    d, R = spec.dim, spec.rungs
    out = np.empty((L, steps, d))
    cnt = api.get_counters()
    for l in range(L):
        n = int(cnt["nsize"][l * R])
        h = api.get_history(l, 0, n - steps, steps, full=False)   # the cold chain may hold a few extra swap appends; the newest `steps`
        out[l] = h["x"]
    return out


def reference_run(oracle_cls, spec, L):
    o = oracle_cls(spec.config(n_ladders=L, rng_mode=2))
    spec.setup(o); o.seed_newran(spec.seed); o.init_from_prior(); o.step(BURN + STEPS)
    return cold_chains(o, spec, L, STEPS)


def engine_run(engine_cls, spec, L, swap_mode=K.SWAP_REFERENCE, seed=0xB2000003):
    e = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=spec.de_ni * spec.dim + 2 * (BURN + STEPS) + 64,
                               swap_mode=swap_mode, seed=seed, record_level=K.RECORD_BASIC))
    spec.setup(e); e.init_from_prior(); e.step(BURN + STEPS); e.synchronize()
    return cold_chains(e, spec, L, STEPS)


@pytest.mark.parametrize("name,spec", CASES, ids=[c[0] for c in CASES])
def test_posterior_kl_and_ess_match_reference(name, spec, oracle_cls, engine_cls):
    L = 256
    ref = reference_run(oracle_cls, spec, L)
    eng = engine_run(engine_cls, spec, L)
    eo = engine_run(engine_cls, spec, L, swap_mode=K.SWAP_EVEN_ODD, seed=0xB2000004)
    thin = lambda c: c[:, ::THIN, :].reshape(-1, spec.dim)
    p_ref, p_eng, p_eo = thin(ref), thin(eng), thin(eo)
    half = len(p_ref) // 2
    floor = abs(knn_kl(p_ref[:half], p_ref[half:]))              # estimator noise between two halves of the reference itself
    kl1, kl2 = knn_kl(p_eng, p_ref), knn_kl(p_ref, p_eng)
    kl_eo = knn_kl(p_eo, p_ref)
    print("%s: KL(engine||ref)=%.4f KL(ref||engine)=%.4f KL(even/odd||ref)=%.4f noise floor=%.4f (n=%d)" % (name, kl1, kl2, kl_eo, floor, len(p_ref)))
    assert abs(kl1) < 0.01 and abs(kl2) < 0.01, (kl1, kl2, floor)
    assert abs(kl_eo) < 0.01, kl_eo
    eps_ref, tau_ref = ess_per_sample(ref)
    eps_eng, tau_eng = ess_per_sample(eng)
    print("%s: ESS/sample reference %.4f (tau %s) engine %.4f (tau %s)" % (name, eps_ref, np.round(tau_ref, 1), eps_eng, np.round(tau_eng, 1)))
    assert abs(eps_eng / eps_ref - 1) < 0.10
    # first two moments as a plain cross-check
    assert np.allclose(p_eng.mean(axis=0), p_ref.mean(axis=0), atol=4 * p_ref.std(axis=0).max() / np.sqrt(len(p_ref) / 4))
