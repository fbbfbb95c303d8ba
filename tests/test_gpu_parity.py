"""Parity tests proper: the CUDA engine (through the C ABI) against the CPU oracle on the same seeded inputs.

Tape parity = BASELINE.json's correctness bar (1): given identical injected uniform / normal draws, accept / reject and
swap decisions, proposal types, history indices and positions are bit-exact with the reference (whose RNG the oracle
restates and whose histories the oracle reproduces bit for bit, tests/test_oracle_ref.py); log-likelihoods and
log-posteriors agree to 1e-12 relative."""
import os
import numpy as np
import pytest
from ptmcmc_b200 import _capi as K
from tests.models import Spec, parity_cases, poly_data, sinusoid_spec, fullcov_spec, engine_dump, compare_dumps
from tests.parity import tape_parity, philox_parity, RTOL

pytestmark = pytest.mark.gpu
CASES = parity_cases()
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,spec,steps,L", CASES, ids=[c[0] for c in CASES])
def test_tape_parity(name, spec, steps, L, oracle_cls, engine_cls):
    assert tape_parity(oracle_cls, engine_cls, spec, steps, L, what=name) == []


@pytest.mark.parametrize("name,spec,steps,L", CASES, ids=[c[0] for c in CASES])
def test_philox_consistency(name, spec, steps, L, oracle_cls, engine_cls):
    assert philox_parity(oracle_cls, engine_cls, spec, min(steps, 400), L, what=name) == []


@pytest.mark.parametrize("name", ["A_gauss2d_default", "C1_sines_d3_R32", "sines_evolve_cut", "B_poly", "D_fullcov_d6", "W_D_fullcov_d100_R8",
                                  "W_D_fullcov_d100_R24_cov1d"])
def test_engine_matches_golden_reference_histories(name, oracle_cls, engine_cls):
    """engine (tape mode, draws recorded from the oracle's reference-RNG run) vs the COMMITTED histories of the unmodified
    reference: cold-chain positions bit-exact, lpost / llike to 1e-12"""
    from tests.parity import record_reference_run
    case = [c for c in CASES if c[0] == name][0]
    _, spec, steps, _L = case
    g = np.load(os.path.join(GOLDEN, "ref_%s.npz" % name))
    if spec.prop in ("cov", "covde"):
        d = spec.dim
        spec.eig = (g["eig"][:d].copy(), g["eig"][d:].reshape(d, d).copy())
    cap = spec.de_ni * spec.dim + 2 * steps + 8
    o, tapes, marks = record_reference_run(oracle_cls, spec, steps, 1, hist_capacity=cap)
    e = engine_cls(spec.config(n_ladders=1, rng_mode=K.RNG_TAPE, hist_capacity=cap))
    spec.setup(e); e.inject_tapes(*tapes); e.inject_tape_marks(*marks); e.init_from_prior(); e.step(steps); e.synchronize()
    n = int(e.get_counters()["nsize"][0])
    assert n == int(g["counters"][0, 0])
    f = int(g["cold_from"])
    h = e.get_history(0, 0, f, n - f)
    assert h["x"].tobytes() == g["cold_x"].tobytes()
    assert (h["type"] == g["cold_type"]).all()
    assert h["acc"].tobytes() == g["cold_acc"].tobytes()
    assert np.allclose(h["llike"], g["cold_llike"], rtol=RTOL, atol=0)
    assert np.allclose(h["lpost"], g["cold_lpost"], rtol=RTOL, atol=0)
    cnt = e.get_counters()
    got = np.stack([cnt[k] for k in ("nsize", "nhist", "ntries", "naccept", "last_type")], axis=1)
    assert (got == g["counters"]).all()
    sw = e.get_swap_stats()
    assert (sw["swap_count"][0] == g["swap_count"]).all() and (sw["swap_accept"][0] == g["swap_accept"]).all()


@pytest.mark.parametrize("name", ["A_gauss2d_default", "sines_evolve_cut", "C1_sines_d3_R32", "unlikely_alpha"])
def test_tape_parity_first_generation_kernel(name, oracle_cls, engine_cls):
    """the shared-memory kernel (used for ladders of more than 32 rungs) forced onto small ladders"""
    class Gen1(engine_cls):
        def __init__(self, cfg):
            super().__init__(cfg)
            self.select_kernel(K.KERNEL_SHARED)
    _, spec, steps, L = [c for c in CASES if c[0] == name][0]
    assert tape_parity(oracle_cls, Gen1, spec, steps, L, what=name) == []


def max_size_cases():
    """the limits of every kernel family: dim 16 (thread-per-chain), dim 128 x 32 rungs (warp-per-chain, 1024-thread CTAs),
    64 rungs (shared-memory kernel, two warps per ladder), 1 rung x dim 1"""
    return [
        ("max_dim16_R8", fullcov_spec(16, 8, Tmax=1e3, de_ni=12), 150, 2),
        ("max_dim128_R32", fullcov_spec(128, 32, Tmax=1e4, de_ni=11, swap_rate=0.05), 40, 1),
        ("max_R64", Spec("gauss", 2, 64, centers=[2, -3], halfwidths=[2, 3], seed=0.55, swap_rate=0.05), 200, 1),
        ("max_swaps_R40", Spec("gauss", 2, 40, centers=[2, -3], halfwidths=[2, 3], seed=0.58, swap_rate=0.5), 150, 1),   # 41 trials per step
        ("min_dim1_R1", Spec("gauss", 1, 1, centers=[0.5], halfwidths=[2.0], prop="gauss", seed=0.66), 500, 3),
    ]


@pytest.mark.parametrize("name,spec,steps,L", max_size_cases(), ids=[c[0] for c in max_size_cases()])
def test_parity_at_size_limits(name, spec, steps, L, oracle_cls, engine_cls):
    assert tape_parity(oracle_cls, engine_cls, spec, steps, L, what=name) == []
    assert philox_parity(oracle_cls, engine_cls, spec, steps, L, what=name) == []


WIDE = [c for c in CASES if c[0].startswith("W_")]


@pytest.mark.parametrize("name,spec,steps,L", WIDE, ids=[c[0] for c in WIDE])
def test_philox_consistency_wide_exact_kernel(name, spec, steps, L, oracle_cls, engine_cls):
    """dim > 16, Philox draws through the exact-summation-order kernel (test_philox_consistency runs the DMMA-batched one)"""
    class Exact(engine_cls):
        def __init__(self, cfg):
            super().__init__(cfg)
            self.select_kernel(K.KERNEL_WARP)
    assert philox_parity(oracle_cls, Exact, spec, steps, L, what=name) == []


def test_tape_parity_with_wrapping_ring(oracle_cls, engine_cls):
    """history ring smaller than the run (H3): DE draws from the newest `capacity` samples; indices must still agree"""
    spec = Spec("gauss", 2, 6, centers=[2, -3], halfwidths=[2, 3])
    o_cfg = dict(hist_capacity=160)
    from tests.parity import record_reference_run, compare_runs
    steps, L = 900, 2
    o, tapes, marks = record_reference_run(oracle_cls, spec, steps, L, **o_cfg)
    g = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_TAPE, trace_steps=steps, hist_capacity=160))
    spec.setup(g); g.inject_tapes(*tapes); g.inject_tape_marks(*marks); g.init_from_prior(); g.step(steps); g.synchronize()
    lo, co = o.get_trace(0, steps); lg, cg = g.get_trace(0, steps)
    assert (co == cg).all()
    co_, cg_ = o.get_current(), g.get_current()
    assert co_["x"].tobytes() == cg_["x"].tobytes()
    cnt_o, cnt_g = o.get_counters(), g.get_counters()
    for k in ("nhist", "nsize", "ntries", "naccept", "last_type"):
        assert (cnt_o[k] == cnt_g[k]).all()
    # the newest 160 samples of every chain are in the ring and equal the oracle's
    for l in range(L):
        for r in range(6):
            n = int(cnt_g["nsize"][l * 6 + r])
            ho = o.get_history(l, r, n - 160, 160); hg = g.get_history(l, r, n - 160, 160)
            assert ho["x"].tobytes() == hg["x"].tobytes()
            assert np.allclose(ho["lpost"], hg["lpost"], rtol=RTOL, atol=0)
    with pytest.raises(K.CApiError, match="overwritten"):
        g.get_history(0, 0, 0, 10)


def eval_cases():
    rng = np.random.default_rng(11)
    out = []
    sp = Spec("sines", 3, 2); out.append(("sines", sp, rng.uniform(-0.2, 1.2, (4000, 3))))
    sp = Spec("gauss", 3, 2, centers=[2, -3, 5], halfwidths=[2, 3, 5]); out.append(("gauss", sp, rng.normal(size=(4000, 3)) * 3))
    sp = Spec("poly", 5, 2, centers=np.zeros(5), halfwidths=np.full(5, 10.0), prop="de", extra=poly_data())
    out.append(("poly", sp, rng.uniform(-10, 10, (500, 5))))
    sp = sinusoid_spec(2, n=10000)
    c = np.array([1, 5, np.pi] * 3)
    out.append(("sinusoid_N1e4", sp, rng.uniform(0, 2, (64, 9)) * c))
    sp = fullcov_spec(16, 2); out.append(("fullcov16", sp, rng.normal(size=(2000, 16)) * 5))
    sp = fullcov_spec(100, 2); out.append(("fullcov100_wide", sp, rng.normal(size=(500, 100)) * 5))
    sp = Spec("gauss", 40, 2, centers=np.zeros(40), halfwidths=np.full(40, 2.0), prior="gaussian", bound="r"); out.append(("gauss40_wide_gaussprior_reflect", sp, rng.normal(size=(500, 40)) * 3))
    sp = Spec("flat", 4, 2, centers=[0, 1, 2, 3], halfwidths=[1, 2, 3, 4], prior="mixed", prior_types=[1, 2, 1, 2], prop="gauss", bound="owrl")
    out.append(("mixed_prior_bounds", sp, rng.normal(size=(4000, 4)) * 4 + 1))
    # the remaining mixed_dist_product factor types, inside and outside their supports (ProbabilityDist.h:107-139,171-232)
    sp = Spec("flat", 4, 2, centers=[1.6, 0.1, 3.0, 0.0], halfwidths=[1.2, 1.0, 2.5, 1.5], prior="mixed", prior_types=[3, 4, 5, 2], prop="gauss")
    out.append(("mixed_polar_copolar_log_gauss", sp, np.column_stack([rng.uniform(0.0, 3.2, 4000), rng.uniform(-1.6, 1.6, 4000), rng.uniform(0.5, 9.0, 4000), rng.normal(size=4000) * 2])))
    sp = Spec("flat", 3, 2, centers=[np.pi / 2, 0.0, 2.0], halfwidths=[np.pi / 2 + 0.3, np.pi / 2 + 0.2, 3.0], prior="mixed", prior_types=[3, 4, 5], prop="gauss")
    out.append(("polar_copolar_clamped_ranges", sp, np.column_stack([rng.uniform(-0.5, 3.6, 3000), rng.uniform(-2.0, 2.0, 3000), rng.uniform(0.3, 7.0, 3000)])))
    # shell likelihoods (example.cc:147-222, 226-421)
    sp = Spec("shell2d", 2, 2, centers=[0, 0], halfwidths=[6, 6]); out.append(("shell2d", sp, rng.uniform(-6, 6, (4000, 2))))
    sp = Spec("shells", 5, 2, centers=np.zeros(5), halfwidths=np.full(5, 6.0), extra=dict(shell_spm=2.0)); out.append(("shells_d5", sp, rng.uniform(-5, 5, (4000, 5))))
    sp = Spec("shells", 2, 2, centers=[20.0, 0.0], halfwidths=[19.0, 6.0], extra=dict(shell_logx=1)); out.append(("shells_logx", sp, np.column_stack([rng.uniform(-2, 300, 3000), rng.uniform(-6, 6, 3000)])))
    return out


@pytest.mark.parametrize("name,spec,x", eval_cases(), ids=[c[0] for c in eval_cases()])
def test_device_functors_match_oracle(name, spec, x, oracle_cls, engine_cls):
    """log-likelihoods (and log-priors after boundary enforcement) to 1e-12 relative, infinities in the same places"""
    o = oracle_cls(spec.config(n_ladders=1)); spec.setup(o)
    g = engine_cls(spec.config(n_ladders=1)); spec.setup(g)
    for fo, fg in ((o.eval_loglike, g.eval_loglike), (o.eval_logprior, g.eval_logprior)):
        a, b = fo(x), fg(x)
        assert (np.isfinite(a) == np.isfinite(b)).all()
        fin = np.isfinite(a)
        assert ((a[~fin] == b[~fin]) | (np.isnan(a[~fin]) & np.isnan(b[~fin]))).all()   # -inf outside the support; NaN where a clamped-range pdf turns negative, on both sides
        assert np.allclose(a[fin], b[fin], rtol=RTOL, atol=1e-300)
    # the sinusoid chi^2 over 1e4 samples sums ~1e4 terms of libm sin: tolerance is on the SUM (relative), as above


def test_init_states_path(oracle_cls, engine_cls):
    """ptg_init_states (chain_init_file path): caller-provided initial histories, then Philox stepping"""
    spec = Spec("gauss", 2, 4, centers=[2, -3], halfwidths=[2, 3])
    L, ninit = 3, 100
    x0 = np.random.default_rng(3).uniform([0, -6], [4, 0], (L * 4 * ninit, 2))
    runs = []
    for cls in (oracle_cls, engine_cls):
        e = cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, trace_steps=50, hist_capacity=400))
        spec.setup(e); e.init_states(x0); e.step(50); runs.append(e)
    from tests.parity import compare_runs, PHILOX_RTOL
    runs[1].synchronize()
    assert compare_runs(runs[0], runs[1], 50, L, "init_states", rtol=PHILOX_RTOL, exact_x=False) == []
