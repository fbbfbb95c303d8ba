"""Size-independent properties of the CUDA path at BASELINE.json's full sizes, plus the C-ABI error behaviour."""
import os
import numpy as np
import pytest
from ptmcmc_b200 import _capi as K
from tests.models import Spec, engine_dump, compare_dumps, poly_data, sinusoid_spec, fullcov_spec

pytestmark = pytest.mark.gpu


def c1(engine_cls, L, steps, chunks=None, ladder_offset=0, cap=512, **kw):
    """config C1 (sines.hh as written, d=3, 32 rungs, default proposal mix)"""
    spec = Spec("sines", 3, 32)
    e = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=cap, ladder_offset=ladder_offset, **kw))
    spec.setup(e); e.init_from_prior()
    for n in (chunks or [steps]):
        e.step(n)
    e.synchronize()
    return e


def same_ladder(a, la, b, lb, n_last=300):
    R, ca, cb = a.cfg.n_rungs, a.get_counters(), b.get_counters()
    for r in range(R):
        ia, ib = la * R + r, lb * R + r
        for k in ("nhist", "nsize", "ntries", "naccept", "last_type"):
            assert ca[k][ia] == cb[k][ib], (k, r)
        n = int(ca["nsize"][ia])
        ha = a.get_history(la, r, n - n_last, n_last); hb = b.get_history(lb, r, n - n_last, n_last)
        for k in ("x", "lpost", "llike", "acc", "beta"):
            assert ha[k].tobytes() == hb[k].tobytes(), (k, r)
        assert (ha["type"] == hb["type"]).all()


def test_full_size_batch_is_invariant_to_sharding(engine_cls):
    """BASELINE config 3 size (4096 ladders x 32 rungs = 131072 chains): ladder g of the batch is bit-identical to the
    same ladder run alone with ladder_offset = g -- results do not depend on batch size, CTA packing or GPU count"""
    steps = 400
    big = c1(engine_cls, 4096, steps)
    assert big.get_total_steps() == big.get_counters()["nhist"].sum()
    for g in (0, 1337, 4095):
        one = c1(engine_cls, 1, steps, ladder_offset=g)
        same_ladder(big, g, one, 0)
    # a shard of 512 ladders starting at global ladder 1024 == the corresponding slice of the full batch
    shard = c1(engine_cls, 512, steps, ladder_offset=1024)
    same_ladder(big, 1024 + 77, shard, 77)


KERNEL_CASES = [
    ("c1", Spec("sines", 3, 32), {}),
    ("c1_evolve_cut", Spec("sines", 3, 32, evolve_rate=0.01, evolve_lpost_cut=0.5), {}),
    ("r8_packed_4_per_warp", Spec("sines", 3, 8, swap_rate=0.3), {}),
    ("r24_even_odd", Spec("sines", 3, 24), dict(swap_mode=K.SWAP_EVEN_ODD)),
    ("r5_ghost_lanes_full_record", Spec("gauss", 2, 5, centers=[2, -3], halfwidths=[2, 3], save_every=3), {}),
    ("bounded_unlikely_alpha", Spec("gauss", 3, 6, centers=[2, -3, 5], halfwidths=[2, 3, 5], bound="w", extra=dict(sigma=3.0, de_unlikely_alpha=0.5)), {}),
    ("gaussian_prior_prior_draw", Spec("gauss", 2, 5, centers=[2, -3], halfwidths=[2, 3], prop="prior", prior="mixed", prior_types=[1, 2]), {}),
    ("single_chain", Spec("sines", 2, 1, prop="de"), {}),
    ("de_not_ready_at_start", Spec("gauss", 2, 4, centers=[2, -3], halfwidths=[2, 3], de_ni=3, Tmax=100), {}),
    # higher dimensions, the data likelihoods of BASELINE configs B / C2 and an eigen-rotated Gaussian proposal
    ("b_poly_d5_R4", Spec("poly", 5, 4, centers=np.zeros(5), halfwidths=np.full(5, 10.0), prop="de", Tmax=1e6, extra=poly_data(n=120)), dict(L=60)),
    ("c2_sinusoid_d9_R4", sinusoid_spec(4, n=150, dt=0.05), dict(L=40)),
    ("fullcov_d16_R6_eigen_rotated", fullcov_spec(16, 6, Tmax=1e3, de_ni=12), dict(L=60)),
    ("gauss_d9_R32_save2", Spec("gauss", 9, 32, centers=np.zeros(9), halfwidths=np.full(9, 4.0), save_every=2), dict(L=40)),
    ("gauss_d5_R32_save_every3_full_record", Spec("gauss", 5, 32, centers=np.zeros(5), halfwidths=np.full(5, 4.0), save_every=3), dict(L=40)),
    ("mixed_polar_copolar_log_d4", Spec("gauss", 4, 5, centers=[np.pi / 2, 0.0, 2.0, 1.0], halfwidths=[np.pi / 2, np.pi / 2, 3.0, 2.0], prior="mixed",
                                        prior_types=[3, 4, 5, 1], Tmax=100, extra=dict(sigma=0.6, x0=np.array([1.2, 0.3, 1.5, 0.5]))), dict(L=60)),
    ("Tpow_hot_prior_draws", Spec("gauss", 2, 6, centers=[2, -3], halfwidths=[2, 3], prop="prior", Tmax=1e3, extra=dict(Tpow=0.5, prior_draw_frac=0.1)), dict(L=60)),
    ("shells_d3", Spec("shells", 3, 6, centers=[0, 0, 0], halfwidths=[6, 6, 6], Tmax=100, extra=dict(shell_spm=1.5)), dict(L=60)),
    ("dim13_R5", Spec("gauss", 13, 5, centers=np.zeros(13), halfwidths=np.full(13, 3.0), de_ni=12), dict(L=30)),
]


@pytest.mark.parametrize("name,spec,kw", KERNEL_CASES, ids=[c[0] for c in KERNEL_CASES])
def test_kernels_agree_bitwise(name, spec, kw, engine_cls):
    """Philox mode: the production kernel (FAST), the tape-capable warp kernel (WARP, bit-exact with the reference under
    injected draws) and the shared-memory kernel (SHARED) produce bit-identical chains on a 300-ladder batch"""
    kw = dict(kw)
    L = kw.pop("L", 300)

    def run(kern):
        e = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=spec.de_ni * spec.dim + 1100, **kw))
        e.select_kernel(kern)
        spec.setup(e); e.init_from_prior(); e.step(77); e.step(423); e.synchronize()
        return e
    es = [run(k) for k in (K.KERNEL_FAST, K.KERNEL_WARP, K.KERNEL_SHARED)]
    assert len({e.get_total_steps() for e in es}) == 1
    for l in (0, 7, L // 2, L - 1):
        d0 = engine_dump(es[0], l)
        for e in es[1:]:
            assert compare_dumps(d0, engine_dump(e, l), rtol=0.0, what=name) == []
    for e in es[1:]:
        assert es[0].get_current()["x"].tobytes() == e.get_current()["x"].tobytes()


STREAMLINED_CASES = [
    ("c1_sines_d3_R32", Spec("sines", 3, 32), 300),
    ("sines_d3_R8_four_ladders_per_warp", Spec("sines", 3, 8, swap_rate=0.3), 300),
    ("sines_d3_R24_ghost_lanes_save3", Spec("sines", 3, 24, save_every=3), 301),
    ("a_gauss_d2_R8", Spec("gauss", 2, 8, centers=[2, -3], halfwidths=[2, 3]), 300),
    ("gauss_d5_R16_high_swap_rate", Spec("gauss", 5, 16, centers=np.zeros(5), halfwidths=np.full(5, 4.0), swap_rate=0.45), 70),
    ("gauss_d9_R32", Spec("gauss", 9, 32, centers=np.zeros(9), halfwidths=np.full(9, 4.0)), 64),
    ("gauss_d16_R12", Spec("gauss", 16, 12, centers=np.zeros(16), halfwidths=np.full(16, 3.0), de_ni=12), 33),
    ("gauss_d11_R7_gauss_only", Spec("gauss", 11, 7, centers=np.zeros(11), halfwidths=np.full(11, 3.0), prop="gauss"), 40),
]


@pytest.mark.parametrize("name,spec,L", STREAMLINED_CASES, ids=[c[0] for c in STREAMLINED_CASES])
def test_streamlined_instantiation_is_bit_identical(name, spec, L, engine_cls):
    """The production kernel's streamlined instantiations (compile-time configuration: reference swap schedule, open space, uniform
    priors, DE + Gaussian members, one likelihood functor; pooled work rounds; deferred swap appends) against its general instantiation
    and the tape-capable warp kernel -- which is bit-exact with the reference under injected draws: same chains to the last bit"""
    def run(kern):
        e = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, record_level=K.RECORD_BASIC, hist_capacity=spec.de_ni * spec.dim + 1300))
        e.select_kernel(kern)
        spec.setup(e); e.init_from_prior(); e.step(77); e.step(423); e.step(1); e.synchronize()
        return e
    es = [run(k) for k in (K.KERNEL_FAST, K.KERNEL_FAST_GENERAL, K.KERNEL_WARP)]
    assert len({e.get_total_steps() for e in es}) == 1
    R = spec.rungs
    c0, cur0, sw0 = es[0].get_counters(), es[0].get_current(), es[0].get_swap_stats()
    for e in es[1:]:
        c, cur, sw = e.get_counters(), e.get_current(), e.get_swap_stats()
        for k in ("nhist", "nsize", "ntries", "naccept", "last_type", "map_lpost"):
            assert np.asarray(c0[k]).tobytes() == np.asarray(c[k]).tobytes(), k
        for k in ("x", "lpost", "llike", "beta"):
            assert cur0[k].tobytes() == cur[k].tobytes(), k
        for k in sw0:
            assert (np.asarray(sw0[k]) == np.asarray(sw[k])).all(), k
        for l in (0, L // 2, L - 1):
            for r in (0, R // 2, R - 1):
                n = int(c0["nsize"][l * R + r])
                h0 = es[0].get_history(l, r, 0, n, full=False); h1 = e.get_history(l, r, 0, n, full=False)
                for k in ("x", "lpost", "llike"):
                    assert h0[k].tobytes() == h1[k].tobytes(), (k, l, r)


def test_step_chunking_and_checkpoint_roundtrip(engine_cls, tmp_path):
    a = c1(engine_cls, 64, 300, evolve_rate=0.01)
    b = c1(engine_cls, 64, 300, chunks=[1, 2, 97, 200], evolve_rate=0.01)
    for l in (0, 63):
        same_ladder(a, l, b, l)
    # checkpoint after 120 steps, restore into a fresh engine, continue: equals the uninterrupted run
    spec = Spec("sines", 3, 32)
    c = c1(engine_cls, 64, 120, evolve_rate=0.01)
    path = os.path.join(str(tmp_path), "ptg.ckpt")
    c.checkpoint(path)
    d = engine_cls(spec.config(n_ladders=64, rng_mode=K.RNG_PHILOX, hist_capacity=512, evolve_rate=0.01))
    spec.setup(d); d.restore(path); d.step(180); d.synchronize()
    for l in (0, 31, 63):
        same_ladder(a, l, d, l)
    assert a.get_current()["x"].tobytes() == d.get_current()["x"].tobytes()


def test_restore_refuses_a_different_run(engine_cls, tmp_path):
    """ptg_restore continues the interrupted run or fails: another seed, save cadence or swap schedule is an error, not a silent fork"""
    spec = Spec("sines", 3, 8)
    mk = lambda **kw: engine_cls(spec.config(n_ladders=4, rng_mode=K.RNG_PHILOX, hist_capacity=400, **kw))
    a = mk(); spec.setup(a); a.init_from_prior(); a.step(50)
    path = os.path.join(str(tmp_path), "a.ckpt")
    a.checkpoint(path)
    for kw, word in ((dict(seed=12345), "seed"), (dict(save_every=2), "save_every"), (dict(swap_rate=0.2), "swap_rate"), (dict(swap_mode=K.SWAP_EVEN_ODD), "swap_mode"),
                     (dict(evolve_rate=0.01), "evolve_rate"), (dict(ladder_offset=4), "ladder_offset")):
        b = mk(**kw); spec.setup(b)
        with pytest.raises(K.CApiError, match=word):
            b.restore(path)
        b.close()
    with open(path, "r+b") as f:       # a truncated file is detected, too
        f.truncate(os.path.getsize(path) // 2)
    b = mk(); spec.setup(b)
    with pytest.raises(K.CApiError, match="truncated|mismatch"):
        b.restore(path)


def test_chi2_workload_is_invariant_to_batch_size(engine_cls):
    """BASELINE config B under automatic kernel selection: a shard of the batch produces the chains the full batch produces for the same
    global ladders -- the kernel a workload runs in must not depend on how many ladders share the GPU (2048 x 16 chains would have crossed
    the old 48 k-chain threshold)"""
    spec = Spec("poly", 5, 16, centers=np.zeros(5), halfwidths=np.full(5, 10.0), prop="de", Tmax=1e6, extra=poly_data(n=100))

    def run(L, off):
        e = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=400, ladder_offset=off, record_level=K.RECORD_BASIC))
        spec.setup(e); e.init_from_prior(); e.step(60); e.synchronize()
        return e
    big, part = run(3200, 0), run(64, 3000)       # 51 200 chains vs 1024 chains
    xb, xp = big.get_current()["x"].reshape(3200, 16, 5), part.get_current()["x"].reshape(64, 16, 5)
    assert xb[3000:3064].tobytes() == xp.tobytes()
    hb = big.get_history(3010, 0, 0, 300, full=False); hp = part.get_history(10, 0, 0, 300, full=False)
    assert hb["x"].tobytes() == hp["x"].tobytes() and hb["llike"].tobytes() == hp["llike"].tobytes()


@pytest.mark.parametrize("name", ["b_poly_N1000", "c2_sinusoid_N10000_uniform_grid", "c2_sinusoid_irregular_grid"])
def test_production_chi2_functors_match_the_reference_arithmetic(name, engine_cls, oracle_cls):
    """BASELINE configs B / C2 under automatic kernel selection run the warp-per-chain production functors (Horner on fused multiply-adds;
    sin / cos advanced by rotations on a uniform time grid, ptg_wide_mma.cuh).  Every log-likelihood the chains STORED must equal the
    reference's unfused arithmetic (the oracle's functor) at the stored position to 1e-12 relative -- BASELINE's bar for log-likelihoods"""
    from tests.models import sinusoid_data
    if name.startswith("b_poly"):
        spec = Spec("poly", 5, 16, centers=np.zeros(5), halfwidths=np.full(5, 10.0), prop="de", Tmax=1e6, extra=poly_data())
    else:
        spec = sinusoid_spec(8, n=10000)
        if "irregular" in name:
            t = np.sort(np.random.default_rng(5).uniform(0, 10.0, 3000))
            spec.extra.update(data_x=t, data_y=np.sin(2 * np.pi * 1.3 * t) + np.random.default_rng(6).normal(size=3000), data_dy=np.ones(3000))
    L = 24
    e = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=spec.de_ni * spec.dim + 400, record_level=K.RECORD_BASIC))
    spec.setup(e); e.init_from_prior(); e.step(150); e.synchronize()
    o = oracle_cls(spec.config(n_ladders=1)); spec.setup(o)
    cnt = e.get_counters()
    R = spec.rungs
    worst = 0.0
    for l, r in ((0, 0), (7, 3), (L - 1, R - 1)):
        n = int(cnt["nsize"][l * R + r])
        h = e.get_history(l, r, n - 120, 120, full=False)       # samples produced by the step kernel (not the start-up draws)
        want = o.eval_loglike(h["x"])
        fin = np.isfinite(want)
        assert (np.isfinite(h["llike"]) == fin).all()
        rel = np.abs(h["llike"][fin] - want[fin]) / np.abs(want[fin])
        worst = max(worst, float(rel.max()))
    print("%s: max relative deviation of stored log-likelihoods from the reference arithmetic: %.3g" % (name, worst))
    assert worst < 1e-12


def test_async_host_blocks_return_every_cold_sample(engine_cls):
    """ptg_step_host_begin / _wait: two blocks in flight, every cold sample of each block lands in its own host buffer"""
    spec = Spec("sines", 3, 8)
    e = engine_cls(spec.config(n_ladders=32, rng_mode=K.RNG_PHILOX, hist_capacity=2048, record_level=K.RECORD_BASIC))
    spec.setup(e); e.init_from_prior(); e.step(40)
    S = 100
    bufs = [(np.empty((32, S, 3)), np.empty((32, S)), np.empty((32, S))) for _ in range(3)]
    for b in bufs:
        e.step_host_begin(S, S, *b)
    e.step_host_wait(); e.synchronize()
    cnt = e.get_counters()
    for l in (0, 17, 31):
        n = int(cnt["nsize"][l * 8])
        h = e.get_history(l, 0, n - S, S, full=False)
        assert bufs[2][0][l].tobytes() == h["x"].tobytes() and bufs[2][1][l].tobytes() == h["lpost"].tobytes()
    # consecutive blocks are consecutive stretches of the cold chain (up to the few extra swap appends of the cold rung)
    assert not np.isnan(bufs[0][0]).any() and not (bufs[0][0][0] == bufs[1][0][0]).all()


def test_step_host_returns_newest_cold_samples(engine_cls):
    spec = Spec("sines", 3, 32)
    L, nout = 256, 16
    e = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=512, record_level=K.RECORD_BASIC))
    spec.setup(e); e.init_from_prior()
    x = np.empty((L, nout, 3)); lp = np.empty((L, nout)); ll = np.empty((L, nout))
    e.step_host(100, nout, x, lp, ll)
    cnt = e.get_counters()
    for l in (0, 100, 255):
        n = int(cnt["nsize"][l * 32])
        h = e.get_history(l, 0, n - nout, nout, full=False)
        assert h["x"].tobytes() == x[l].tobytes() and h["lpost"].tobytes() == lp[l].tobytes() and h["llike"].tobytes() == ll[l].tobytes()


def test_posterior_statistics_full_size(engine_cls):
    """sines 2x2x2 surface: peak occupation of the cold chains -> exp(-sum idx ln2) weights (testMH.cpp:186,195)"""
    e = c1(engine_cls, 2048, 1500, cap=1024)
    spec_d, nout = 3, 200
    x = np.empty((2048, nout, spec_d)); lp = np.empty((2048, nout)); ll = np.empty((2048, nout))
    e.step_host(0, nout, x, lp, ll)
    idx = (x.reshape(-1, 3) * 2).astype(int).clip(0, 1).sum(axis=1)
    p = np.bincount(idx, minlength=4) / idx.size
    w = np.array([1, 3 * 0.5, 3 * 0.25, 0.125]); w /= w.sum()
    assert np.allclose(p, w, atol=0.02), (p, w)


def test_swap_modes_agree_statistically(engine_cls):
    """even/odd performance mode vs the reference swap schedule: same cold-chain peak occupation"""
    spec = Spec("sines", 3, 16, Tmax=1e4)
    ps = []
    for mode in (K.SWAP_REFERENCE, K.SWAP_EVEN_ODD):
        e = engine_cls(spec.config(n_ladders=2048, rng_mode=K.RNG_PHILOX, hist_capacity=1024, swap_mode=mode, record_level=K.RECORD_BASIC))
        spec.setup(e); e.init_from_prior()
        nout = 200
        x = np.empty((2048, nout, 3)); lp = np.empty((2048, nout)); ll = np.empty((2048, nout))
        e.step_host(1500, nout, x, lp, ll)
        idx = (x.reshape(-1, 3) * 2).astype(int).clip(0, 1).sum(axis=1)
        ps.append(np.bincount(idx, minlength=4) / idx.size)
    assert np.allclose(ps[0], ps[1], atol=0.02), ps


def test_api_edge_cases(engine_cls):
    """zero-length steps, default / too-small ring capacity, pure-DE set before its history is ready"""
    spec = Spec("gauss", 2, 4, centers=[2, -3], halfwidths=[2, 3])
    e = engine_cls(spec.config(n_ladders=2, hist_capacity=10))   # smaller than n_init = 100: raised to n_init
    spec.setup(e); e.init_from_prior()
    e.step(0); e.synchronize()
    assert e.get_total_steps() == 0 and (e.get_counters()["nsize"] == 100).all()
    e.step(30); e.synchronize()
    cnt = e.get_counters()
    assert (cnt["nhist"] >= 30).all()
    h = e.get_history(0, 0, int(cnt["nsize"][0]) - 100, 100)     # the newest `capacity` samples are readable
    assert np.isfinite(h["x"]).all()
    # a bare differential_evolution before 10 d samples exist: the reference's set would find no ready member
    spec2 = Spec("gauss", 2, 4, centers=[2, -3], halfwidths=[2, 3], de_ni=3, extra=dict())
    e2 = engine_cls(spec2.config(n_ladders=1))
    spec2.setup(e2)
    e2.set_proposals([dict(kind=K.PROP_DE, share=1.0)], wrap_in_set=True)
    e2.init_from_prior(); e2.step(5)
    with pytest.raises(K.CApiError, match="no member ready"):
        e2.synchronize()


def test_error_behaviour(engine_cls):
    spec = Spec("gauss", 2, 4, centers=[2, -3], halfwidths=[2, 3])
    e = engine_cls(spec.config(n_ladders=1))
    with pytest.raises(K.CApiError, match="before initializing|initialis"):
        e.step(1)
    with pytest.raises(K.CApiError, match="set prior"):
        e.init_from_prior()
    spec.setup(e); e.init_from_prior()
    with pytest.raises(K.CApiError, match="already initialised"):
        e.init_from_prior()
    # tape mode without tapes / exhausted tapes
    t = engine_cls(spec.config(n_ladders=1, rng_mode=K.RNG_TAPE)); spec.setup(t)
    with pytest.raises(K.CApiError, match="inject tapes"):
        t.init_from_prior()
    ns = 5
    t.inject_tapes(np.full(10, 0.5), np.arange(ns + 1) * 2, np.zeros(10), np.arange(ns + 1) * 2)
    with pytest.raises(K.CApiError, match="tape exhausted"):
        t.init_from_prior()
    with pytest.raises(K.CApiError, match="dim out of range"):
        engine_cls(K.make_config(1, 4, 129))
    for d in (11, 13, 14, 15):     # no holes in the dimension range (states.hh is any d)
        engine_cls(K.make_config(1, 4, d)).close()
    with pytest.raises(K.CApiError, match="dim = 2"):
        s2 = Spec("shell2d", 3, 2, centers=[0, 0, 0], halfwidths=[6, 6, 6]); e3 = engine_cls(s2.config(n_ladders=1)); s2.setup(e3)


def test_set_current_roundtrip(engine_cls):
    spec = Spec("gauss", 2, 4, centers=[2, -3], halfwidths=[2, 3])
    e = engine_cls(spec.config(n_ladders=8)); spec.setup(e); e.init_from_prior(); e.step(10)
    cur = e.get_current(); lpr = e.get_lprior()
    x2 = cur["x"][::-1].copy()
    e.set_current(x2, cur["lpost"][::-1].copy(), cur["llike"][::-1].copy(), lpr[::-1].copy())
    e.synchronize()
    got = e.get_current()
    assert got["x"].tobytes() == x2.tobytes() and got["llike"].tobytes() == cur["llike"][::-1].tobytes()


def test_device_evidence_matches_reference_formula_and_analytic_value(engine_cls):
    """thermodynamic-integration evidence from the device history ring: equals the reference's formula (chain.cc:1582-1600,
    1984-2012) applied on the host to the same histories, and approaches the analytic ln Z = -ln(prior volume) of a normalised
    Gaussian likelihood inside a uniform box (example.cc:102-105: -ln(8*2*3*5) in 3-D; here 2-D, volume 4 x 6)"""
    spec = Spec("gauss", 2, 24, centers=[2, -3], halfwidths=[2, 3], Tmax=1e6)
    L, steps, n_last = 128, 5000, 3000
    e = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=8192, record_level=K.RECORD_BASIC))
    spec.setup(e); e.init_from_prior(); e.step(steps); e.synchronize()
    ev = e.get_log_evidence(n_last)
    ml = e.get_mean_loglike(n_last).reshape(L, 24)
    beta = e.get_current()["beta"].reshape(L, 24)
    cnt = e.get_counters()
    # host restatement on ladder 0 from its histories
    host_ml = []
    for r in range(24):
        n = int(cnt["nsize"][r]); host_ml.append(e.get_history(0, r, n - n_last, n_last, full=False)["llike"].mean())
    host_ml = np.array(host_ml)
    assert np.allclose(ml[0], host_ml, rtol=1e-12)
    b = beta[0]
    up = host_ml[1:] * (b[:-1] - b[1:]); down = -(host_ml[:-1] * (b[1:] - b[:-1]))
    want = ((up + down) / 2).sum() + (up[-1] + down[-1]) / 2 / (b[-2] / b[-1] - 1)
    assert ev[0] == pytest.approx(want, rel=1e-12)
    assert abs(ev.mean() - (-np.log(24.0))) < 0.15, ev.mean()


def test_device_act_matches_host_estimator(engine_cls):
    """ptg_get_act (one CTA per ladder on the device) == the same Sokal-windowed estimator applied on the host to the same history"""
    from ptmcmc_b200.analysis import integrated_act
    spec = Spec("gauss", 2, 8, centers=[2, -3], halfwidths=[2, 3])
    L, n_last, max_lag = 64, 3000, 600
    e = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=4096, record_level=K.RECORD_BASIC))
    spec.setup(e); e.init_from_prior(); e.step(5000); e.synchronize()
    tau = e.get_act(0, n_last, max_lag)
    cnt = e.get_counters()
    for l in (0, 17, 63):
        n = int(cnt["nsize"][l * 8])
        x = e.get_history(l, 0, n - n_last, n_last, full=False)["x"]
        for j in range(2):
            assert tau[l, j] == pytest.approx(integrated_act(x[:, j]), rel=1e-9)
    assert 8 < np.median(tau) < 30   # tau ~ 16 for this configuration (tests/test_gpu_statistical.py)


def test_host_callback_likelihood_equals_device_functor(engine_cls):
    """ptg_register_evaluate_log: a likelihood that lives on the host (here the isotropic Gaussian of example.cc:116-143 in numpy,
    same operation order) gives bit-identical chains to the fused run with the device functor: the swap phase, proposals, prior,
    Metropolis test and history are the same device code, only log L comes from the caller"""
    spec = Spec("gauss", 2, 6, centers=[2, -3], halfwidths=[2, 3], evolve_rate=0.01)
    L, steps = 40, 150
    x0 = np.array([2.0, -3.0]); tw = 2 * 0.5 * 0.5; lnnorm = -0.5 * 2 * np.log(np.pi * tw)
    calls = []

    def loglike(x):
        calls.append(len(x))
        dx = x - x0
        r2 = dx[:, 0] * dx[:, 0] + dx[:, 1] * dx[:, 1]
        return lnnorm - r2 / tw

    a = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=1000, trace_steps=steps))
    spec.setup(a); a.select_kernel(K.KERNEL_SHARED); a.init_from_prior(); a.step(steps); a.synchronize()
    b = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=1000, trace_steps=steps))
    spec.setup(b); b.register_evaluate_log(loglike); b.init_from_prior(); b.step(steps); b.synchronize()
    assert a.get_total_steps() == b.get_total_steps()
    for l in (0, 13, 39):
        assert compare_dumps(engine_dump(a, l), engine_dump(b, l), rtol=0.0, what="callback") == []
    la, ca = a.get_trace(0, steps); lb, cb = b.get_trace(0, steps)
    assert (ca == cb).all() and la.tobytes() == lb.tobytes()
    assert len(calls) >= 100 + steps and max(calls) <= L * 6          # one batched call per init round and per PT iteration
    assert np.allclose(b.eval_loglike(np.array([[2.0, -3.0]])), [lnnorm])


def test_device_ess_recipe_matches_host_recipe(engine_cls):
    """ptg_get_autocovar_windows + the batched combination = analysis.report_effective_samples (the restatement pinned to the reference
    build) applied to each cold chain's history, whole-run and ring-window forms, with thinning."""
    from ptmcmc_b200.analysis import report_effective_samples
    from tests.models import Spec
    for se, steps in ((1, 9000), (3, 21000)):
        spec = Spec("gauss", 2, 4, centers=[2, -3], halfwidths=[2, 3], save_every=se)
        L = 12
        e = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=spec.de_ni * 2 + steps // se * 6 // 5 + 600, record_level=K.RECORD_BASIC, save_every=se))
        spec.setup(e); e.init_from_prior(); e.step(steps); e.synchronize()
        ess, length = e.report_effective_samples_all()
        cnt = e.get_counters()
        for l in range(L):
            h_ess, h_len = e.report_effective_samples(ladder=l)
            assert length[l] == h_len and h_len > 0
            assert abs(ess[l] - h_ess) <= 1e-9 * h_ess, (l, ess[l], h_ess)
        nr = 4000
        ess_w, len_w = e.report_effective_samples_all(window_records=nr)
        for l in range(0, L, 5):
            n = int(cnt["nsize"][l * 4])
            x = e.get_history(l, 0, n - nr, nr, full=False)["x"]
            h_ess, h_len = report_effective_samples(x, nr * se, n_init=0, add_every=se, width=se * 1000, every=se)
            assert len_w[l] == h_len and abs(ess_w[l] - h_ess) <= 1e-9 * h_ess
        with pytest.raises(Exception):
            e.get_autocovar_windows(0, 1000, 50, [0, 1], 2)       # windows older than the chain
        e.close()


@pytest.mark.parametrize("d,R,L,kw", [(100, 24, 5, {}), (100, 24, 3, dict(evolve_rate=0.01, swap_rate=0.3)), (40, 6, 7, {}), (20, 4, 9, dict(save_every=3)),
                                       (64, 15, 3, {}), (100, 31, 2, dict(swap_rate=0.05))])
def test_pipelined_fullcov_kernel_is_bit_identical(d, R, L, kw, engine_cls):
    """the pipelined full-covariance kernel (ptg_wide_pipe.cuh: matrices resident in shared memory, ladder warp, rows of the contractions
    picked by indirection) against the round-1 L1-streamed DMMA kernel (pinned with KERNEL_SHARED): same chains bit for bit"""
    from tests.models import fullcov_spec
    spec = fullcov_spec(d, R, Tmax=1e4, de_ni=11, prop="covde", **kw)
    spec.extra["gauss_1d_frac"] = 0.2
    runs = []
    for kern in (K.KERNEL_AUTO, K.KERNEL_SHARED):
        e = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=11 * d + 700, record_level=K.RECORD_FULL, trace_steps=300))
        e.select_kernel(kern); spec.setup(e); e.init_from_prior(); e.step(120); e.step(180); e.synchronize()
        cur, cnt = e.get_current(), e.get_counters()
        n = int(cnt["nsize"][R - 1])
        h = e.get_history(L - 1, R - 1, 0, int(cnt["nsize"][(L - 1) * R + R - 1]))
        runs.append((cur["x"].tobytes(), cur["lpost"].tobytes(), cur["beta"].tobytes(), cnt["naccept"].tobytes(), cnt["nsize"].tobytes(), h["x"].tobytes(),
                     h["lpost"].tobytes(), e.get_trace(0, 300)[1].tobytes(), e.get_swap_stats()["swap_accept"].tobytes()))
        e.close()
    for a, b in zip(*runs):
        assert a == b


@pytest.mark.parametrize("model,R,d,n", [("poly", 16, 5, 1000), ("poly", 8, 5, 1000), ("poly", 5, 4, 203), ("poly", 3, 2, 37), ("poly", 24, 5, 500),
                                         ("poly", 32, 3, 64), ("poly", 1, 5, 100), ("sinusoid", 32, 9, 1000), ("sinusoid", 6, 9, 517), ("sinusoid", 16, 9, 130)])
def test_compacted_data_likelihood_matches_the_reference_arithmetic(model, R, d, n, engine_cls, oracle_cls):
    """the production data-chi^2 functors run one ladder per warp and spread the data sums of the chains that passed the prior gate over
    all 32 lanes (32 / n lanes per chain, slices added in slice order; ragged slices, one to 32 wanting chains, ladders of 1 ... 32 rungs):
    every STORED log-likelihood equals the reference's unfused arithmetic at the stored position to 1e-12, and the decisions agree with the
    unfused one-lane-per-chain kernel's over the first steps"""
    if model == "poly":
        spec = Spec("poly", d, R, centers=np.zeros(d), halfwidths=np.full(d, 10.0), prop="de", Tmax=1e6, de_ni=12, extra=poly_data(n=n, d=d))
    else:
        spec = sinusoid_spec(R, n=n, dt=0.01, de_ni=12)
    L = 37
    mk = lambda: engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=spec.de_ni * d + 300, record_level=K.RECORD_BASIC, trace_steps=40))
    e = mk(); spec.setup(e); e.init_from_prior(); e.step(120); e.synchronize()
    o = oracle_cls(spec.config(n_ladders=1)); spec.setup(o)
    cnt = e.get_counters()
    for l, r in ((0, 0), (L // 2, R // 2), (L - 1, R - 1)):
        nrec = int(cnt["nsize"][l * R + r])
        h = e.get_history(l, r, nrec - 100, 100, full=False)
        want = o.eval_loglike(h["x"])
        fin = np.isfinite(want)
        assert (np.isfinite(h["llike"]) == fin).all()
        assert (np.abs(h["llike"][fin] - want[fin]) <= 1e-12 * np.abs(want[fin])).all()
    # the unfused functor (pinned PTG_KERNEL_FAST, several ladders per warp where they fit) takes the same decisions over the first steps
    f = mk(); f.select_kernel(K.KERNEL_FAST); spec.setup(f); f.init_from_prior(); f.step(120); f.synchronize()
    ce, cf = e.get_trace(0, 40)[1], f.get_trace(0, 40)[1]
    assert (ce == cf).mean() > 0.999
