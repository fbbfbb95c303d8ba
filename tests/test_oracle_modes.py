"""Host-side properties of the hot path that need no GPU, checked on the oracle: tape replay, Philox addressing
invariances (the properties the multi-GPU sharding relies on), draw-order contract, history / counter bookkeeping."""
import numpy as np
import pytest
from ptmcmc_b200 import _capi as K
from tests.models import Spec, engine_dump, compare_dumps, parity_cases
from tests.oracle_binding import Oracle
from tests.parity import record_reference_run

CASES = [c for c in parity_cases() if c[0] in ("A_gauss2d_default", "sines_evolve_cut", "unlikely_alpha", "prior_draw_mixed")]


@pytest.mark.parametrize("name,spec,steps,L", CASES, ids=[c[0] for c in CASES])
def test_tape_replay_reproduces_reference_rng_run(name, spec, steps, L):
    """SURVEY.md 8c draw-order contract: replaying the recorded uniforms / normals stream by stream gives the same run"""
    o, tapes, marks = record_reference_run(Oracle, spec, steps, L)
    o2 = Oracle(spec.config(n_ladders=L, rng_mode=K.RNG_TAPE, trace_steps=steps))
    spec.setup(o2); o2.inject_tapes(*tapes); o2.init_from_prior(); o2.step(steps)
    for l in range(L):
        assert compare_dumps(engine_dump(o, l), engine_dump(o2, l), rtol=0.0, what=name) == []
    lo, co = o.get_trace(0, steps); l2, c2 = o2.get_trace(0, steps)
    assert (co == c2).all() and lo.tobytes() == l2.tobytes()
    # marks are monotone cursors inside each stream's tape segment
    um, zm = marks
    assert (np.diff(um, axis=0) >= 0).all() and (np.diff(zm, axis=0) >= 0).all()
    assert (um[0] >= tapes[1][:-1]).all() and (um[-1] <= tapes[1][1:]).all()


def run_philox(spec, L, steps, chunks=None, ladder_offset=0, seed=0xB2000003, **kw):
    o = Oracle(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, seed=seed, ladder_offset=ladder_offset, **kw))
    spec.setup(o); o.init_from_prior()
    for n in (chunks or [steps]):
        o.step(n)
    return o


def test_philox_run_is_invariant_to_batching_and_offset():
    """ladder g of a batch == ladder 0 of an engine created with ladder_offset = g: the property that makes results
    independent of how ladders are sharded over GPUs (mirrors per-chain RNG invariance, chain.hh:45,66-67)"""
    spec = Spec("sines", 3, 8, seed=0.1234)
    big = run_philox(spec, 5, 300)
    for g in (0, 3, 4):
        one = run_philox(spec, 1, 300, ladder_offset=g)
        assert compare_dumps(engine_dump(big, g), engine_dump(one, 0), rtol=0.0) == []


def test_philox_run_is_invariant_to_step_chunking():
    spec = Spec("gauss", 2, 6, centers=[2, -3], halfwidths=[2, 3], evolve_rate=0.01)
    a = run_philox(spec, 2, 400)
    b = run_philox(spec, 2, 400, chunks=[1, 99, 250, 50])
    for l in range(2):
        assert compare_dumps(engine_dump(a, l), engine_dump(b, l), rtol=0.0) == []


def test_different_seeds_differ():
    spec = Spec("gauss", 2, 4, centers=[2, -3], halfwidths=[2, 3])
    a = run_philox(spec, 1, 100, seed=1); b = run_philox(spec, 1, 100, seed=2)
    assert compare_dumps(engine_dump(a, 0), engine_dump(b, 0), rtol=0.0) != []


def test_history_bookkeeping_reference_swaps():
    """H4: a rung can take part in two swap trials per step and then gets two appends; swapped rungs skip their MH step;
    total appends = sum of Nhist; Nsize = Ninit + ceil-ish(Nhist / save_every)"""
    spec = Spec("sines", 2, 8, seed=0.5, swap_rate=0.4, save_every=3)
    steps = 600
    o = run_philox(spec, 3, steps, trace_steps=steps)
    cnt = o.get_counters(); R = 8
    assert o.get_total_steps() == cnt["nhist"].sum()
    assert (cnt["nhist"] >= steps).all() and (cnt["nhist"] > steps).any()
    assert (cnt["nsize"] == spec.de_ni * spec.dim + (cnt["nhist"] + 2) // 3).all()
    _, code = o.get_trace(0, steps)
    swapped = (code & K.TRACE_SWAPPED) != 0
    # MH tries are counted only on steps where the rung was not in a swap trial (Ntries starts at 1, chain.cc:649)
    assert (cnt["ntries"] - 1 == (~swapped).sum(axis=0)).all()
    sw = o.get_swap_stats()
    assert (sw["swap_accept"] <= sw["swap_count"]).all() and sw["swap_count"].sum() > 0
    # every ladder keeps a permutation of instances
    assert (np.sort(sw["instances"], axis=1) == np.arange(R)).all()


def test_even_odd_mode_properties():
    """even/odd performance mode: pair (i,i+1) tried only when i has the step's parity; never two appends per step"""
    spec = Spec("sines", 2, 8, seed=0.5, swap_rate=0.5)
    steps = 200
    o = run_philox(spec, 2, steps, trace_steps=steps, swap_mode=K.SWAP_EVEN_ODD)
    cnt = o.get_counters()
    assert (cnt["nhist"] == steps).all()
    _, code = o.get_trace(0, steps)
    swapped = ((code & K.TRACE_SWAPPED) != 0).reshape(steps, 2, 8)
    for s in range(steps):
        par = s & 1
        for l in range(2):
            idx = np.flatnonzero(swapped[s, l])
            # swapped rungs come in adjacent pairs starting at a rung of the step's parity
            assert len(idx) % 2 == 0
            assert all(idx[k] % 2 == par and idx[k + 1] == idx[k] + 1 for k in range(0, len(idx), 2))


def test_posterior_moments_gaussian():
    """statistical sanity of the restated algorithm: cold-chain mean / variance of a 2-D Gaussian target"""
    spec = Spec("gauss", 2, 6, centers=[2, -3], halfwidths=[2, 3], Tmax=100)
    L, steps = 8, 4000
    o = run_philox(spec, L, steps)
    xs = []
    for l in range(L):
        n = int(o.get_counters()["nsize"][l * 6])
        xs.append(o.get_history(l, 0, n - 3000, 3000)["x"])
    x = np.concatenate(xs)
    assert np.allclose(x.mean(axis=0), [2, -3], atol=0.03)
    assert np.allclose(x.var(axis=0), [0.25, 0.25], rtol=0.1)
