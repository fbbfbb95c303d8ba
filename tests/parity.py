"""Parity harness shared by the GPU tests, `__graft_entry__.smoke()` and scratch scripts.

`tape_parity`   : the oracle runs a configuration with the reference's own RNG (newran restatement) and records every
                  uniform / normal each stream consumes plus the cursor of every stream at each PT-step boundary; the CUDA
                  engine replays those tapes (PTG_RNG_TAPE) and must reproduce decisions, states and histories.
`philox_parity` : both run the published Philox draw layout (include/ptmcmc_b200_rng.h) from the same key.

Tolerances (BASELINE.json north_star): decisions, proposal types, history indices and positions bit-exact;
log-likelihood / log-posterior values to 1e-12 relative (CUDA libm vs glibc differ in the last ulp of sin/log/exp).
"""
import numpy as np
from ptmcmc_b200 import _capi as K
from .models import engine_dump, compare_dumps

RTOL = 1e-12
PAD = 8  # spare tape entries per stream: a near-tie step may consume one draw more than the recorded run did


def record_reference_run(oracle_cls, spec, steps, L=1, **cfg_kw):
    """oracle with the reference RNG; returns (oracle, padded tapes, absolute step marks)"""
    o = oracle_cls(spec.config(n_ladders=L, rng_mode=2, trace_steps=steps, **cfg_kw))
    spec.setup(o)
    o.seed_newran(spec.seed)
    o.record_tapes(True)
    o.init_from_prior()
    o.step(steps)
    u, uo, z, zo = o.get_tapes()
    um, zm = o.get_tape_marks()

    def pad(t, off, fill):
        ns = len(off) - 1
        cnt = np.diff(off)
        noff = np.concatenate([[0], np.cumsum(cnt + PAD)]).astype(np.int64)
        out = np.full(int(noff[-1]), fill)
        for s in range(ns):
            out[noff[s]:noff[s] + cnt[s]] = t[off[s]:off[s + 1]]
        return out, noff

    up, uop = pad(u, uo, 0.5)
    zp, zop = pad(z, zo, 0.0)
    return o, (up, uop, zp, zop), (um + uop[:-1][None, :], zm + zop[:-1][None, :])


def lhr_close(lo, lg, scale):
    """|lhr_o - lhr_g| <= RTOL * scale, where scale bounds the magnitudes that were subtracted to form lhr"""
    both_inf = (lo == lg) | (np.isnan(lo) & np.isnan(lg))
    with np.errstate(invalid="ignore"):
        return both_inf | (np.abs(lo - lg) <= RTOL * scale)


def compare_runs(o, g, steps, L, what, lpost_scale=None, rtol=RTOL, exact_x=True):
    """returns list of mismatch descriptions (empty = parity)"""
    bad = []
    for l in range(L):
        bad += compare_dumps(engine_dump(o, l), engine_dump(g, l), rtol=rtol, what="%s ladder %d" % (what, l), exact_x=exact_x)
    lo, co = o.get_trace(0, steps)
    lg, cg = g.get_trace(0, steps)
    if not (co == cg).all():
        st, chn = np.argwhere(co != cg)[0]
        bad.append("%s: first decision mismatch at step %d chain %d: oracle code %#x lhr %r | engine code %#x lhr %r"
                   % (what, st, chn, co[st, chn], lo[st, chn], cg[st, chn], lg[st, chn]))
    # lhr = hastings + newlpost - lpost: tolerance relative to the size of the posteriors that were subtracted
    cur = o.get_current()
    scale = max(1.0, float(np.nanmax(np.abs(np.where(np.isfinite(cur["lpost"]), cur["lpost"], 0.0))))) if lpost_scale is None else lpost_scale
    hist_scale = 1.0
    for l in range(L):
        for r in range(o.cfg.n_rungs):
            n = int(o.get_counters()["nsize"][l * o.cfg.n_rungs + r])
            h = o.get_history(l, r, 0, n)
            v = np.abs(h["lpost"][np.isfinite(h["lpost"])])
            if v.size:
                hist_scale = max(hist_scale, float(v.max()))
    ok = lhr_close(lo, lg, max(scale, hist_scale) * (rtol / RTOL))
    if not ok.all():
        st, chn = np.argwhere(~ok)[0]
        bad.append("%s: lhr differs beyond %g at step %d chain %d: %r vs %r" % (what, rtol, st, chn, lo[st, chn], lg[st, chn]))
    return bad


def tape_parity(oracle_cls, engine_cls, spec, steps, L=1, what="", **cfg_kw):
    if "hist_capacity" not in cfg_kw:  # big enough that the ring never wraps: the reference history is unbounded (H3)
        cfg_kw["hist_capacity"] = spec.de_ni * spec.dim + 2 * steps + 8
    o, tapes, marks = record_reference_run(oracle_cls, spec, steps, L, **cfg_kw)
    cap = cfg_kw.pop("hist_capacity")
    g = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_TAPE, trace_steps=steps, hist_capacity=cap, **cfg_kw))
    spec.setup(g)
    g.inject_tapes(*tapes)
    g.inject_tape_marks(*marks)
    g.init_from_prior()
    g.step(steps)
    g.synchronize()
    # Prior factors whose inverse cdf goes through libm (polar: acos, co-polar: asin, logarithmic: exp; ProbabilityDist.h:131-134,202-205,
    # 248-251) make the DRAWN POSITIONS differ between glibc and CUDA libm in the last ulp, like the log-likelihoods: those cases compare
    # positions to 1e-12 relative; decisions, types and counters stay exact.  Every other case: positions bit-exact.
    libm_positions = spec.prior == "mixed" and any(int(t) in (K.PRIOR_POLAR, K.PRIOR_COPOLAR, K.PRIOR_LOG) for t in spec.prior_types)
    bad = compare_runs(o, g, steps, L, what, exact_x=not libm_positions)
    o.close(); g.close()
    return bad


# Philox mode: the normals come from Box-Muller through libm (log, sincospi on the device; log, sin, cos in glibc), so
# positions agree to rounding, not bitwise, and the run is compared over a short horizon with a loose tolerance;
# decisions must still be identical.
PHILOX_RTOL = 1e-8


def philox_parity(oracle_cls, engine_cls, spec, steps, L=1, what="", seed=0xB2000003, chunks=None, **cfg_kw):
    cap = cfg_kw.pop("hist_capacity", spec.de_ni * spec.dim + 2 * steps + 8)
    mk = lambda: spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, trace_steps=steps, hist_capacity=cap, seed=seed, **cfg_kw)
    o = oracle_cls(mk()); spec.setup(o); o.init_from_prior(); o.step(steps)
    g = engine_cls(mk()); spec.setup(g); g.init_from_prior()
    for n in (chunks or [steps]):
        g.step(n)
    g.synchronize()
    bad = compare_runs(o, g, steps, L, what, rtol=PHILOX_RTOL, exact_x=False)
    o.close(); g.close()
    return bad
