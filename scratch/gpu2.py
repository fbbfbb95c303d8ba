import sys, numpy as np
sys.path.insert(0, '/root/repo')
from tests.models import *
from tests.oracle_binding import Oracle
from ptmcmc_b200.engine import Engine
from ptmcmc_b200 import _capi as K
np.set_printoptions(precision=17, linewidth=200)
spec = Spec("sines", 2, 8, seed=0.012556, evolve_rate=0.01, evolve_lpost_cut=0.5)
steps, L = 200, 2
o = Oracle(spec.config(n_ladders=L, rng_mode=2, trace_steps=steps))
spec.setup(o); o.seed_newran(spec.seed); o.record_tapes(True); o.init_from_prior(); o.step(steps)
u, uo, z, zo = o.get_tapes()
o2 = Oracle(spec.config(n_ladders=L, rng_mode=K.RNG_TAPE, trace_steps=steps))
spec.setup(o2); o2.inject_tapes(u, uo, z, zo); o2.init_from_prior()
g = Engine(spec.config(n_ladders=L, rng_mode=K.RNG_TAPE, trace_steps=steps, hist_capacity=1000))
spec.setup(g); g.inject_tapes(u, uo, z, zo); g.init_from_prior()
prev = None
for s in range(steps):
    co, cg = o2.get_current(), g.get_current()
    so, sg = o2.get_swap_stats(), g.get_swap_stats()
    def far(a, b): return (np.abs(a-b) > 1e-9*np.maximum(1e-300, np.abs(a))).any()
    if far(co["beta"], cg["beta"]) or far(co["x"], cg["x"]) or far(co["lpost"], cg["lpost"]) or far(co["llike"], cg["llike"]) or not (so["swap_accept"] == sg["swap_accept"]).all():
        print("x o", co["x"][8:].T, "\nx g", cg["x"][8:].T); print("prev x", prev[0]["x"][8:].T)
        cnto, cntg = o2.get_counters(), g.get_counters(); print("nsize o", cnto["nsize"][8:], "g", cntg["nsize"][8:])
        do, dg = engine_dump(o2, 1), engine_dump(g, 1)
        for r in range(8):
            a, b = do["rungs"][r], dg["rungs"][r]
            n = min(a["nsize"], b["nsize"])
            bad = np.argwhere((a["x"][:n] != b["x"][:n]).any(axis=1)).ravel()
            if len(bad): print("rung", r, "hist x first mismatch idx", bad[:6], "of", n, "\n o", a["x"][bad[0]-1:bad[0]+2], "\n g", b["x"][bad[0]-1:bad[0]+2], "beta", a["hbeta"][bad[0]-1:bad[0]+2], b["hbeta"][bad[0]-1:bad[0]+2], "types", a["htype"][bad[0]-1:bad[0]+2])
        lo, c_o = o2.get_trace(0, s); lg, c_g = g.get_trace(0, s)
        for st in range(s-6, s):
            print("step", st, "codes o", [hex(v) for v in c_o[st, 8:]], "g", [hex(v) for v in c_g[st, 8:]]); print("   lhr o", lo[st, 8:]); print("   lhr g", lg[st, 8:])
        print("beta/swap mismatch after", s, "steps")
        sys.exit(0)
        print("before: beta", prev[0]["beta"][8:], "\n lpost o", prev[0]["lpost"][8:], "\n lpost g", prev[1]["lpost"][8:], "\n llike o", prev[0]["llike"][8:], "\n llike g", prev[1]["llike"][8:])
        print("after: beta o", co["beta"][8:], "\n beta g", cg["beta"][8:], "\n lpost o", co["lpost"][8:], "\n lpost g", cg["lpost"][8:])
        print("swap_count o", so["swap_count"][1], "g", sg["swap_count"][1]); print("swap_accept o", so["swap_accept"][1], "g", sg["swap_accept"][1])
        print("prev swap_count", prev[2]["swap_count"][1], "acc", prev[2]["swap_accept"][1])
        lo, c_o = o2.get_trace(s-1, 1); print("codes", [hex(v) for v in c_o[0, 8:]])
        break
    prev = (co, cg, so)
    o2.step(1); g.step(1)
