import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from tests.models import *
from tests.oracle_binding import Oracle
from ptmcmc_b200.engine import Engine
from ptmcmc_b200 import _capi as K

def parity(spec, steps, L=1, what=""):
    # oracle with the reference RNG, recording tapes
    o = Oracle(spec.config(n_ladders=L, rng_mode=2, trace_steps=steps))
    spec.setup(o); o.seed_newran(spec.seed); o.record_tapes(True); o.init_from_prior(); o.step(steps)
    u, uo, z, zo = o.get_tapes()
    g = Engine(spec.config(n_ladders=L, rng_mode=K.RNG_TAPE, trace_steps=steps, hist_capacity=spec.de_ni*spec.dim+steps+steps//4+8))
    spec.setup(g); g.inject_tapes(u, uo, z, zo); g.init_from_prior(); g.step(steps)
    try: g.synchronize()
    except Exception as e: print("   !!", e)
    nbad = 0
    for l in range(L):
        bad = compare_dumps(engine_dump(o, l), engine_dump(g, l), rtol=1e-12, what="%s L%d" % (what, l))
        nbad += len(bad)
        for b in bad[:3]: print("   ", b)
    lo, co = o.get_trace(0, steps); lg, cg = g.get_trace(0, steps)
    if not (co == cg).all():
        st, chn = np.argwhere(co != cg)[0]
        print("   first decision mismatch at step %d chain %d: oracle code %x lhr %r | gpu code %x lhr %r" % (st, chn, co[st,chn], lo[st,chn], cg[st,chn], lg[st,chn]))
        for s2 in range(max(0,st-2), st+1):
            print("     step", s2, "codes o", [hex(v) for v in co[s2]], "g", [hex(v) for v in cg[s2]])
            print("     lhr o", lo[s2], "g", lg[s2])
    print("%-28s %s  decisions equal: %s  max|dlhr| %.3g" % (what, "OK" if nbad == 0 else "MISMATCH", bool((co == cg).all()),
          np.nanmax(np.abs(np.where(np.isfinite(lo) & np.isfinite(lg), lo - lg, 0)))))
    return nbad == 0

rng = np.random.default_rng(5)
ok = True
ok &= parity(Spec("sines", 2, 8, seed=0.012556, evolve_rate=0.01), 1500, L=1, what="sines evolve")
ok &= parity(Spec("sines", 2, 8, seed=0.012556, evolve_rate=0.01, evolve_lpost_cut=0.5), 1500, L=1, what="sines evolve+cut L1")
ok &= parity(Spec("gauss", 2, 8, centers=[2,-3], halfwidths=[2,3]), 1000, L=3, what="A gauss2d default")
ok &= parity(Spec("sines", 3, 32, seed=0.1234), 500, L=2, what="C1 sines d3 R32")
ok &= parity(Spec("sines", 2, 8, seed=0.012556, evolve_rate=0.01, evolve_lpost_cut=0.5), 1500, L=2, what="sines evolve+cut")
ok &= parity(Spec("gauss", 3, 6, centers=[2,-3,5], halfwidths=[2,3,5], prop="de", save_every=3), 1500, what="gauss3 de save3")
ok &= parity(Spec("gauss", 2, 5, centers=[2,-3], halfwidths=[2,3], prop="default", bound="w", extra=dict(sigma=3.0)), 1000, what="wrap")
ok &= parity(Spec("gauss", 2, 5, centers=[2,-3], halfwidths=[2,3], prop="default", bound="r", extra=dict(sigma=3.0)), 1000, what="reflect")
ok &= parity(Spec("gauss", 3, 6, centers=[2,-3,5], halfwidths=[2,3,5], prop="gauss", bound="l"), 1000, what="limit gauss")
ok &= parity(Spec("gauss", 2, 5, centers=[2,-3], halfwidths=[2,3], prop="prior", prior="mixed", prior_types=[1,2]), 1000, what="prior-draw mixed")
ok &= parity(Spec("gauss", 2, 5, centers=[2,-3], halfwidths=[2,3], prop="default", extra=dict(de_unlikely_alpha=0.5)), 1000, what="unlikely_alpha")
xs = -10 + 0.02*(np.arange(1000)+0.5); truth = rng.uniform(-10,10,5)
ys = sum(truth[j]*xs**j for j in range(5)) + rng.normal(size=1000)
ok &= parity(Spec("poly", 5, 4, centers=np.zeros(5), halfwidths=np.full(5,10.0), prop="de", Tmax=1e6, extra=dict(data_x=xs, data_y=ys, data_dy=np.ones(1000))), 200, what="B poly")
t = np.arange(500)*1e-3*20; A=[1,0.7,0.4]; f=[1.3,3.1,7.7]; ph=[0.3,1.1,2.0]
y = sum(A[k]*np.sin(2*np.pi*f[k]*t+ph[k]) for k in range(3)) + rng.normal(size=500)
c = np.array([1,5,np.pi]*3, dtype=float)
ok &= parity(Spec("sinusoid", 9, 4, centers=c, halfwidths=c.copy(), prop="default", bound="oowoowoow", extra=dict(data_x=t, data_y=y, data_dy=np.ones(500))), 150, what="C2 sinusoid")
d=6; Amat = rng.normal(size=(d+5,d)); Cm = Amat.T@Amat; Cinv=np.linalg.inv(Cm)
sp = Spec("fullcov", d, 4, centers=np.zeros(d), halfwidths=100*np.sqrt(np.diag(Cm)), prop="covde", Tmax=100, extra=dict(cinv=Cinv.ravel(), like0=-3.0))
w, V = np.linalg.eigh(2.38**2/d*Cm); sp.eig = (np.sqrt(w), V)
ok &= parity(sp, 800, what="fullcov d6 covde")
print("ALL OK" if ok else "SOME FAILED")

# quick throughput probe, config C1
spec = Spec("sines", 3, 32)
for L in (4096,):
    g = Engine(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, record_level=K.RECORD_BASIC, hist_capacity=1024))
    spec.setup(g); g.init_from_prior(); g.step(50); g.synchronize()
    t0 = time.time(); g.step(500); g.synchronize(); dt = time.time()-t0
    print("C1 L=%d: %.3g tempered chain-steps/s (%.2f ms/PT step)" % (L, L*32*500/dt, dt/500*1e3))
