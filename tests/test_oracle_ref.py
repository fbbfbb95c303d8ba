"""Pins the CPU oracle (oracle/, TEST INFRASTRUCTURE) to the reference.

 * against the committed golden fixtures of tests/golden/ (made by tests/golden/make_golden.py from the unmodified reference
   run through oracle/_ref/ref_trace) -- bit-exact, every rung, every stored sample;
 * where oracle/_ref/ref_trace is present (build container, and the GPU box via gpurun), against the live reference too;
 * the restatement of the reference's RNG (newran MotherOfAll + Normal) against raw draws of the real library.
"""
import ctypes as C
import hashlib
import os
import subprocess
import numpy as np
import pytest
from tests.models import parity_cases, read_ref_trace, engine_dump, compare_dumps, fullcov_spec
from tests.oracle_binding import Oracle, oracle_lib, have_ref, run_ref_trace, REF_DIR

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = parity_cases()


def rung_digest(r):
    h = hashlib.sha256()
    for k in ("x", "hlpost", "hllike", "hacc", "hbeta"):
        h.update(np.ascontiguousarray(r[k], dtype=np.float64).tobytes())
    h.update(np.ascontiguousarray(r["htype"], dtype=np.int32).tobytes())
    return h.hexdigest()


def run_oracle_newran(spec, steps):
    o = Oracle(spec.config(n_ladders=1, rng_mode=2))
    spec.setup(o)
    o.seed_newran(spec.seed)
    o.init_from_prior()
    o.step(steps)
    return o


@pytest.mark.parametrize("name,spec,steps,L", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_golden(name, spec, steps, L):
    g = np.load(os.path.join(GOLDEN, "ref_%s.npz" % name))
    assert int(g["steps"]) == steps
    if spec.prop in ("cov", "covde"):
        d = spec.dim
        spec.eig = (g["eig"][:d].copy(), g["eig"][d:].reshape(d, d).copy())
    o = run_oracle_newran(spec, steps)
    dump = engine_dump(o, 0)
    counters = np.array([[r[k] for k in ("nsize", "nhist", "ntries", "naccept", "last_type")] for r in dump["rungs"]], dtype=np.int64)
    assert (counters == g["counters"]).all()
    finals = np.array([[r[k] for k in ("beta", "lpost", "llike", "map_lpost")] for r in dump["rungs"]])
    assert finals.tobytes() == g["finals"].tobytes()
    cold = dump["rungs"][0]
    f = int(g["cold_from"])
    for a, b in (("x", "cold_x"), ("hlpost", "cold_lpost"), ("hllike", "cold_llike"), ("hacc", "cold_acc"), ("hbeta", "cold_beta")):
        assert np.ascontiguousarray(cold[a][f:]).tobytes() == g[b].tobytes(), a
    assert (cold["htype"][f:] == g["cold_type"]).all()
    assert [rung_digest(r) for r in dump["rungs"]] == list(g["digests"])
    for k in ("swap_count", "swap_accept", "directions", "ups", "downs", "instances"):
        assert (np.asarray(dump[k]) == g[k]).all(), k


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/ref_trace not built (needs /root/reference)")
@pytest.mark.parametrize("name,spec,steps,L", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_live_reference(name, spec, steps, L, tmp_path):
    out = run_ref_trace(spec, steps, tmp_path)
    ref = read_ref_trace(out)
    if spec.prop in ("cov", "covde"):
        e = np.fromfile(out + ".eig"); d = spec.dim
        spec.eig = (e[:d].copy(), e[d:].reshape(d, d).copy())
    o = run_oracle_newran(spec, steps)
    assert compare_dumps(ref, engine_dump(o, 0), rtol=0.0, what=name) == []


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/ref_trace not built (needs /root/reference)")
@pytest.mark.parametrize("d", [3, 5, 7, 8, 12, 16])
def test_eigen_rotated_proposal_matches_live_reference(d, tmp_path):
    """`vec = diagTransform*vec` (proposal_distribution.hh:212): Eigen's GEMV summation order, odd and even dimensions"""
    spec = fullcov_spec(d, 3, Tmax=100, prop="cov")
    out = run_ref_trace(spec, 400, tmp_path)
    ref = read_ref_trace(out)
    e = np.fromfile(out + ".eig")
    spec.eig = (e[:d].copy(), e[d:].reshape(d, d).copy())
    o = run_oracle_newran(spec, 400)
    assert compare_dumps(ref, engine_dump(o, 0), rtol=0.0, what="cov d=%d" % d) == []


class Mother(C.Structure):
    _fields_ = [("m1", C.c_int16 * 10), ("m2", C.c_int16 * 10), ("started", C.c_int), ("seed", C.c_uint64)]


def newran_port_draws(seed, n):
    lib = oracle_lib()
    lib.mother_next.restype = C.c_double
    lib.newran_normal.restype = C.c_double
    a, b = Mother(), Mother()
    lib.mother_init(C.byref(a), C.c_double(seed))
    u = [lib.mother_next(C.byref(a)) for _ in range(n)]
    lib.mother_init(C.byref(b), C.c_double(seed))
    z = [lib.newran_normal(C.byref(b)) for _ in range(n)]
    # chain::chain() seeding (chain.hh:58-59): MotherOfAll(master.Next())
    m = Mother(); lib.mother_init(C.byref(m), C.c_double(seed))
    sub = []
    for _ in range(4):
        c = Mother(); lib.mother_init(C.byref(c), C.c_double(lib.mother_next(C.byref(m))))
        sub.append(lib.mother_next(C.byref(c)))
    return u + z + sub


def golden_newran():
    blocks, cur = {}, None
    for line in open(os.path.join(GOLDEN, "newran.txt")):
        if line.startswith("# seed"):
            cur = line.split()[2]; blocks[cur] = []
        elif line.strip():
            blocks[cur].append(float(line))
    return blocks


@pytest.mark.parametrize("seed", ["0.224", "0.012556", "0.1234"])
def test_newran_port_matches_golden(seed):
    want = golden_newran()[seed]
    got = newran_port_draws(float(seed), 64)
    assert len(want) == len(got) == 132
    assert np.array(want).tobytes() == np.array(got).tobytes()


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "ref_rng")), reason="oracle/_ref/ref_rng not built")
def test_newran_port_matches_live_reference():
    out = subprocess.check_output([os.path.join(REF_DIR, "ref_rng"), "0.3141", "200"]).decode().split()
    want = np.array([float(v) for v in out])
    got = np.array(newran_port_draws(0.3141, 200))
    assert want.tobytes() == got.tobytes()
