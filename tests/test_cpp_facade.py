"""The C++ host facade (include/ptmcmc_b200.hh: the reference's class interface over the C ABI).
CPU: the header compiles as C++11 and a driver written in the reference's style links against the engine library.
GPU: the driver's chain files (reference dumpChain format) equal the histories the ctypes mirror reads for the same run."""
import os
import subprocess
import numpy as np
import pytest
from ptmcmc_b200 import _capi as K
from ptmcmc_b200._lib import CSRC, SO
from tests.models import Spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER_SRC = os.path.join(ROOT, "tests", "cpp", "facade_driver.cc")


def build_driver(outdir):
    exe = os.path.join(str(outdir), "facade_driver")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), DRIVER_SRC, "-o", exe,
                           "-L", CSRC, "-lptmcmc_b200", "-Wl,-rpath," + CSRC])
    return exe


def test_facade_header_compiles_and_driver_links(tmp_path):
    assert os.path.exists(SO)
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Wextra", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c++",
                           os.path.join(ROOT, "include", "ptmcmc_b200.hh")])
    exe = build_driver(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


def read_chain_file(path):
    rows = []
    for line in open(path):
        if line.startswith("#"):
            continue
        head, tail = line.split(":")
        h = head.split(); t = tail.split()
        rows.append([float(v) for v in h] + [float(v) for v in t])
    return np.array(rows)


@pytest.mark.gpu
@pytest.mark.parametrize("model,dim,R,steps,L,evolve", [("sines", 3, 8, 400, 5, 0.0), ("gauss", 2, 6, 300, 3, 0.01), ("hostgauss", 2, 6, 120, 3, 0.01)])
def test_facade_driver_matches_engine(model, dim, R, steps, L, evolve, tmp_path, engine_cls):
    exe = build_driver(tmp_path)
    out = os.path.join(str(tmp_path), "chain.dat")
    r = subprocess.run([exe, model, str(dim), str(R), str(steps), str(L), out, repr(evolve)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert open(out).readline().startswith("#Ninit=%d, Nburn=%d" % (50 * dim, -50 * dim))
    assert open(out).readlines()[1].startswith("#eval: log(posterior) log(likelihood) acceptance_ratio prop_type: p0 p1")
    # the same run through the ctypes mirror
    if model == "sines":
        spec = Spec("sines", dim, R, evolve_rate=evolve)
    else:
        c = np.array([2.0 - 5.0 * (i % 2) for i in range(dim)]); hw = np.array([2.0 + i for i in range(dim)])
        spec = Spec("gauss", dim, R, centers=c, halfwidths=hw, evolve_rate=evolve)
    e = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=50 * dim + 2 * steps + 8, seed=0xB2000003))
    spec.setup(e)
    if model == "hostgauss":
        e.select_kernel(K.KERNEL_SHARED)   # the C++ driver's host likelihood (evaluate_log(state&)) against the fused device functor
    e.init_from_prior(); e.step(steps); e.synchronize()
    cnt = e.get_counters(); cur = e.get_current()
    for path, ladder, with_init in ((out, 0, True), (out + ".last", L - 1, False)):
        f = read_chain_file(path)
        c = ladder * R
        n = int(cnt["nsize"][c]); ninit = 50 * dim
        h = e.get_history(ladder, 0, 0, n)
        first = 0 if with_init else ninit
        assert len(f) == n - first
        assert (f[:, 0] == np.arange(first - ninit, n - ninit)).all()
        assert f[:, 1].tobytes() == h["lpost"][first:].tobytes()
        assert f[:, 2].tobytes() == h["llike"][first:].tobytes()
        assert f[:, 3].tobytes() == h["acc"][first:].tobytes()
        assert (f[:, 4] == h["type"][first:]).all()
        assert f[:, 5:5 + dim].tobytes() == h["x"][first:].tobytes()
        assert (f[:, 5 + dim] == cur["beta"][c]).all()
    summary = dict(kv.split("=") for kv in r.stdout.split(":", 1)[1].split())
    assert int(summary["size"]) == int(cnt["nsize"][0]) and int(summary["step"]) == int(cnt["nhist"][0])
    assert int(summary["total"]) == e.get_total_steps()
    assert float(summary["lpost"]) == cur["lpost"][0] and float(summary["x0"]) == cur["x"][0, 0]
    assert float(summary["invtemp_hot"]) == cur["beta"][R - 1]
    assert float(summary["first"]) == e.get_history(0, 0, 50 * dim, 1)["x"][0, 0]


@pytest.mark.gpu
def test_facade_effective_samples_match_python_recipe(tmp_path, engine_cls):
    """gpu_parallel_tempering_chains::report_effective_samples (device lag statistics + the C++ combination of chain.cc:340-416) against
    Engine.report_effective_samples (the numpy restatement pinned to the reference build) on the same Philox run"""
    exe = build_driver(tmp_path)
    out = os.path.join(str(tmp_path), "chain.dat")
    dim, R, steps, L = 2, 4, 7000, 2
    r = subprocess.run([exe, "gauss", str(dim), str(R), str(steps), str(L), out, "0.0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    line = [x for x in r.stdout.splitlines() if x.startswith("facade_ess:")][0]
    kv = dict(t.split("=") for t in line.split(":", 1)[1].split())
    c = np.array([2.0 - 5.0 * (i % 2) for i in range(dim)]); hw = np.array([2.0 + i for i in range(dim)])
    spec = Spec("gauss", dim, R, centers=c, halfwidths=hw)
    e = engine_cls(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=50 * dim + 2 * steps + 8, seed=0xB2000003, record_level=K.RECORD_BASIC))
    spec.setup(e); e.init_from_prior(); e.step(steps); e.synchronize()
    ess, length = e.report_effective_samples(ladder=0)
    assert int(kv["length"]) == length and length > 0
    assert abs(float(kv["ess"]) - ess) <= 1e-9 * ess
