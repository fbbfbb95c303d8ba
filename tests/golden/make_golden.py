"""Generates the golden fixtures of tests/golden/ from the UNMODIFIED reference.

Run in the build container (where /root/reference exists) after `oracle/build_ref.sh`:
    python tests/golden/make_golden.py
For every case of tests.models.parity_cases() it runs oracle/_ref/ref_trace (our driver, compiled against the reference
sources in place; it steps parallel_tempering_chains through the reference's own public API with the reference's own RNG)
and stores, per case, in ref_<case>.npz:
    * the complete raw history of the cold rung (x, lpost, llike, acceptance ratio, invtemp, proposal type),
    * for every rung the final counters and the SHA-256 of its full history record block (bit-exact pin),
    * the swap statistics, and for eigen-rotated Gaussian proposals the eigen-decomposition the reference used.
newran.txt holds raw draws of the reference's RNG layer (oracle/_ref/ref_rng) for the seeds the tests use.
ref_ess.npz (`python tests/golden/make_golden.py ess`) holds, for a few longer runs, the cold chain's parameter history and the
(ess, length) pair the reference's own chain::report_effective_samples returned for it (called as ptmcmc.cc:645 does).
"""
import hashlib
import os
import subprocess
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.models import parity_cases, read_ref_trace  # noqa: E402
from tests.oracle_binding import REF_DIR, run_ref_trace  # noqa: E402


def rung_digest(r):
    h = hashlib.sha256()
    for k in ("x", "hlpost", "hllike", "hacc", "hbeta"):
        h.update(np.ascontiguousarray(r[k], dtype=np.float64).tobytes())
    h.update(np.ascontiguousarray(r["htype"], dtype=np.int32).tobytes())
    return h.hexdigest()


def main():
    for name, spec, steps, _L in parity_cases():
        with tempfile.TemporaryDirectory() as td:
            out = run_ref_trace(spec, steps, td)
            ref = read_ref_trace(out)
            eig = np.fromfile(out + ".eig") if spec.prop in ("cov", "covde") else np.zeros(0)
        cold = ref["rungs"][0]
        keep = 150 if spec.dim > 16 else cold["nsize"]   # wide cases: newest 150 cold samples + digests of everything
        cold_from = max(0, cold["nsize"] - keep)
        counters = np.array([[r[k] for k in ("nsize", "nhist", "ntries", "naccept", "last_type")] for r in ref["rungs"]], dtype=np.int64)
        finals = np.array([[r[k] for k in ("beta", "lpost", "llike", "map_lpost")] for r in ref["rungs"]])
        np.savez_compressed(os.path.join(HERE, "ref_%s.npz" % name), steps=steps, counters=counters, finals=finals,
                            digests=np.array([rung_digest(r) for r in ref["rungs"]]), cold_from=cold_from, cold_x=cold["x"][cold_from:],
                            cold_lpost=cold["hlpost"][cold_from:], cold_llike=cold["hllike"][cold_from:], cold_acc=cold["hacc"][cold_from:],
                            cold_beta=cold["hbeta"][cold_from:], cold_type=cold["htype"][cold_from:],
                            swap_count=ref["swap_count"], swap_accept=ref["swap_accept"], directions=ref["directions"],
                            ups=ref["ups"], downs=ref["downs"], instances=ref["instances"], eig=eig)
        print("golden:", name, "cold_nsize", cold["nsize"])
    with open(os.path.join(HERE, "newran.txt"), "w") as f:
        for seed in ("0.224", "0.012556", "0.1234"):
            out = subprocess.check_output([os.path.join(REF_DIR, "ref_rng"), seed, "64"]).decode()
            f.write("# seed %s\n%s" % (seed, out))


ESS_CASES = [  # (parity case, steps, esslimit)
    ("A_gauss2d_default", 12000, -1),
    ("C1_sines_d3_R32", 12000, -1),
    ("gauss3_de_save3", 40000, -1),
    ("A_gauss2d_default", 30000, 1000),
]


def run_ref_ess(spec, steps, esslimit, td):
    """-> (ref trace dict, ess, length) from oracle/_ref/ref_trace ... ess=1"""
    import re
    out = os.path.join(td, "ref.bin")
    args = [os.path.join(REF_DIR, "ref_trace")] + spec.ref_args(td, steps, out) + ["ess=1", "esslimit=%d" % esslimit]
    m = re.search(r"ref_ess: ess=(\S+) length=(\d+)", subprocess.check_output(args).decode())
    return read_ref_trace(out), float(m.group(1)), int(m.group(2))


def main_ess():
    cases = {c[0]: c[1] for c in parity_cases()}
    blob = {}
    for i, (name, steps, esslimit) in enumerate(ESS_CASES):
        with tempfile.TemporaryDirectory() as td:
            ref, ess, length = run_ref_ess(cases[name], steps, esslimit, td)
        cold = ref["rungs"][0]
        blob["x%d" % i] = cold["x"]
        blob["meta%d" % i] = np.array([cold["nhist"], ref["ninit"], ref["save_every"], esslimit, ess, length], dtype=np.float64)
        print("golden ess:", name, steps, esslimit, "->", ess, length)
    np.savez_compressed(os.path.join(HERE, "ref_ess.npz"), **blob)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ess":
        main_ess()
    else:
        main()
