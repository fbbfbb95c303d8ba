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


if __name__ == "__main__":
    main()
