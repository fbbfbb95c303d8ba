"""Locates / builds / loads libptmcmc_b200.so (the CUDA engine).  Fails loudly: there is no CPU fallback."""
import ctypes as C
import os
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
SO = os.environ.get("PTMCMC_B200_LIB") or os.path.join(CSRC, "libptmcmc_b200.so")  # env override: developer builds
_lib = None


def build(jobs=8, verbose=False):
    """compile the engine for sm_100a in-tree (nvcc cross-compiles without a GPU)"""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC, "-j%d" % jobs], stdout=out)
    return SO


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise RuntimeError("ptmcmc_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the engine is CUDA-only; there is no CPU fallback)" % SO)
        _lib = C.CDLL(SO)
    return _lib
