"""Loads the CPU oracle (TEST INFRASTRUCTURE, oracle/libptmcmc_oracle.so) behind the same ctypes class as the product."""
import ctypes as C
import os
import subprocess
import numpy as np
from ptmcmc_b200 import _capi as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "libptmcmc_oracle.so")
        src = [os.path.join(ORACLE_DIR, f) for f in ("ptmcmc_oracle.c", "newran_port.c", "ptmcmc_oracle.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "CC=gcc"], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(so)
    return _lib


class Oracle(K.CApi):
    def __init__(self, cfg):
        super().__init__(oracle_lib(), "pto_", cfg)

    def seed_newran(self, seed):
        self._call("seed_newran", self.h, C.c_double(seed))

    def record_tapes(self, on=True):
        self._call("record_tapes", self.h, C.c_int(1 if on else 0))

    def get_tapes(self):
        ns = self.n_chains + self.cfg.n_ladders
        uc = np.zeros(ns, dtype=np.int64); zc = np.zeros(ns, dtype=np.int64)
        self._call("get_tape_sizes", self.h, K._ip(uc, C.c_int64), K._ip(zc, C.c_int64))
        u = np.zeros(max(int(uc.sum()), 1)); z = np.zeros(max(int(zc.sum()), 1))
        self._call("get_tapes", self.h, K._dp(u), K._dp(z))
        u_off = np.concatenate([[0], np.cumsum(uc)]).astype(np.int64)
        z_off = np.concatenate([[0], np.cumsum(zc)]).astype(np.int64)
        return u[:int(uc.sum())], u_off, z[:int(zc.sum())], z_off

    def get_tape_marks(self):
        """draws each stream had consumed at the start of every recorded PT step: (u_mark, z_mark), [n_steps][n_streams]"""
        ns = self.n_chains + self.cfg.n_ladders
        n = C.c_int64()
        self._call("get_tape_marks", self.h, C.byref(n), None, None)
        um = np.zeros((max(n.value, 1), ns), dtype=np.int64); zm = np.zeros_like(um)
        self._call("get_tape_marks", self.h, None, K._ip(um, C.c_int64), K._ip(zm, C.c_int64))
        return um[:n.value], zm[:n.value]

    def eval_loglike(self, x):
        x = K._f64(x).reshape(-1, self.dim); out = np.empty(len(x))
        self._call("eval_loglike", self.h, K._dp(x), C.c_int64(len(x)), K._dp(out)); return out

    def eval_logprior(self, x):
        x = K._f64(x).reshape(-1, self.dim); out = np.empty(len(x))
        self._call("eval_logprior", self.h, K._dp(x), C.c_int64(len(x)), K._dp(out)); return out


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "ref_trace"))


def run_ref_trace(spec, steps, tmpdir):
    out = os.path.join(str(tmpdir), "ref.bin")
    args = [os.path.join(REF_DIR, "ref_trace")] + spec.ref_args(str(tmpdir), steps, out)
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return out
