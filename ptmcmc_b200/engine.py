"""`Engine`: one ptg handle (one GPU) behind the C ABI, numpy in / numpy out."""
import ctypes as C
import numpy as np
from . import _capi as K
from ._lib import load


class Engine(K.CApi):
    """The CUDA engine.  Raises if the shared library or a CUDA device is missing (no CPU fallback)."""

    def __init__(self, cfg):
        super().__init__(load(), "ptg_", cfg)

    def register_evaluate_log(self, fn):
        """host-callback likelihood: fn(x[n, dim]) -> loglike[n], called once per PT iteration with every gated proposal
        (the batched form of bayes_likelihood::register_evaluate_log, bayesian.hh:544-552)"""
        CB = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double))
        d = self.dim

        def trampoline(_user, xp, n, outp):
            x = np.ctypeslib.as_array(xp, shape=(n, d))
            out = np.ctypeslib.as_array(outp, shape=(n,))
            out[:] = fn(x)

        self._cb = CB(trampoline)  # keep alive
        self._call("register_evaluate_log", self.h, self._cb, None)

    def synchronize(self):
        self._call("synchronize", self.h)

    def select_kernel(self, kernel):
        """K.KERNEL_AUTO / KERNEL_SHARED / KERNEL_WARP / KERNEL_FAST (include/ptmcmc_b200.h)"""
        self._call("select_kernel", self.h, C.c_int32(kernel))

    def launch_count(self):
        n = C.c_int64(); self._call("get_launch_count", self.h, C.byref(n)); return n.value

    def set_stream(self, cuda_stream_ptr):
        self._call("set_stream", self.h, C.c_void_p(cuda_stream_ptr))

    def inject_tape_marks(self, u_mark, z_mark):
        um = np.ascontiguousarray(u_mark, dtype=np.int64); zm = np.ascontiguousarray(z_mark, dtype=np.int64)
        assert um.shape == zm.shape and um.shape[1] == self.n_chains + self.cfg.n_ladders
        self._call("inject_tape_marks", self.h, C.c_int64(um.shape[0]), K._ip(um, C.c_int64), K._ip(zm, C.c_int64))

    def eval_loglike(self, x):
        x = K._f64(x).reshape(-1, self.dim); out = np.empty(len(x))
        self._call("eval", self.h, K._dp(x), C.c_int64(len(x)), K._dp(out), None)
        return out

    def eval_logprior(self, x):
        x = K._f64(x).reshape(-1, self.dim); out = np.empty(len(x))
        self._call("eval", self.h, K._dp(x), C.c_int64(len(x)), None, K._dp(out))
        return out

    def get_act(self, rung, n_last, max_lag):
        """integrated autocorrelation times [n_ladders, dim] of rung `rung`, computed on the device"""
        out = np.empty((self.cfg.n_ladders, self.dim))
        self._call("get_act", self.h, C.c_int32(rung), C.c_int32(n_last), C.c_int32(max_lag), K._dp(out))
        return out

    def report_effective_samples(self, ladder=0, rung=0, esslimit=-1, imax=-1):
        """(ess, useful length) of one chain by the reference's own recipe, called the way the run loop calls it
        (chain::report_effective_samples(-1, save_every*1000, save_every, esslimit), ptmcmc.cc:645; restated in
        analysis.report_effective_samples and pinned to the reference build's outputs).  Host-side analysis of the history the
        ring still holds: the whole run must be resident (hist_capacity >= n_init + steps / save_every)."""
        from .analysis import report_effective_samples
        c = self.get_counters()
        i = ladder * self.cfg.n_rungs + rung
        nsize, nhist = int(c["nsize"][i]), int(c["nhist"][i])
        cap = self.hist_capacity()
        if nsize > cap:
            raise RuntimeError("report_effective_samples: the ring has wrapped (%d records > capacity %d)" % (nsize, cap))
        x = self.get_history(ladder, rung, 0, nsize, full=False)["x"]
        se = self.cfg.save_every
        return report_effective_samples(x, nhist, n_init=self.cfg.n_init, add_every=se,
                                        width=se * 1000, every=se, esslimit=esslimit, imax=imax)

    def get_autocovar_windows(self, rung, swidth, n_win, lag_rec, n_feat, end_rec=None):
        """device part of the reference's ESS recipe -> means, covar [n_ladders, n_feat, n_win, n_lag] (ptg_get_autocovar_windows)"""
        L = self.cfg.n_ladders
        lag_rec = np.ascontiguousarray(lag_rec, dtype=np.int32)
        shape = (L, n_feat, n_win, len(lag_rec))
        means, covar = np.empty(shape), np.empty(shape)
        end = None if end_rec is None else np.ascontiguousarray(end_rec, dtype=np.int64)
        self._call("get_autocovar_windows", self.h, C.c_int32(rung), C.c_int32(swidth), C.c_int32(n_win), C.c_int32(len(lag_rec)),
                   lag_rec.ctypes.data_as(C.POINTER(C.c_int32)), None if end is None else end.ctypes.data_as(C.POINTER(C.c_int64)),
                   C.c_int32(n_feat), K._dp(means), K._dp(covar))
        return means, covar

    def report_effective_samples_all(self, rung=0, window_records=None, imax=-1):
        """(ess[n_ladders], useful length[n_ladders]) of rung `rung` of EVERY ladder by the reference's recipe
        (chain::report_effective_samples(-1, 1000 save_every, save_every), esslimit < 0): window statistics on the device, in the
        reference's summation order, combination on the host.  window_records: treat the newest that many stored records of each
        chain as the chain (a wrapped ring); default: the whole run, which must still be resident."""
        from .analysis import recipe_geometry, effective_samples_from_windows
        L, R, se = self.cfg.n_ladders, self.cfg.n_rungs, self.cfg.save_every
        nf = self.dim if imax < 0 else min(imax, self.dim)
        nf = min(nf, 20)
        c = self.get_counters()
        idx = np.arange(L) * R + rung
        if window_records is not None:
            nstep = np.full(L, int(window_records) * se, dtype=np.int64)
            end = None
        else:
            nstep = c["nhist"][idx].astype(np.int64)
            end = self.cfg.n_init + nstep // se
        ess, length = np.zeros(L), np.zeros(L, dtype=np.int64)
        geo = {}
        for l in range(L):
            geo.setdefault(recipe_geometry(int(nstep[l]), se * 1000, se)[:3], []).append(l)
        for (width, swidth, n_win), members in geo.items():
            if n_win < 1:
                continue
            lags = recipe_geometry(int(nstep[members[0]]), se * 1000, se)[3]
            means, covar = self.get_autocovar_windows(rung, swidth, n_win, [g // se for g in lags], nf, end)
            e, nw = effective_samples_from_windows(covar[members], means[members], lags, width, se, swidth)
            ess[members], length[members] = e, nw * width
        return ess, length

    def get_mean_loglike(self, n_last):
        out = np.empty(self.n_chains)
        self._call("get_mean_loglike", self.h, C.c_int32(n_last), K._dp(out))
        return out

    def get_log_evidence(self, n_last):
        out = np.empty(self.cfg.n_ladders)
        self._call("get_log_evidence", self.h, C.c_int32(n_last), K._dp(out))
        return out

    def get_lprior(self):
        out = np.empty(self.n_chains)
        self._call("get_lprior", self.h, K._dp(out))
        return out

    def set_current(self, x, lpost, llike, lprior):
        x, lpost, llike, lprior = K._f64(x), K._f64(lpost), K._f64(llike), K._f64(lprior)
        self._keep = [x, lpost, llike, lprior]  # async copy: keep the buffers alive
        self._call("set_current", self.h, K._dp(x), K._dp(lpost), K._dp(llike), K._dp(lprior))

    def step_host(self, n_steps, n_out, x_out, lpost_out, llike_out):
        self._call("step_host", self.h, C.c_int64(n_steps), C.c_int32(n_out), K._dp(x_out), K._dp(lpost_out), K._dp(llike_out))

    def step_host_begin(self, n_steps, n_out, x_out, lpost_out, llike_out):
        """enqueue n_steps iterations + the copy of every ladder's newest n_out cold samples to (pinned) host arrays; returns at once"""
        self._call("step_host_begin", self.h, C.c_int64(n_steps), C.c_int32(n_out), K._dp(x_out), K._dp(lpost_out), K._dp(llike_out))

    def step_host_wait(self):
        self._call("step_host_wait", self.h)

    def hist_capacity(self):
        """effective ring capacity (a config value of 0 means n_init + 1024)"""
        c = C.c_int32(); self._call("get_hist_capacity", self.h, C.byref(c)); return c.value

    def checkpoint(self, path):
        self._call("checkpoint", self.h, path.encode())

    def restore(self, path):
        self._call("restore", self.h, path.encode())
