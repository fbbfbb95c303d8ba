"""ctypes binding of the C ABI declared in include/ptmcmc_b200.h.

`CApi(lib, prefix)` wraps a loaded shared library whose entry points are `<prefix>create`, `<prefix>step`, ...
The product uses prefix ``ptg_`` (ptmcmc_b200/csrc/libptmcmc_b200.so, the CUDA engine).  The test-suite binds
the CPU oracle (oracle/libptmcmc_oracle.so, prefix ``pto_``) through the same class so that parity tests run
one call sequence on both; nothing in this package ever loads the oracle.
"""
import ctypes as C
import numpy as np

ABI_VERSION = 1
MAX_DIM = 128
MAX_PROPOSALS = 16
MAX_RUNGS = 64

# enums (include/ptmcmc_b200.h)
BOUND_OPEN, BOUND_LIMIT, BOUND_REFLECT, BOUND_WRAP = 0, 1, 2, 3
PRIOR_UNIFORM, PRIOR_GAUSSIAN, PRIOR_POLAR, PRIOR_COPOLAR, PRIOR_LOG, PRIOR_GAUSSIAN_WRAPPED = 1, 2, 3, 4, 5, 6
LIKE_FLAT, LIKE_GAUSS_ISO, LIKE_SINES, LIKE_POLY_CHI2, LIKE_SINUSOID_CHI2, LIKE_GAUSS_FULLCOV, LIKE_HOST_CALLBACK, LIKE_SHELL2D, LIKE_SHELLS = 0, 1, 2, 3, 4, 5, 6, 7, 8
PROP_DE, PROP_GAUSS, PROP_PRIOR_DRAW = 1, 2, 3
SWAP_REFERENCE, SWAP_EVEN_ODD = 0, 1
RNG_PHILOX, RNG_TAPE = 0, 1
RECORD_BASIC, RECORD_FULL = 0, 1
KERNEL_AUTO, KERNEL_SHARED, KERNEL_WARP, KERNEL_FAST, KERNEL_FAST_GENERAL = 0, 1, 2, 3, 4
TRACE_TYPE_MASK, TRACE_ACCEPT, TRACE_INVALID, TRACE_SWAPPED, TRACE_NOLIKE = 0xFF, 0x100, 0x200, 0x400, 0x800


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("n_ladders", C.c_int32), ("n_rungs", C.c_int32),
        ("dim", C.c_int32), ("save_every", C.c_int32), ("hist_capacity", C.c_int32), ("n_init", C.c_int32),
        ("swap_mode", C.c_int32), ("rng_mode", C.c_int32), ("record_level", C.c_int32), ("trace_steps", C.c_int32),
        ("swap_rate", C.c_double), ("Tmax", C.c_double), ("dprior_min", C.c_double), ("evolve_rate", C.c_double),
        ("evolve_lpost_cut", C.c_double), ("seed", C.c_uint64), ("ladder_offset", C.c_int64),
    ]


class Proposal(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("reserved", C.c_int32), ("share", C.c_double), ("hot_share", C.c_double),
        ("snooker", C.c_double), ("gamma_one_frac", C.c_double), ("b_small", C.c_double), ("ignore_frac", C.c_double),
        ("unlikely_alpha", C.c_double), ("reduce_gamma", C.c_double), ("one_d_frac", C.c_double),
        ("sigmas", C.POINTER(C.c_double)), ("transform", C.POINTER(C.c_double)),
    ]


def make_config(n_ladders, n_rungs, dim, *, n_init=None, save_every=1, hist_capacity=0, swap_mode=SWAP_REFERENCE,
                rng_mode=RNG_PHILOX, record_level=RECORD_FULL, trace_steps=0, swap_rate=0.1, Tmax=1e9,
                dprior_min=-30.0, evolve_rate=0.0, evolve_lpost_cut=-1.0, seed=0xB2000003, ladder_offset=0, device=0):
    if n_init is None:
        n_init = 50 * dim  # de_ni * Npar, ptmcmc.cc:86
    return Config(ABI_VERSION, device, n_ladders, n_rungs, dim, save_every, hist_capacity, n_init, swap_mode, rng_mode,
                  record_level, trace_steps, swap_rate, Tmax, dprior_min, evolve_rate, evolve_lpost_cut, seed,
                  ladder_offset)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class CApiError(RuntimeError):
    pass


class CApi:
    """One engine handle behind the C ABI.  Arrays in/out are numpy, host side."""

    def __init__(self, lib, prefix, cfg):
        self.lib, self.prefix, self.cfg = lib, prefix, cfg
        self._keep = []
        self.h = C.c_void_p()
        f = self._fn("last_error"); f.restype = C.c_char_p
        self._call("create", C.byref(cfg), C.byref(self.h))
        self.n_chains = cfg.n_ladders * cfg.n_rungs
        self.dim = cfg.dim
        self.n_props = 0

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _call(self, name, *args):
        f = self._fn(name)
        f.restype = C.c_int
        rc = f(*args)
        if rc != 0:
            raise CApiError("%s%s failed (%d): %s" % (self.prefix, name, rc, self._fn("last_error")().decode()))
        return rc

    def close(self):
        if self.h:
            self._call("destroy", self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- model set-up
    def set_space(self, lower, upper, xmin, xmax):
        lo = np.ascontiguousarray(lower, dtype=np.int32); up = np.ascontiguousarray(upper, dtype=np.int32)
        a, b = _f64(xmin), _f64(xmax)
        self._call("set_space", self.h, _ip(lo, C.c_int32), _ip(up, C.c_int32), _dp(a), _dp(b))

    def set_prior(self, types, a, b):
        t = np.ascontiguousarray(types, dtype=np.int32); a, b = _f64(a), _f64(b)
        self._call("set_prior", self.h, _ip(t, C.c_int32), _dp(a), _dp(b))

    def set_likelihood(self, kind, params, data=None):
        p = _f64(params).ravel()
        d = _f64(data).ravel() if data is not None else np.zeros(0)
        self._call("set_likelihood", self.h, C.c_int32(kind), _dp(p), C.c_int32(p.size), _dp(d) if d.size else None,
                   C.c_int64(d.size))

    def set_proposals(self, props, Tpow=0.0, wrap_in_set=True):
        """props: list of dicts(kind=..., share=..., [DE fields] / [sigmas, one_d_frac, transform])"""
        arr = (Proposal * len(props))()
        self.n_props = len(props)
        self.n_nested = 0
        for i, p in enumerate(props):
            q = arr[i]
            q.kind = p["kind"]; q.share = p.get("share", 1.0); q.hot_share = p.get("hot_share", 0.0)
            q.snooker = p.get("snooker", 0.1); q.gamma_one_frac = p.get("gamma_one_frac", 0.3)
            q.b_small = p.get("b_small", 1e-4); q.ignore_frac = p.get("ignore_frac", 0.0)
            q.unlikely_alpha = p.get("unlikely_alpha", 0.0); q.reduce_gamma = p.get("reduce_gamma", 4.0)
            q.one_d_frac = p.get("one_d_frac", 0.0)
            if p["kind"] == PROP_GAUSS:
                s = _f64(p["sigmas"]); self._keep.append(s); q.sigmas = _dp(s)
                if p.get("transform") is not None:
                    t = _f64(p["transform"]); self._keep.append(t); q.transform = _dp(t)
        self._call("set_proposals", self.h, C.c_int32(len(props)), arr, C.c_double(Tpow), C.c_int32(1 if wrap_in_set else 0))

    def set_proposal_options(self, adapt_rate=0.0, de_mixing=False, de_Tmix=1.0):
        """adaptive shares of the set (proposal_distribution.cc:132-166) / temperature mixing of a bare DE proposal (:594-741)"""
        self._call("set_proposal_options", self.h, C.c_double(adapt_rate), C.c_int32(1 if de_mixing else 0), C.c_double(de_Tmix))

    def set_nested_set(self, first, count, share, hot_share=0.0, adapt_rate=0.0):
        """members [first, first + count) form one nested proposal_distribution_set in a single top-level slot (ptmcmc.cc:123-143)"""
        self._call("set_nested_set", self.h, C.c_int32(first), C.c_int32(count), C.c_double(share), C.c_double(hot_share), C.c_double(adapt_rate))
        self.n_nested = count

    def get_proposal_shares(self):
        """[n_chains, top-level slots + nested members]"""
        nn = getattr(self, "n_nested", 0)
        sh = np.empty((self.n_chains, max(self.n_props + (1 if nn else 0), 1)))
        self._call("get_proposal_shares", self.h, _dp(sh))
        return sh

    def set_betas(self, betas):
        b = _f64(betas); self._call("set_betas", self.h, _dp(b))

    def seed(self, seed):
        self._call("seed", self.h, C.c_uint64(seed))

    def inject_tapes(self, u, u_off, z, z_off):
        u, z = _f64(u), _f64(z)
        uo = np.ascontiguousarray(u_off, dtype=np.int64); zo = np.ascontiguousarray(z_off, dtype=np.int64)
        if u.size == 0: u = np.zeros(1)
        if z.size == 0: z = np.zeros(1)
        self._call("inject_tapes", self.h, _dp(u), _ip(uo, C.c_int64), _dp(z), _ip(zo, C.c_int64))

    # ---- run
    def init_from_prior(self):
        self._call("init_from_prior", self.h)

    def init_states(self, x):
        x = _f64(x); assert x.size == self.n_chains * self.cfg.n_init * self.dim
        self._call("init_states", self.h, _dp(x))

    def step(self, n):
        self._call("step", self.h, C.c_int64(n))

    # ---- read-back
    def get_current(self):
        n, d = self.n_chains, self.dim
        x = np.empty((n, d)); lp = np.empty(n); ll = np.empty(n); b = np.empty(n)
        self._call("get_current", self.h, _dp(x), _dp(lp), _dp(ll), _dp(b))
        return dict(x=x, lpost=lp, llike=ll, beta=b)

    def get_counters(self):
        n = self.n_chains
        nh, ns, nt, na = (np.empty(n, dtype=np.int64) for _ in range(4))
        lt = np.empty(n, dtype=np.int32); mp = np.empty(n)
        self._call("get_counters", self.h, _ip(nh, C.c_int64), _ip(ns, C.c_int64), _ip(nt, C.c_int64), _ip(na, C.c_int64),
                   _ip(lt, C.c_int32), _dp(mp))
        return dict(nhist=nh, nsize=ns, ntries=nt, naccept=na, last_type=lt, map_lpost=mp)

    def get_history(self, ladder, rung, first, count, full=True):
        d = self.dim
        x = np.empty((count, d)); lp = np.empty(count); ll = np.empty(count)
        acc = np.empty(count) if full else None; beta = np.empty(count) if full else None
        typ = np.empty(count, dtype=np.int32) if full else None
        self._call("get_history", self.h, C.c_int32(ladder), C.c_int32(rung), C.c_int64(first), C.c_int64(count), _dp(x),
                   _dp(lp), _dp(ll), _dp(acc), _dp(beta), _ip(typ, C.c_int32))
        return dict(x=x, lpost=lp, llike=ll, acc=acc, beta=beta, type=typ)

    def get_swap_stats(self):
        L, R = self.cfg.n_ladders, self.cfg.n_rungs
        sc = np.zeros((L, max(R - 1, 1)), dtype=np.int64); sa = np.zeros_like(sc)
        di, up, dn, ins = (np.zeros((L, R), dtype=np.int32) for _ in range(4))
        self._call("get_swap_stats", self.h, _ip(sc, C.c_int64), _ip(sa, C.c_int64), _ip(di, C.c_int32), _ip(up, C.c_int32),
                   _ip(dn, C.c_int32), _ip(ins, C.c_int32))
        return dict(swap_count=sc[:, :R - 1], swap_accept=sa[:, :R - 1], directions=di, ups=up, downs=dn, instances=ins)

    def get_trace(self, first, count):
        lhr = np.empty((count, self.n_chains)); code = np.empty((count, self.n_chains), dtype=np.int32)
        self._call("get_trace", self.h, C.c_int64(first), C.c_int64(count), _dp(lhr), _ip(code, C.c_int32))
        return lhr, code

    # ---- rung-sharded ladders (buffers are raw addresses: device memory for the engine, host memory for the test oracle)
    def boundary_pack(self, rung, out_ptr):
        self._call("boundary_pack", self.h, C.c_int32(rung), C.c_void_p(out_ptr))

    def boundary_swap(self, my_rung, neighbour_pack_ptr, i_am_lower, shared_seed, boundary_id, exchange_index):
        self._call("boundary_swap", self.h, C.c_int32(my_rung), C.c_void_p(neighbour_pack_ptr), C.c_int32(1 if i_am_lower else 0),
                   C.c_uint64(shared_seed), C.c_int64(boundary_id), C.c_int64(exchange_index))

    # ---- the same exchange fused into the production step kernel over peer memory (engine only)
    def xchg_export(self):
        """-> (64-byte CUDA IPC handle, device pointer) of this engine's exchange area"""
        hd = C.create_string_buffer(64); ptr = C.c_void_p()
        self._call("xchg_export", self.h, hd, C.byref(ptr))
        return hd.raw, ptr.value

    def xchg_connect(self, colder, hotter, shared_seed, colder_boundary_id, hotter_boundary_id):
        """colder / hotter: the neighbour's IPC handle (bytes, another process), its device pointer (int, same process) or None"""
        ipc = isinstance(colder, (bytes, bytearray)) or isinstance(hotter, (bytes, bytearray))
        def arg(v):
            if v is None:
                return None
            return C.c_char_p(bytes(v)) if ipc else C.c_void_p(v)
        self._keep_xchg = (arg(colder), arg(hotter))
        self._call("xchg_connect", self.h, self._keep_xchg[0], self._keep_xchg[1], C.c_int32(1 if ipc else 0), C.c_uint64(shared_seed),
                   C.c_int64(colder_boundary_id), C.c_int64(hotter_boundary_id))

    def xchg_abort(self):
        """watchdog: every boundary wait of this engine gives up (callable from another thread while synchronize blocks)"""
        self._call("xchg_abort", self.h)

    def step_exchange(self, n_steps, apply_pending, publish, every=0):
        self._call("step_exchange", self.h, C.c_int64(n_steps), C.c_int32(every), C.c_int32(1 if apply_pending else 0), C.c_int32(1 if publish else 0))

    def get_total_steps(self):
        t = C.c_int64(); self._call("get_total_steps", self.h, C.byref(t)); return t.value
