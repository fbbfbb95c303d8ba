#!/usr/bin/env python
"""bench.py -- tempered chain-steps/s of the PT hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1_sines] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path over the whole resident batch: `--pt-steps` iterations of
parallel_tempering_chains::step (chain.cc:1393) for every ladder on the GPU, in ONE launch of the fused step kernel.
`value`  = tempered chain-steps (history appends = Nhist increments) of all ranks / device time (CUDA
           events on the engine's stream, max over ranks), state resident in HBM.
`e2e`    = the same metric through the C-ABI calls a host facade makes per block with HOST buffers inside the timed
           region: ptg_set_current (H2D of every chain's state from pinned memory) -> ptg_step_host_begin (steps + D2H of
           EVERY cold-chain sample the block produced, what the reference's run loop dumps, ptmcmc.cc:601-616) with two
           pinned buffers, so the copy of block k overlaps the kernel of block k+1.
`sustained` = the device-resident measurement repeated over >= 3 s with the clocks sampled: the number a warm part holds.
`workloads` (1 GPU) = short runs of BASELINE.json's other configs (A, B, C2, D), each with value, e2e and both rooflines;
`config5`, `rung_sharded` (N > 1) = config 5's batch (2048 ladders x 32 rungs per GPU) and the rung-sharded layout with the
           boundary exchange fused into the step kernel over NVLink peer memory (K = 10 iterations per exchange).
The reference arm (`--impl reference`) times the UNMODIFIED reference (oracle/_ref/ref_trace: our driver compiled
against the reference sources, stepping parallel_tempering_chains through its public API) on the host cores, one
ladder per process, all cores busy -- the only CPU parallelism that scales for this code.
Weak scaling: every rank owns `ladders` whole ladders (ladder_offset = rank * ladders); no data-path collective.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---------------------------------------------------------------------------------------------------- workloads
WORKLOADS = {
    # BASELINE.json configs[2] / SURVEY 8d C1: sines.hh as written, d=3, ks=(2,2,2), default select_proposal mix
    # hist = 8192 slots per chain (43 GB): a short ring makes DE adapt to the chain's own recent past and biases the sampler
    # (measured: variance -6 % at 1024 slots); the reference's unbounded history is the limit of a long ring
    "c1_sines": dict(model="sines", dim=3, rungs=32, ladders=4096, pt_steps=1000, hist=8192, save_every=1, f_de=0.8, f_sn=0.1,
                     desc="sines.hh sin^4 surface d=3, 4096 ladders x 32 rungs, default proposal mix (80% DE / 20% 6-scale Gaussian), swap_rate 0.1"),
    # configs[1] / B: polynomial chi^2, d=5, N=1000, DE only
    "b_poly": dict(model="poly", dim=5, rungs=16, ladders=1024, pt_steps=20, hist=1024, f_de=1.0, f_sn=0.1,
                   flops=dict(per_chain_step=1000 * 12.0, peak_key="fp64_dfma_tflops", what="1000 points x (Horner on FMA 2(d-1) + residual 2 + weighted accumulate 2)"),
                   limiter="fp64 pipe + L1 loads of the data block (warp-per-chain production functor, ptg_wide_mma.cuh)",
                   desc="5-coefficient polynomial chi^2 over 1000 points, 1024 ladders x 16 rungs, DE proposals"),
    # C2: 3-sinusoid chi^2 over 1e4 samples, d=9
    "c2_sinusoid": dict(model="sinusoid", dim=9, rungs=32, ladders=4096, pt_steps=1, hist=512, f_de=0.8, f_sn=0.1,
                        flops=dict(per_chain_step=1e4 * 28.0, peak_key="fp64_dfma_tflops", what="1e4 samples x (3 components x (rotation 6 + amplitude 2) + residual 4): sin / cos advance by rotations on the uniform grid, re-anchored every 64 samples"),
                        limiter="fp64 pipe (rotation recurrences of the sinusoid functor, ptg_wide_mma.cuh)",
                        desc="3-sinusoid chi^2 fit to 1e4 samples d=9, 4096 ladders x 32 rungs, default proposal mix"),
    # configs[3] / D: correlated Gaussian d=100, full covariance; "65536 chains x 24 rungs" read as 65 544 chains in total
    # (2731 ladders x 24 rungs; the 1.57 M-chain reading leaves < 100 history slots per chain in 180 GB, SURVEY.md 8d, and the
    # reference's DE member needs >= 10 d = 1000 stored samples to be ready).  Ninit = 11 d prior draws per chain (de_ni = 11).
    # save_every = 8: the 1280-slot ring then spans 10 240 PT iterations, which removes the short-window bias at d = 100 without
    # 8x the memory ("D must run with a short ring and/or save_every >> 1")
    # Prior box +-10 sqrt(C_ii), not the example's +-100: at d = 100 the reference's log(prod_i pdf_i) underflows to -inf for the wider box
    # and the likelihood is never evaluated (ptmcmc_b200/workloads.py fullcov_spec); burn_in PT iterations run before the warm-up so that
    # the timed region samples the posterior (DE proposals from a posterior history pass the prior gate) instead of the start-up transient.
    "d_fullcov": dict(model="fullcov", dim=100, rungs=24, ladders=2731, pt_steps=50, hist=1280, save_every=8, f_de=0.5, f_sn=0.1, burn_in=4000,
                      flops=dict(per_chain_step=2 * 100 * 100 * 1.5, peak_key="fp64_dmma_tflops", what="quadratic form 2 d^2 + eigen-rotation 2 d^2 on the 50 % Gaussian proposals"),
                      desc="correlated Gaussian d=100 full covariance (DMMA batched quadratic form + proposal rotation), 2731 ladders x 24 rungs, "
                           "50% eigen-rotated Gaussian proposal (2.38^2/d C) + 50% DE"),
    # configs[0] / A as a throughput batch
    "a_gauss": dict(model="gauss", dim=2, rungs=8, ladders=16384, pt_steps=1000, hist=1024, f_de=0.8, f_sn=0.1,
                    desc="2-D isotropic Gaussian, 16384 ladders x 8 rungs, default proposal mix"),
}


def make_spec(w):
    spec = _make_spec(w)
    spec.save_every = w.get("save_every", 1)
    return spec


def _make_spec(w):
    from ptmcmc_b200.workloads import Spec, poly_data, sinusoid_spec
    if w["model"] == "sines":
        return Spec("sines", w["dim"], w["rungs"])
    if w["model"] == "poly":
        return Spec("poly", 5, w["rungs"], centers=np.zeros(5), halfwidths=np.full(5, 10.0), prop="de", Tmax=1e6, extra=poly_data())
    if w["model"] == "sinusoid":
        return sinusoid_spec(w["rungs"], n=10000)
    if w["model"] == "gauss":
        return Spec("gauss", 2, w["rungs"], centers=[2, -3], halfwidths=[2, 3])
    if w["model"] == "fullcov":
        from ptmcmc_b200.workloads import fullcov_spec
        return fullcov_spec(w["dim"], w["rungs"], de_ni=11)
    raise ValueError(w["model"])


def algorithmic_bytes_per_chain_step(w):
    """history append 8(d+2)/s [x, lpost, llike] + DE gathers f_DE (2 + f_sn) 8 d"""
    d, save_every = w["dim"], w.get("save_every", 1)
    return 8.0 * (d + 2) / save_every + w["f_de"] * (2 + w["f_sn"]) * 8.0 * d


# ---------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.p = index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.p:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(2)
        except subprocess.TimeoutExpired:
            self.p.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)), samples=len(sm), reasons=sorted(reasons))


# ---------------------------------------------------------------------------------------------------- CPU reference arm
def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_sample(w, pt_steps, procs, ess=False):
    """`procs` independent processes, each one ladder of the workload stepped `pt_steps` times by the reference itself
    (oracle/_ref/ref_trace) or, when that binary is absent, by the oracle port.
    Returns dict(chain_steps, seconds [max over processes of the time inside the stepping loop], wall, kind, ess [sum over ladders of
    the reference's own chain::report_effective_samples, or None])."""
    spec = make_spec(w)
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_trace")
    if os.path.exists(ref):
        with tempfile.TemporaryDirectory() as td:
            cmds = []
            for i in range(procs):
                spec.seed = 0.05 + 0.9 * (i + 0.5) / procs
                pd = os.path.join(td, "p%d" % i); os.makedirs(pd)
                cmds.append([ref] + spec.ref_args(pd, pt_steps, "/dev/null") + (["ess=1"] if ess else []))
            t0 = time.perf_counter()
            ps = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for c in cmds]
            outs = [p.communicate()[0] for p in ps]
            wall = time.perf_counter() - t0
        total, secs, ess_sum, ess_n = 0, 0.0, 0.0, 0
        for o, p in zip(outs, ps):
            if p.returncode != 0:
                raise RuntimeError("ref_trace failed")
            total += int(o.split("total_Nhist=")[1].split()[0])
            secs = max(secs, float(o.split("step_seconds=")[1].split()[0]))
            if ess and "ref_ess: ess=" in o:
                v = float(o.split("ref_ess: ess=")[1].split()[0])
                if v == v and v > 0:
                    ess_sum += v; ess_n += 1
        return dict(chain_steps=total, seconds=secs, wall=wall, kind="reference", ess=ess_sum if ess_n else None, ess_ladders=ess_n)
    import multiprocessing as mp
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(procs) as pool:
        totals = pool.map(_port_worker, [(w, pt_steps, 0.05 + 0.9 * (i + 0.5) / procs) for i in range(procs)])
    dt = time.perf_counter() - t0
    return dict(chain_steps=sum(totals), seconds=dt, wall=dt, kind="port", ess=None, ess_ladders=0)


def _port_worker(args):
    w, pt_steps, seed = args
    from tests.oracle_binding import Oracle
    spec = make_spec(w)
    o = Oracle(spec.config(n_ladders=1, rng_mode=2))
    spec.setup(o); o.seed_newran(seed); o.init_from_prior(); o.step(pt_steps)
    return o.get_total_steps()


def reference_pt_steps(w, seconds=8.0):
    """PT iterations per process for roughly `seconds` of CPU work (1.6e5 chain-steps/s/core measured for A/C1-like
    models; chi^2 models scale with the data size)"""
    per_chain_step = {"sines": 6e-6, "gauss": 6e-6, "poly": 1.2e-5, "sinusoid": 3e-4, "fullcov": 6e-5}[w["model"]]
    return max(20, int(seconds / (per_chain_step * w["rungs"])))


def cpu_baseline(w, seconds, procs, unit):
    """the unmodified reference on `procs` host cores (one ladder per process), chain-steps/s and -- from the reference's own
    report_effective_samples on every process's cold chain -- ESS/s"""
    n_pt = reference_pt_steps(w, seconds)
    r = run_reference_sample(w, n_pt, procs, ess=True)
    out = dict(value=r["chain_steps"] / r["seconds"], unit=unit, cores=procs, kind=r["kind"],
               sample="%d processes x 1 ladder x %d rungs x %d PT iterations of the same workload (timed: the stepping loop, max over processes)" % (procs, w["rungs"], n_pt))
    if r["ess"] is not None:
        out["ess"] = dict(value=r["ess"] / r["seconds"], unit="ESS/s", ladders_with_estimate=r["ess_ladders"],
                          estimator="chain::report_effective_samples(-1, 1000 save_every, save_every) of the reference itself on each process's cold chain (ref_trace ess=1, chain.cc:549-643), summed")
    return out


# ---------------------------------------------------------------------------------------------------- one workload on this rank's GPU
def peaks():
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        return float(json.load(open(pp))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def fp64_peaks():
    for name in ("r02_fp64_peaks.json", "r01_fp64_peaks.json"):
        pp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(pp):
            return json.load(open(pp)), "measured: profiles/" + name
    return None, None


def measure(name, w, steps, warmup, rank, world, local, swap_mode="reference", e2e_steps=None, sustain_s=0.0, with_ess=False, with_e2e=True):
    """device-resident throughput, optional >= sustain_s seconds repeat, ESS, end-to-end blocks and rooflines of one workload"""
    import torch
    import torch.distributed as dist
    from ptmcmc_b200 import _capi as K
    from ptmcmc_b200.engine import Engine
    from ptmcmc_b200.sharding import max_over_ranks, sum_over_ranks

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
    spec = make_spec(w)
    L, R, d, S = w["ladders"], w["rungs"], w["dim"], w["pt_steps"]
    sm = K.SWAP_REFERENCE if swap_mode == "reference" else K.SWAP_EVEN_ODD
    eng = Engine(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, record_level=K.RECORD_BASIC, hist_capacity=w["hist"], save_every=w["save_every"], swap_mode=sm,
                             device=local, ladder_offset=rank * L, seed=0xB2000003))
    spec.setup(eng)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    eng.init_from_prior()
    eng.synchronize()
    res = dict(name=name)
    with torch.cuda.stream(stream):
        if w.get("burn_in"):
            eng.step(int(w["burn_in"]))
        for _ in range(warmup):
            eng.step(S)
        eng.synchronize()
        l0 = eng.launch_count()
        n0 = eng.get_total_steps()
        clocks = ClockSampler(local); clocks.start()
        barrier(); torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(steps):
            eng.step(S)
        ev1.record(stream)
        torch.cuda.synchronize(); barrier()
        ms = ev0.elapsed_time(ev1)
        clk = clocks.stop()
        eng.synchronize()
        n1 = eng.get_total_steps()
        launches = eng.launch_count() - l0
        ms_max = max_over_ranks(ms)
        res.update(value=sum_over_ranks(n1 - n0) / (ms_max * 1e-3), ms_per_step=ms_max / steps, clocks=clk, gpu_launches=int(launches), steps=steps,
                   chain_steps_per_launch=(n1 - n0) / max(launches, 1), launch_ms=ms / max(launches, 1))
        # ---- the same loop over >= sustain_s seconds: what the part holds once it is warm
        if sustain_s > 0:
            reps = max(steps, int(np.ceil(sustain_s * 1e3 / (ms / steps))))
            m0 = eng.get_total_steps()
            cs = ClockSampler(local); cs.start()
            barrier(); torch.cuda.synchronize()
            ev0.record(stream)
            for _ in range(reps):
                eng.step(S)
            ev1.record(stream)
            torch.cuda.synchronize(); barrier()
            sms = max_over_ranks(ev0.elapsed_time(ev1))
            sclk = cs.stop()
            eng.synchronize()
            m1 = eng.get_total_steps()
            res["sustained"] = dict(value=sum_over_ranks(m1 - m0) / (sms * 1e-3), unit="chain-steps/s", seconds=sms * 1e-3, steps=reps, clocks=sclk)
    if with_ess:
        res["ess"] = ess_report(eng, w, L * world, res, world)
    # ---- end to end through the C ABI with host buffers: H2D of every chain's state, steps, D2H of every cold sample of the block
    if with_e2e:
        n_chains = L * R
        n_out = max(1, min(S // w["save_every"], w["hist"]))
        pin = lambda *shape: torch.empty(*shape, dtype=torch.float64).pin_memory().numpy()
        hx, hlp, hll, hpr = pin(n_chains, d), pin(n_chains), pin(n_chains), pin(n_chains)
        bufs = [(pin(L, n_out, d), pin(L, n_out), pin(L, n_out)) for _ in range(2)]
        cur = eng.get_current(); hx[:] = cur["x"]; hlp[:] = cur["lpost"]; hll[:] = cur["llike"]; hpr[:] = eng.get_lprior()
        ne = e2e_steps or max(3, min(steps, 30))
        with torch.cuda.stream(stream):
            for i in range(2):
                eng.set_current(hx, hlp, hll, hpr); eng.step_host_begin(S, n_out, *bufs[i])
            eng.step_host_wait(); eng.synchronize()
            m0 = eng.get_total_steps()
            barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter(); ev0.record(stream)
            for i in range(ne):
                eng.set_current(hx, hlp, hll, hpr)
                eng.step_host_begin(S, n_out, *bufs[i & 1])
            eng.step_host_wait()
            ev1.record(stream); torch.cuda.synchronize(); barrier()
            e2e_ms = max(ev0.elapsed_time(ev1), 1e3 * (time.perf_counter() - t0))
            eng.synchronize()
            m1 = eng.get_total_steps()
        assert not np.isnan(bufs[(ne - 1) & 1][0]).any(), "cold samples missing from the end-to-end read-back"
        res["e2e"] = dict(value=sum_over_ranks(m1 - m0) / (max_over_ranks(e2e_ms) * 1e-3), unit="chain-steps/s", h2d_bytes_per_step=n_chains * (d + 3) * 8,
                          d2h_bytes_per_step=L * n_out * (d + 2) * 8, steps=ne, cold_samples_per_ladder_per_step=n_out,
                          how="ptg_set_current + ptg_step_host_begin per block, two pinned buffers, copy of block k overlaps kernel of block k+1")
        res["_last_cold"] = bufs[(ne - 1) & 1][0]
    # ---- rooflines of the dominant kernel (one launch per bench step)
    peak, peak_src = peaks()
    bpcs = algorithmic_bytes_per_chain_step(w)
    achieved = bpcs * res["chain_steps_per_launch"] / (res["launch_ms"] * 1e-3) / 1e9
    traffic, traffic_src = None, None
    for tname in ("r02_traffic_%s.json" % name, "traffic_%s.json" % name):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            t = json.load(open(tpath))  # ncu --set full capture of the same kernel, dram bytes per chain-step, scaled to this launch
            if "dram_bytes_per_chain_step" in t:
                traffic = t["dram_bytes_per_chain_step"] * res["chain_steps_per_launch"]
            else:
                traffic = t["dram_bytes_per_pt_iteration"] * S * (L * R) / float(t.get("chains_profiled", 131072))
            traffic_src = "profiles/" + tname
            break
    kernel_name = "ptg_xpstep_kernel" if w["model"] == "fullcov" and d > 16 else ("ptg_xstep_kernel" if d > 16 else ("ptg_fstep_kernel" if R <= 32 else "ptg_step_kernel"))
    res["roofline"] = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic, traffic_source=traffic_src, kernel=kernel_name,
                           algorithmic_bytes_per_chain_step=bpcs, chain_steps_per_launch=res["chain_steps_per_launch"], launch_ms=res["launch_ms"], peak_source=peak_src,
                           binding_limiter=w.get("limiter", "instruction issue (fp64 libm + Philox integer work), see profiles/README.md"))
    fl = w.get("flops")
    pk, pk_src = fp64_peaks()
    if fl and pk:
        ach = fl["per_chain_step"] * res["chain_steps_per_launch"] / (res["launch_ms"] * 1e-3) / 1e12
        res["roofline_fp64"] = dict(bound="fp64", achieved=ach, peak=float(pk[fl["peak_key"]]), unit="TFLOP/s", frac=ach / float(pk[fl["peak_key"]]),
                                    flops_per_chain_step=fl["per_chain_step"], counted=fl["what"], peak_source=pk_src + " " + fl["peak_key"])
    eng.close()
    return res


def ess_report(eng, w, ladders_total, res, world):
    """ESS/s (second half of BASELINE.json's metric): the reference's own estimator (chain::report_effective_samples as the run loop calls
    it, ptmcmc.cc:645) for EVERY ladder's cold chain over the newest ring window -- lag statistics on the device in the reference's
    summation order (ptg_get_autocovar_windows), combination on the host -- with a Sokal-window cross-check"""
    try:
        L, R = w["ladders"], w["rungs"]
        cnt = eng.get_counters()
        pt_iter_per_s = w["pt_steps"] / (res["ms_per_step"] * 1e-3)
        nh = int(min(w["hist"] - 8, 2000, cnt["nsize"][::R].min()))
        tau_dev = eng.get_act(0, nh, min(nh // 2, 1000))
        eps_dev = float(np.mean(1.0 / tau_dev.max(axis=1)))
        nr = int(min(w["hist"] - 8, 8000, cnt["nsize"][::R].min()))
        eps_recipe, n_recipe = None, 0
        ess_l, len_l = eng.report_effective_samples_all(rung=0, window_records=nr)
        ok = len_l > 0
        n_recipe = int(ok.sum())
        if n_recipe:
            eps_recipe = float(np.mean(ess_l[ok] / len_l[ok]))      # ESS per PT iteration, mean over ladders
        sokal = eps_dev / w["save_every"] * pt_iter_per_s * ladders_total
        out = dict(value=(eps_recipe * pt_iter_per_s * ladders_total) if eps_recipe else sokal, unit="ESS/s",
                   estimator=("the reference's chain::report_effective_samples(-1, 1000 save_every, save_every) (chain.cc:457-643) on every cold chain's newest "
                              "%d records: lag statistics on the device (ptg_get_autocovar_windows), min over parameters, summed over ladders" % nr) if eps_recipe
                   else "Sokal window (see sokal_device); the ring window is too short for the reference recipe",
                   ess_per_pt_iteration=eps_recipe, ladders=ladders_total, ladders_with_estimate=n_recipe * world, pt_iterations_per_s=pt_iter_per_s,
                   sokal_device=dict(value=sokal, tau_pt_iterations=float(np.median(tau_dev.max(axis=1))) * w["save_every"], window=nh,
                                     estimator="ptg_get_act: per cold chain N/tau, Sokal-windowed integrated autocorrelation time, min over parameters, mean over ladders"))
        if nh > int(cnt["nsize"][::R].min()) - eng.cfg.n_init:
            out["note"] = "the run is shorter than the analysis window: it still contains start-up prior draws, the ESS figure is not meaningful"
        return out
    except Exception as exc:  # analysis is not part of the timed path
        return dict(value=None, error=repr(exc))


def brief(r):
    """what a sub-measurement contributes to the main JSON line"""
    keep = ("value", "ms_per_step", "steps", "gpu_launches", "e2e", "roofline", "roofline_fp64", "clocks", "sustained", "ess", "config")
    return {k: r[k] for k in keep if k in r}


def workload_config(name, w, swap_mode="reference"):
    extra = {}
    if w.get("burn_in"):
        extra = dict(burn_in_pt_iterations=w["burn_in"], prior="uniform box +-10 sqrt(C_ii): at +-100 the reference's log(prod pdf) underflows to -inf at d = 100 and "
                                                                 "the likelihood is never evaluated")
    return dict(workload="%s: %s" % (name, w["desc"]), ladders_per_gpu=w["ladders"], rungs=w["rungs"], dim=w["dim"], **extra,
                chains_per_gpu=w["ladders"] * w["rungs"], pt_iterations_per_step=w["pt_steps"], save_every=w["save_every"], hist_capacity=w["hist"],
                swap_mode=swap_mode, rng="philox4x32-10", parallelism="ladders sharded, %d per GPU, no data-path collective" % w["ladders"],
                l2="history ring (%.1f GB per GPU) is larger than L2" % (w["ladders"] * w["rungs"] * w["hist"] * 8.0 * (hx(w["dim"]) + 2) / 1e9))


def hx(d):
    return (d + 3) // 4 * 4 if d <= 16 else d


# ---------------------------------------------------------------------------------------------------- rung-sharded layout (N > 1)
def measure_rung_sharded(w, K_ex, steps, warmup, rank, world, local, fused=True, in_launch=False, watchdog_s=120.0):
    """ONE (world x rungs)-rung ladder family sharded by rung blocks (ptmcmc_b200/rung_sharding.py, the reference's MPI scheme
    chain.cc:1290-1311 with block assignment): boundary swaps every K_ex PT iterations, fused into the step kernel over NVLink peer
    memory (or over NCCL send/recv with fused=False)"""
    import threading
    import torch
    import torch.distributed as dist
    from ptmcmc_b200 import _capi as K
    from ptmcmc_b200.engine import Engine
    from ptmcmc_b200.sharding import max_over_ranks, sum_over_ranks
    from ptmcmc_b200.rung_sharding import RungShardedLadders, FusedRungShardedLadders, rank_betas, rank_chain_seed
    spec = make_spec(w)
    L, R, d, S = w["ladders"], w["rungs"], w["dim"], w["pt_steps"]
    eng = Engine(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, record_level=K.RECORD_BASIC, hist_capacity=w["hist"], save_every=w["save_every"], device=local,
                             seed=rank_chain_seed(0xB2000003, rank)))
    spec.setup(eng)
    eng.set_betas(rank_betas(L, R, rank, world, spec.Tmax))
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
    with torch.cuda.stream(stream):
        eng.init_from_prior(); eng.synchronize()
        if fused:
            drv = FusedRungShardedLadders(eng, rank, world, 0xB2005EED, exchange_every=K_ex, in_launch=in_launch, max_launch=S)
        else:
            drv = RungShardedLadders(eng, rank, world, 0xB2005EED, exchange_every=K_ex, device="cuda:%d" % local, stream_ordered=True)
        dog = threading.Timer(watchdog_s, eng.xchg_abort) if fused else None   # a neighbour that never launches must not hang this rank
        if dog:
            dog.daemon = True; dog.start()
        for _ in range(warmup):
            drv.run(S)
        eng.synchronize(); n0 = eng.get_total_steps(); l0 = eng.launch_count()
        barrier(); torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); ev0.record(stream)
        for _ in range(steps):
            drv.run(S)
        drv.finish()
        ev1.record(stream)
        eng.synchronize(); torch.cuda.synchronize(); barrier()
        dt = max(ev0.elapsed_time(ev1) * 1e-3, time.perf_counter() - t0) if not fused else ev0.elapsed_time(ev1) * 1e-3
        if dog:
            dog.cancel()
        n1 = eng.get_total_steps(); launches = eng.launch_count() - l0
        # the same launches without any exchange (what a cycle costs on its own)
        torch.cuda.synchronize(); ev0.record(stream)
        for _ in range(100):
            eng.step(K_ex)
        ev1.record(stream); eng.synchronize(); torch.cuda.synchronize()
        bare_ms = ev0.elapsed_time(ev1) / 100
    cycles = steps * ((S + K_ex - 1) // K_ex)
    dt_max = max_over_ranks(dt)
    out = dict(value=sum_over_ranks(n1 - n0) / dt_max, unit="chain-steps/s", steps=steps, pt_iterations_per_step=S, exchange_every=K_ex,
               layout="%d ladders x (%d GPUs x %d rungs): rank g owns rungs [%d g, %d (g+1)) of every ladder" % (L, world, R, R, R),
               exchange=("fused into the step kernel over NVLink peer memory (CUDA IPC, per-ladder flags): peer loads in the prologue of the next launch, no collective"
                         + (", inside %d-iteration launches" % S if in_launch else ", one launch per exchange")) if fused else "NCCL neighbour send/recv, stream-ordered with the pack / swap kernels",
               ms_per_cycle=1e3 * dt_max / cycles, ms_per_cycle_without_exchange=max_over_ranks(bare_ms), gpu_launches=int(launches),
               nvlink_bytes_per_exchange_per_rank=int((2 if 0 < rank < world - 1 else 1) * L * (d + 3) * 8),
               nvlink_bytes_per_exchange_interior_rank=int(2 * L * (d + 3) * 8))
    over = out["ms_per_cycle"] - out["ms_per_cycle_without_exchange"]
    out["exchange_overhead_ms_per_cycle"] = over
    out["limiter"] = ("launch boundary: one launch per exchange costs the launch latency plus the wait for the slower neighbour's previous launch; the peer loads themselves "
                      "(%d bytes per ladder edge) are latency-, not bandwidth-bound" % ((d + 3) * 8))
    eng.close()
    return out


# ---------------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed bench steps (default: enough for a >= 3 s timed region)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1_sines", choices=sorted(WORKLOADS))
    ap.add_argument("--pt-steps", type=int, default=0, help="PT iterations per bench step (default: per workload)")
    ap.add_argument("--ladders", type=int, default=0, help="ladders per GPU (default: per workload)")
    ap.add_argument("--hist", type=int, default=0, help="history ring slots per chain (default: per workload)")
    ap.add_argument("--save-every", type=int, default=0, help="add_every_N: store every N-th sample (default: per workload)")
    ap.add_argument("--swap-mode", default="reference", choices=["reference", "even_odd"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the main workload: no `workloads` / `config5` / `rung_sharded` objects, no sustained repeat")
    ap.add_argument("--in-launch", action="store_true", help="with --fused-exchange: exchange inside long launches (whole grid must be resident)")
    ap.add_argument("--fused-exchange", action="store_true", help="with --rung-sharded: exchange fused into the step kernel over NVLink peer memory (no collective)")
    ap.add_argument("--rung-sharded", type=int, default=0, metavar="K",
                    help="main line = the rung-sharded layout: ONE (n_gpus x rungs)-rung ladder family sharded by rung blocks, cross-GPU boundary swaps every K PT iterations")
    ap.add_argument("--peaks", action="store_true", help="measure the FP64 peaks (DFMA, DMUL+DADD, DMMA) of cuda:0, print them as JSON and exit")
    args = ap.parse_args()
    if args.peaks:
        import ctypes as C
        from ptmcmc_b200._lib import load
        out = (C.c_double * 4)()
        rc = load().ptg_measure_fp64_peaks(C.c_int32(0), out)
        assert rc == 0, rc
        print(json.dumps(dict(fp64_dfma_tflops=out[0], fp64_dmul_dadd_tflops=out[1], fp64_dmma_tflops=out[2], sm_count=int(out[3]),
                              how="ptmcmc_b200/csrc/ptg_peaks.cu: 8 independent chains per thread x 20000 iterations, 148x8 CTAs x 256 threads, best of 3, CUDA events")))
        return 0
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    w = dict(WORKLOADS[args.workload])
    if args.pt_steps: w["pt_steps"] = args.pt_steps
    if args.ladders: w["ladders"] = args.ladders
    if args.hist: w["hist"] = args.hist
    if args.save_every: w["save_every"] = args.save_every
    w.setdefault("save_every", 1)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    config = workload_config(args.workload, w, args.swap_mode)
    metric, unit = "tempered chain-steps/s", "chain-steps/s"

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = args.steps or 5
        procs = cpu_threads()
        n_pt = reference_pt_steps(w, 6.0)
        for _ in range(min(args.warmup, 1)):
            run_reference_sample(w, max(20, n_pt // 10), procs)
        tot, secs, ess_sum, ess_ok, kind = 0, 0.0, 0.0, True, "reference"
        for _ in range(steps):
            r = run_reference_sample(w, n_pt, procs, ess=True)
            tot += r["chain_steps"]; secs += r["seconds"]; kind = r["kind"]
            if r["ess"] is None: ess_ok = False
            else: ess_sum += r["ess"]
        val = tot / secs
        sample = "%d processes x 1 ladder x %d rungs x %d PT iterations per step (timed: the stepping loop, max over processes)" % (procs, w["rungs"], n_pt)
        line = dict(metric=metric, value=val, unit=unit, n_gpus=args.gpus, steps=steps, warmup=args.warmup,
                    ms_per_step=1e3 * secs / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                    data="synthetic", impl="reference", config=config,
                    cpu_baseline=dict(value=val, unit=unit, cores=procs, kind=kind, sample=sample),
                    e2e=dict(value=val, unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        if ess_ok and ess_sum > 0:
            line["ess"] = dict(value=ess_sum / secs, unit="ESS/s", estimator="chain::report_effective_samples of the reference itself on every process's cold chain (ref_trace ess=1), summed over processes and steps")
            line["cpu_baseline"]["ess"] = line["ess"]
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    if args.rung_sharded:
        steps = args.steps or 20
        r = measure_rung_sharded(w, args.rung_sharded, steps, args.warmup, rank, world, local, fused=args.fused_exchange, in_launch=args.in_launch)
        config["parallelism"] = "rung-sharded: " + r["layout"] + "; " + r["exchange"]
        if rank == 0:
            print(json.dumps(dict(metric=metric, value=r["value"], unit=unit, n_gpus=world, steps=steps, warmup=args.warmup, ms_per_step=r["ms_per_cycle"] * ((w["pt_steps"] + args.rung_sharded - 1) // args.rung_sharded),
                                  higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", config=config, exchange=r, gpu_launches=r["gpu_launches"])))
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- the main workload
    est_ms = dict(c1_sines=10.5, a_gauss=11.0, b_poly=1.3, c2_sinusoid=11.0, d_fullcov=16.0).get(args.workload, 10.0)
    steps = args.steps or int(np.ceil(3000.0 / est_ms))
    r = measure(args.workload, w, steps, args.warmup, rank, world, local, swap_mode=args.swap_mode, sustain_s=0.0 if args.no_extras else 3.0, with_ess=True)
    out = dict(metric=metric, value=r["value"], unit=unit, n_gpus=world, steps=steps, warmup=args.warmup, ms_per_step=r["ms_per_step"],
               higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", config=config, clocks=r["clocks"],
               e2e=r["e2e"], gpu_launches=r["gpu_launches"], roofline=r["roofline"], ess=r.get("ess"))
    for k in ("roofline_fp64", "sustained"):
        if k in r:
            out[k] = r[k]
    # ---- N > 1: the one collective of the ladder-sharded layout, the gather of cold-chain samples to rank 0 (off the hot path)
    if world > 1:
        from ptmcmc_b200.sharding import gather_cold_samples
        ox = r["_last_cold"]
        local_t = torch.from_numpy(np.ascontiguousarray(ox[:, -64:, :])).cuda(local)
        torch.cuda.synchronize(); tg = time.perf_counter()
        full = gather_cold_samples(local_t, w["ladders"] * world, device=torch.device("cuda", local))
        torch.cuda.synchronize(); tg = time.perf_counter() - tg
        if rank == 0:
            assert tuple(full.shape) == (w["ladders"] * world, local_t.shape[1], w["dim"])
            out["cold_gather"] = dict(ms=1e3 * tg, bytes=int(full.numel() * 8), backend="nccl", samples_per_ladder=int(local_t.shape[1]))
    if not args.no_extras:
        torch.cuda.empty_cache()
        if world == 1:
            # BASELINE.json's other configs, short runs: value, e2e and both rooflines each
            subs = {}
            short = dict(a_gauss=30, b_poly=200, c2_sinusoid=30, d_fullcov=12)
            for name in ("a_gauss", "b_poly", "c2_sinusoid", "d_fullcov"):
                if name == args.workload:
                    continue
                ws = dict(WORKLOADS[name]); ws.setdefault("save_every", 1)
                try:
                    rr = measure(name, ws, short[name], 3, rank, world, local, e2e_steps=5)
                    rr["config"] = workload_config(name, ws)
                    subs[name] = brief(rr)
                except Exception as exc:
                    subs[name] = dict(error=repr(exc))
                torch.cuda.empty_cache()
            out["workloads"] = subs
        else:
            # BASELINE config 5: 65 536 chains per GPU (2048 ladders x 32 rungs), ladder-sharded and rung-sharded with cross-GPU swaps
            w5 = dict(WORKLOADS["c1_sines"]); w5.setdefault("save_every", 1); w5["ladders"] = 2048
            try:
                r5 = measure("c1_sines", w5, 60, 3, rank, world, local, e2e_steps=5)
                r5["config"] = workload_config("c1_sines", w5)
                out["config5"] = brief(r5)
            except Exception as exc:
                out["config5"] = dict(error=repr(exc))
            torch.cuda.empty_cache()
            wr = dict(WORKLOADS["c1_sines"]); wr.setdefault("save_every", 1)
            try:
                out["rung_sharded"] = measure_rung_sharded(wr, 10, 10, 3, rank, world, local, fused=True)
                wr5 = dict(wr); wr5["ladders"] = 2048
                out["rung_sharded_config5"] = measure_rung_sharded(wr5, 10, 10, 3, rank, world, local, fused=True)
            except Exception as exc:
                out["rung_sharded"] = dict(error=repr(exc))
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(w, 10.0, cpu_threads(), unit)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
