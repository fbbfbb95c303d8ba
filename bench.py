#!/usr/bin/env python
"""bench.py -- tempered chain-steps/s of the PT hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1_sines] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path over the whole resident batch: `--pt-steps` iterations of
parallel_tempering_chains::step (chain.cc:1393) for every ladder on the GPU, in ONE launch of the fused step kernel.
`value`  = tempered chain-steps (history appends = Nhist increments) of all ranks / device time (CUDA
           events on the engine's stream, max over ranks), state resident in HBM.
`e2e`    = the same metric through the C-ABI calls a host facade makes per block with HOST buffers inside the timed
           region: ptg_set_current (H2D of every chain's state from pinned memory) -> ptg_step_host (steps + D2H of
           the cold chains' newest samples).
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
                   flops=dict(per_chain_step=1000 * 26.0, peak_key="fp64_dmul_dadd_tflops", what="1000 points x (Horner 2(d-1) + residual 3 + division 15)"),
                   desc="5-coefficient polynomial chi^2 over 1000 points, 1024 ladders x 16 rungs, DE proposals"),
    # C2: 3-sinusoid chi^2 over 1e4 samples, d=9
    "c2_sinusoid": dict(model="sinusoid", dim=9, rungs=32, ladders=4096, pt_steps=1, hist=512, f_de=0.8, f_sn=0.1,
                        flops=dict(per_chain_step=1e4 * 89.0, peak_key="fp64_dmul_dadd_tflops", what="1e4 samples x (3 sin at 24 flop + 17), SURVEY.md 8(d) convention"),
                        desc="3-sinusoid chi^2 fit to 1e4 samples d=9, 4096 ladders x 32 rungs, default proposal mix"),
    # configs[3] / D: correlated Gaussian d=100, full covariance; "65536 chains x 24 rungs" read as 65 544 chains in total
    # (2731 ladders x 24 rungs; the 1.57 M-chain reading leaves < 100 history slots per chain in 180 GB, SURVEY.md 8d, and the
    # reference's DE member needs >= 10 d = 1000 stored samples to be ready).  Ninit = 11 d prior draws per chain (de_ni = 11).
    # save_every = 8: the 1280-slot ring then spans 10 240 PT iterations, which removes the short-window bias at d = 100 without
    # 8x the memory ("D must run with a short ring and/or save_every >> 1")
    "d_fullcov": dict(model="fullcov", dim=100, rungs=24, ladders=2731, pt_steps=50, hist=1280, save_every=8, f_de=0.5, f_sn=0.1,
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


def run_reference_sample(w, pt_steps, procs):
    """`procs` independent processes, each one ladder of the workload stepped `pt_steps` times by the reference itself
    (oracle/_ref/ref_trace) or, when that binary is absent, by the oracle port.  Returns (chain_steps, seconds, kind)."""
    spec = make_spec(w)
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_trace")
    if os.path.exists(ref):
        with tempfile.TemporaryDirectory() as td:
            cmds = []
            for i in range(procs):
                spec.seed = 0.05 + 0.9 * (i + 0.5) / procs
                pd = os.path.join(td, "p%d" % i); os.makedirs(pd)
                cmds.append([ref] + spec.ref_args(pd, pt_steps, "/dev/null"))
            t0 = time.perf_counter()
            ps = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for c in cmds]
            outs = [p.communicate()[0] for p in ps]
            dt = time.perf_counter() - t0
        total = 0
        for o, p in zip(outs, ps):
            if p.returncode != 0:
                raise RuntimeError("ref_trace failed")
            total += int(o.split("total_Nhist=")[1].split()[0])
        return total, dt, "reference"
    import multiprocessing as mp
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(procs) as pool:
        totals = pool.map(_port_worker, [(w, pt_steps, 0.05 + 0.9 * (i + 0.5) / procs) for i in range(procs)])
    return sum(totals), time.perf_counter() - t0, "port"


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


# ---------------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1_sines", choices=sorted(WORKLOADS))
    ap.add_argument("--pt-steps", type=int, default=0, help="PT iterations per bench step (default: per workload)")
    ap.add_argument("--ladders", type=int, default=0, help="ladders per GPU (default: per workload)")
    ap.add_argument("--hist", type=int, default=0, help="history ring slots per chain (default: per workload)")
    ap.add_argument("--save-every", type=int, default=0, help="add_every_N: store every N-th sample (default: per workload)")
    ap.add_argument("--swap-mode", default="reference", choices=["reference", "even_odd"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--in-launch", action="store_true", help="with --fused-exchange: exchange inside long launches (whole grid must be resident)")
    ap.add_argument("--fused-exchange", action="store_true", help="with --rung-sharded: exchange fused into the step kernel over NVLink peer memory (no collective)")
    ap.add_argument("--rung-sharded", type=int, default=0, metavar="K",
                    help="optional layout: ONE (n_gpus x rungs)-rung ladder family sharded by rung blocks, cross-GPU boundary swaps over NCCL every K PT iterations")
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
    config = dict(workload="%s: %s" % (args.workload, w["desc"]), ladders_per_gpu=w["ladders"], rungs=w["rungs"], dim=w["dim"],
                  chains_per_gpu=w["ladders"] * w["rungs"], pt_iterations_per_step=w["pt_steps"], save_every=w["save_every"], hist_capacity=w["hist"],
                  swap_mode=args.swap_mode, rng="philox4x32-10", parallelism="ladders sharded, %d per GPU, no data-path collective" % w["ladders"],
                  l2="history ring (%.1f GB per GPU) is larger than L2" % (w["ladders"] * w["rungs"] * w["hist"] * 8.0 * (w["dim"] + 2) / 1e9))
    metric, unit = "tempered chain-steps/s", "chain-steps/s"

    if args.impl == "reference":
        if rank != 0:
            return 0
        procs = cpu_threads()
        n_pt = reference_pt_steps(w, 6.0)
        for _ in range(min(args.warmup, 1)):
            run_reference_sample(w, max(20, n_pt // 10), procs)
        tot, secs = 0, 0.0
        for _ in range(args.steps):
            t, dt, kind = run_reference_sample(w, n_pt, procs)
            tot += t; secs += dt
        val = tot / secs
        sample = "%d processes x 1 ladder x %d rungs x %d PT iterations per step" % (procs, w["rungs"], n_pt)
        print(json.dumps(dict(metric=metric, value=val, unit=unit, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                              ms_per_step=1e3 * secs / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                              data="synthetic", impl="reference", config=config,
                              cpu_baseline=dict(value=val, unit=unit, cores=procs, kind=kind, sample=sample),
                              e2e=dict(value=val, unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0))))
        return 0

    import torch
    import torch.distributed as dist
    from ptmcmc_b200 import _capi as K
    from ptmcmc_b200.engine import Engine
    from ptmcmc_b200.sharding import max_over_ranks, sum_over_ranks
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])

    spec = make_spec(w)
    L, R, d, S = w["ladders"], w["rungs"], w["dim"], w["pt_steps"]
    if args.rung_sharded:
        # ---- optional rung-sharded layout (ptmcmc_b200/rung_sharding.py): the same ladders on every rank, a different rung block
        from ptmcmc_b200.rung_sharding import RungShardedLadders, FusedRungShardedLadders, rank_betas
        eng = Engine(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, record_level=K.RECORD_BASIC, hist_capacity=w["hist"], save_every=w["save_every"], device=local,
                                 seed=0xB2000003 + 977 * rank))
        spec.setup(eng)
        eng.set_betas(rank_betas(L, R, rank, world, spec.Tmax))
        stream = torch.cuda.Stream()
        eng.set_stream(stream.cuda_stream)
        with torch.cuda.stream(stream):
            eng.init_from_prior(); eng.synchronize()
            if args.fused_exchange:
                drv = FusedRungShardedLadders(eng, rank, world, 0xB2005EED, exchange_every=args.rung_sharded, in_launch=args.in_launch, max_launch=S)
            else:
                drv = RungShardedLadders(eng, rank, world, 0xB2005EED, exchange_every=args.rung_sharded, device="cuda:%d" % local, stream_ordered=True)
            for _ in range(args.warmup):
                drv.run(S)
            eng.synchronize(); n0 = eng.get_total_steps()
            barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                drv.run(S)
            drv.finish()
            eng.synchronize(); torch.cuda.synchronize(); barrier()
            dt = time.perf_counter() - t0
            n1 = eng.get_total_steps()
            # the same launches without any exchange (what a cycle costs on its own), and the exchange step alone
            torch.cuda.synchronize(); t1 = time.perf_counter()
            for _ in range(100):
                eng.step(args.rung_sharded)
            eng.synchronize(); torch.cuda.synchronize(); bare_ms = 1e3 * (time.perf_counter() - t1) / 100
            ex_ms = None
            if not args.fused_exchange:
                torch.cuda.synchronize(); t1 = time.perf_counter()
                for _ in range(20):
                    drv.exchange()
                eng.synchronize(); torch.cuda.synchronize(); ex_ms = 1e3 * (time.perf_counter() - t1) / 20
        val = sum_over_ranks(n1 - n0) / max_over_ranks(dt)
        config["parallelism"] = "rung-sharded: %d ladders x (%d GPUs x %d rungs), boundary swaps %s every %d PT iterations" % (
            L, world, R, ("fused into the step kernel over NVLink peer memory" + (", inside %d-iteration launches" % S if args.in_launch else ", one launch per exchange")) if args.fused_exchange else "over NCCL", args.rung_sharded)
        ms_step = 1e3 * max_over_ranks(dt) / args.steps
        if rank == 0:
            print(json.dumps(dict(metric=metric, value=val, unit=unit, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_step,
                                  higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", config=config,
                                  exchange=dict(ms=ex_ms, ms_per_cycle=ms_step / ((S + args.rung_sharded - 1) // args.rung_sharded), ms_per_cycle_without_exchange=bare_ms,
                                                bytes_per_rank=int((2 if 0 < rank < world - 1 else 1) * L * (d + 3) * 8),
                                                collective="none: peer-memory loads in the step kernel's prologue, per-ladder flags" if args.fused_exchange else
                                                "neighbour send/recv pairs (NCCL batch_isend_irecv), stream-ordered with the pack / swap kernels"),
                                  gpu_launches=args.steps * ((S + args.rung_sharded - 1) // args.rung_sharded) * (1 if args.fused_exchange else 5))))
        eng.close()
        if world > 1:
            dist.destroy_process_group()
        return 0
    swap_mode = K.SWAP_REFERENCE if args.swap_mode == "reference" else K.SWAP_EVEN_ODD
    eng = Engine(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, record_level=K.RECORD_BASIC, hist_capacity=w["hist"], save_every=w["save_every"], swap_mode=swap_mode,
                             device=local, ladder_offset=rank * L, seed=0xB2000003))
    spec.setup(eng)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    eng.init_from_prior()
    eng.synchronize()

    # ---- device-resident throughput
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            eng.step(S)
        eng.synchronize()
        n0 = eng.get_total_steps()
        clocks = ClockSampler(local); clocks.start()
        barrier(); torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            eng.step(S)
        ev1.record(stream)
        torch.cuda.synchronize(); barrier()
        ms = ev0.elapsed_time(ev1)
        clk = clocks.stop()
        eng.synchronize()
        n1 = eng.get_total_steps()
    ms_max = max_over_ranks(ms)
    chain_steps = sum_over_ranks(n1 - n0)
    value = chain_steps / (ms_max * 1e-3)

    # ---- end to end through the C ABI with host buffers
    n_chains = L * R
    n_out = min(S, 64)
    pin = lambda *shape: torch.empty(*shape, dtype=torch.float64).pin_memory().numpy()
    hx, hlp, hll, hpr = pin(n_chains, d), pin(n_chains), pin(n_chains), pin(n_chains)
    ox, olp, oll = pin(L, n_out, d), pin(L, n_out), pin(L, n_out)
    cur = eng.get_current(); hx[:] = cur["x"]; hlp[:] = cur["lpost"]; hll[:] = cur["llike"]; hpr[:] = eng.get_lprior()
    e2e_steps = max(3, min(args.steps, 10))
    with torch.cuda.stream(stream):
        for _ in range(2):
            eng.set_current(hx, hlp, hll, hpr); eng.step_host(S, n_out, ox, olp, oll)
        m0 = eng.get_total_steps()
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter(); ev0.record(stream)
        for _ in range(e2e_steps):
            eng.set_current(hx, hlp, hll, hpr)
            eng.step_host(S, n_out, ox, olp, oll)
        ev1.record(stream); torch.cuda.synchronize(); barrier()
        e2e_ms = max(ev0.elapsed_time(ev1), 1e3 * (time.perf_counter() - t0))
        m1 = eng.get_total_steps()
    e2e_val = sum_over_ranks(m1 - m0) / (max_over_ranks(e2e_ms) * 1e-3)
    h2d = n_chains * (d + 3) * 8
    d2h = L * n_out * (d + 2) * 8

    # ---- ESS/s (second half of BASELINE.json's metric): integrated autocorrelation time of the cold chains over the newest
    # history-ring window of a sample of ladders; ESS/s = (cold-chain steps per second over all ladders) / tau
    ess = None
    try:
        from ptmcmc_b200.analysis import ess_per_sample
        cnt = eng.get_counters()
        nh = int(min(w["hist"] - 8, 2000, cnt["nsize"][::R].min()))      # never more than the cold chains have stored so far
        tau_dev = eng.get_act(0, nh, min(nh // 2, 1000))                 # device: every ladder's cold chain, per parameter
        eps_dev = float(np.mean(1.0 / tau_dev.max(axis=1)))              # ESS per stored sample, mean over ladders
        nl = min(L, 256)                                                 # host cross-check on a sample of ladders (ACFs averaged, then windowed)
        cold = np.stack([eng.get_history(l, 0, int(cnt["nsize"][l * R]) - nh, nh, full=False)["x"] for l in range(nl)])
        eps_host, taus = ess_per_sample(cold)
        pt_iter_per_s = args.steps * S / (ms_max * 1e-3)
        # the reference's own estimator (chain::report_effective_samples as the run loop calls it, ptmcmc.cc:645) for EVERY ladder's cold
        # chain over the newest ring window: lag statistics on the device in the reference's summation order
        # (ptg_get_autocovar_windows), combination on the host; analysis.py restates the recipe and tests pin it to the reference build
        nr = int(min(w["hist"] - 8, 8000, cnt["nsize"][::R].min()))
        se = w["save_every"]
        eps_recipe, n_recipe = None, 0
        try:
            ess_l, len_l = eng.report_effective_samples_all(rung=0, window_records=nr)
            ok = len_l > 0
            n_recipe = int(ok.sum())
            if n_recipe:
                eps_recipe = float(np.mean(ess_l[ok] / len_l[ok]))      # ESS per PT iteration, mean over ladders
        except Exception as exc:
            recipe_error = repr(exc)
        sokal = eps_dev / w["save_every"] * pt_iter_per_s * L * world
        ess = dict(value=(eps_recipe * pt_iter_per_s * L * world) if eps_recipe else sokal, unit="ESS/s",
                   estimator=("the reference's chain::report_effective_samples(-1, 1000 save_every, save_every) (chain.cc:457-643) on every cold chain's newest "
                              "%d records: lag statistics on the device (ptg_get_autocovar_windows), min over parameters, summed over ladders" % nr) if eps_recipe
                   else "Sokal window (see sokal_device); the ring window is too short for the reference recipe",
                   ess_per_pt_iteration=eps_recipe, ladders=L, ladders_with_estimate=n_recipe,
                   sokal_device=dict(value=sokal, tau_pt_iterations=float(np.median(tau_dev.max(axis=1))) * w["save_every"], window=nh,
                                     estimator="ptg_get_act: per cold chain N/tau, Sokal-windowed integrated autocorrelation time, min over parameters, mean over ladders"), host_check=dict(value=eps_host / w["save_every"] * pt_iter_per_s * L * world,
                   tau_pt_iterations=float(taus.max()) * w["save_every"], ladders_sampled=nl))
        if nh > int(cnt["nsize"][::R].min()) - eng.cfg.n_init:
            ess["note"] = "the run is shorter than the analysis window: it still contains start-up prior draws, the ESS figure is not meaningful"
    except Exception as exc:  # analysis is not part of the timed path
        ess = dict(value=None, error=repr(exc))

    # ---- roofline of the dominant kernel (ptg_step_kernel: one launch per bench step)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bpcs = algorithmic_bytes_per_chain_step(w)
    launch_s = ms * 1e-3 / args.steps
    achieved = bpcs * ((n1 - n0) / args.steps) / launch_s / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
    if os.path.exists(tpath):
        t = json.load(open(tpath))  # ncu capture of the same kernel at 100 PT iterations per launch, scaled to this launch
        traffic = t["dram_bytes_per_pt_iteration"] * S * (L * R) / float(t.get("chains_profiled", 131072))
    kernel_name = "ptg_xmstep_kernel" if d > 16 else ("ptg_fstep_kernel" if R <= 32 else "ptg_step_kernel")
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic, kernel=kernel_name,
                    algorithmic_bytes_per_chain_step=bpcs, chain_steps_per_launch=(n1 - n0) / args.steps, launch_ms=launch_s * 1e3, peak_source=peak_src,
                    note="latency/ALU-bound fp64+Philox kernel: HBM is the stated roofline, see DESIGN.md section 5")

    # compute-side view for the workloads with a dense fp64 flop count: algorithmic flops against the FP64 peaks measured by
    # `bench.py --peaks` on this pool (profiles/r01_fp64_peaks.json; MEASURED_PEAKS.json carries HBM and bf16 only)
    roofline_fp64 = None
    fl = w.get("flops")
    ppath = os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")
    if fl and os.path.exists(ppath):
        pk = float(json.load(open(ppath))[fl["peak_key"]])
        ach = fl["per_chain_step"] * ((n1 - n0) / args.steps) / launch_s / 1e12
        roofline_fp64 = dict(bound="fp64", achieved=ach, peak=pk, unit="TFLOP/s", frac=ach / pk, flops_per_chain_step=fl["per_chain_step"],
                             counted=fl["what"], peak_source="measured: profiles/r01_fp64_peaks.json " + fl["peak_key"])

    out = dict(metric=metric, value=value, unit=unit, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_max / args.steps,
               higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", config=config, clocks=clk,
               e2e=dict(value=e2e_val, unit=unit, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, steps=e2e_steps),
               gpu_launches=args.steps, roofline=roofline, ess=ess)
    if roofline_fp64:
        out["roofline_fp64"] = roofline_fp64
    # ---- N > 1: the one collective of the ladder-sharded layout, the gather of cold-chain samples to rank 0 (off the hot path)
    if world > 1:
        from ptmcmc_b200.sharding import gather_cold_samples
        local_t = torch.from_numpy(np.ascontiguousarray(ox)).cuda(local)
        torch.cuda.synchronize(); tg = time.perf_counter()
        full = gather_cold_samples(local_t, L * world, device=torch.device("cuda", local))
        torch.cuda.synchronize(); tg = time.perf_counter() - tg
        if rank == 0:
            assert tuple(full.shape) == (L * world, n_out, d)
            out["cold_gather"] = dict(ms=1e3 * tg, bytes=int(full.numel() * 8), backend="nccl", samples_per_ladder=n_out)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        procs = cpu_threads()
        n_pt = reference_pt_steps(w, 10.0)
        tot, secs, kind = run_reference_sample(w, n_pt, procs)
        out["cpu_baseline"] = dict(value=tot / secs, unit=unit, cores=procs, kind=kind,
                                   sample="%d processes x 1 ladder x %d rungs x %d PT iterations of the same workload" % (procs, R, n_pt))
    if rank == 0:
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
