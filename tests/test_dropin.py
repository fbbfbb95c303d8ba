"""The drop-in claim, run for real: oracle/_ref/gpu_dropin is a driver in the style of the reference's example.cc, compiled by
oracle/build_ref.sh against the UNMODIFIED reference sources with include/gpu_parallel_tempering_chains.hh (the reference-side binding of
the engine: a parallel_tempering_chains whose step() is ptg_step and whose MH_chain objects mirror the device histories).  The reference's
own run loop (ptmcmc_sampler::run, ptmcmc.cc:563-661) then drives the engine: dumpChain, status, report_prop, report_effective_samples
and checkpoint are the reference's code, and the checkpoint files (PTchain.cp / MHchain.cp / chain.cp, chain.cc:656-731, 1213-1239) are
read back by the reference's own restart."""
import os
import re
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "gpu_dropin")
ARGS = ["--seed=0.5", "--pt=8", "--nsteps=20000", "--nevery=5000"]
needs_exe = pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/gpu_dropin not built (needs /root/reference at build time)")


def run(cwd, outname, extra=(), gpu=False, nsteps=None):
    env = dict(os.environ)
    env.pop("PTMCMC_GPU", None)
    if gpu:
        env["PTMCMC_GPU"] = "1"
    args = [a if nsteps is None or not a.startswith("--nsteps") else "--nsteps=%d" % nsteps for a in ARGS]
    r = subprocess.run([EXE] + args + ["--outname=" + outname] + list(extra), cwd=str(cwd), env=env, capture_output=True, text=True, timeout=900)
    return r


def read_chain(path):
    """rows of a reference chain file (MH_chain::dumpChain, chain.cc:1112-1135): step lpost llike acc type: p0 p1 p2 invtemp"""
    rows = []
    for line in open(path):
        if not line.strip() or line.startswith("#"):
            continue
        head, tail = line.split(":")
        h, t = head.split(), tail.split()
        rows.append([float(h[0]), float(h[1]), float(h[2])] + [float(v) for v in t[:3]])
    return np.array(rows)


@needs_exe
def test_dropin_driver_reference_arm_and_restart_roundtrip(tmp_path):
    """without a GPU: the driver is the reference (same binary, PTMCMC_GPU unset); its checkpoint / restart cycle works"""
    r = run(tmp_path, "cpu", nsteps=6000, extra=["--checkp_at_step=3001", "--nevery=1500"])
    assert r.returncode == 0, r.stdout[-800:]
    assert "parallel_tempering_chains" in r.stdout and os.path.exists(os.path.join(str(tmp_path), "step_3001-cp", "chain0-cp", "PTchain.cp"))
    r = run(tmp_path, "cpu", nsteps=6000, extra=["--restart_dir=step_3001-cp/", "--nevery=1500"])
    assert r.returncode == 0 and "Finished running chain 0" in r.stdout, r.stdout[-800:]
    c = read_chain(os.path.join(str(tmp_path), "cpu_t0.dat"))
    assert c[-1, 0] >= 5990


@pytest.mark.gpu
@needs_exe
def test_reference_run_loop_drives_the_engine(tmp_path):
    cpu = run(tmp_path, "cpu")
    assert cpu.returncode == 0, cpu.stdout[-800:]
    gpu = run(tmp_path, "gpu", gpu=True)
    assert gpu.returncode == 0, gpu.stdout[-1500:] + gpu.stderr[-500:]
    assert "chains stepped by the ptg engine" in gpu.stdout
    # the reference's own status / proposal report / effective-sample report ran on the mirrored chains
    for word in ("Effective sample size test", "useful chain length is", "acceptance report", "Finished running chain 0"):
        assert word in gpu.stdout, word
    ess_cpu = float(re.findall(r"Over 3 pars: ess=(\S+)", cpu.stdout)[-1]); ess_gpu = float(re.findall(r"Over 3 pars: ess=(\S+)", gpu.stdout)[-1])
    c, g = read_chain(os.path.join(str(tmp_path), "cpu_t0.dat")), read_chain(os.path.join(str(tmp_path), "gpu_t0.dat"))
    # same dump cadence and extent: start-up samples (negative steps) then every Nskip-th step up to nsteps
    assert abs(len(c) - len(g)) <= 0.02 * len(c) and g[0, 0] == c[0, 0] and abs(g[-1, 0] - c[-1, 0]) <= 10
    pc, pg = c[c[:, 0] > 2000][:, 3:], g[g[:, 0] > 2000][:, 3:]
    print("cold chain after burn-in: reference mean %s std %s | engine mean %s std %s | ESS reference %.0f engine %.0f" %
          (np.round(pc.mean(0), 3), np.round(pc.std(0), 3), np.round(pg.mean(0), 3), np.round(pg.std(0), 3), ess_cpu, ess_gpu))
    # target: N((2,-3,5), 0.5^2) inside the prior box
    for p in (pc, pg):
        assert np.abs(p.mean(0) - np.array([2.0, -3.0, 5.0])).max() < 0.08
        assert np.abs(p.std(0) - 0.5).max() < 0.06
    assert 0.5 < ess_gpu / ess_cpu < 2.0
    # log-posterior column is consistent with the positions: lpost = llike + log prior (uniform box of volume 8*2*3*5)
    lp = -1.5 * np.log(np.pi * 0.5) - ((g[:, 3:] - np.array([2.0, -3.0, 5.0])) ** 2).sum(1) / 0.5
    assert np.allclose(g[:, 2], lp, rtol=1e-9, atol=1e-9)
    assert np.allclose(g[:, 1], lp - np.log(8 * 2 * 3 * 5.0), rtol=1e-9, atol=1e-9)


@pytest.mark.gpu
@needs_exe
def test_engine_checkpoint_is_read_by_the_reference_restart(tmp_path):
    """the engine-backed run checkpoints through the reference's own writers; the REFERENCE (CPU arm) restarts from those files"""
    g = run(tmp_path, "mix", gpu=True, nsteps=12000, extra=["--checkp_at_step=6001", "--nevery=3000"])
    assert g.returncode == 0 and "Writing checkpoint files" in g.stdout, g.stdout[-800:]
    cp = os.path.join(str(tmp_path), "step_6001-cp")
    assert os.path.exists(os.path.join(cp, "chain0-cp", "PTchain.cp")) and os.path.exists(os.path.join(cp, "chain1-cp", "MHchain.cp"))
    before = read_chain(os.path.join(str(tmp_path), "mix_t0.dat"))
    r = run(tmp_path, "mix", nsteps=12000, extra=["--restart_dir=step_6001-cp/", "--nevery=3000"])
    assert r.returncode == 0 and "Finished running chain 0" in r.stdout, r.stdout[-1200:]
    after = read_chain(os.path.join(str(tmp_path), "mix_t0.dat"))
    assert len(after) > len(before) and after[-1, 0] >= 11990
    tail = after[after[:, 0] > 7000][:, 3:]
    assert np.abs(tail.mean(0) - np.array([2.0, -3.0, 5.0])).max() < 0.12 and np.abs(tail.std(0) - 0.5).max() < 0.08


@pytest.mark.gpu
@needs_exe
def test_sampler_adaptive_proposal_options_run_on_the_engine(tmp_path):
    """--prop_adapt_rate --prop_adapt_more, the options the reference's own exampleLISA test runs with (test/exampleLISA/Makefile:7): the
    sampler builds a nested adaptive set of Gaussian scales under an adaptive top level (ptmcmc.cc:70-72,123-143); the drop-in class flattens
    it for the engine (ptg_set_nested_set / ptg_set_proposal_options) and the reference's run loop reports the adapted shares"""
    extra = ["--prop_adapt_rate=0.01", "--prop_adapt_more"]
    cpu = run(tmp_path, "acpu", extra=extra)
    assert cpu.returncode == 0, cpu.stdout[-800:]
    gpu = run(tmp_path, "agpu", gpu=True, extra=extra)
    assert gpu.returncode == 0, gpu.stdout[-1500:] + gpu.stderr[-500:]
    assert "chains stepped by the ptg engine" in gpu.stdout
    c, g = read_chain(os.path.join(str(tmp_path), "acpu_t0.dat")), read_chain(os.path.join(str(tmp_path), "agpu_t0.dat"))
    pc, pg = c[c[:, 0] > 2000][:, 3:], g[g[:, 0] > 2000][:, 3:]
    for p in (pc, pg):
        assert np.abs(p.mean(0) - np.array([2.0, -3.0, 5.0])).max() < 0.08
        assert np.abs(p.std(0) - 0.5).max() < 0.06
