"""Rung-sharded ladders (ptmcmc_b200/rung_sharding.py): host logic with world_size-2 gloo on the CPU oracle, and engine-vs-oracle
parity of the boundary exchange on one GPU (two engines stand for two ranks; no collective is needed to check the arithmetic)."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from ptmcmc_b200 import _capi as K
from ptmcmc_b200.rung_sharding import RungShardedLadders, global_betas, rank_betas
from tests.models import Spec
from tests.oracle_binding import Oracle

L, RPR, DIM, TMAX, SHARED = 24, 4, 2, 1e3, 0xB2005EED


def make_rank(cls, rank, world, device=0):
    spec = Spec("gauss", DIM, RPR, centers=[2, -3], halfwidths=[2, 3], Tmax=TMAX)
    cfg = spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=4000, seed=0xB2000003 + 977 * rank, device=device)
    api = cls(cfg)
    spec.setup(api)
    api.set_betas(rank_betas(L, RPR, rank, world, TMAX))
    api.init_from_prior()
    return api


def test_global_ladder_is_the_reference_geometric_ladder():
    b = global_betas(8, 1e6)
    assert b[0] == 1.0 and b[-1] == pytest.approx(1e-6, rel=1e-12)
    assert np.allclose(b[:-1] / b[1:], 1e6 ** (1 / 7), rtol=1e-12)
    rb = rank_betas(3, 4, 1, 2, 1e6)
    assert rb.shape == (12,) and (rb[:4] == b[4:]).all() and (rb[4:8] == b[4:]).all()


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        api = make_rank(Oracle, rank, world)
        drv = RungShardedLadders(api, rank, world, SHARED, exchange_every=5)
        before = api.get_current()
        drv.exchange()                      # one exchange on the initial states: check the data movement exactly
        after = api.get_current()
        drv.run(1500)
        cnt = api.get_counters()
        n = int(cnt["nsize"][0])
        cold = np.stack([api.get_history(l, 0, n - 1000, 1000)["x"] for l in range(L)]) if rank == 0 else None
        q.put((rank, before, after, api.get_total_steps(), cold, drv.n_exchanges))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_on_oracle():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = {}
    for _ in range(2):
        r = q.get(); res[r[0]] = r
    for p in procs:
        p.join(180); assert p.exitcode == 0
    (_, b0, a0, t0, cold, nex), (_, b1, a1, t1, _, _) = res[0], res[1]
    top0 = np.arange(L) * RPR + RPR - 1          # rank 0's hottest rung of every ladder
    bot1 = np.arange(L) * RPR                    # rank 1's coldest rung
    swapped = (a0["x"][top0] != b0["x"][top0]).any(axis=1)
    assert swapped.any() and (~swapped).any()     # some ladders swapped, some did not
    # an accepted trial exchanges the two states exactly, on both ranks, with the same decision
    assert (a0["x"][top0][swapped] == b1["x"][bot1][swapped]).all() and (a1["x"][bot1][swapped] == b0["x"][top0][swapped]).all()
    assert (a1["x"][bot1][~swapped] == b1["x"][bot1][~swapped]).all()
    assert (a0["llike"][top0][swapped] == b1["llike"][bot1][swapped]).all()
    # nothing else moved
    other0 = np.setdiff1d(np.arange(L * RPR), top0)
    assert (a0["x"][other0] == b0["x"][other0]).all()
    assert nex == 1 + 300
    # the cold chains of the 8-rung sharded ladder still sample the posterior (2-D Gaussian, sigma 0.5)
    x = cold.reshape(-1, DIM)
    assert np.allclose(x.mean(axis=0), [2, -3], atol=0.04) and np.allclose(x.var(axis=0), [0.25, 0.25], rtol=0.12)


@pytest.mark.gpu
def test_boundary_exchange_engine_matches_oracle(engine_cls):
    """two engines on one GPU stand for two ranks; the same exchange through the oracle must give the same states"""
    eng = [make_rank(engine_cls, r, 2) for r in range(2)]
    ora = [make_rank(Oracle, r, 2) for r in range(2)]
    for a in eng + ora:
        a.step(40)
    for e in eng:
        e.synchronize()
    dev = torch.device("cuda", 0)
    for it in range(6):
        packs_e = [[torch.zeros((L, DIM + 3), dtype=torch.float64, device=dev) for _ in range(2)] for _ in range(2)]
        packs_o = [[torch.zeros((L, DIM + 3), dtype=torch.float64) for _ in range(2)] for _ in range(2)]
        for r in range(2):
            eng[r].boundary_pack(0, packs_e[r][0].data_ptr()); eng[r].boundary_pack(RPR - 1, packs_e[r][1].data_ptr()); eng[r].synchronize()
            ora[r].boundary_pack(0, packs_o[r][0].data_ptr()); ora[r].boundary_pack(RPR - 1, packs_o[r][1].data_ptr())
            assert np.allclose(packs_e[r][0].cpu().numpy(), packs_o[r][0].numpy(), rtol=1e-9) and np.allclose(packs_e[r][1].cpu().numpy(), packs_o[r][1].numpy(), rtol=1e-9)
        eng[0].boundary_swap(RPR - 1, packs_e[1][0].data_ptr(), True, SHARED, 0, it); eng[1].boundary_swap(0, packs_e[0][1].data_ptr(), False, SHARED, 0, it)
        ora[0].boundary_swap(RPR - 1, packs_o[1][0].data_ptr(), True, SHARED, 0, it); ora[1].boundary_swap(0, packs_o[0][1].data_ptr(), False, SHARED, 0, it)
        for a in eng + ora:
            a.step(7)
        for e in eng:
            e.synchronize()
    for r in range(2):
        ce, co = eng[r].get_current(), ora[r].get_current()
        assert np.allclose(ce["x"], co["x"], rtol=1e-8) and np.allclose(ce["lpost"], co["lpost"], rtol=1e-8)
        ne, no = eng[r].get_counters(), ora[r].get_counters()
        assert (ne["nhist"] == no["nhist"]).all() and (ne["nsize"] == no["nsize"]).all() and (ne["naccept"] == no["naccept"]).all()
    assert eng[0].get_total_steps() == ora[0].get_total_steps()


def _final(api):
    api.synchronize()
    c, n = api.get_current(), api.get_counters()
    k = int(n["nsize"][0])
    return c["x"].copy(), c["lpost"].copy(), c["llike"].copy(), n["nhist"].copy(), n["nsize"].copy(), n["naccept"].copy(), api.get_history(0, 0, 0, k)["x"].copy(), \
        api.get_swap_stats()["swap_count"].copy(), api.get_swap_stats()["swap_accept"].copy()


@pytest.mark.gpu
def test_fused_exchange_reproduces_the_unfused_exchange(engine_cls):
    """three engines on one GPU stand for three ranks (the middle one has two boundaries).  The exchange fused into the step kernel
    (peer pointers, per-ladder flags) uses the Philox addresses of ptg_boundary_swap, so the chains are bit-identical to the run
    with separate pack / swap launches.  Launches are strictly sequential here: engines that spin on each other share this GPU."""
    from ptmcmc_b200.rung_sharding import FusedRungShardedLadders
    W, K_EVERY, CYCLES = 3, 7, 9
    dev = torch.device("cuda", 0)
    # --- unfused
    ref = [make_rank(engine_cls, r, W) for r in range(W)]
    packs = [[torch.zeros((L, DIM + 3), dtype=torch.float64, device=dev) for _ in range(2)] for _ in range(W)]
    for c in range(CYCLES):
        for e in ref:
            e.step(K_EVERY); e.synchronize()
        for r, e in enumerate(ref):
            e.boundary_pack(0, packs[r][0].data_ptr()); e.boundary_pack(RPR - 1, packs[r][1].data_ptr()); e.synchronize()
        for r, e in enumerate(ref):
            if r + 1 < W:
                e.boundary_swap(RPR - 1, packs[r + 1][0].data_ptr(), True, SHARED, r, c)
            if r > 0:
                e.boundary_swap(0, packs[r - 1][1].data_ptr(), False, SHARED, r - 1, c)
            e.synchronize()
    # --- fused
    eng = [make_rank(engine_cls, r, W) for r in range(W)]
    ptrs = {r: e.xchg_export()[1] for r, e in enumerate(eng)}
    drv = [FusedRungShardedLadders(e, r, W, SHARED, exchange_every=K_EVERY, peers=ptrs) for r, e in enumerate(eng)]
    for c in range(CYCLES):
        for d, e in zip(drv, eng):
            d.run(K_EVERY); e.synchronize()
    for d, e in zip(drv, eng):
        d.finish(); e.synchronize()
    for r in range(W):
        a, b = _final(ref[r]), _final(eng[r])
        for u, v in zip(a, b):
            assert np.array_equal(u, v), r
    # every exchange appends one record to each edge chain on top of the K_EVERY per-iteration records
    assert eng[1].get_total_steps() == ref[1].get_total_steps() == L * RPR * K_EVERY * CYCLES + 2 * L * CYCLES
    for e in ref + eng:
        e.close()


@pytest.mark.gpu
def test_fused_exchange_across_gpus():
    """one rank per GPU over CUDA-IPC peer pointers (tests/mgpu_fused_exchange.py): chains identical to the NCCL exchange"""
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    import subprocess, sys
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mgpu_fused_exchange.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), script], capture_output=True, text=True, timeout=600, env=dict(os.environ, MGPU_LADDERS="256"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "fused == nccl on every rank: True" in r.stdout
