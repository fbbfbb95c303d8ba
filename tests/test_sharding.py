"""N>1 host logic on CPU: world_size-2 gloo.  Each rank runs ITS shard of the ladders (through the oracle, which stands
in for the engine here -- same C ABI, same Philox addressing) and the gathered cold chains must equal a single-process
run of all ladders: results are invariant to the number of ranks."""
import os
import socket
import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp
from ptmcmc_b200 import _capi as K
from ptmcmc_b200.sharding import ladder_shard, shard_config, gather_cold_samples, consensus_stop, max_over_ranks, sum_over_ranks
from tests.models import Spec
from tests.oracle_binding import Oracle


def test_ladder_shard_partitions():
    for total in (1, 7, 8, 4096, 65537):
        for world in (1, 2, 3, 8):
            spans = [ladder_shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _c1) in zip(spans, spans[1:]):
                assert f1 == f0 + c0
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        ladder_shard(4, 2, 2)


SPEC_ARGS = dict(model="sines", dim=3, rungs=6)
TOTAL, STEPS, NOUT = 5, 120, 40


def cold_block(o, n_local):
    out = []
    for l in range(n_local):
        n = int(o.get_counters()["nsize"][l * SPEC_ARGS["rungs"]])
        h = o.get_history(l, 0, n - NOUT, NOUT)
        out.append(np.concatenate([h["x"], h["lpost"][:, None], h["llike"][:, None]], axis=1))
    return np.stack(out)


def run_shard(rank, world):
    spec = Spec(SPEC_ARGS["model"], SPEC_ARGS["dim"], SPEC_ARGS["rungs"], seed=0.1234)
    cfg = shard_config(spec.config, TOTAL, rank, world, rng_mode=K.RNG_PHILOX, seed=0xB2000003)
    o = Oracle(cfg); spec.setup(o); o.init_from_prior(); o.step(STEPS)
    return cold_block(o, cfg.n_ladders), o.get_total_steps()


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local, nsteps = run_shard(rank, world)
        full = gather_cold_samples(local, TOTAL)
        total_steps = sum_over_ranks(nsteps)
        slowest = max_over_ranks(1.0 + rank)
        stop = consensus_stop(rank == 1)
        if rank == 0:
            q.put((full, total_steps, slowest, stop))
        else:
            assert full is None and stop
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    full, total_steps, slowest, stop = q.get()
    for p in procs:
        p.join(120); assert p.exitcode == 0
    single, single_steps = run_shard(0, 1)
    assert full.shape == single.shape == (TOTAL, NOUT, SPEC_ARGS["dim"] + 2)
    assert full.tobytes() == single.tobytes()
    assert total_steps == single_steps
    assert slowest == 2.0 and stop is True
