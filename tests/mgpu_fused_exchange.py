"""Multi-GPU check of the fused rung-boundary exchange (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_fused_exchange.py
Every rank runs the rung-sharded ladder twice from the same start -- NCCL neighbour exchange between launches
(RungShardedLadders), the exchange fused into the step kernel over CUDA-IPC peer pointers at launch boundaries and INSIDE long
launches (FusedRungShardedLadders) -- and
asserts bit-identical chains; then reports the time per cycle of both.  tests/test_rung_sharding.py drives it when >= 2 GPUs exist."""
import os
import sys
import time
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ptmcmc_b200 import _capi as K  # noqa: E402
from ptmcmc_b200.engine import Engine  # noqa: E402
from ptmcmc_b200.rung_sharding import RungShardedLadders, FusedRungShardedLadders, rank_betas  # noqa: E402
from ptmcmc_b200.workloads import Spec  # noqa: E402

L, R, DIM, TMAX, SHARED, EVERY, CYCLES = int(os.environ.get("MGPU_LADDERS", 512)), 32, 3, 1e6, 0xB2005EED, 10, 40


def make(rank, world, local):
    spec = Spec("sines", DIM, R, Tmax=TMAX)
    e = Engine(spec.config(n_ladders=L, rng_mode=K.RNG_PHILOX, hist_capacity=2048, seed=0xB2000003 + 977 * rank, device=local, record_level=K.RECORD_BASIC))
    spec.setup(e)
    e.set_betas(rank_betas(L, R, rank, world, TMAX))
    return e


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {}
    for mode in ("nccl", "fused", "in_launch"):
        e = make(rank, world, local)
        stream = torch.cuda.Stream()
        e.set_stream(stream.cuda_stream)
        with torch.cuda.stream(stream):
            e.init_from_prior(); e.synchronize()
            drv = RungShardedLadders(e, rank, world, SHARED, exchange_every=EVERY, device="cuda:%d" % local, stream_ordered=True) if mode == "nccl" \
                else FusedRungShardedLadders(e, rank, world, SHARED, exchange_every=EVERY, in_launch=(mode == "in_launch"), max_launch=200)
            drv.run(EVERY * 3); drv.finish(); e.synchronize()
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            drv.run(EVERY * CYCLES); drv.finish(); e.synchronize()
            torch.cuda.synchronize(); dist.barrier()
            dt = time.perf_counter() - t0
        c, n = e.get_current(), e.get_counters()
        out[mode] = (c["x"].copy(), c["lpost"].copy(), n["nhist"].copy(), n["naccept"].copy(), e.get_history(0, 0, 0, int(n["nsize"][0]), full=False)["x"].copy(), dt)
        e.close()
    same = all(np.array_equal(a, b) for a, b in zip(out["nccl"][:5], out["fused"][:5])) and all(np.array_equal(a, b) for a, b in zip(out["nccl"][:5], out["in_launch"][:5]))
    flags = [None] * world
    dist.all_gather_object(flags, bool(same))
    if rank == 0:
        print("fused == nccl on every rank:", all(flags), "| ms per cycle of %d iterations: nccl %.4f fused %.4f fused in-launch %.4f (%d ladders x %d rungs per GPU, %d GPUs)"
              % (EVERY, 1e3 * out["nccl"][5] / CYCLES, 1e3 * out["fused"][5] / CYCLES, 1e3 * out["in_launch"][5] / CYCLES, L, R, world))
    dist.barrier()
    dist.destroy_process_group()
    assert all(flags)


if __name__ == "__main__":
    main()
