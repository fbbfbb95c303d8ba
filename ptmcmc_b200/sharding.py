"""Multi-GPU host logic: temperature ladders are independent units (no cross-ladder coupling anywhere in chain.cc; the
reference runs exactly one ladder per process), so they are dealt to the ranks in contiguous blocks with NO data-path
collective.  Each rank's engine is created with `ladder_offset` = global id of its first ladder; every Philox draw is
addressed by the GLOBAL ladder id (include/ptmcmc_b200_rng.h), so the samples of ladder g do not depend on the number
of GPUs -- the property the reference gets from per-chain generators (chain.hh:45,66-67) and tests with
`mpirun -np 2` vs serial diffs (test/exampleLISA/Makefile:25-31).

Collectives (torch.distributed; NCCL on GPUs, gloo in the CPU tests) appear only off the hot path:
  * gather_cold_samples: the final / per-dump gather of cold-chain samples to rank 0 (replaces the reference's
    per-step MPI_Allgather of all states, chain.cc:1905, which the ladder-sharded layout does not need);
  * consensus_stop: 1-int max-reduce replacing MPI_Allreduce(LOR) / MPI_Bcast(stop) (ptmcmc.cc:589,657).
"""
import numpy as np
import torch
import torch.distributed as dist


def ladder_shard(n_ladders_total, rank, world):
    """contiguous block of ladders owned by `rank`: (global id of first ladder, count); remainders go to the low ranks"""
    if world < 1 or not (0 <= rank < world) or n_ladders_total < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(n_ladders_total, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def shard_config(make_cfg, n_ladders_total, rank, world, **kw):
    """engine config of this rank: `make_cfg(n_ladders=..., ladder_offset=..., **kw)`"""
    first, count = ladder_shard(n_ladders_total, rank, world)
    if count == 0:
        raise ValueError("rank %d of %d has no ladder (n_ladders_total=%d)" % (rank, world, n_ladders_total))
    return make_cfg(n_ladders=count, ladder_offset=first, **kw)


def _as_tensor(a, device):
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    return t.to(device)


def gather_cold_samples(local, n_ladders_total, device=None, dst=0):
    """local: [n_local_ladders, ...] array of this rank's cold-chain samples (leading axis = ladder).
    Returns on rank `dst` the [n_ladders_total, ...] array in global ladder order, None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local) if not isinstance(local, torch.Tensor) else local
    world, rank = dist.get_world_size(), dist.get_rank()
    device = device or ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = _as_tensor(local, device)
    counts = [ladder_shard(n_ladders_total, r, world)[1] for r in range(world)]
    assert t.shape[0] == counts[rank], "local ladder count does not match the shard layout"
    pad = max(counts)
    buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
    buf[: t.shape[0]] = t
    out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    if dist.get_backend() == "nccl":  # NCCL gather = all ranks send to dst
        dist.gather(buf, out, dst=dst)
    else:
        dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None
    full = torch.cat([out[r][: counts[r]] for r in range(world)], dim=0)
    return full if isinstance(local, torch.Tensor) else full.cpu().numpy()


def consensus_stop(flag, device=None):
    """True on every rank iff any rank raises `flag` (stop / checkpoint consensus)"""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return bool(flag)
    device = device or ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(t.item())


def max_over_ranks(value, device=None):
    """max of a python float over ranks (device-timed step durations are reported as the max over ranks)"""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    device = device or ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    device = device or ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
