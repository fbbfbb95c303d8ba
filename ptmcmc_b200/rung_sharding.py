"""Rung-sharded temperature ladders: the optional multi-GPU layout with an exchange step (BASELINE config 5, "cross-GPU rung
swaps over NVLink").  It mirrors the reference's MPI scheme (chain.cc:1290-1311: rungs dealt to ranks, every rank replays the same
swap decisions from the shared seed, only data is communicated, chain.cc:1433-1435) with BLOCK rung assignment, so that only
world-1 rung boundaries cross NVLink:

    rank g owns rungs [g*R, (g+1)*R) of EVERY ladder of a (world*R)-rung geometric ladder (ptg_set_betas);
    swaps inside a block run in the step kernels (reference schedule over the local rungs);
    every `exchange_every` PT iterations neighbouring ranks exchange their edge rungs -- (x, llike, lprior, beta) per ladder,
    [n_ladders, dim+3] doubles each way over NCCL send/recv -- and both sides of a boundary evaluate the same swap trial
    (ptg_boundary_swap: acceptance draw from the ladder's Philox stream under the shared key), so decisions never travel.

The ladder-sharded layout (sharding.py) needs no exchange step and is the default; this one exists for ladders too long for
one GPU's warp/CTA (more than 32 rungs per ladder at full speed) and to measure the collective's cost.
"""
import numpy as np
import torch
import torch.distributed as dist


def global_betas(n_rungs_total, Tmax):
    """inverse temperatures of the geometric ladder (chain.cc:1181-1183,1339): temps[i] = temps[i-1] * Tmax^(1/(N-1))"""
    tratio = np.exp(np.log(Tmax) / (n_rungs_total - 1)) if n_rungs_total > 1 else 1.0
    temps = np.empty(n_rungs_total)
    t = 1.0
    for i in range(n_rungs_total):
        if i > 0:
            t = t * tratio
        temps[i] = t
    return 1.0 / temps


def rank_betas(n_ladders, rungs_per_rank, rank, world, Tmax):
    """[n_ladders * rungs_per_rank] betas of this rank's block of every ladder"""
    b = global_betas(rungs_per_rank * world, Tmax)[rank * rungs_per_rank:(rank + 1) * rungs_per_rank]
    return np.tile(b, n_ladders)


def rank_chain_seed(base_seed, rank):
    """Philox key of rank `rank`'s chain streams.  Every rank numbers its ladders and LOCAL rungs identically, so with one shared key the
    rung k of rank 0 and the rung k of rank 1 would consume identical streams: each rank steps its chains under its own key (the boundary
    trials use the SHARED key given to the drivers)."""
    return (int(base_seed) + 977 * int(rank)) & 0xFFFFFFFFFFFFFFFF


def _check_distinct_chain_seeds(api, world, peers):
    """the drivers refuse a family whose ranks share a chain key (correlated proposals across rung blocks)"""
    if world <= 1 or peers is not None or not dist.is_initialized():
        return
    seeds = [None] * world
    dist.all_gather_object(seeds, int(api.cfg.seed))
    if len(set(seeds)) != world:
        raise ValueError("rung-sharded ladders: every rank needs its own chain seed (use rank_chain_seed(base, rank)); got %r" % (seeds,))


class RungShardedLadders:
    """Drives one rank's engine (anything with the C-ABI methods step / boundary_pack / boundary_swap) in the rung-sharded layout."""

    def __init__(self, api, rank, world, shared_seed, exchange_every=10, device="cpu", stream_ordered=False):
        """stream_ordered: the engine launches on torch's CURRENT CUDA stream (Engine.set_stream(torch stream) inside
        `with torch.cuda.stream(...)`), so packs, the NCCL all_gather and the swap kernels are ordered on the device and the host
        never waits inside the loop.  Otherwise the host synchronises around the collective."""
        self.api, self.rank, self.world, self.shared_seed, self.exchange_every = api, rank, world, int(shared_seed), exchange_every
        _check_distinct_chain_seeds(api, world, None)
        self.device = torch.device(device)
        self.stream_ordered = bool(stream_ordered) and self.device.type == "cuda"
        L, d = api.cfg.n_ladders, api.cfg.dim
        self.R = api.cfg.n_rungs
        # edges[0] = my coldest rung (bottom), edges[1] = my hottest rung (top); nbr[0] = the hotter neighbour's bottom, nbr[1] = the
        # colder neighbour's top.  Only neighbours talk: two send/recv pairs per rank per exchange, whatever the world size.
        self.edges = torch.zeros((2, L, d + 3), dtype=torch.float64, device=self.device)
        self.nbr = torch.zeros((2, L, d + 3), dtype=torch.float64, device=self.device)
        self.n_exchanges = 0

    def _neighbour_exchange(self):
        g, ops = self.rank, []
        if g + 1 < self.world:
            ops += [dist.P2POp(dist.isend, self.edges[1], g + 1), dist.P2POp(dist.irecv, self.nbr[0], g + 1)]
        if g > 0:
            ops += [dist.P2POp(dist.isend, self.edges[0], g - 1), dist.P2POp(dist.irecv, self.nbr[1], g - 1)]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()  # NCCL: orders the current stream after the transfer, the host does not block

    def exchange(self):
        """one cross-boundary swap trial per ladder and per rung boundary"""
        api, R, g = self.api, self.R, self.rank
        api.boundary_pack(0, self.edges[0].data_ptr())
        api.boundary_pack(R - 1, self.edges[1].data_ptr())
        if self.device.type == "cuda" and not self.stream_ordered:
            api.synchronize()  # packs were written on the engine's stream
        self._neighbour_exchange()
        if self.device.type == "cuda" and not self.stream_ordered:
            torch.cuda.current_stream().synchronize()
        if g + 1 < self.world:   # my hottest rung with the neighbour's coldest: I hold the lower index of the pair
            api.boundary_swap(R - 1, self.nbr[0].data_ptr(), True, self.shared_seed, g, self.n_exchanges)
        if g > 0:
            api.boundary_swap(0, self.nbr[1].data_ptr(), False, self.shared_seed, g - 1, self.n_exchanges)
        self.n_exchanges += 1

    def run(self, n_steps):
        """n_steps PT iterations with an exchange after every `exchange_every`"""
        done = 0
        while done < n_steps:
            k = min(self.exchange_every, n_steps - done)
            self.api.step(k)
            self.exchange()
            done += k

    def finish(self):
        pass


class FusedRungShardedLadders:
    """The same layout and the same chains (same Philox addresses, so bit-identical states) with the exchange step fused into the
    production step kernel over NVLink peer memory (ptg_xchg_* / ptg_step_exchange): the epilogue of each launch publishes the edge
    rungs in this rank's own HBM, the prologue of the next launch reads the neighbours' records through CUDA-IPC peer pointers once
    their per-ladder flags are up.  No collective, no extra launch and no host synchronisation inside the loop; one rank per GPU."""

    def __init__(self, api, rank, world, shared_seed, exchange_every=10, peers=None, in_launch=False, max_launch=1000):
        """peers: {rank: device pointer} for engines of the SAME process (tests); otherwise handles travel over torch.distributed.
        in_launch: run up to `max_launch` iterations per launch with the exchange INSIDE the kernel every `exchange_every`
        iterations (needs one rank per GPU and the whole grid resident); otherwise one launch per exchange."""
        self.api, self.rank, self.world, self.exchange_every = api, rank, world, exchange_every
        self.in_launch, self.max_launch = bool(in_launch), max(exchange_every, max_launch - max_launch % exchange_every)
        _check_distinct_chain_seeds(api, world, peers)
        handle, ptr = api.xchg_export()
        if peers is None:
            handles = [None] * world
            if world > 1:
                dist.all_gather_object(handles, handle)
            lo = handles[rank - 1] if rank > 0 else None
            hi = handles[rank + 1] if rank + 1 < world else None
        else:
            lo = peers.get(rank - 1) if rank > 0 else None
            hi = peers.get(rank + 1) if rank + 1 < world else None
        api.xchg_connect(lo, hi, int(shared_seed), rank - 1, rank)
        self.pending = False
        self.n_exchanges = 0

    def run(self, n_steps):
        """n_steps PT iterations, edges published after every `exchange_every`; each launch first applies the exchange that the
        previous launch left pending (finish() applies the last one)"""
        done = 0
        while done < n_steps:
            if self.in_launch and n_steps - done >= self.exchange_every:
                k = min(self.max_launch, (n_steps - done) // self.exchange_every * self.exchange_every)
                self.api.step_exchange(k, self.pending, True, every=self.exchange_every)
                self.n_exchanges += k // self.exchange_every
            else:
                k = min(self.exchange_every, n_steps - done)
                self.api.step_exchange(k, self.pending, True)
                self.n_exchanges += 1
            self.pending = True
            done += k

    def finish(self):
        if self.pending:
            self.api.step_exchange(0, True, False)
            self.pending = False
