"""Sharding of independent Markov chains (walkers) over ranks and the one collective of the path.

Chains are independent units: rank r evaluates a contiguous block of chains
with no data-path communication; the only exchange is one small all-reduce per measurement of
[sum log|det G|, sum Re sgn, sum Im sgn, number of chains] (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_chains(total: int, world: int, rank: int) -> range:
    """Contiguous, balanced partition of ``range(total)``: the first ``total % world`` ranks get one extra."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def weak_chain_ids(chains_per_rank: int, rank: int) -> range:
    """Weak scaling: every rank owns ``chains_per_rank`` chains; ids are global so that a chain's
    Hubbard-Stratonovich fields do not depend on the number of ranks."""
    return range(rank * chains_per_rank, (rank + 1) * chains_per_rank)


def reduce_walkers(logdet: torch.Tensor, sign: torch.Tensor, out: torch.Tensor | None = None, group=None):
    """Cross-walker sums [sum logdet, sum Re(sign), sum Im(sign), count], all-reduced over the ranks.
    Runs on the tensors' device and stream (NCCL orders the collective on the current stream).  On a GPU the local sums
    are ONE library kernel (``sla_walker_sums``) writing the 4-double buffer that goes straight into the all-reduce."""
    if out is None:
        out = torch.zeros(4, dtype=torch.float64, device=logdet.device)
    if logdet.is_cuda:
        from . import api
        api.walker_sums(logdet, sign, out)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(out, group=group)
        return out
    s = sign.sum()
    vals = [logdet.sum(), s.real if sign.is_complex() else s,
            s.imag if sign.is_complex() else torch.zeros((), dtype=torch.float64, device=logdet.device),
            torch.tensor(float(logdet.numel()), dtype=torch.float64, device=logdet.device)]
    torch.stack([v.to(torch.float64) for v in vals], out=out)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, group=group)
    return out
