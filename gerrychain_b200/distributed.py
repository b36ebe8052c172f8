"""Multi-GPU plumbing: chains shard over ranks, nothing is exchanged per step.

The reference parallelises only by running whole chains in separate processes and summing a
statistic afterwards (/root/reference/tests/test_region_aware.py:89-95).  Here rank r owns
chains [r * n / world, (r + 1) * n / world) with Philox subsequences equal to their GLOBAL
chain id (RcConfig.chain_offset), so an ensemble does not depend on how many GPUs ran it;
the only collective is one all-reduce(SUM, int64) of the ensemble histograms (NCCL over
NVLink on GPUs, gloo in the CPU tests).
"""

from __future__ import annotations

from typing import Tuple


def shard_chains(n_chains_total: int, rank: int, world: int) -> Tuple[int, int]:
    """(chain_offset, n_local) of this rank; the first ``n % world`` ranks take one extra."""
    base, extra = divmod(int(n_chains_total), int(world))
    n_local = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, n_local


def all_reduce_histograms(*tensors):
    """In-place SUM over ranks; no-op when torch.distributed is not initialised."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in tensors:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return tensors


def gather_chain_scalars(t):
    """All ranks' per-chain values (e.g. cut-edge counts at a fixed step) on every rank, in
    global chain order; used by the ensemble (KS) checks."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    sizes = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(dist.get_world_size())]
    dist.all_gather(sizes, torch.tensor([t.numel()], dtype=torch.int64, device=t.device))
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
    pad[: t.numel()] = t
    out = [torch.zeros_like(pad) for _ in sizes]
    dist.all_gather(out, pad)
    return torch.cat([o[: int(s.item())] for o, s in zip(out, sizes)])
