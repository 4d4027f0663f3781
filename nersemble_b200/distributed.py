"""Ray sharding across ranks (one process per GPU).  Rays are independent units: each rank renders a contiguous
slice with a full parameter replica; there is no data-path collective.  The only exchange is the final gather of
per-ray outputs to every rank (config 4: 25 MB per 1088x1920 frame).  Works with NCCL (CUDA tensors) and gloo (CPU)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced: the first n % world ranks get one extra ray."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(tensors: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    n = next(iter(tensors.values())).shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return {k: v[lo:hi] for k, v in tensors.items()}


def gather_rays(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """all_gather of ragged per-ray outputs [n_local, C] -> [n_total, C] (same order as before sharding)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    max_n = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)


def allreduce_gradients(params, average: bool = True) -> None:
    """The training collective: ONE all-reduce of the flat gradient buffer per step
    (403 M table gradients + MLP/embedding gradients), then divide by the world size.  Parameters whose .grad is
    None on this rank contribute zeros so that every rank issues the same collective."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= dist.get_world_size()
    ofs = 0
    for p in params:
        n = p.numel()
        g = flat[ofs:ofs + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        ofs += n
