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


def shard_rows_round_robin(n_rows: int, rank: int, world: int, device=None) -> torch.Tensor:
    """Row indices of an image for this rank, dealt round-robin (row r -> rank r % world; n_rows % world == 0).
    For camera frames: the subject sits in the middle of the image, and with contiguous row blocks the central ranks march
    several times the samples of the outer ones (8 GPUs: 2.85x instead of 7.36x)."""
    assert n_rows % world == 0, "frames are sharded by rows: the row count must divide by the world size"
    return torch.arange(rank, n_rows, world, device=device)


def gather_rows_round_robin(local: torch.Tensor, out: torch.Tensor = None, scratch: torch.Tensor = None) -> torch.Tensor:
    """Inverse of shard_rows_round_robin for a per-pixel output: local [rows, W, C] of every rank -> [rows * world, W, C]
    in image order on every rank (one all_gather_into_tensor + one strided copy).  `out` / `scratch` may be passed to
    avoid allocations in a frame loop (shapes [rows * world, W, C] and [world, rows, W, C])."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        if out is None:
            return local
        out.copy_(local)
        return out
    world = dist.get_world_size()
    rows, w, c = local.shape
    if scratch is None:
        scratch = torch.empty((world, rows, w, c), dtype=local.dtype, device=local.device)
    if out is None:
        out = torch.empty((rows * world, w, c), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(scratch.view(-1), local.reshape(-1))
    out.view(rows, world, w, c).copy_(scratch.permute(1, 0, 2, 3))          # image row r * world + k came from rank k
    return out


def allreduce_pending(he, average: bool = True) -> None:
    """All-reduce of a HashEnsemble's parked rank-1 table gradient (fused optimiser), once: a reduced gradient is marked.
    Same slot map and blend weights on every rank (n_timesteps <= 32: slot = timestep, cw_slots from the replicated
    time embedding) -> the [slots][entries][2] workspace is summed in
    place and the 1/world factor folded into the optimiser step; otherwise, or when a dense `.grad` exists as well
    (gradient accumulation: the dense part must be averaged too), it becomes a dense `.grad` for the dense reduction."""
    pend = he.pending_table_grad
    if pend is not None and pend.get("sync_work") is not None:
        pend.pop("sync_work").wait()          # overlapped reduction (overlap_table_allreduce): the current stream waits for it
    if pend is None or pend.get("reduced") or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    world = dist.get_world_size()
    if pend.get("slots_are_timesteps") and he.tables.grad is None:
        dist.all_reduce(pend["g_rank1"], op=dist.ReduceOp.SUM)
        pend["scale"] = float(pend.get("scale", 1.0)) / (world if average else 1)
        pend["reduced"] = True
    else:
        he.materialize_pending()


def overlap_table_allreduce(hash_ensemble, average: bool = True) -> None:
    """Overlap the table-gradient collective with the rest of the backward ("overlap ... as backward
    finishes").  The parked rank-1 workspace [slots][entries][2] (1.2 GB at T = 24) is complete once nsb_field_backward
    is enqueued -- a third of the way into the backward; the deformation-field backward (the largest backward kernel)
    and the dW reductions still follow.  This installs a hook that the training backward calls at that point: the
    all-reduce is issued on a side stream and runs concurrently with those kernels; FusedFieldsAdam.step() /
    allreduce_gradients() wait for it.  No-op without a process group or when the slot map is not rank-invariant."""
    import torch
    state = {"stream": None}

    def hook(he) -> None:
        pend = he.pending_table_grad
        if (pend is None or pend.get("reduced") or not dist.is_initialized() or dist.get_world_size() == 1
                or not pend.get("slots_are_timesteps") or he.tables.grad is not None):
            return
        g = pend["g_rank1"]
        if g.is_cuda:
            if state["stream"] is None:
                state["stream"] = torch.cuda.Stream(g.device)
            side = state["stream"]
            side.wait_stream(torch.cuda.current_stream(g.device))
            with torch.cuda.stream(side):
                pend["sync_work"] = dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True)
            g.record_stream(side)
        else:                                   # gloo (CPU tests): asynchronous work object, no streams
            pend["sync_work"] = dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True)
        pend["scale"] = float(pend.get("scale", 1.0)) / (dist.get_world_size() if average else 1)
        pend["reduced"] = True

    hash_ensemble.table_grad_hook = hook


_BIG = 1 << 24   # elements: tensors this large are reduced in place, not copied into the flat bucket


def allreduce_gradients(params, average: bool = True, hash_ensembles=(), reduce_pending: bool = True) -> None:
    """The training collective: gradients summed over ranks once per step, then divided by the world size.
    Small gradients (MLPs, embeddings) travel in ONE flat bucket; the table gradient (1.6 GB dense) is reduced in place
    -- copying it into a bucket and back costs three extra passes over it.  With the fused optimiser
    (nersemble_b200.optim) the table gradient is still in rank-1 form [slots][entries][2] when this runs: when the
    slot map is the same on every rank (n_timesteps <= 32: slot = timestep) that workspace is reduced directly
    (n_timesteps/32 of the dense volume) and the 1/world factor is folded into the optimiser step; otherwise it is
    expanded to a dense .grad first.  Parameters whose .grad is None on this rank contribute zeros so that every rank
    issues the same collectives."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    world = dist.get_world_size()
    for he in hash_ensembles:
        if reduce_pending:               # False: FusedFieldsAdam(shard_tables=True) reduce-scatters the parked gradient itself
            allreduce_pending(he, average)
    params = [p for p in params if p.requires_grad]
    deferred = {id(he.tables) for he in hash_ensembles if he.pending_table_grad is not None}
    params = [p for p in params if id(p) not in deferred]
    if not params:
        return
    small = []
    for p in params:
        if p.numel() >= _BIG:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
            if average:
                p.grad /= world
        else:
            small.append(p)
    if not small:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in small])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= world
    ofs = 0
    for p in small:
        n = p.numel()
        g = flat[ofs:ofs + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        ofs += n
