"""world_size-2 gloo test of the N>1 host logic (ray sharding + output gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nersemble_b200.distributed import gather_rays, shard_bounds, shard_rays


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 5, 4096, 2088960):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def _worker(rank, world, port, n):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n * 3, dtype=torch.float32).view(n, 3)
    local = shard_rays({"rgb": full}, rank, world)["rgb"]
    out = gather_rays(local * 2.0, n)          # "render" = x2 on the local shard
    assert torch.equal(out, full * 2.0)
    dist.destroy_process_group()


def _grad_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nersemble_b200.distributed import allreduce_gradients
    a = torch.nn.Parameter(torch.zeros(5)); b = torch.nn.Parameter(torch.zeros(2, 3)); c = torch.nn.Parameter(torch.zeros(4))
    a.grad = torch.full((5,), float(rank + 1)); b.grad = torch.arange(6.0).view(2, 3) * (rank + 1)
    if rank == 0:
        c.grad = torch.ones(4)            # rank 1 has no gradient for c
    allreduce_gradients([a, b, c])
    assert torch.allclose(a.grad, torch.full((5,), 1.5)) and torch.allclose(b.grad, torch.arange(6.0).view(2, 3) * 1.5)
    assert torch.allclose(c.grad, torch.full((4,), 0.5))
    dist.destroy_process_group()


def test_gloo_world2_gradient_allreduce():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_grad_worker, args=(2, port), nprocs=2, join=True)


def test_gloo_world2_shard_and_gather():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, 101), nprocs=2, join=True)


def _pending_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nersemble_b200.distributed import allreduce_gradients

    class FakeEnsemble:      # the protocol allreduce_gradients relies on (plugin/components.py: HashEnsemble)
        def __init__(self):
            self.tables = torch.nn.Parameter(torch.zeros(5, 32, 2))
            self.pending_table_grad = {"g_rank1": torch.full((3, 5, 2), float(rank + 1)), "cw_slots": torch.ones(3, 32),
                                       "n_slots": 3, "slots_are_timesteps": True}
    he = FakeEnsemble()
    w = torch.nn.Parameter(torch.zeros(4))
    w.grad = torch.full((4,), float(rank))
    allreduce_gradients([he.tables, w], hash_ensembles=[he])
    assert torch.equal(he.pending_table_grad["g_rank1"], torch.full((3, 5, 2), 3.0))
    assert he.pending_table_grad["scale"] == 0.5 and he.tables.grad is None
    assert torch.equal(w.grad, torch.full((4,), 0.5))
    dist.destroy_process_group()


def test_allreduce_reduces_the_deferred_rank1_table_gradient_in_place():
    """world_size 2, gloo: the parked [slots][entries][2] workspace is summed across ranks, the 1/world factor is left
    to the fused optimiser step, and no dense table gradient is created."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_pending_worker, args=(2, port), nprocs=2, join=True)


def _overlap_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nersemble_b200.distributed import allreduce_gradients, overlap_table_allreduce

    class FakeEnsemble:
        def __init__(self):
            self.tables = torch.nn.Parameter(torch.zeros(5, 32, 2))
            self.pending_table_grad = {"g_rank1": torch.full((3, 5, 2), float(rank + 1)), "cw_slots": torch.ones(3, 32),
                                       "n_slots": 3, "slots_are_timesteps": True}
            self.table_grad_hook = None
    he = FakeEnsemble()
    overlap_table_allreduce(he)
    he.table_grad_hook(he)                       # what the training backward calls once the gradient is parked
    assert he.pending_table_grad["reduced"] and he.pending_table_grad["scale"] == 0.5
    w = torch.nn.Parameter(torch.zeros(4)); w.grad = torch.full((4,), float(rank))
    allreduce_gradients([he.tables, w], hash_ensembles=[he])      # waits for the overlapped work, does not reduce twice
    assert "sync_work" not in he.pending_table_grad
    assert torch.equal(he.pending_table_grad["g_rank1"], torch.full((3, 5, 2), 3.0)) and he.pending_table_grad["scale"] == 0.5
    assert torch.equal(w.grad, torch.full((4,), 0.5)) and he.tables.grad is None
    dist.destroy_process_group()


def test_overlapped_table_gradient_allreduce_hook():
    """world_size 2, gloo: the hook the training backward calls issues the reduction asynchronously; the later
    allreduce_gradients / optimiser step only waits for it."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_overlap_worker, args=(2, port), nprocs=2, join=True)


def _shard_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nersemble_b200 import ops
    from nersemble_b200.optim import FusedFieldsAdam
    E, T = 12, 3
    calls = []

    def fake_step(tables, m, v, shadow, *, step, lr, betas, eps, weight_decay, grad=None, pending=None, grad_scale=1.0):
        """plain SGD stand-in with the kernel's interface: p -= lr * scale * sum_slots cw x g (what the shard must see)"""
        g = torch.einsum("sm,sef->emf", pending["cw_slots"], pending["g_rank1"]) * grad_scale
        tables -= lr * g
        shadow.copy_(tables.half())
        calls.append((tuple(tables.shape), float(grad_scale)))
    ops.table_adam_step = fake_step

    class FakeEnsemble:
        def __init__(self):
            self.tables = torch.nn.Parameter(torch.zeros(E, 32, 2))
            self.defer_table_grad = False
            self.pending_table_grad = None
            self._shadow = torch.zeros(E, 32, 2, dtype=torch.float16)
        def shadow_buffer(self): return self._shadow
        def set_native_tables(self, s): self.native = s
        def materialize_pending(self, scale=1.0): raise AssertionError("dense path must not run")
    he = FakeEnsemble()
    import weakref
    he.tables._nsb_hash_ensemble = weakref.ref(he)
    opt = FusedFieldsAdam([he.tables], lr=1.0)
    opt.shard_tables = True
    gen = torch.Generator().manual_seed(0)
    g_all = [torch.randn(T, E, 2, generator=gen) for _ in range(world)]           # every rank knows every rank's gradient
    cw = torch.rand(T, 32, generator=gen)
    he.pending_table_grad = {"g_rank1": g_all[rank].clone(), "cw_slots": cw, "n_slots": T, "slots_are_timesteps": True}
    with torch.no_grad():
        opt.step()
    want = -torch.einsum("sm,sef->emf", cw, sum(g_all) / world)                   # averaged gradient, lr = 1
    n = E // world
    assert calls == [((n, 32, 2), 1.0 / world)]                                   # the kernel saw only this rank's entries
    # fp16 table: complete on every rank; fp32 master: the owner's range only, until consolidate()
    torch.testing.assert_close(he._shadow.float(), want, rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(he.tables.detach()[rank * n:(rank + 1) * n], want[rank * n:(rank + 1) * n], rtol=1e-6, atol=1e-6)
    other = 1 - rank
    assert torch.equal(he.tables.detach()[other * n:(other + 1) * n], torch.zeros(n, 32, 2))
    opt.consolidate()
    torch.testing.assert_close(he.tables.detach(), want, rtol=1e-6, atol=1e-6)
    assert he.pending_table_grad is None
    dist.destroy_process_group()


def test_sharded_table_optimiser_reduce_scatter_step_all_gather():
    """world_size 2, gloo: FusedFieldsAdam(shard_tables=True) hands the kernel only this rank's 1/N of the entries with
    the reduced gradient, all-gathers the fp16 table, and consolidate() completes the fp32 master."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_shard_worker, args=(2, port), nprocs=2, join=True)


def _rows_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nersemble_b200.distributed import gather_rows_round_robin, shard_rows_round_robin
    H, W = 12, 5
    image = torch.arange(H * W * 3, dtype=torch.float32).view(H, W, 3)
    rows = shard_rows_round_robin(H, rank, world)
    assert rows.tolist() == list(range(rank, H, world))
    local = image[rows] * 2.0                        # "render" = x2 on this rank's rows
    out = gather_rows_round_robin(local)
    assert torch.equal(out, image * 2.0)             # image order restored on every rank
    frame, scratch = torch.empty((H, W, 3)), torch.empty((world, H // world, W, 3))
    assert gather_rows_round_robin(local, out=frame, scratch=scratch) is frame and torch.equal(frame, image * 2.0)
    dist.destroy_process_group()


def test_gloo_world2_round_robin_frame_rows():
    """Frames are sharded by rows dealt round-robin (load balance: the subject is in the middle of the image); the
    all-gather + strided copy puts the rows back in image order on every rank.  world_size 1 is the identity."""
    from nersemble_b200.distributed import gather_rows_round_robin, shard_rows_round_robin
    assert shard_rows_round_robin(6, 0, 1).tolist() == [0, 1, 2, 3, 4, 5]
    x = torch.rand(6, 4, 3)
    assert gather_rows_round_robin(x) is x
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_rows_worker, args=(2, port), nprocs=2, join=True)
