"""Fused hash-table optimiser (nsb_table_adam_step / nsb_rank1_expand, nersemble_b200/optim.py) against torch.optim.Adam --
the optimiser the reference's recipe runs over the 8 tcnn grid tensors (train_nersemble.py: AdamOptimizerConfig, eps 1e-15)."""
import os, sys
import pytest, torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand_pending(E, n_slots, gen):
    g1 = torch.zeros(n_slots, E, 2)
    touched = torch.rand(n_slots, E, generator=gen) < 0.3          # most (slot, line) pairs untouched, like a real batch
    g1[touched] = torch.randn(int(touched.sum()), 2, generator=gen) * 1e-3
    cw = torch.rand(n_slots, 32, generator=gen).half().float()
    return {"g_rank1": g1.to(DEV), "cw_slots": cw.to(DEV), "n_slots": n_slots}, torch.einsum("sm,sef->emf", cw, g1)


@pytest.mark.parametrize("E,n_slots", [(1000, 24), (37, 1), (4099, 32)])
def test_rank1_expand_matches_einsum(E, n_slots):
    from nersemble_b200 import ops
    gen = torch.Generator().manual_seed(E)
    pend, dense = _rand_pending(E, n_slots, gen)
    base = torch.randn(E, 32, 2, generator=gen)
    out = ops.rank1_expand(pend, E, grad_scale=0.5, out=base.clone().to(DEV))
    torch.testing.assert_close(out.cpu(), base + 0.5 * dense, rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("mode", ["dense", "rank1", "both"])
def test_table_adam_step_matches_torch_adam(mode):
    from nersemble_b200 import ops
    E, n_slots = 2053, 24
    gen = torch.Generator().manual_seed(7)
    p0 = (torch.rand(E, 32, 2, generator=gen) * 2 - 1) * 1e-2
    ref = torch.nn.Parameter(p0.clone().to(DEV))
    opt = torch.optim.Adam([ref], lr=5e-3, eps=1e-15, foreach=False)
    p = p0.clone().to(DEV)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    shadow = torch.empty(p.shape, dtype=torch.float16, device=DEV)
    for step in range(1, 6):
        pend, dense_r1 = _rand_pending(E, n_slots, gen)
        dense = torch.randn(E, 32, 2, generator=gen) * 1e-3
        dense[torch.rand(E, generator=gen) < 0.5] = 0          # lines with an exactly-zero gradient still decay/move
        total = {"dense": dense, "rank1": dense_r1, "both": dense + dense_r1}[mode]
        ref.grad = (0.25 * total).to(DEV)
        opt.step()
        ops.table_adam_step(p, m, v, shadow, step=step, lr=5e-3, eps=1e-15,
                            grad=dense.to(DEV) if mode != "rank1" else None,
                            pending=pend if mode != "dense" else None, grad_scale=0.25)
        st = opt.state[ref]
        torch.testing.assert_close(m, st["exp_avg"], rtol=2e-5, atol=1e-10)      # values ~1e-4: lerp-form rounding
        torch.testing.assert_close(v, st["exp_avg_sq"], rtol=2e-5, atol=1e-20)
        # a step is at most lr per element; allow 1e-4 of that
        assert (p - ref.detach()).abs().max().item() < 5e-3 * 1e-4 * step
        assert torch.equal(shadow, p.half())


def test_fused_fields_adam_tracks_torch_adam_on_the_model():
    """Same model, same batches: FusedFieldsAdam (deferred rank-1 gradient, fused step, fp16 shadow) vs torch Adam on
    the dense gradient.  Checks the parameters after 3 steps and that the render uses the refreshed tables."""
    from test_plugin_cpu import make_model
    from oracle.gen_golden import ring_rays, blob_grid
    from nersemble_b200.nerfstudio_shim import RayBundle
    from nersemble_b200.optim import FusedFieldsAdam

    def run(fused):
        torch.manual_seed(0)
        m = make_model(T=4, log2T=14, lambda_near_loss=0, lambda_empty_loss=0, lambda_depth_loss=0).to(DEV).train()
        with torch.no_grad():
            m.field.hash_ensemble.tables.uniform_(-0.5, 0.5)
            m.time_embedding.weight.normal_(0, 0.18)
        occ = blob_grid(3)
        m.occupancy_grid.binaries[0] = occ.to(DEV)
        m.occupancy_grid.occs.copy_((occ.flatten().float() * 0.05).to(DEV))
        m.sampler.eval()
        t0 = m.field.hash_ensemble.tables.detach().clone()
        groups = m.get_param_groups()
        # eps 1e-8 here: the scatter uses float atomics, so two runs differ by rounding order, and with the recipe's
        # eps = 1e-15 Adam turns the sign of a cancelling ~1e-12 gradient into a full lr step (true of the reference's
        # tcnn backward as well).  Exact agreement at eps = 1e-15 is the identical-gradient test above.
        fields = (FusedFieldsAdam if fused else torch.optim.Adam)(groups["fields"], lr=5e-3, eps=1e-8)
        rest = torch.optim.Adam(groups["embeddings"], lr=5e-3, eps=1e-8)
        losses = []
        for it in range(3):
            o, d, times, cams = ring_rays(64, 21 + it)
            gen = torch.Generator().manual_seed(it)
            batch = {"image": torch.rand((64, 3), generator=gen), "alpha_map": torch.randint(0, 256, (64, 1), generator=gen).float()}
            rb = RayBundle(origins=o.to(DEV), directions=d.to(DEV), pixel_area=torch.ones(64, 1, device=DEV),
                           camera_indices=cams.to(DEV), times=times.to(DEV))
            fields.zero_grad(); rest.zero_grad()
            loss = sum(m.get_loss_dict(m.get_outputs(rb), batch).values())
            loss.backward()
            if fused:
                assert m.field.hash_ensemble.tables.grad is None and m.field.hash_ensemble.pending_table_grad is not None
            fields.step(); rest.step()
            losses.append(loss.item())
        he = m.field.hash_ensemble
        assert torch.equal(he.native_tables(), he.tables.detach().half())      # the cache the kernels read is current
        return losses, he.tables.detach() - t0, m.field.mlp_base.params.detach().clone()

    l_f, t_f, b_f = run(True)
    l_t, t_t, b_t = run(False)
    for a, b in zip(l_f, l_t):
        assert abs(a - b) < 1e-4 * abs(b) + 1e-6, (l_f, l_t)
    # the table UPDATES of the two paths agree (element-wise equality is not defined: float atomics + Adam's normalisation)
    cos = torch.nn.functional.cosine_similarity(t_f.reshape(1, -1), t_t.reshape(1, -1)).item()
    assert cos > 0.995, cos
    assert (t_f != 0).float().mean().item() > 0.01            # and they are real updates
    assert (b_f - b_t).abs().max().item() < 5e-3 * 0.05


@pytest.mark.parametrize("training", [True, False])
def test_fused_losses_match_the_torch_formulation_values_and_gradients(training):
    """nsb_losses_forward/backward vs the plugin's torch formulation of models/base.py:90-249 (itself pinned to the
    reference glue's golden in tests/test_plugin_cpu.py): six loss values and the gradients w.r.t. rgb / accumulation /
    depth / per-sample weights, ragged rays including empty ones, upstream gradients != 1."""
    from test_plugin_cpu import make_model
    from nersemble_b200.nerfstudio_shim import Frustums, RaySamples
    g = torch.Generator().manual_seed(11)
    R = 300
    cnt = torch.randint(0, 70, (R,), generator=g); cnt[5] = 0; cnt[R - 1] = 0; cnt[17] = 64
    S = int(cnt.sum())
    ri = torch.repeat_interleave(torch.arange(R), cnt)
    ts = torch.cat([torch.sort(torch.rand(int(c), generator=g) * 6 + 5)[0] for c in cnt]) if S else torch.zeros(0)
    te = ts + 0.011
    m = make_model(T=4, log2T=12, dist_loss_max_rays=250).to(DEV)
    m.train(training)
    m.sched_eps_depth.value = 0.35
    w = (torch.rand(S, generator=g) * 0.05).to(DEV).requires_grad_(True)
    rgb = torch.rand(R, 3, generator=g).to(DEV).requires_grad_(True)
    acc = torch.rand(R, 1, generator=g).to(DEV).requires_grad_(True)
    depth = (torch.rand(R, 1, generator=g) * 6 + 5).to(DEV).requires_grad_(True)
    rs = RaySamples(Frustums(torch.zeros(S, 3, device=DEV), torch.zeros(S, 3, device=DEV), ts[:, None].to(DEV), te[:, None].to(DEV), torch.zeros(S, 1, device=DEV)))
    outputs = {"rgb": rgb, "accumulation": acc, "depth": depth, "ray_samples": (rs,), "ray_indices": (ri.to(DEV),),
               "weights": (w[:, None],), "num_samples_per_ray": cnt.to(DEV)}
    batch = {"image": torch.rand(R, 3, generator=g), "alpha_map": torch.randint(0, 256, (R, 1), generator=g).float(),
             "depth_maps": (torch.rand(R, generator=g) * 6 + 5) * (torch.rand(R, generator=g) > 0.3)}
    batch["alpha_map"][:20] = 255.0; batch["alpha_map"][20:30] = 0.0
    up = {k: 0.5 + i * 0.37 for i, k in enumerate(("rgb_loss", "alpha_loss", "empty_loss", "near_loss", "depth_loss", "dist_loss"))}

    def run(fused):
        m.fused_losses = fused
        for t in (w, rgb, acc, depth):
            t.grad = None
        ld = m.get_loss_dict(outputs, batch)
        sum(up[k] * v for k, v in ld.items()).backward()
        return {k: v.item() for k, v in ld.items()}, [torch.zeros_like(t) if t.grad is None else t.grad.clone() for t in (w, rgb, acc, depth)]

    v_f, g_f = run(True)
    v_t, g_t = run(False)
    assert set(v_f) == set(v_t) and (len(v_f) == 6 if training else set(v_f) == {"rgb_loss", "alpha_loss", "dist_loss"})
    for k in v_t:
        assert abs(v_f[k] - v_t[k]) <= 2e-4 * abs(v_t[k]) + 1e-9, (k, v_f[k], v_t[k])
    for a, b, name in zip(g_f, g_t, ("weights", "rgb", "acc", "depth")):
        scale = b.abs().max().item() + 1e-12
        assert (a - b).abs().max().item() <= 2e-4 * scale, (name, (a - b).abs().max().item(), scale)


def test_fused_fields_adam_cooperates_with_grad_scaler():
    """engine/nersemble_trainer.py:182-186 trains with `scaler.scale(loss).backward(); scaler.step(opt); scaler.update()`.
    The parked rank-1 table gradient is no `.grad`: FusedFieldsAdam must unscale it itself (moments at the TRUE gradient
    scale whatever the scaler does), and a non-finite parked gradient must skip the whole step and back the scale off."""
    from nersemble_b200.optim import FusedFieldsAdam
    from nersemble_b200.plugin.components import HashEnsemble, HashEnsembleConfig, TCNNHashEncodingConfig
    gen = torch.Generator().manual_seed(3)

    def make():
        he = HashEnsemble(HashEnsembleConfig(32, TCNNHashEncodingConfig(log2_hashmap_size=12), True, True), seed=1).to(DEV)
        other = torch.nn.Parameter(torch.ones(7, device=DEV))
        return he, other, FusedFieldsAdam([he.tables, other], lr=5e-3, eps=1e-15)

    E = make()[0].tables.shape[0]
    pends = [_rand_pending(E, 4, gen)[0] for _ in range(3)]
    og = [torch.randn(7, generator=gen).to(DEV) for _ in range(3)]

    def run(scale_seq, use_scaler):
        he, other, opt = make()
        scaler = torch.amp.GradScaler("cuda", init_scale=scale_seq[0], growth_interval=10 ** 9) if use_scaler else None
        if use_scaler:
            scaler.scale(torch.zeros(1, device=DEV))         # lazy initialisation of the scale tensor
        for it in range(3):
            s = scale_seq[it]
            opt.zero_grad()
            he.pending_table_grad = {**pends[it], "g_rank1": pends[it]["g_rank1"] * s, "slots_are_timesteps": True}
            other.grad = og[it] * s
            if use_scaler:
                scaler._scale.fill_(s)                       # what backoff/growth would do between steps
                scaler.step(opt); scaler.update()
                assert not opt.last_step_skipped
            else:
                opt.step()
        st = opt.state[he.tables]
        return he.tables.detach().clone(), st["exp_avg_sq"].clone(), other.detach().clone()

    t_ref, v_ref, o_ref = run([1.0, 1.0, 1.0], False)
    t_amp, v_amp, o_amp = run([65536.0, 1024.0, 32768.0], True)          # a scale that changes between steps
    torch.testing.assert_close(v_amp, v_ref, rtol=1e-4, atol=1e-30)      # second moments at the true gradient scale
    assert (t_amp - t_ref).abs().max().item() < 5e-3 * 1e-3
    torch.testing.assert_close(o_amp, o_ref, rtol=1e-5, atol=1e-7)

    # a non-finite gradient (the training backward poisons mlp_base's .grad when the table side overflows, see
    # test_plugin_gpu.py): the WHOLE step is skipped -- table, moments, the other parameter untouched -- and the scale halves
    he, other, opt = make()
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    scaler.scale(torch.zeros(1, device=DEV))
    t0 = he.tables.detach().clone()
    he.pending_table_grad = {**pends[0], "slots_are_timesteps": True}
    other.grad = og[0] * 1024.0
    other.grad[3] = float("nan")
    scaler.step(opt); scaler.update()
    assert opt.last_step_skipped and torch.equal(he.tables.detach(), t0) and torch.equal(other.detach(), torch.ones(7, device=DEV))
    assert len(opt.state[he.tables]) == 0 and scaler.get_scale() == 512.0 and he.pending_table_grad is None
