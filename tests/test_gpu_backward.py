"""Backward kernels vs autograd through the CPU oracle ("kernel" precision mode: straight-through fp16 roundings)."""
import pytest
import torch

from conftest import native_from_oracle, oracle_params
from oracle import pipeline as pl
from oracle.tp import nerfacc_cpu
from oracle.tp.tcnn_cpu import Precision

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TRAINED = dict(seed=19980801, n_timesteps=4, log2_hashmap_size=14, table_scale=0.5, time_std_scale=100.0,
               deform_last_scale=1e-3)


@pytest.fixture(autouse=True)
def _mode():
    Precision.mode = "kernel"; Precision.autocast = False
    yield
    Precision.mode = "reference"


def _relerr(got, want):
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()


def test_composite_backward_vs_autograd():
    from nersemble_b200 import ops
    g = torch.Generator().manual_seed(3)
    R = 37
    cnt = torch.randint(0, 80, (R,), generator=g); cnt[5] = 0; cnt[11] = 1
    ri = torch.repeat_interleave(torch.arange(R), cnt)
    n = ri.numel()
    info = nerfacc_cpu.pack_info(ri, R)
    base = torch.rand((R,), generator=g) * 3 + 5
    k = torch.arange(n) - info[:, 0][ri]
    ts = base[ri] + k * 0.011; te = ts + 0.011
    sigma = (torch.rand((n,), generator=g) * 20).requires_grad_()
    rgb = torch.rand((n, 3), generator=g).requires_grad_()
    g_rgb = torch.randn((R, 3), generator=g); g_acc = torch.randn((R,), generator=g)
    g_dep = torch.randn((R,), generator=g); g_w = torch.randn((n,), generator=g)
    w = nerfacc_cpu.render_weight_from_density(ts, te, sigma, info)[0]
    comp = nerfacc_cpu.accumulate_along_rays(w, rgb, ri, R)
    acc = nerfacc_cpu.accumulate_along_rays(w, None, ri, R)
    comp = comp + (1.0 - acc)
    mid = ((ts + te) / 2)[:, None]
    depth = nerfacc_cpu.accumulate_along_rays(w, mid, ri, R) / (acc + 1e-10)
    depth = torch.clip(depth, mid.min(), mid.max())
    loss = (comp * g_rgb).sum() + (acc[:, 0] * g_acc).sum() + (depth[:, 0] * g_dep).sum() + (w * g_w).sum()
    loss.backward()
    fwd = ops.composite(info.to(DEV), ts.to(DEV), te.to(DEV), sigma.detach().to(DEV), rgb.detach().to(DEV), training=True)
    torch.testing.assert_close(fwd["rgb"].cpu(), comp.detach(), rtol=1e-4, atol=1e-5)
    d_sigma, d_rgb = ops.composite_backward(info.to(DEV), ts.to(DEV), te.to(DEV), sigma.detach().to(DEV), rgb.detach().to(DEV),
                                            fwd["workspace"], g_rgb.to(DEV), g_acc.to(DEV), g_dep.to(DEV), g_w.to(DEV))
    torch.testing.assert_close(d_rgb.cpu(), rgb.grad, rtol=1e-4, atol=1e-6)
    assert _relerr(d_sigma.cpu(), sigma.grad) < 2e-4


@pytest.mark.parametrize("rank1", [True, False, "saved_corners", "deferred"])
def test_field_backward_vs_autograd(rank1):
    """Gradients of the hash tables, time codes and both tiny MLPs (no deformation field); rank1 = the
    per-timestep 2-vector scatter + dense expansion, else the direct 64-float-per-line scatter;
    "saved_corners" = rank-1 scatter from the corner values the training forward saved (no table re-gather, merged
    atomics); "deferred" = additionally the table gradient stays in rank-1 form and is expanded by nsb_rank1_expand."""
    cv = rank1 in ("saved_corners", "deferred")
    deferred = rank1 == "deferred"
    rank1 = bool(rank1)
    from nersemble_b200 import ops
    P = pl.random_params(**TRAINED)
    NP = native_from_oracle(P, DEV)
    g = torch.Generator().manual_seed(5)
    n = 700                                              # ragged: 5 full tiles + 60
    lo, hi = P.aabb[0], P.aabb[1]
    pos = lo + (torch.rand((n, 3), generator=g) * 1.1 - 0.05) * (hi - lo)     # a few outside the box
    dirs = torch.randn((n, 3), generator=g); dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    tsteps = torch.sort(torch.randint(0, 4, (n,), generator=g))[0]
    times = tsteps.float()[:, None] / 3
    w_hash = 20.25
    P.requires_grad_(True)
    pos = pos.clone().requires_grad_(True)
    sigma, geo = pl.field_density(P, pos, P.time_emb[tsteps], w_hash)
    rgb = pl.field_rgb(P, dirs, geo)
    g_sigma = torch.randn((n,), generator=g) * 0.1
    g_rgb = torch.randn((n, 3), generator=g)
    ((sigma[:, 0] * g_sigma).sum() + (rgb * g_rgb).sum()).backward()

    kw = dict(positions=pos.detach().to(DEV), sample_times=times.to(DEV), sample_directions=dirs.to(DEV))
    saved = ops.field_forward(NP, window_hash=w_hash, use_deformation=False,
                              want=("sigma", "rgb", "feat", "xs") + (("corner_vals",) if cv else ()), **kw)
    torch.testing.assert_close(saved["sigma"].cpu(), sigma[:, 0].detach(), rtol=5e-3, atol=1e-5)
    grads = ops.field_backward(NP, saved, g_sigma.to(DEV), g_rgb.to(DEV), window_hash=w_hash, loss_scale=128.0, want_dx=True,
                               rank1=rank1, defer_tables=deferred, **kw)
    if deferred:
        assert "d_tables" not in grads
        grads["d_tables"] = ops.rank1_expand(grads["pending"], P.tables.shape[0])
    # dL/d(world position) = dL/d(normalised) / aabb size   (tcnn kernel_grid_backward_input)
    dpos = grads["d_xs"].cpu() / (hi - lo)
    assert _relerr(dpos, pos.grad) < 3e-2
    assert (dpos[saved["xs"].cpu()[:, 3] == 0] == 0).all()
    gb = torch.cat([w.grad.reshape(-1) for w in P.base_w]); gh = torch.cat([w.grad.reshape(-1) for w in P.head_w])
    # deltas are fp16 MMA operands (2^-11 relative each) -> percent-level agreement on the reduced gradients
    assert _relerr(grads["d_head_w"].cpu(), gh) < 2e-2
    assert _relerr(grads["d_base_w"].cpu(), gb) < 2e-2
    assert _relerr(grads["d_blend_codes"].cpu(), P.time_emb.grad) < 2e-2
    dt = grads["d_tables"].cpu()
    # per-entry gradients are sums of only a few fp16-delta contributions: no averaging -> looser max-norm bound
    assert _relerr(dt, P.tables.grad) < 8e-2
    assert ((dt != 0) == (P.tables.grad != 0)).float().mean() > 0.999        # same entries touched
    cos = torch.nn.functional.cosine_similarity(dt.reshape(1, -1), P.tables.grad.reshape(1, -1)).item()
    assert cos > 0.9995, cos


def test_deform_backward_vs_autograd():
    """SE(3) deformation field: gradients of the 6 stem layers, mlp_r/mlp_v and the warp-code embedding."""
    from nersemble_b200 import ops
    knobs = dict(TRAINED); knobs["deform_last_scale"] = 0.05        # |r| above the 1e-2 clamp: exercises d theta / d r
    P = pl.random_params(**knobs)
    NP = native_from_oracle(P, DEV)
    g = torch.Generator().manual_seed(9)
    n = 333
    lo, hi = P.aabb[0], P.aabb[1]
    pos = lo + (torch.rand((n, 3), generator=g) * 0.9 + 0.05) * (hi - lo)
    tsteps = torch.sort(torch.randint(0, 4, (n,), generator=g))[0]
    times = tsteps.float()[:, None] / 3
    w_deform = 5.5
    P.requires_grad_(True)
    off = pl.compute_offsets(P, pos, P.time_emb_deform[tsteps], w_deform)
    g_off = torch.randn((n, 3), generator=g)
    (off * g_off).sum().backward()

    kw = dict(positions=pos.to(DEV), sample_times=times.to(DEV))
    saved = ops.field_forward(NP, window_hash=None, window_deform=w_deform, use_deformation=True,
                              want=("offsets", "deform_acts"), **kw)
    torch.testing.assert_close(saved["offsets"].cpu(), off.detach(), rtol=5e-3, atol=5e-5)
    d_xs = (g_off * (hi - lo)).to(DEV)                               # offsets enter the hash input as offset / size
    gr = ops.deform_backward(NP, saved, d_xs, window_deform=w_deform, loss_scale=64.0, **kw)
    for l in range(6):
        assert _relerr(gr["d_stem_w"][l].cpu(), P.deform_w[l].grad) < 4e-2, l
        assert _relerr(gr["d_stem_b"][l].cpu(), P.deform_b[l].grad) < 4e-2, l
    assert _relerr(gr["d_r_w"].cpu(), P.r_w.grad) < 3e-2 and _relerr(gr["d_v_w"].cpu(), P.v_w.grad) < 3e-2
    assert _relerr(gr["d_r_b"].cpu(), P.r_b.grad) < 3e-2 and _relerr(gr["d_v_b"].cpu(), P.v_b.grad) < 3e-2
    assert _relerr(gr["d_warp_codes"].cpu(), P.time_emb_deform.grad) < 4e-2
