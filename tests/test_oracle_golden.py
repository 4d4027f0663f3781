"""Oracle restatement (oracle/pipeline.py) vs the goldens produced by the REAL reference glue
(oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import load_golden, oracle_params
from oracle import pipeline as pl
from oracle.tp.tcnn_cpu import Precision

RENDER_CASES = ["config1_init_reference", "config1_trained_none", "fixed_trained_reference",
                "fixed_trained_kernel", "fixed_trained_none", "fixed_trained_autocast",
                "occ_eval_soft", "occ_eval_whash1", "occ_train_prepass", "losses_train_render"]


def _blob(seed):
    from oracle.gen_golden import blob_grid
    return blob_grid(seed)


def _samples(P, g, meta):
    o, d, times = g["origins"], g["directions"], g["times"]
    if meta["sampler"] == "fixed":
        return pl.fixed_samples(o, d, P.aabb, meta["n_fixed"], 0.011, near=0.2)
    occ = _blob(meta["grid_seed"])
    jitter = None
    if meta["training"]:
        torch.manual_seed(meta["jitter_seed"])
        jitter = torch.rand(meta["R"])
    return pl.sample_occupancy(P, o, d, times, occ[None], float((occ.flatten().float() * 0.05).mean()),
                               render_step_size=0.011, near_plane=0.2, far_plane=1e3, alpha_thre=1e-2,
                               early_stop_eps=0.0, training=meta["training"], jitter=jitter,
                               window_hash=meta["w_hash"], window_deform=meta["w_deform"])


@pytest.mark.parametrize("name", RENDER_CASES)
def test_render_matches_reference_glue(name):
    g, meta = load_golden(name)
    Precision.mode = meta["mode"]; Precision.autocast = meta.get("autocast", False)
    P = oracle_params(meta["knobs"])
    with torch.no_grad():
        ts, te, ri = _samples(P, g, meta)
        # samples: bit-exact
        assert ts.shape == g["t_starts"].shape, (ts.shape, g["t_starts"].shape)
        assert torch.equal(ri, g["ray_indices"])
        assert torch.equal(ts, g["t_starts"]) and torch.equal(te, g["t_ends"])
        out = pl.render(P, g["origins"], g["directions"], g["times"], ts, te, ri,
                        window_hash=meta["w_hash"], window_deform=meta["w_deform"], training=meta["training"])
    assert torch.equal(out["num_samples_per_ray"], g["num_samples_per_ray"])
    # same arithmetic, different summation order inside torch ops -> tiny fp32 noise only
    tol = dict(rtol=2e-4, atol=2e-5) if meta["mode"] != "none" else dict(rtol=1e-4, atol=2e-6)
    for k in ("offsets", "weights", "rgb", "accumulation", "depth", "deformation"):
        torch.testing.assert_close(out[k], g[k], **tol, msg=lambda m, k=k: f"{name}:{k}: {m}")
    Precision.mode = "reference"; Precision.autocast = False


@pytest.mark.parametrize("name", ["density_fn_none", "density_fn_kernel"])
def test_density_fn(name):
    g, meta = load_golden(name)
    Precision.mode = meta["mode"]; Precision.autocast = False
    P = oracle_params(meta["knobs"])
    with torch.no_grad():
        sig = pl.field_density_fn(P, g["positions"], g["times"], meta["w_hash"], meta["w_deform"])
    assert (g["density"] == 0).any(), "golden should contain out-of-box points (selector)"
    torch.testing.assert_close(sig, g["density"], rtol=2e-4, atol=1e-6)
    Precision.mode = "reference"


def test_losses():
    g, meta = load_golden("losses_train")
    r, rmeta = load_golden(meta["render_case"])
    out = {k: r[k] for k in ("rgb", "accumulation", "depth", "weights")}
    batch = {k[len("batch_"):]: v for k, v in g.items() if k.startswith("batch_")}
    ld = pl.loss_dict(out, r["t_starts"], r["t_ends"], r["ray_indices"], batch, eps_depth=meta["eps_depth"])
    want = {k[len("loss_"):]: v for k, v in g.items() if k.startswith("loss_")}
    assert set(ld) == set(want)
    for k in want:
        torch.testing.assert_close(ld[k].float(), want[k].float(), rtol=1e-4, atol=1e-9, msg=lambda m, k=k: f"{k}: {m}")


def test_precision_modes_within_rgb_tolerance():
    """The fp16 roundings of the reference ("reference") and of the B200 kernels ("kernel") both stay
    within the north-star tolerance (RGB L2 <= 1e-3) of the fp32 evaluation ("none")."""
    ref, _ = load_golden("fixed_trained_reference")
    ker, _ = load_golden("fixed_trained_kernel")
    non, _ = load_golden("fixed_trained_none")
    for a in (ref, ker):
        l2 = (a["rgb"] - non["rgb"]).norm(dim=-1)
        assert l2.max() < 1e-3, l2.max()
