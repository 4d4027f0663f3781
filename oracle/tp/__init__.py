"""Third-party restatements + a stub installer.

``install_stubs()`` registers the restatements in ``sys.modules`` under the third-party
package names (tinycudann, nerfacc, nerfstudio.*, torch_efficient_distloss, and inert
placeholders for dreifus / torchmetrics), so that the REAL reference glue in
/root/reference/src can be imported and executed on CPU (oracle/gen_golden.py).
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import sys
import types

import torch

from . import nerfacc_cpu, nerfstudio_cpu, tcnn_cpu


def flatten_eff_distloss(w, m, interval, ray_id):
    """torch_efficient_distloss.flatten_eff_distloss [3P-mem] (sunset1995/torch_efficient_distloss,
    torch_efficient_distloss/eff_distloss.py: FlattenEffDistLoss.forward; unpinned in environment.yml:26):
    loss = (sum_i 1/3*interval_i*w_i^2 + sum_i 2*w_i*(m_i*W_<i - WM_<i)) / n_rays,
    n_rays = ray_id.max()+1; prefixes are per-ray exclusive sums."""
    if w.numel() == 0:
        return w.sum()
    n_rays = int(ray_id.max()) + 1
    info = nerfacc_cpu.pack_info(ray_id, n_rays)
    w_prefix = nerfacc_cpu.exclusive_sum(w, info)
    wm_prefix = nerfacc_cpu.exclusive_sum(w * m, info)
    loss_uni = (1.0 / 3.0) * (interval * w * w).sum()
    loss_bi = 2.0 * (w * (m * w_prefix - wm_prefix)).sum()
    return (loss_uni + loss_bi) / n_rays


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        if parent not in sys.modules:
            _mod(parent)
        setattr(sys.modules[parent], child, m)
    return m


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise NotImplementedError("inert stub (off the hot path)")


def install_stubs() -> None:
    ns = nerfstudio_cpu
    _mod("tinycudann", Encoding=tcnn_cpu.Encoding, Network=tcnn_cpu.Network,
         NetworkWithInputEncoding=tcnn_cpu.NetworkWithInputEncoding)
    _mod("nerfacc", OccGridEstimator=nerfacc_cpu.OccGridEstimator, pack_info=nerfacc_cpu.pack_info,
         render_weight_from_density=nerfacc_cpu.render_weight_from_density,
         accumulate_along_rays=nerfacc_cpu.accumulate_along_rays)
    _mod("torch_efficient_distloss", flatten_eff_distloss=flatten_eff_distloss)
    _mod("nerfstudio")
    _mod("nerfstudio.cameras")
    _mod("nerfstudio.cameras.rays", RayBundle=ns.RayBundle, RaySamples=ns.RaySamples, Frustums=ns.Frustums)
    _mod("nerfstudio.data")
    _mod("nerfstudio.data.scene_box", SceneBox=ns.SceneBox)
    _mod("nerfstudio.field_components", MLP=ns.MLP)
    _mod("nerfstudio.field_components.encodings", NeRFEncoding=ns.NeRFEncoding)
    _mod("nerfstudio.field_components.activations", trunc_exp=ns.trunc_exp)
    _mod("nerfstudio.field_components.embedding", Embedding=_Inert)
    _mod("nerfstudio.field_components.field_heads", FieldHeadNames=ns.FieldHeadNames,
         PredNormalsFieldHead=_Inert, SemanticFieldHead=_Inert, TransientDensityFieldHead=_Inert,
         TransientRGBFieldHead=_Inert, UncertaintyFieldHead=_Inert)
    _mod("nerfstudio.field_components.spatial_distortions", SpatialDistortion=_Inert, SceneContraction=_Inert)
    _mod("nerfstudio.fields")
    _mod("nerfstudio.fields.base_field", shift_directions_for_tcnn=ns.shift_directions_for_tcnn, Field=ns.Field)
    _mod("nerfstudio.fields.nerfacto_field", TCNNNerfactoField=ns.TCNNNerfactoField)
    _mod("nerfstudio.utils")
    _mod("nerfstudio.utils.math", expected_sin=ns.expected_sin)
    _mod("nerfstudio.utils.writer", put_scalar=lambda *a, **k: None)
    _mod("nerfstudio.utils.colormaps", ColormapOptions=_Inert, apply_colormap=_Inert(),
         apply_depth_colormap=_Inert())
    sys.modules["nerfstudio.utils"].writer = sys.modules["nerfstudio.utils.writer"]
    sys.modules["nerfstudio.utils"].colormaps = sys.modules["nerfstudio.utils.colormaps"]
    _mod("nerfstudio.engine")
    _mod("nerfstudio.engine.callbacks", TrainingCallback=ns.TrainingCallback,
         TrainingCallbackAttributes=ns.TrainingCallbackAttributes,
         TrainingCallbackLocation=ns.TrainingCallbackLocation)
    _mod("nerfstudio.model_components")
    _mod("nerfstudio.model_components.losses", MSELoss=ns.MSELoss)
    _mod("nerfstudio.model_components.renderers", RGBRenderer=ns.RGBRenderer, DepthRenderer=ns.DepthRenderer,
         AccumulationRenderer=ns.AccumulationRenderer)
    _mod("nerfstudio.model_components.ray_samplers", VolumetricSampler=ns.VolumetricSampler, DensityFn=ns.DensityFn)
    _mod("nerfstudio.models")
    _mod("nerfstudio.models.base_model", Model=ns.Model, ModelConfig=ns.ModelConfig)
    _mod("nerfstudio.models.instant_ngp", NGPModel=ns.NGPModel, InstantNGPModelConfig=ns.InstantNGPModelConfig)
    _mod("dreifus")
    _mod("dreifus.util")
    _mod("dreifus.util.colormap", apply_scene_flow_colormap=_Inert())
    _mod("torchmetrics", PeakSignalNoiseRatio=_Inert)
    _mod("torchmetrics.functional", structural_similarity_index_measure=_Inert())
    _mod("torchmetrics.image")
    _mod("torchmetrics.image.lpip", LearnedPerceptualImagePatchSimilarity=_Inert)
