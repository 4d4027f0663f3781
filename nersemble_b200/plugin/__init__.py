"""Host-side mirror of the reference's nerfstudio plugin surface for the render hot path.

Same class names, constructor/forward signatures, config dataclasses, output keys and
state_dict keys as /root/reference/src/nersemble/nerfstudio/** (cited per class), with all
arithmetic behind libnsb (CUDA, sm_100a).  Forward / inference path (eval render, density_fn,
sampler incl. the no-grad training pre-pass, occupancy update).  The backward kernels are
round-2 work: calling these modules with autograd enabled raises instead of silently
returning graph-less tensors.
"""
from .components import (HashEnsemble, HashEnsembleConfig, SE3DeformationField, SE3DeformationFieldConfig,
                         TCNNHashEncodingConfig, GenericScheduler)
from .field import NeRSembleNeRFactoField
from .sampler import NeRSembleVolumetricSampler, OccGridEstimator
from .model import NeRSembleNGPModel, NeRSembleNGPModelConfig, BaseModelConfig

__all__ = ["HashEnsemble", "HashEnsembleConfig", "SE3DeformationField", "SE3DeformationFieldConfig",
           "TCNNHashEncodingConfig", "GenericScheduler", "NeRSembleNeRFactoField", "NeRSembleVolumetricSampler",
           "OccGridEstimator", "NeRSembleNGPModel", "NeRSembleNGPModelConfig", "BaseModelConfig"]
