"""On-GPU training-batch assembly: the image cache lives in HBM, pixel sampling and ray generation are
one CUDA launch, and next_train() returns a device-resident (RayBundle, batch) without the reference's `.cpu()` round
trips.

Mirrors, for the fields the NeRSemble model consumes:
  data/nersemble_pixel_sampler.py:23-69   NeRSemblePixelSampler.collate_image_dataset_batch (uniform (image, y, x)
                                          samples; `image` / `alpha_map` / `depth_maps` gathered at those pixels;
                                          per-image metadata -- the timestep -- gathered at the image index)
  datamanager/nersemble_datamanager.py:68-81   _add_metadata_to_ray_bundle
  nerfstudio 0.3.1 [3P-mem]: PixelSampler.sample_method (floor(rand * [N, H, W])), RayGenerator.forward,
  Cameras._generate_rays_from_coords (perspective cameras, no distortion).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from .. import _lib, ops
from ..nerfstudio_shim import RayBundle


class DeviceImageCache:
    """Images of the training window resident in device memory: rgb uint8 [N,H,W,3], alpha uint8 [N,H,W] (optional),
    depth float32 [N,H,W] (optional; 0 = no depth), plus per-image camera index and time, per-camera intrinsics
    (fx, fy, cx, cy) and camera_to_world [3,4].  180 GB of HBM hold ~30 000 frames of 1100 x 1604."""

    def __init__(self, images: Tensor, image_camera: Tensor, image_times: Optional[Tensor], intrinsics: Tensor,
                 camera_to_world: Tensor, alpha_maps: Optional[Tensor] = None, depth_maps: Optional[Tensor] = None,
                 device="cuda"):
        dev = torch.device(device)
        assert images.dtype == torch.uint8 and images.dim() == 4 and images.shape[-1] == 3
        self.images = images.to(dev).contiguous()
        self.n_images, self.height, self.width = (int(v) for v in images.shape[:3])
        self.image_camera = image_camera.to(dev, torch.int64).contiguous()
        self.image_times = None if image_times is None else image_times.to(dev, torch.float32).contiguous()
        self.intrinsics = intrinsics.to(dev, torch.float32).reshape(-1, 4).contiguous()
        self.camera_to_world = camera_to_world.to(dev, torch.float32).reshape(-1, 3, 4).contiguous()
        self.alpha_maps = None if alpha_maps is None else alpha_maps.to(dev).reshape(self.n_images, self.height, self.width).contiguous()
        self.depth_maps = None if depth_maps is None else depth_maps.to(dev, torch.float32).reshape(self.n_images, self.height, self.width).contiguous()
        assert self.alpha_maps is None or self.alpha_maps.dtype == torch.uint8


def ray_batch(cache: DeviceImageCache, indices: Tensor) -> Tuple[RayBundle, Dict[str, Tensor]]:
    """(RayBundle, batch) for pixel indices [R,3] = (image, y, x) on the cache's device: nsb_ray_batch, one launch."""
    lib = _lib.load()
    dev = cache.images.device
    idx = indices.to(dev, torch.int64).contiguous()
    R = int(idx.shape[0])
    f32 = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    out = dict(origins=f32(R, 3), directions=f32(R, 3), pixel_area=f32(R, 1), directions_norm=f32(R, 1), times=f32(R, 1),
               camera_indices=torch.empty((R, 1), dtype=torch.int64, device=dev), image=f32(R, 3))
    a = _lib.RayBatchArgs()
    a.n_rays, a.indices, a.height, a.width = R, idx.data_ptr(), cache.height, cache.width
    a.image_camera = cache.image_camera.data_ptr()
    a.image_times = None if cache.image_times is None else cache.image_times.data_ptr()
    a.intrinsics, a.camera_to_world, a.images = cache.intrinsics.data_ptr(), cache.camera_to_world.data_ptr(), cache.images.data_ptr()
    for k in ("origins", "directions", "pixel_area", "directions_norm", "times", "camera_indices"):
        setattr(a, k, out[k].data_ptr())
    a.out_image = out["image"].data_ptr()
    batch = {"image": out["image"], "indices": idx}
    if cache.alpha_maps is not None:
        batch["alpha_map"] = f32(R, 1); a.alpha_maps, a.out_alpha = cache.alpha_maps.data_ptr(), batch["alpha_map"].data_ptr()
    if cache.depth_maps is not None:
        batch["depth_maps"] = f32(R); a.depth_maps, a.out_depth = cache.depth_maps.data_ptr(), batch["depth_maps"].data_ptr()
    if R > 0:
        _lib.check(lib.nsb_ray_batch(C.byref(a), ops._stream()), "nsb_ray_batch")
    rb = RayBundle(origins=out["origins"], directions=out["directions"], pixel_area=out["pixel_area"],
                   camera_indices=out["camera_indices"], times=out["times"] if cache.image_times is not None else None,
                   metadata={"directions_norm": out["directions_norm"]})
    return rb, batch


class DeviceRaySampler:
    """next_train(step) -> (RayBundle, batch), everything on the device: uniform pixel sampling over the cached images
    (PixelSampler.sample_method: floor(rand(R, 3) * [N, H, W])) followed by ray_batch()."""

    def __init__(self, cache: DeviceImageCache, num_rays_per_batch: int = 4096, seed: Optional[int] = None):
        self.cache, self.num_rays_per_batch = cache, num_rays_per_batch
        self.generator = torch.Generator(device=cache.images.device)
        if seed is not None:
            self.generator.manual_seed(seed)

    def sample_indices(self, n: Optional[int] = None) -> Tensor:
        c = self.cache
        n = self.num_rays_per_batch if n is None else n
        scale = torch.tensor([c.n_images, c.height, c.width], dtype=torch.float32, device=c.images.device)
        return torch.floor(torch.rand((n, 3), generator=self.generator, device=c.images.device) * scale).long()

    def next_train(self, step: int = 0) -> Tuple[RayBundle, Dict[str, Tensor]]:
        return ray_batch(self.cache, self.sample_indices())

    def camera_ray_bundle(self, image_index: int) -> Tuple[RayBundle, Dict[str, Tensor]]:
        """All pixels of one cached image as an [H, W] bundle (next_eval_image / util/render.py:35-70 input)."""
        c = self.cache
        ys, xs = torch.meshgrid(torch.arange(c.height, device=c.images.device), torch.arange(c.width, device=c.images.device), indexing="ij")
        idx = torch.stack([torch.full_like(ys, image_index), ys, xs], -1).reshape(-1, 3)
        rb, batch = ray_batch(c, idx)
        v = lambda t: None if t is None else t.view(c.height, c.width, -1)
        rb = RayBundle(origins=v(rb.origins), directions=v(rb.directions), pixel_area=v(rb.pixel_area),
                       camera_indices=v(rb.camera_indices), times=v(rb.times), metadata={k: v(t) for k, t in rb.metadata.items()})
        return rb, {k: (t.view(c.height, c.width, -1) if k != "indices" else t) for k, t in batch.items()}
