"""OccGridEstimator + NeRSembleVolumetricSampler -- mirrors of nerfacc 0.5.2 estimators/occ_grid.py [3P] and
model_components/nersemble_volumetric_sampler.py:13-135, marching through libnsb."""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
from torch import Tensor, nn

from .. import ops
from ..nerfstudio_shim import Frustums, RayBundle, RaySamples


def _enlarge_aabb(aabb: Tensor, factor: float) -> Tensor:
    center = (aabb[:3] + aabb[3:]) / 2
    extent = (aabb[3:] - aabb[:3]) / 2
    return torch.cat([center - extent * factor, center + extent * factor])


class OccGridEstimator(nn.Module):
    """Same buffers (state_dict keys `occs`, `binaries`, `resolution`, `aabbs`) and methods as nerfacc's."""

    def __init__(self, roi_aabb, resolution=128, levels: int = 1, **kwargs):
        super().__init__()
        if isinstance(resolution, int):
            resolution = [resolution] * 3
        resolution = torch.tensor(resolution, dtype=torch.int32)
        roi_aabb = torch.as_tensor(roi_aabb, dtype=torch.float32).flatten()
        aabbs = torch.stack([_enlarge_aabb(roi_aabb, 2 ** i) for i in range(levels)], 0)
        self.cells_per_lvl = int(resolution.prod().item())
        self.levels = levels
        self.register_buffer("resolution", resolution)
        self.register_buffer("aabbs", aabbs)
        self.register_buffer("occs", torch.zeros(self.levels * self.cells_per_lvl))
        self.register_buffer("binaries", torch.zeros([levels] + resolution.tolist(), dtype=torch.bool))
        gx, gy, gz = torch.meshgrid(*[torch.arange(int(r)) for r in resolution], indexing="ij")
        self.register_buffer("grid_coords", torch.stack([gx, gy, gz], -1).reshape(self.cells_per_lvl, 3), persistent=False)
        self.register_buffer("grid_indices", torch.arange(self.cells_per_lvl), persistent=False)

    @property
    def device(self):
        return self.occs.device

    @torch.no_grad()
    def sampling(self, rays_o, rays_d, sigma_fn: Optional[Callable] = None, alpha_fn=None, near_plane: float = 0.0,
                 far_plane: float = 1e10, t_min=None, t_max=None, render_step_size: float = 1e-3,
                 early_stop_eps: float = 1e-4, alpha_thre: float = 0.0, stratified: bool = False,
                 cone_angle: float = 0.0, jitter: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
        """nerfacc OccGridEstimator.sampling.  `jitter` [n_rays] in [0,1) overrides torch.rand_like (tests)."""
        near_planes = torch.full_like(rays_o[..., 0], fill_value=near_plane)
        far_planes = torch.full_like(rays_o[..., 0], fill_value=far_plane)
        if t_min is not None:
            near_planes = torch.clamp(near_planes, min=t_min)
        if t_max is not None:
            far_planes = torch.clamp(far_planes, max=t_max)
        if stratified:
            near_planes += (torch.rand_like(near_planes) if jitter is None else jitter.to(near_planes)) * render_step_size
        t_starts, t_ends, ray_indices, packed_info = ops.march_occupancy(
            rays_o, rays_d, near_planes, far_planes, self.binaries, self.aabbs, render_step_size, cone_angle)
        ray_indices = ray_indices.long()
        if (alpha_thre > 0.0 or early_stop_eps > 0.0) and (sigma_fn is not None or alpha_fn is not None):
            alpha_thre = min(alpha_thre, self.occs.mean().item())
            if t_starts.shape[0] != 0:
                sigmas = sigma_fn(t_starts, t_ends, ray_indices)
                assert sigmas.shape == t_starts.shape, "sigmas must have shape of (N,)! Got {}".format(sigmas.shape)
                masks, _ = ops.visibility_mask(packed_info, t_starts, t_ends, sigmas, early_stop_eps, alpha_thre)
                ray_indices, t_starts, t_ends = ray_indices[masks], t_starts[masks], t_ends[masks]
        return ray_indices, t_starts, t_ends

    @torch.no_grad()
    def update_every_n_steps(self, step: int, occ_eval_fn: Callable, occ_thre: float = 1e-2, ema_decay: float = 0.95,
                             warmup_steps: int = 256, n: int = 16) -> None:
        if not self.training:
            raise RuntimeError("You should only call this function only during training. "
                               "Please call _update() directly if you want to update the field during inference.")
        if step % n == 0 and self.training:
            self._update(step=step, occ_eval_fn=occ_eval_fn, occ_thre=occ_thre, ema_decay=ema_decay, warmup_steps=warmup_steps)

    @torch.no_grad()
    def _get_all_cells(self) -> List[Tensor]:
        return [self.grid_indices] * self.levels

    @torch.no_grad()
    def _sample_uniform_and_occupied_cells(self, n: int) -> List[Tensor]:
        lvl_indices = []
        for lvl in range(self.levels):
            uniform_indices = torch.randint(self.cells_per_lvl, (n,), device=self.device)
            occupied_indices = torch.nonzero(self.binaries[lvl].flatten())[:, 0]
            if n < len(occupied_indices):
                selector = torch.randint(len(occupied_indices), (n,), device=self.device)
                occupied_indices = occupied_indices[selector]
            lvl_indices.append(torch.cat([uniform_indices, occupied_indices], dim=0))
        return lvl_indices

    @torch.no_grad()
    def _update(self, step: int, occ_eval_fn: Callable, occ_thre: float = 0.01, ema_decay: float = 0.95,
                warmup_steps: int = 256) -> None:
        lvl_indices = self._get_all_cells() if step < warmup_steps else self._sample_uniform_and_occupied_cells(self.cells_per_lvl // 4)
        ids, vals = [], []
        for lvl, indices in enumerate(lvl_indices):
            grid_coords = self.grid_coords[indices]
            x = (grid_coords + torch.rand_like(grid_coords, dtype=torch.float32)) / self.resolution
            x = self.aabbs[lvl, :3] + x * (self.aabbs[lvl, 3:] - self.aabbs[lvl, :3])
            vals.append(occ_eval_fn(x).squeeze(-1))
            ids.append(lvl * self.cells_per_lvl + indices)
        # occs[cell] = max(occs[cell] * decay, occ); thre = min(mean(occs[occs >= 0]), occ_thre); binaries = occs > thre
        # -- one C-ABI call (nsb_occ_update): no boolean-index host sync, duplicates resolved deterministically
        if not self.binaries.is_contiguous():
            self.binaries = self.binaries.contiguous()
        ops.occ_update(self.occs, self.binaries, torch.cat(ids), torch.cat(vals), ema_decay, occ_thre)


def frustum_cull_grid(normals: Tensor, offsets: Tensor, scene_aabb: Tensor, resolution, min_views: int,
                      device=None, chunk: int = 1 << 18) -> Tensor:
    """View-frustum-cull mask of the occupancy grid, built on `device` (model_components/nersemble_volumetric_sampler.py:
    28-40 over frustum.py:43-53): cell (i, j, k) -- sampled at linspace(aabb_min, aabb_max, res) like upstream -- is kept
    when at least `min_views` camera frustums contain it; a frustum = 4 half spaces, inside iff n . (p - o) >= 0 for all
    four (normals are normalised like TorchHalfSpace3D).  normals / offsets: [n_cameras, 4, 3].  Returns bool [rx, ry, rz].
    The reference evaluates one camera at a time with a [B*4, 1, 3] @ [B*4, 3, 1] bmm; here all cameras of a chunk of
    points are one einsum on the GPU."""
    dev = torch.device(device) if device is not None else normals.device
    n = normals.to(dev, torch.float32)
    n = n / n.norm(dim=-1, keepdim=True)
    o = offsets.to(dev, torch.float32)
    res = [int(r) for r in (resolution.tolist() if torch.is_tensor(resolution) else ([resolution] * 3 if isinstance(resolution, int) else resolution))]
    ab = scene_aabb.to(dev, torch.float32).reshape(2, 3)
    gx, gy, gz = torch.meshgrid(*[torch.linspace(float(ab[0][k]), float(ab[1][k]), steps=res[k], device=dev) for k in range(3)], indexing="ij")
    pts = torch.stack([gx, gy, gz], dim=-1).view(-1, 3)
    d_const = (n * o).sum(-1)                                   # n . o  [C, 4]
    out = torch.empty((pts.shape[0],), dtype=torch.bool, device=dev)
    for i in range(0, pts.shape[0], chunk):
        sd = torch.einsum("bk,cfk->bcf", pts[i:i + chunk], n) - d_const[None]        # n . p - n . o
        out[i:i + chunk] = (sd >= 0).all(dim=-1).sum(dim=-1) >= min_views
    return out.view(*res)


class NeRSembleVolumetricSampler(nn.Module):
    """model_components/nersemble_volumetric_sampler.py:13-135."""

    def __init__(self, occupancy_grid: OccGridEstimator, density_fn=None, scene_aabb: Optional[Tensor] = None,
                 camera_frustums=None, view_frustum_culling: Optional[int] = None):
        super().__init__()
        assert occupancy_grid is not None
        self.density_fn = density_fn
        self.occupancy_grid = occupancy_grid
        self.camera_frustums = camera_frustums
        self.view_frustum_culling = view_frustum_culling
        if camera_frustums is not None and view_frustum_culling is not None and isinstance(camera_frustums, (tuple, dict)):
            # raw half-space arrays (normals [C,4,3], offsets [C,4,3]): the mask is built on the grid's device
            nrm, off = (camera_frustums["normals"], camera_frustums["offsets"]) if isinstance(camera_frustums, dict) else camera_frustums
            self.camera_frustum_grid = frustum_cull_grid(nrm, off, scene_aabb, self.occupancy_grid.resolution, view_frustum_culling,
                                                         device=self.occupancy_grid.occs.device)
        elif camera_frustums is not None and view_frustum_culling is not None:
            res = self.occupancy_grid.resolution
            gx, gy, gz = torch.meshgrid(torch.linspace(scene_aabb[0][0], scene_aabb[1][0], steps=int(res[0])),
                                        torch.linspace(scene_aabb[0][1], scene_aabb[1][1], steps=int(res[1])),
                                        torch.linspace(scene_aabb[0][2], scene_aabb[1][2], steps=int(res[2])), indexing="ij")
            pts = torch.stack([gx, gy, gz], dim=-1).view(-1, 3)
            masks = [f.contains_points(pts.to(f._half_space_collection.offsets.device)) for f in camera_frustums]
            vis = torch.stack(masks).sum(dim=0) >= view_frustum_culling
            self.camera_frustum_grid = vis.view(*[int(r) for r in res])
        else:
            self.camera_frustum_grid = None

    def get_sigma_fn(self, origins, directions, times=None):
        """nerfstudio VolumetricSampler.get_sigma_fn: None unless density_fn is set AND training."""
        if self.density_fn is None or not self.training:
            return None
        density_fn = self.density_fn

        def sigma_fn(t_starts, t_ends, ray_indices):
            t_origins = origins[ray_indices]
            t_dirs = directions[ray_indices]
            positions = t_origins + t_dirs * (t_starts + t_ends)[:, None] / 2.0
            if times is None:
                return density_fn(positions).squeeze(-1)
            return density_fn(positions, times[ray_indices]).squeeze(-1)

        return sigma_fn

    def eval_planes(self, ray_bundle: RayBundle, near_plane: float = 0.0, far_plane: Optional[float] = None) -> Tuple[Tensor, Tensor]:
        """Per-ray near / far planes exactly as forward() + OccGridEstimator.sampling() build them in eval mode (no
        stratified jitter), and the frustum-cull AND of the grid: the inputs of the fused render (ops.render_rays)."""
        rays_o = ray_bundle.origins.reshape(-1, 3)
        if far_plane is None:
            far_plane = 1e10
        near_planes = torch.full_like(rays_o[..., 0], fill_value=near_plane)
        far_planes = torch.full_like(rays_o[..., 0], fill_value=far_plane)
        if ray_bundle.nears is not None and ray_bundle.fars is not None:
            near_planes = torch.clamp(near_planes, min=ray_bundle.nears.contiguous().reshape(-1))
            far_planes = torch.clamp(far_planes, max=ray_bundle.fars.contiguous().reshape(-1))
        if self.camera_frustum_grid is not None:
            self.occupancy_grid.binaries[0] = self.occupancy_grid.binaries[0] & self.camera_frustum_grid.to(self.occupancy_grid.binaries.device)
        return near_planes, far_planes

    @torch.no_grad()
    def sample_packed(self, ray_bundle: RayBundle, render_step_size: float, near_plane: float = 0.0,
                      far_plane: Optional[float] = None, alpha_thre: float = 0.01, cone_angle: float = 0.0,
                      early_stop_eps: float = 1e-4, jitter: Optional[Tensor] = None, sigma_packed_fn=None):
        """forward() for the model's own use, with ONE host synchronisation instead of four: the nerfacc march as a
        cooperative launch whose count stays on the device (ops.march_occupancy_packed), the density of the candidates
        from `sigma_packed_fn(candidates) -> (sigma [capacity], payload | None)` evaluated with a device-side count,
        then visibility filter + packing in one launch (ops.visibility_compact, alpha_thre capped by occs.mean() on the
        device).  Same kernels' arithmetic as forward(): the kept samples are identical.  Returns a dict with the packed
        t_starts / t_ends / ray_indices (int32) [n], packed_info [R,2], optional payload rows, and n (a Python int: the
        one sync, needed because the plugin contract hands out exact-size per-sample tensors)."""
        near_planes, far_planes = self.eval_planes(ray_bundle, near_plane, far_plane)
        rays_o = ray_bundle.origins.reshape(-1, 3).contiguous()
        rays_d = ray_bundle.directions.reshape(-1, 3).contiguous()
        og = self.occupancy_grid
        if self.training:                                          # stratified = self.training (forward())
            u = torch.rand_like(near_planes) if jitter is None else jitter.to(near_planes)
            near_planes = near_planes + u * render_step_size
        cand = ops.march_occupancy_packed(rays_o, rays_d, near_planes, far_planes, og.binaries, og.aabbs, render_step_size, cone_angle)
        kept = cand
        if (alpha_thre > 0.0 or early_stop_eps > 0.0) and self.density_fn is not None and self.training and sigma_packed_fn is not None:
            sigma, payload = sigma_packed_fn(cand)
            kept = ops.visibility_compact(cand, sigma, early_stop_eps, alpha_thre, alpha_thre_cap=og.occs.mean(), payload=payload)
        head = kept["header"][:3].tolist()                        # the one host synchronisation: (.., status, n_total)
        n, status = int(head[2]), int(head[1]) >> 32
        if status != 0 or n > kept["capacity"]:
            raise RuntimeError(f"sampler: the march produced more samples than its workspace holds ({n} > {kept['capacity']})")
        out = {k: kept[k][:n] for k in ("t_starts", "t_ends", "ray_indices")}
        for k in ("feat", "xs", "corner_vals"):
            if k in kept:
                out[k] = kept[k][:n]
        out["packed_info"], out["n"] = kept["packed_info"], n
        if n == 0:      # zero-sample guard (nersemble_volumetric_sampler.py:110-114): one fake sample on ray 0
            dev = rays_o.device
            out["t_starts"] = torch.ones((1,), dtype=torch.float32, device=dev)
            out["t_ends"] = torch.ones((1,), dtype=torch.float32, device=dev)
            out["ray_indices"] = torch.zeros((1,), dtype=torch.int32, device=dev)
            pi = torch.zeros_like(kept["packed_info"]); pi[0, 1] = 1
            out["packed_info"], out["n"] = pi, 1
            for k in ("feat", "xs", "corner_vals"):
                out.pop(k, None)
        return out

    def forward(self, ray_bundle: RayBundle, render_step_size: float, near_plane: float = 0.0,
                far_plane: Optional[float] = None, alpha_thre: float = 0.01, cone_angle: float = 0.0,
                early_stop_eps: float = 1e-4, jitter: Optional[Tensor] = None) -> Tuple[RaySamples, Tensor]:
        rays_o = ray_bundle.origins.contiguous()
        rays_d = ray_bundle.directions.contiguous()
        times = ray_bundle.times
        if ray_bundle.nears is not None and ray_bundle.fars is not None:
            t_min = ray_bundle.nears.contiguous().reshape(-1)
            t_max = ray_bundle.fars.contiguous().reshape(-1)
        else:
            t_min = t_max = None
        if far_plane is None:
            far_plane = 1e10
        camera_indices = ray_bundle.camera_indices.contiguous() if ray_bundle.camera_indices is not None else None
        if self.camera_frustum_grid is not None:
            self.occupancy_grid.binaries[0] = self.occupancy_grid.binaries[0] & self.camera_frustum_grid.to(self.occupancy_grid.binaries.device)
        ray_indices, starts, ends = self.occupancy_grid.sampling(
            rays_o=rays_o, rays_d=rays_d, t_min=t_min, t_max=t_max, sigma_fn=self.get_sigma_fn(rays_o, rays_d, times),
            render_step_size=render_step_size, near_plane=near_plane, far_plane=far_plane, stratified=self.training,
            cone_angle=cone_angle, alpha_thre=alpha_thre, early_stop_eps=early_stop_eps, jitter=jitter)
        if starts.shape[0] == 0:
            ray_indices = torch.zeros((1,), dtype=torch.long, device=rays_o.device)
            starts = torch.ones((1,), dtype=starts.dtype, device=rays_o.device)
            ends = torch.ones((1,), dtype=ends.dtype, device=rays_o.device)
        origins = rays_o[ray_indices]
        dirs = rays_d[ray_indices]
        if camera_indices is not None:
            camera_indices = camera_indices[ray_indices]
        ray_samples = RaySamples(frustums=Frustums(origins=origins, directions=dirs, starts=starts[..., None],
                                                   ends=ends[..., None], pixel_area=torch.zeros_like(origins[:, :1])),
                                 camera_indices=camera_indices)
        if ray_bundle.times is not None:
            ray_samples.times = ray_bundle.times[ray_indices]
        return ray_samples, ray_indices
