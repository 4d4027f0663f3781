"""On-GPU batch assembly (nsb_ray_batch, plugin/datamanager.py) against a torch restatement of nerfstudio 0.3.1's
Cameras._generate_rays_from_coords (perspective) [3P-mem] and of NeRSemblePixelSampler's gathers
(data/nersemble_pixel_sampler.py:47-62)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cameras(n_cam, g):
    ang = torch.rand(n_cam, generator=g) * 2 * math.pi
    eye = torch.stack([9 * torch.sin(ang), 0.5 * torch.rand(n_cam, generator=g), 9 * torch.cos(ang)], -1)
    fwd = -eye / eye.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 1.0, 0.0]).expand_as(fwd)
    right = torch.linalg.cross(fwd, up); right = right / right.norm(dim=-1, keepdim=True)
    up2 = torch.linalg.cross(right, fwd)
    c2w = torch.cat([torch.stack([right, up2, -fwd], -1), eye[:, :, None]], -1)          # camera looks along -z
    intr = torch.stack([torch.full((n_cam,), 800.0) + torch.rand(n_cam, generator=g) * 50, torch.full((n_cam,), 790.0),
                        torch.full((n_cam,), 40.3), torch.full((n_cam,), 29.6)], -1)
    return c2w.float(), intr.float()


def _reference_rays(idx, image_camera, image_times, intr, c2w):
    """nerfstudio Cameras._generate_rays_from_coords restated (float32 torch, CPU)."""
    cam = image_camera[idx[:, 0]]
    y, x = idx[:, 1].float() + 0.5, idx[:, 2].float() + 0.5
    fx, fy, cx, cy = (intr[cam, k] for k in range(4))
    coord = torch.stack([(x - cx) / fx, -(y - cy) / fy], -1)
    cx_off = torch.stack([(x - cx + 1) / fx, -(y - cy) / fy], -1)
    cy_off = torch.stack([(x - cx) / fx, -(y - cy + 1) / fy], -1)
    stack = torch.stack([coord, cx_off, cy_off], 0)
    stack = torch.cat([stack, -torch.ones_like(stack[..., :1])], -1)                   # [3, R, 3]
    rot = c2w[cam][:, :3, :3]
    stack = torch.sum(stack[..., None, :] * rot, dim=-1)
    norm = stack.norm(dim=-1, keepdim=True)
    stack = stack / torch.clamp(norm, min=torch.finfo(torch.float32).eps)
    dx = torch.sqrt(((stack[0] - stack[1]) ** 2).sum(-1)); dy = torch.sqrt(((stack[0] - stack[2]) ** 2).sum(-1))
    return dict(origins=c2w[cam][:, :3, 3], directions=stack[0], pixel_area=(dx * dy)[:, None], directions_norm=norm[0],
                camera_indices=cam[:, None], times=image_times[idx[:, 0]][:, None])


def test_ray_batch_matches_the_nerfstudio_camera_model_and_the_pixel_gathers():
    from nersemble_b200.plugin import DeviceImageCache, DeviceRaySampler, ray_batch
    g = torch.Generator().manual_seed(0)
    n_cam, T, H, W = 5, 3, 60, 80
    c2w, intr = _cameras(n_cam, g)
    n_img = n_cam * T
    image_camera = torch.arange(n_img) % n_cam
    image_times = (torch.arange(n_img) // n_cam).float() / (T - 1)
    images = torch.randint(0, 256, (n_img, H, W, 3), generator=g, dtype=torch.uint8)
    alpha = torch.randint(0, 256, (n_img, H, W), generator=g, dtype=torch.uint8)
    depth = torch.rand((n_img, H, W), generator=g) * 5 + 5
    cache = DeviceImageCache(images, image_camera, image_times, intr, c2w, alpha_maps=alpha, depth_maps=depth, device=DEV)
    sampler = DeviceRaySampler(cache, num_rays_per_batch=4096, seed=3)
    idx = sampler.sample_indices()
    assert idx.shape == (4096, 3) and int(idx[:, 0].max()) < n_img and int(idx[:, 1].max()) < H and int(idx[:, 2].max()) < W
    idx[0] = torch.tensor([0, 0, 0]); idx[1] = torch.tensor([n_img - 1, H - 1, W - 1])
    rb, batch = ray_batch(cache, idx)
    want = _reference_rays(idx.cpu(), image_camera, image_times, intr, c2w)
    assert torch.equal(rb.origins.cpu(), want["origins"]) and torch.equal(rb.camera_indices.cpu(), want["camera_indices"])
    assert torch.equal(rb.times.cpu(), want["times"])
    torch.testing.assert_close(rb.directions.cpu(), want["directions"], rtol=0, atol=2e-7)
    torch.testing.assert_close(rb.metadata["directions_norm"].cpu(), want["directions_norm"], rtol=2e-7, atol=0)
    torch.testing.assert_close(rb.pixel_area.cpu(), want["pixel_area"], rtol=2e-3, atol=0)     # differences of nearly equal unit vectors
    c, y, x = idx.cpu().unbind(-1)
    assert torch.equal(batch["image"].cpu(), images[c, y, x].float() / 255.0)
    assert torch.equal(batch["alpha_map"].cpu()[:, 0], alpha[c, y, x].float())
    assert torch.equal(batch["depth_maps"].cpu(), depth[c, y, x])
    assert (rb.directions.norm(dim=-1) - 1).abs().max() < 1e-6
    # a full image as a camera ray bundle + the sampler's next_train
    full, fb = sampler.camera_ray_bundle(7)
    assert full.origins.shape == (H, W, 3) and fb["image"].shape == (H, W, 3)
    assert torch.equal(fb["image"].cpu(), images[7].float() / 255.0)
    rb2, b2 = sampler.next_train(0)
    assert rb2.origins.shape == (4096, 3) and rb2.origins.is_cuda and b2["image"].is_cuda
