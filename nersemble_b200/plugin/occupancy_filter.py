"""Largest-connected-component filter of the occupancy grid (evaluation-time floater removal).

Mirror of the reference's CPU utility `util/connected_components.py:102-139` (`filter_occupancy_grid`) and of
`extract_top_k_connected_component` (`:29-99`) which does the work.  The reference labels components with `cc3d`
(not in this image); here `scipy.ndimage.label` with the same 6-connectivity does it.  Semantics, step by step:

  1. occs [res^3] -> sigmoid -> ((s - 0.5) * 2 * 255) as uint8                (:56-58)
  2. gaussian blur of the uint8 grid with sigma_thinning (breaks thin bridges; the blur stays in uint8)   (:61)
  3. binarise at 255 * threshold                                               (:65-66)
  4. label 6-connected components, keep the K largest, ordered by size         (:80-84: cc3d.largest_k labels the
     components 1..K by INCREASING voxel count, so label K is the largest)
  5. the K-th (largest) component is dilated: gaussian blur of (mask * 100) with sigma_erosion, > 0   (:90-95)
  6. binaries[0] &= mask of component 1 of the returned list                   (:127-139; with K = 1 that is the
     largest component)

This is host code on a 128^3 grid that runs once before an evaluation; it stays on the CPU like the reference's.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch


def extract_top_k_connected_component(density_grid: np.ndarray, threshold: float = 0.6, sigma_thinning: float = 1,
                                      sigma_erosion: float = 2, K: int = 1) -> List[np.ndarray]:
    from scipy import ndimage
    x = np.asarray(density_grid)                # the sigmoid runs in the grid's own dtype (float32 occs) like :11-12, :53
    s = 1 / (1 + np.exp(-x))
    g = ((s - 0.5) * 2 * 255).astype(np.uint8)
    g = ndimage.gaussian_filter(g, sigma=sigma_thinning)
    binary = g >= 255 * threshold
    labels, n = ndimage.label(binary, structure=ndimage.generate_binary_structure(3, 1))    # faces only = 6-connectivity
    sizes = np.bincount(labels.ravel(), minlength=n + 1)[1:]
    order = np.argsort(sizes, kind="stable")[::-1][:K] + 1       # labels of the K largest, largest first
    order = order[::-1]                                          # cc3d.largest_k numbering: 1 = smallest of them, K = largest
    out = []
    for k, lab in enumerate(order, start=1):
        cc = labels == lab
        if k == K:
            cc = ndimage.gaussian_filter(cc * 100, sigma=sigma_erosion) > 0
        out.append(cc.astype(np.int64))
    while len(out) < K:      # fewer than K components exist
        out.insert(0, np.zeros(labels.shape, dtype=np.int64))
    return out


@torch.no_grad()
def filter_occupancy_grid(occupancy_grid, threshold: float = 0.6, sigma_thinning: float = 1, sigma_erosion: float = 5) -> None:
    """In place: `occupancy_grid.binaries[0] &= largest connected component of the thresholded occs`."""
    res = occupancy_grid.resolution
    res = [int(v) for v in (res.tolist() if torch.is_tensor(res) else (res if hasattr(res, "__iter__") else [res] * 3))]
    dens = occupancy_grid.occs.detach().reshape(-1)[: res[0] * res[1] * res[2]].reshape(*res).cpu().numpy()
    keep = extract_top_k_connected_component(dens, threshold=threshold, sigma_thinning=sigma_thinning, sigma_erosion=sigma_erosion)[0] > 0
    mask = torch.as_tensor(keep, device=occupancy_grid.binaries.device, dtype=occupancy_grid.binaries.dtype)
    occupancy_grid.binaries[0] = occupancy_grid.binaries[0] & mask
