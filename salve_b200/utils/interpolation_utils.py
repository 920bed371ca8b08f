"""Sparse -> dense interpolation and hallucination mask (mirrors reference
salve/utils/interpolation_utils.py:14-122).  scipy.interpolate.griddata(method="linear") is
replaced by the GPU Delaunay (parallel Lawson flips) + exact barycentric rasteriser."""

from __future__ import annotations

import numpy as np

from .. import _ctx
from ..renderer import IMG_COLLINEAR, IMG_DEGENERATE

DEFAULT_KERNEL_SZ = 11
MIN_REQUIRED_POINTS_SIMPLEX = 4

try:  # callers that catch the reference's failure mode keep working
    from scipy.spatial import QhullError as _QhullBase
except Exception:  # pragma: no cover
    _QhullBase = RuntimeError


class QhullError(_QhullBase):  # type: ignore[misc, valid-type]
    """Raised where the reference's Qhull call raises: all sites on one oblique line."""


def is_collinear(points: np.ndarray) -> bool:
    """Cheap check of the reference (:57-71): all points share their first, or their second, coordinate."""
    return bool(np.allclose(points[:, 0], points[0, 0]) or np.allclose(points[:, 1], points[0, 1]))


def interp_dense_grid_from_sparse(bev_img: np.ndarray, points: np.ndarray, rgb_values: np.ndarray, grid_h: int, grid_w: int,
                                  is_semantics: bool) -> np.ndarray:
    """Fill `bev_img` (grid_h, grid_w, 3) by linear interpolation of `rgb_values` at integer `points` (x, y);
    pixels outside the convex hull become 0.  Mutates and returns `bev_img`; returns it untouched for
    fewer than 4 points or axis-aligned collinear input (:37-42)."""
    if is_semantics:
        raise NotImplementedError("semantic (nearest) interpolation is dead code in the reference (bev_rendering_utils.py:555)")
    if points.shape[0] < MIN_REQUIRED_POINTS_SIMPLEX or is_collinear(points):
        return bev_img
    # Limits of the CUDA entry point (include/salve_bev.h, salve_bev_interp_dense): values are interpolated as the uint8 colours
    # the path produces (the reference interpolates float64 and truncates afterwards: the same for integral values in 0..255,
    # which is all the render path ever passes), and every point must lie inside the grid.
    vals = np.asarray(rgb_values, np.float64)
    if not (np.all(vals == np.floor(vals)) and vals.min() >= 0 and vals.max() <= 255):
        raise NotImplementedError("interp_dense_grid_from_sparse: only integral colour values in [0, 255] are supported by the CUDA path")
    pts = np.asarray(points)[:, :2]
    if (pts < 0).any() or (pts[:, 0] >= grid_w).any() or (pts[:, 1] >= grid_h).any():
        raise NotImplementedError("interp_dense_grid_from_sparse: points outside the (grid_h, grid_w) grid are not supported by the CUDA path")
    r = _ctx.get(grid_h=max(grid_h, 2), grid_w=max(grid_w, 2))
    img, _, status = r.interp_dense(pts, vals, grid_h, grid_w)
    if status == IMG_DEGENERATE:
        return bev_img
    if status == IMG_COLLINEAR or img is None:
        raise QhullError("initial simplex is flat: all sites are collinear")
    bev_img[:] = img
    return bev_img


def remove_hallucinated_content(sparse_bev_img: np.ndarray, interp_bev_img: np.ndarray, K: int = DEFAULT_KERNEL_SZ) -> np.ndarray:
    """Zero interpolated pixels with no sparse sample in their KxK neighbourhood (:74-122).
    'Non-empty' is r*g*b > 0 evaluated in the dtype of `sparse_bev_img` -- for uint8 the product wraps mod 256."""
    sparse = np.asarray(sparse_bev_img)
    if sparse.dtype != np.uint8:
        mul = sparse[:, :, 0] * sparse[:, :, 1] * sparse[:, :, 2]
        sparse = np.repeat((mul > 0).astype(np.uint8)[:, :, None], 3, axis=2)
    interp = np.asarray(interp_bev_img).astype(np.uint8)
    return _ctx.get().remove_hallucinated(sparse, interp, K)
