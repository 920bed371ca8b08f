"""Process-wide cache of render contexts for the function-style (drop-in) entry points."""

from __future__ import annotations

import os
from typing import Dict, Tuple

from .renderer import BevRenderer

_CACHE: Dict[Tuple, BevRenderer] = {}


def device() -> int:
    return int(os.environ.get("SALVE_BEV_DEVICE", "0"))


def get(pano_h: int = 512, pano_w: int = 1024, grid_h: int = 501, grid_w: int = 501, xlims=(-5.0, 5.0), ylims=(-5.0, 5.0),
        px_per_m: float = 50.0, kernel_sz: int = 11, crop_ratio: float = 80 / 512, depth_scale: float = 0.001,
        max_images: int = 4, max_panos: int = 2) -> BevRenderer:
    key = (device(), pano_h, pano_w, grid_h, grid_w, tuple(xlims), tuple(ylims), px_per_m, kernel_sz, crop_ratio, depth_scale, max_images, max_panos)
    r = _CACHE.get(key)
    if r is None:
        if len(_CACHE) >= 6:  # scratch is large: keep only a few contexts alive
            _, old = _CACHE.popitem()
            old.close()
        r = BevRenderer(pano_h=pano_h, pano_w=pano_w, max_panos=max_panos, max_images=max_images, device=device(), grid_h=grid_h,
                        grid_w=grid_w, xlims=xlims, ylims=ylims, px_per_m=px_per_m, kernel_sz=kernel_sz, crop_ratio=crop_ratio,
                        depth_scale=depth_scale)
        _CACHE[key] = r
    return r


def clear() -> None:
    for r in _CACHE.values():
        r.close()
    _CACHE.clear()
