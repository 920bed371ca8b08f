"""Process-wide cache of render contexts for the function-style (drop-in) entry points.

A context is not thread-safe (include/salve_bev.h), so every cached renderer carries a lock: the drop-in functions hold it for
the whole upload -> (set bands) -> render -> (restore bands) sequence.  The cache is a small LRU; an evicted renderer is only
dropped from the cache -- whoever still holds it keeps a live context, and the context is destroyed when the last reference goes.
"""

from __future__ import annotations

import os
import threading
from collections import OrderedDict
from typing import Tuple

from .renderer import BevRenderer

MAX_CONTEXTS = 6  # scratch is large: keep only a few contexts alive
_CACHE: "OrderedDict[Tuple, BevRenderer]" = OrderedDict()
_LOCK = threading.Lock()


def device() -> int:
    return int(os.environ.get("SALVE_BEV_DEVICE", "0"))


def get(pano_h: int = 512, pano_w: int = 1024, grid_h: int = 501, grid_w: int = 501, xlims=(-5.0, 5.0), ylims=(-5.0, 5.0),
        px_per_m: float = 50.0, kernel_sz: int = 11, crop_ratio: float = 80 / 512, depth_scale: float = 0.001,
        max_images: int = 4, max_panos: int = 2) -> BevRenderer:
    """Renderer for this configuration (created on first use).  Use `with r.lock:` around a sequence of calls on it."""
    key = (device(), pano_h, pano_w, grid_h, grid_w, tuple(xlims), tuple(ylims), px_per_m, kernel_sz, crop_ratio, depth_scale, max_images, max_panos)
    with _LOCK:
        r = _CACHE.get(key)
        if r is not None:
            _CACHE.move_to_end(key)
            return r
        while len(_CACHE) >= MAX_CONTEXTS:
            _CACHE.popitem(last=False)  # least recently used; not closed here: another caller may still hold it (BevRenderer.__del__ closes)
        r = BevRenderer(pano_h=pano_h, pano_w=pano_w, max_panos=max_panos, max_images=max_images, device=device(), grid_h=grid_h,
                        grid_w=grid_w, xlims=xlims, ylims=ylims, px_per_m=px_per_m, kernel_sz=kernel_sz, crop_ratio=crop_ratio,
                        depth_scale=depth_scale)
        _CACHE[key] = r
        return r


def clear() -> None:
    with _LOCK:
        _CACHE.clear()
