"""Unit-sphere directions of an equirectangular pano (mirrors reference
salve/utils/hohonet_pano_utils.py:10-44).  The table is assembled on the GPU from numpy's trig
factors, so it is bit-identical to the reference's."""

import numpy as np

from .. import _ctx


def get_uni_sphere_xyz(H: int, W: int) -> np.ndarray:
    """(H,W,3) float64: x = cos(phi)cos(theta), y = cos(phi)sin(theta), z = -sin(phi); -x points at the pano centre."""
    if W % 4:
        raise ValueError("pano width must be a multiple of 4")
    return _ctx.get(pano_h=H, pano_w=W).uni_sphere_xyz()
