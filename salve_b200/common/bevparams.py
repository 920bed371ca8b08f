"""BEV grid parameters (mirrors the reference's salve/common/bevparams.py:17-99)."""

from __future__ import annotations

import numpy as np

from .sim2 import Sim2

DEFAULT_BEV_IMG_H_PX = 500
DEFAULT_BEV_IMG_W_PX = 500
DEFAULT_METERS_PER_PX = 0.02
FULL_RES_METERS_PER_PX = 0.005
FULL_RES_LINE_WIDTH_PX = 30


class BEVParams:
    """img_h x img_w pixels at meters_per_px; limits are truncated to whole metres (:52-61)."""

    def __init__(self, img_h: int = DEFAULT_BEV_IMG_H_PX, img_w: int = DEFAULT_BEV_IMG_W_PX,
                 meters_per_px: float = DEFAULT_METERS_PER_PX) -> None:
        self.img_h, self.img_w, self.meters_per_px = img_h, img_w, meters_per_px
        half_w_m = int((img_w / 2) * meters_per_px)
        half_h_m = int((img_h / 2) * meters_per_px)
        self.xlims = [-half_w_m, half_w_m]
        self.ylims = [-half_h_m, half_h_m]

    @property
    def bevimg_Sim2_world(self) -> Sim2:
        """p_bevimg = bevimg_Sim2_world * p_world (:69-78)."""
        return Sim2(R=np.eye(2), t=np.array([-self.xlims[0], -self.ylims[0]]), s=1 / self.meters_per_px)


def get_line_width_by_resolution(resolution: float) -> int:
    """Polyline width in px for a rendering resolution (:81-99)."""
    return max(round(FULL_RES_LINE_WIDTH_PX / (resolution / FULL_RES_METERS_PER_PX)), 1)
