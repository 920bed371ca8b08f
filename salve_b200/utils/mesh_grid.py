"""Regular grid as an (N,2) point list, x fastest (mirrors reference salve/utils/mesh_grid.py:11-36).
Defines the query order / output layout of the dense grid; not on the compute path."""

import numpy as np


def get_mesh_grid_as_point_cloud(min_x: int, max_x: int, min_y: int, max_y: int, downsample_factor: float = 1.0) -> np.ndarray:
    xs = np.linspace(min_x, max_x, int((max_x - min_x + 1) / downsample_factor))
    ys = np.linspace(min_y, max_y, int((max_y - min_y + 1) / downsample_factor))
    xg, yg = np.meshgrid(xs, ys)
    return np.stack([xg.ravel(), yg.ravel()], axis=1)
