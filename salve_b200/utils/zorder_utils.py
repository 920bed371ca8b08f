"""Z-order de-duplication (mirrors reference salve/utils/zorder_utils.py:10-83) on the GPU:
one deterministic atomicMax of a packed (z-slice, index) key per point."""

import numpy as np

from .. import _ctx

DUMMY_VAL = np.iinfo(np.uint64).max


def choose_elevated_repeated_vals(x: np.ndarray, y: np.ndarray, z: np.ndarray, zmin: float = -2, zmax: float = 2,
                                  num_slices: int = 4) -> np.ndarray:
    """valid[i] is True iff point i is the winner of pixel (x[i], y[i]): the last point of the highest
    occupied z-slice [zmin, zmax) split into num_slices bins; points outside the range never win."""
    return _ctx.get().choose_elevated(x, y, z, zmin, zmax, num_slices)
