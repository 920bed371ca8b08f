"""Sim(2): 2-d similarity transform, host-side value type.

Mirrors the public interface of the reference's salve/common/sim2.py:23-199 (names, argument
meaning, error behaviour, and the float32 storage of R and t at :50-52 that the renderer's
arithmetic depends on).  Seven numbers per hypothesis; nothing here is hot.
"""

from __future__ import annotations

import json
import os
from typing import Union

import numpy as np

_PathLike = Union[str, "os.PathLike[str]"]


class Sim2:
    def __init__(self, R: np.ndarray, t: np.ndarray, s: Union[int, float]) -> None:
        for name, arr, shape in (("R", R, (2, 2)), ("t", t, (2,))):
            if not isinstance(arr, np.ndarray):
                raise ValueError(f"Input array `{name}` must be a Numpy n-d array.")
            if arr.shape != shape:
                raise ValueError(f"Input array `{name}` must have shape {shape}.")
        assert isinstance(s, (float, int))
        if np.isclose(s, 0.0):
            raise ZeroDivisionError("3x3 matrix formation would require division by zero")
        self.R_ = R.astype(np.float32)
        self.t_ = t.astype(np.float32)
        self.s_ = float(s)

    # ---- accessors ---------------------------------------------------------------------------
    @property
    def rotation(self) -> np.ndarray:
        return self.R_

    @property
    def translation(self) -> np.ndarray:
        return self.t_

    @property
    def scale(self) -> float:
        return self.s_

    @property
    def theta_deg(self) -> float:
        return float(np.rad2deg(np.arctan2(self.R_[1, 0], self.R_[0, 0])))

    @property
    def matrix(self) -> np.ndarray:
        T = np.zeros((3, 3))
        T[:2, :2] = self.R_
        T[:2, 2] = self.t_
        T[2, 2] = 1 / self.s_
        return T

    def __repr__(self) -> str:
        return f"Angle (deg.): {self.theta_deg:.1f}, Trans.: {np.round(self.t_,2)}, Scale: {self.s_:.1f}"

    def __eq__(self, other: object) -> bool:
        return (
            isinstance(other, Sim2)
            and bool(np.isclose(self.scale, other.scale))
            and bool(np.allclose(self.rotation, other.rotation))
            and bool(np.allclose(self.translation, other.translation))
        )

    # ---- group operations ------------------------------------------------------------------------
    def compose(self, S: "Sim2") -> "Sim2":
        return Sim2(R=self.R_ @ S.R_, t=self.R_ @ S.t_ + ((1.0 / S.s_) * self.t_), s=self.s_ * S.s_)

    def inverse(self) -> "Sim2":
        Rt = self.R_.T
        return Sim2(Rt, -Rt @ (self.s_ * self.t_), 1.0 / self.s_)

    def transform_from(self, point_cloud: np.ndarray) -> np.ndarray:
        """p_b = s * (R p_a + t) on an (N,2) array."""
        if not isinstance(point_cloud, np.ndarray):
            raise ValueError("Input `point_cloud` must be a Numpy n-d array.")
        if point_cloud.ndim != 2:
            raise ValueError("Input point cloud is not 2-dimensional.")
        if point_cloud.shape[1] != 2:
            raise ValueError("Input `point_cloud` must have shape (N,2).")
        return ((point_cloud @ self.R_.T) + self.t_) * self.s_

    def transform_point_cloud(self, point_cloud: np.ndarray) -> np.ndarray:
        return self.transform_from(point_cloud)

    # ---- (de)serialisation: {"R": [4], "t": [2], "s": float}, row-major ---------------------------
    def save_as_json(self, save_fpath: _PathLike) -> None:
        d = {"R": self.rotation.flatten().tolist(), "t": self.translation.flatten().tolist(), "s": self.scale}
        os.makedirs(os.path.dirname(os.path.abspath(save_fpath)), exist_ok=True)
        with open(save_fpath, "w") as f:
            json.dump(d, f)

    @classmethod
    def from_json(cls, json_fpath: _PathLike) -> "Sim2":
        with open(json_fpath, "r") as f:
            d = json.load(f)
        return cls(np.array(d["R"]).reshape(2, 2), np.array(d["t"]).reshape(2), float(d["s"]))

    @classmethod
    def from_matrix(cls, T: np.ndarray) -> "Sim2":
        if np.isclose(T[2, 2], 0.0):
            raise ZeroDivisionError("Sim(2) scale calculation would lead to division by zero.")
        return cls(T[:2, :2], T[:2, 2], 1 / T[2, 2])
