"""Seeded synthetic inputs for the BEV rendering path.

TEST/BENCH INFRASTRUCTURE.  Shared by tests/, bench.py and __graft_entry__.smoke();
it only *produces inputs* (no rendering arithmetic lives here).

Pano k -> seed k.  Hypothesis j -> seed 10_000 + j.

Input formats follow the reference hot path:
  * depth: uint16 millimetres, shape (H, W)          (salve/utils/infer_depth.py:59-62,
                                                      salve/utils/bev_rendering_utils.py:367)
  * rgb:   uint8, shape (H, W, 3)                    (salve/utils/bev_rendering_utils.py:370-375)
  * pose:  float32 R (2,2), t (2,), s = 1.0          (salve/common/sim2.py:50-52)
"""

from __future__ import annotations

import numpy as np

# Box room in the HoHoNet frame (camera at origin, z up).
FLOOR_Z = -1.5
CEIL_Z = 1.2
WALL_X = 3.0
WALL_Y = 2.5


def _unit_sphere(H: int, W: int) -> np.ndarray:
    """Ray directions per equirect pixel, same convention as hohonet_pano_utils.py:27-43."""
    v = np.arange(H, dtype=np.float64)
    u = np.arange(W, dtype=np.float64)
    theta = -(u + 0.5) / W * (2 * np.pi)
    phi = ((v + 0.5) / H - 0.5) * np.pi
    z = -np.sin(phi)
    r = np.cos(phi)
    x = r[:, None] * np.cos(theta)[None, :]
    y = r[:, None] * np.sin(theta)[None, :]
    zz = np.broadcast_to(z[:, None], (H, W))
    return np.stack([x, y, zz], -1)


def synth_depth(H: int, W: int, seed: int, jitter: float = 0.0) -> np.ndarray:
    """Ray-cast box-room depth + N(0, 5 mm) noise, uint16 mm."""
    rng = np.random.default_rng(seed)
    s = 1.0
    if jitter > 0:
        s = 1.0 + jitter * (2 * rng.random() - 1)
    d = _unit_sphere(H, W)
    with np.errstate(divide="ignore", invalid="ignore"):
        t_floor = np.where(d[..., 2] < 0, FLOOR_Z / d[..., 2], np.inf)
        t_ceil = np.where(d[..., 2] > 0, CEIL_Z / d[..., 2], np.inf)
        t_wx = (WALL_X * s) / np.abs(d[..., 0])
        t_wy = (WALL_Y * s) / np.abs(d[..., 1])
    t = np.minimum(np.minimum(t_floor, t_ceil), np.minimum(t_wx, t_wy))
    mm = t * 1000.0 + rng.normal(0.0, 5.0, size=(H, W))
    return np.clip(np.rint(mm), 1, 65535).astype(np.uint16)


def synth_rgb(H: int, W: int, seed: int, texture: str = "iid") -> np.ndarray:
    """uint8 RGB in [1, 255].  texture: 'iid' (worst case for Delaunay tie-breaks) or 'smooth'."""
    rng = np.random.default_rng(seed + 500_000)
    if texture == "iid":
        return rng.integers(1, 256, size=(H, W, 3), dtype=np.int64).astype(np.uint8)
    if texture == "smooth":
        # low-res random field, separable linear upsample (pure numpy so it is portable), + noise
        h0, w0 = max(H // 32, 2), max(W // 32, 2)
        low = rng.uniform(20, 235, size=(h0, w0, 3))
        yi = np.linspace(0, h0 - 1, H)
        xi = np.linspace(0, w0 - 1, W)
        y0 = np.floor(yi).astype(int).clip(0, h0 - 2)
        x0 = np.floor(xi).astype(int).clip(0, w0 - 2)
        fy = (yi - y0)[:, None, None]
        fx = (xi - x0)[None, :, None]
        a = low[y0][:, x0]
        b = low[y0][:, x0 + 1]
        c = low[y0 + 1][:, x0]
        d = low[y0 + 1][:, x0 + 1]
        img = (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy
        img = img + rng.uniform(-6, 6, size=(H, W, 3))
        return np.clip(np.rint(img), 1, 255).astype(np.uint8)
    raise ValueError(texture)


def synth_pano(H: int, W: int, seed: int, texture: str = "iid", jitter: float = 0.0):
    """(rgb u8 (H,W,3), depth u16 (H,W)) for pano `seed`."""
    return synth_rgb(H, W, seed, texture), synth_depth(H, W, seed, jitter)


def synth_pose(j: int):
    """Hypothesis j -> (R float32 (2,2), t float32 (2,)), SE(2) with s = 1."""
    rng = np.random.default_rng(10_000 + j)
    th = np.deg2rad(rng.uniform(-180.0, 180.0))
    t = rng.uniform(-2.0, 2.0, size=2)
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    return R.astype(np.float32), t.astype(np.float32)


def synth_building(n_panos: int, n_hyp: int, H: int, W: int, seed: int = 0, texture: str = "iid"):
    """One synthetic building: panos + hypothesis list (i1, i2, R, t).

    Pairs are drawn i1 < i2 deterministically from `seed`.
    """
    rgbs = np.empty((n_panos, H, W, 3), np.uint8)
    depths = np.empty((n_panos, H, W), np.uint16)
    for k in range(n_panos):
        rgbs[k], depths[k] = synth_pano(H, W, seed * 1000 + k, texture, jitter=0.2)
    rng = np.random.default_rng(seed + 77)
    i1 = rng.integers(0, n_panos, size=n_hyp)
    i2 = rng.integers(0, n_panos - 1, size=n_hyp)
    i2 = np.where(i2 >= i1, i2 + 1, i2)
    lo = np.minimum(i1, i2).astype(np.int32)
    hi = np.maximum(i1, i2).astype(np.int32)
    R = np.empty((n_hyp, 2, 2), np.float32)
    t = np.empty((n_hyp, 2), np.float32)
    for j in range(n_hyp):
        R[j], t[j] = synth_pose(seed * 100_000 + j)
    return rgbs, depths, lo, hi, R, t
