"""CPU restatement (numpy + SciPy) of SALVe's BEV texture-map rendering path.

TEST INFRASTRUCTURE: this is the *checker* for the CUDA path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
The product (salve_b200/) never does.

Pinned: scripts/make_golden.py checks this file bit-for-bit against the imported, unmodified
reference (oracle/ref_import.py) at 512x1024 -- final images, and every stage that the
reference exposes -- and freezes the results under tests/golden/.  The densification
arithmetic itself lives in a third-party dependency that is NOT under /root/reference:
scipy.interpolate.griddata -> LinearNDInterpolator -> scipy.spatial.Delaunay (Qhull,
options "Qbb Qc Qz Q12" + "Qt") + scipy's _interpnd barycentric evaluation.  The reference
does not pin SciPy (setup.py:43 install_requires=[]); this container has scipy 1.18.1.
We call SciPy exactly as salve/utils/interpolation_utils.py:46-48 does.

Differences from the reference, all deliberate:
  * (H, W) are parameters (reference hard-codes 512x1024, bev_rendering_utils.py:373-375);
    cv2.resize to the same size is the identity, so it is omitted.
  * inputs are arrays, not file paths.
  * every intermediate stage is returned (crop mask, pixel indices, winner keys, ...).
"""

from __future__ import annotations

import warnings
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

# ---- constants of the path -------------------------------------------------------------
GRID_PX = 500  # salve/common/bevparams.py:17-18
METERS_PER_PX = 0.02  # salve/common/bevparams.py:19
IMG = GRID_PX + 1  # bev_rendering_utils.py:292-293
KERNEL_SZ = 11  # interpolation_utils.py:15
MIN_PTS = 4  # interpolation_utils.py:18
SCALE = 0.001  # bev_rendering_utils.py:611
CROP_RATIO = 80 / 512  # bev_rendering_utils.py:613
HOHO_S_ZIND = 1.5  # bev_rendering_utils.py:448
BANDS = {  # bev_rendering_utils.py:560-566  (lo, hi]  keep lo < z <= hi
    "floor": (-float("inf"), -1.0),
    "ceiling": (0.5, float("inf")),
}


def rotmat2d(theta_deg: float) -> np.ndarray:
    """salve/utils/rotation_utils.py:14-29."""
    th = np.deg2rad(theta_deg)
    s, c = np.sin(th), np.cos(th)
    return np.array([[c, -s], [s, c]])


def sphere_tables(H: int, W: int):
    """Separable factors of get_uni_sphere_xyz (salve/utils/hohonet_pano_utils.py:27-43).

    Returns (cos_phi[H], neg_sin_phi[H], cos_theta[W], sin_theta[W]) float64, such that
    x = cos_phi[v]*cos_theta[u], y = cos_phi[v]*sin_theta[u], z = neg_sin_phi[v].
    """
    u = np.arange(W)
    v = np.arange(H)
    theta = -(u + 0.5) / W
    theta *= 2 * np.pi
    phi = (v + 0.5) / H
    phi -= 0.5
    phi *= np.pi
    return np.cos(phi), -np.sin(phi), np.cos(theta), np.sin(theta)


def uni_sphere_xyz(H: int, W: int) -> np.ndarray:
    cphi, nsphi, cth, sth = sphere_tables(H, W)
    x = cphi[:, None] * cth[None, :]
    y = cphi[:, None] * sth[None, :]
    z = np.broadcast_to(nsphi[:, None], (H, W))
    return np.stack([x, y, z], -1)


def backproject(rgb: np.ndarray, depth: np.ndarray, band, crop_ratio: float = CROP_RATIO, scale: float = SCALE):
    """bev_rendering_utils.py:347-414 on arrays.  Returns (xyzrgb (N,6) f64, src_lin (N,) i64, crop_mask)."""
    H, W = depth.shape
    d = depth[..., None].astype(np.float32) * scale  # f32 * python float -> f32   (:367)
    xyz = d * uni_sphere_xyz(H, W)  # f32 * f64 -> f64                             (:392)
    xyzrgb = np.concatenate([xyz, rgb / 255.0], 2)  #                               (:394)
    src = np.arange(H * W, dtype=np.int64).reshape(H, W)
    if crop_ratio > 0:
        assert crop_ratio < 1
        c = int(H * crop_ratio)  #                                                  (:399)
        xyzrgb = xyzrgb[c:-c]
        src = src[c:-c]
    xyzrgb = xyzrgb.reshape(-1, 6)
    src = src.reshape(-1)
    within = np.logical_and(xyzrgb[:, 2] > band[0], xyzrgb[:, 2] <= band[1])  #    (:408-411)
    return xyzrgb[within], src[within], within


def to_zind_frame(xyzrgb: np.ndarray) -> None:
    """bev_rendering_utils.py:443-446, in place."""
    R = rotmat2d(-90)
    xyzrgb[:, :2] = xyzrgb[:, :2] @ R.T


def apply_pose(xyzrgb: np.ndarray, R32: np.ndarray, t32: np.ndarray) -> None:
    """bev_rendering_utils.py:451, in place.  R32, t32 are float32 (salve/common/sim2.py:50-52)."""
    R32 = np.asarray(R32, np.float32)
    t32 = np.asarray(t32, np.float32)
    xyzrgb[:, :2] = (xyzrgb[:, :2] @ R32.T) + (t32 * HOHO_S_ZIND)


def choose_elevated(x, y, z, zmin=-2.0, zmax=2.0, num_slices=4) -> np.ndarray:
    """Winner rule of salve/utils/zorder_utils.py:49-65, closed form:
    winner(pixel) = argmax over points with z in [zmin, zmax) of (slice, index)."""
    n = x.shape[0]
    planes = np.linspace(zmin, zmax, num_slices + 1)
    sl = np.full(n, -1, np.int64)
    for k in range(num_slices):
        sl[np.logical_and(z >= planes[k], z < planes[k + 1])] = k
    ok = sl >= 0
    w = int(x.max()) + 1 if n else 1
    pix = y.astype(np.int64) * w + x.astype(np.int64)
    key = sl * n + np.arange(n)
    valid = np.zeros(n, bool)
    if not ok.any():
        return valid
    order = np.lexsort((key[ok], pix[ok]))
    p_sorted = pix[ok][order]
    idx_sorted = np.nonzero(ok)[0][order]
    last = np.r_[p_sorted[1:] != p_sorted[:-1], True]
    valid[idx_sorted[last]] = True
    return valid


@dataclass
class Stages:
    """Every intermediate of one rendered image (bev_rendering_utils.py:254-328)."""

    count_crop: int = 0
    count_bbox: int = 0
    src_bbox: Optional[np.ndarray] = None  # (count_bbox,) source linear index of points inside the bbox
    col: Optional[np.ndarray] = None  # (count_bbox,) pixel column
    row: Optional[np.ndarray] = None  # (count_bbox,) pixel row
    z: Optional[np.ndarray] = None  # (count_bbox,)
    key_grid: Optional[np.ndarray] = None  # (IMG, IMG) int64: (slice<<29 | src_lin) + 1 of the winner, 0 = empty
    site_rc: Optional[np.ndarray] = None  # (S,2) row, col -- ascending source index order (reference order)
    site_rgb: Optional[np.ndarray] = None  # (S,3) u8
    sparse: Optional[np.ndarray] = None  # (IMG,IMG,3) u8
    nonempty: Optional[np.ndarray] = None  # (IMG,IMG) bool
    degenerate: bool = False
    interp: Optional[np.ndarray] = None  # (IMG,IMG,3) u8, before mask / flip
    interp_f64: Optional[np.ndarray] = None  # (IMG,IMG,3) f64 with NaN outside the hull
    hull: Optional[np.ndarray] = None  # (IMG,IMG) bool
    keep: Optional[np.ndarray] = None  # (IMG,IMG) bool
    final: Optional[np.ndarray] = None  # (IMG,IMG,3) u8 (flipped) or None when the cloud is empty


def keep_mask(nonempty: np.ndarray, K: int = KERNEL_SZ) -> np.ndarray:
    """counts>0 of a KxK zero-padded box filter == OR over the (K//2)-Chebyshev ball
    (interpolation_utils.py:101-115)."""
    H, W = nonempty.shape
    r = K // 2
    pad = np.zeros((H + 2 * r, W + 2 * r), bool)
    pad[r : r + H, r : r + W] = nonempty
    rows = np.zeros_like(pad)
    for d in range(-r, K - r):
        rows |= np.roll(pad, d, axis=1)
    out = np.zeros_like(pad)
    for d in range(-r, K - r):
        out |= np.roll(rows, d, axis=0)
    return out[r : r + H, r : r + W]


def nonempty_mask(sparse: np.ndarray) -> np.ndarray:
    """uint8 product wraps mod 256 (interpolation_utils.py:95-98)."""
    mul = sparse[:, :, 0] * sparse[:, :, 1] * sparse[:, :, 2]
    return mul > 0


def griddata_linear(site_xy: np.ndarray, values: np.ndarray, grid_h: int, grid_w: int) -> np.ndarray:
    """interpolation_utils.py:44-48: float64 (grid_h*grid_w, C) with NaN outside the hull."""
    import scipy.interpolate

    x = np.linspace(0, grid_w - 1, grid_w)
    y = np.linspace(0, grid_h - 1, grid_h)
    xg, yg = np.meshgrid(x, y)
    xi = np.hstack([xg.flatten()[:, None], yg.flatten()[:, None]])
    return scipy.interpolate.griddata(points=site_xy, values=values, xi=xi, method="linear")


def is_degenerate(site_xy: np.ndarray) -> bool:
    """interpolation_utils.py:37-42, 57-71."""
    if site_xy.shape[0] < MIN_PTS:
        return True
    if np.allclose(site_xy[:, 0], site_xy[0, 0]):
        return True
    if np.allclose(site_xy[:, 1], site_xy[0, 1]):
        return True
    return False


def render_image(xyzrgb: np.ndarray, src: np.ndarray, W_pano: int, densify: bool = True) -> Stages:
    """render_bev_image (bev_rendering_utils.py:254-328) with all stages exposed.

    `xyzrgb` already in the frame to render; `src` = source pano linear index per point.
    """
    st = Stages()
    st.count_crop = xyzrgb.shape[0]
    xyz = xyzrgb[:, :3]
    rgb = xyzrgb[:, 3:] * 255
    lim = int((GRID_PX / 2) * METERS_PER_PX)  # bevparams.py:52-61 -> 5
    x, y = xyz[:, 0], xyz[:, 1]
    ok = np.logical_and.reduce([-lim <= x, x <= lim, -lim <= y, y <= lim])  # :38-45
    xyz, rgb, src = xyz[ok], rgb[ok], src[ok]
    st.count_bbox = xyz.shape[0]
    st.src_bbox = src
    if st.count_bbox == 0:  # :279-280
        return st
    R = np.eye(2).astype(np.float32)  # bevparams.py:78 / sim2.py:50
    t = np.array([lim, lim]).astype(np.float32)
    s = float(1 / METERS_PER_PX)
    img_xy = ((xyz[:, :2] @ R.T) + t) * s  # sim2.py:157-160
    img_xy = np.round(img_xy).astype(np.int64)  # :287
    col, row, z = img_xy[:, 0], img_xy[:, 1], xyz[:, 2]
    st.col, st.row, st.z = col, row, z
    valid = choose_elevated(col, row, z)  # :298-305
    planes = np.linspace(-2.0, 2.0, 5)
    sl = np.searchsorted(planes, z[valid], side="right") - 1
    st.key_grid = np.zeros((IMG, IMG), np.int64)
    st.key_grid[row[valid], col[valid]] = ((sl << 29) | src[valid]) + 1
    site_xy = img_xy[valid]
    site_rgb = rgb[valid]
    st.site_rc = np.stack([site_xy[:, 1], site_xy[:, 0]], 1)
    sparse = np.zeros((IMG, IMG, 3), np.uint8)
    with np.errstate(invalid="ignore"):
        sparse[site_xy[:, 1], site_xy[:, 0]] = site_rgb  # :307-308
    st.sparse = sparse
    st.site_rgb = sparse[site_xy[:, 1], site_xy[:, 0]]
    st.nonempty = nonempty_mask(sparse)
    st.keep = keep_mask(st.nonempty)
    st.degenerate = is_degenerate(site_xy)
    if not densify:
        return st
    interp = np.zeros((IMG, IMG, 3), np.uint8)
    if not st.degenerate:
        vals = griddata_linear(site_xy, site_rgb, IMG, IMG)
        st.interp_f64 = vals.reshape(IMG, IMG, 3)
        st.hull = ~np.isnan(st.interp_f64[:, :, 0])
        with warnings.catch_warnings(), np.errstate(invalid="ignore"):
            warnings.simplefilter("ignore")
            interp[:] = st.interp_f64  # f64 -> u8 truncation, NaN -> 0   (:51-53)
    else:
        st.hull = np.zeros((IMG, IMG), bool)
    st.interp = interp
    mask3 = np.repeat(st.keep[:, :, None], 3, 2).astype(np.float32)
    st.final = np.flipud((mask3 * interp).astype(np.uint8))  # interpolation_utils.py:118-121, :319
    return st


def render_pair(rgb1, depth1, rgb2, depth2, R32, t32, surface: str, densify: bool = True):
    """render_bev_pair (bev_rendering_utils.py:417-480): (Stages pano1-posed, Stages pano2)."""
    band = BANDS[surface]
    a, sa, _ = backproject(rgb1, depth1, band)
    b, sb, _ = backproject(rgb2, depth2, band)
    to_zind_frame(a)
    to_zind_frame(b)
    apply_pose(a, R32, t32)
    W = depth1.shape[1]
    s1 = render_image(a, sa, W, densify)
    s2 = render_image(b, sb, W, densify)
    return s1, s2


def render_pair_images(rgb1, depth1, rgb2, depth2, R32, t32, surface: str):
    """Final images only, with the reference's (None, None) rule (bev_rendering_utils.py:457-458)."""
    s1, s2 = render_pair(rgb1, depth1, rgb2, depth2, R32, t32, surface)
    if s1.final is None or s2.final is None:
        return None, None
    return s1.final, s2.final
