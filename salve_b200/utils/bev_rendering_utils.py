"""BEV texture-map rendering entry points (mirror of the reference's
salve/utils/bev_rendering_utils.py texture functions, :38-45 and :254-663), backed by
libsalve_bev.so.  Signatures, return conventions (None for an empty cloud, (None, None) for a
pair) and error behaviour follow the reference; the arithmetic runs on the GPU.

Out of scope here: the layout modality (rasterize_room_layout_pair and the
cv2 polygon helpers, reference :48-251) and the dead `is_semantics=True` branches.
"""

from __future__ import annotations

import os
from argparse import Namespace
from pathlib import Path
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from .. import _ctx
from ..common.bevparams import DEFAULT_METERS_PER_PX, BEVParams  # noqa: F401  (re-exported like the reference)
from ..common.sim2 import Sim2
from ..renderer import IMG_COLLINEAR, IMG_EMPTY
from .interpolation_utils import QhullError

HOHO_S_ZIND_SCALE_FACTOR = 1.5
PANO_W, PANO_H = 1024, 512  # the reference resizes every pano to this (:373-375)


# ---- image I/O (the reference uses imageio; cv2 is the fallback reader/writer, not a compute path) ----
def _imread(path: str) -> np.ndarray:
    try:
        import imageio

        return np.asarray(imageio.imread(path))
    except ImportError:
        import cv2

        img = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
        if img is None:
            raise FileNotFoundError(path)
        if img.ndim == 3:
            img = img[:, :, ::-1]
        return np.ascontiguousarray(img)


def _imwrite(path: str, img: np.ndarray) -> None:
    try:
        import imageio

        imageio.imwrite(path, img)
    except ImportError:
        import cv2

        cv2.imwrite(str(path), np.ascontiguousarray(img[:, :, ::-1]))


def prune_to_2d_bbox(pts: np.ndarray, rgb: np.ndarray, xmin: float, ymin: float, xmax: float, ymax: float):
    """Keep points with xmin <= x <= xmax and ymin <= y <= ymax, boundaries included (:38-45).
    Public helper kept for callers; the renderer applies the same test inside the splat kernel."""
    x, y = pts[:, 0], pts[:, 1]
    ok = (xmin <= x) & (x <= xmax) & (ymin <= y) & (y <= ymax)
    return pts[ok], rgb[ok]


def _renderer_for(bev_params: BEVParams, **kw):
    return _ctx.get(
        grid_h=bev_params.img_h + 1, grid_w=bev_params.img_w + 1, xlims=tuple(float(v) for v in bev_params.xlims),
        ylims=tuple(float(v) for v in bev_params.ylims), px_per_m=1 / bev_params.meters_per_px, **kw,
    )


def render_bev_image(bev_params: BEVParams, xyzrgb: np.ndarray, is_semantics: bool) -> Optional[np.ndarray]:
    """Coloured cloud (N,6) [x,y,z in world metres, rgb in 0..1] -> dense (img_h+1, img_w+1, 3) uint8 texture
    map, or None when no point falls inside the BEV box (:254-328)."""
    if is_semantics:
        raise NotImplementedError("is_semantics=True is dead code in the reference (hard-coded False at :555)")
    img, counts, status = _renderer_for(bev_params).render_cloud(xyzrgb)
    print(f"Rendering {counts[1]/1e6} million points")
    if status == IMG_EMPTY:
        return None
    if status == IMG_COLLINEAR:
        raise QhullError("initial simplex is flat: all sites are collinear")
    return img


def _check_args(args) -> None:
    if "crop_ratio" not in args.__dict__:
        raise ValueError("Crop ratio for panorama top and bottom must be provided as `args.crop_ratio`.")
    if "crop_z_range" not in args.__dict__:
        raise ValueError("Z-coordinate range for cropping must be provided as `args.crop_z_range`.")


def _load_pano(depth_fpath: str, rgb_fpath: str):
    """uint16 depth + uint8 RGB resized to 1024x512 exactly as the reference does (:367-384, non-semantic)."""
    import cv2

    depth = _imread(depth_fpath)
    rgb = _imread(rgb_fpath)
    rgb = cv2.resize(rgb, (PANO_W, PANO_H), interpolation=cv2.INTER_LINEAR)
    if rgb.ndim == 2:
        rgb = np.repeat(rgb[:, :, None], 3, axis=2)
    if depth.shape != (PANO_H, PANO_W):
        raise ValueError(f"depth map must be {PANO_H}x{PANO_W}, got {depth.shape}")
    return np.ascontiguousarray(rgb[:, :, :3], np.uint8), np.ascontiguousarray(depth, np.uint16)


def _pano_renderer(args, **kw):
    assert args.crop_ratio < 1
    return _ctx.get(pano_h=PANO_H, pano_w=PANO_W, crop_ratio=float(args.crop_ratio), depth_scale=float(getattr(args, "scale", 0.001)), **kw)


def get_xyzrgb_from_depth(args: Union[SimpleNamespace, Namespace], depth_fpath: str, rgb_fpath: str, is_semantics: bool) -> np.ndarray:
    """Back-project a pano through its depth map; keep rows outside the top/bottom crop and points with
    crop_z_range[0] < z <= crop_z_range[1].  Returns (N,6) float64, rgb in [0,1] (:347-414)."""
    _check_args(args)
    if is_semantics:
        raise NotImplementedError("is_semantics=True is dead code in the reference")
    rgb, depth = _load_pano(depth_fpath, rgb_fpath)
    r = _pano_renderer(args)
    with r.lock:
        r.upload_pano(0, rgb, depth)
        return r.backproject(0, args.crop_z_range[0], args.crop_z_range[1], frame=0)


def get_bev_pair_xyzrgb(args, building_id: str, floor_id: str, i1: int, i2: int, i2Ti1: Sim2, is_semantics: bool):
    """The two clouds of a pair in pano 2's frame, without rendering (:483-522)."""
    _check_args(args)
    if is_semantics:
        raise NotImplementedError("is_semantics=True is dead code in the reference")
    r = _pano_renderer(args)
    rgb1, d1 = _load_pano(args.depth_i1, args.img_i1)
    rgb2, d2 = _load_pano(args.depth_i2, args.img_i2)
    print(i2Ti1)
    lo, hi = args.crop_z_range
    with r.lock:
        r.upload_pano(0, rgb1, d1)
        r.upload_pano(1, rgb2, d2)
        return r.backproject(0, lo, hi, frame=2, R=i2Ti1.rotation, t=i2Ti1.translation), r.backproject(1, lo, hi, frame=1)


def render_bev_pair(args, building_id: str, floor_id: str, i1: int, i2: int, i2Ti1: Sim2, is_semantics: bool):
    """(img1, img2): pano 1 rendered in pano 2's frame, and pano 2; (None, None) if either cloud is empty (:417-480)."""
    _check_args(args)
    if is_semantics:
        raise NotImplementedError("is_semantics=True is dead code in the reference")
    rgb1, d1 = _load_pano(args.depth_i1, args.img_i1)
    rgb2, d2 = _load_pano(args.depth_i2, args.img_i2)
    return render_bev_pair_arrays(rgb1, d1, rgb2, d2, i2Ti1, args.crop_z_range, crop_ratio=args.crop_ratio, scale=getattr(args, "scale", 0.001))


def render_bev_pair_arrays(rgb1, depth1, rgb2, depth2, i2Ti1: Sim2, crop_z_range: Sequence[float], crop_ratio: float = 80 / 512,
                           scale: float = 0.001):
    """render_bev_pair on in-memory panos of any (H, W) (W a multiple of 4)."""
    H, W = depth1.shape
    r = _ctx.get(pano_h=H, pano_w=W, crop_ratio=float(crop_ratio), depth_scale=float(scale))
    with r.lock:  # the band is a property of the (shared, cached) context: set, render and restore under its lock
        r.upload_pano(0, rgb1, depth1)
        r.upload_pano(1, rgb2, depth2)
        r.set_bands(a=(crop_z_range[0], crop_z_range[1]))
        try:
            imgs, counts, status = r.render_hypotheses([0], [1], i2Ti1.rotation[None], i2Ti1.translation[None], surfaces=("floor",))
        finally:
            r.set_bands()
    for k in range(2):
        print(f"Rendering {counts[0, 0, k, 1]/1e6} million points")
    if (status == IMG_EMPTY).any():
        return None, None
    if (status == IMG_COLLINEAR).any():
        raise QhullError("initial simplex is flat: all sites are collinear")
    return imgs[0, 0, 0].copy(), imgs[0, 0, 1].copy()


def _surface_band(surface_type: str) -> List[float]:
    if surface_type == "floor":
        return [-float("inf"), -1.0]  # everything 1 m and more below the camera (:560-562)
    if surface_type == "ceiling":
        return [0.5, float("inf")]  # everything 50 cm and more above the camera (:564-566)
    raise ValueError(f"unknown surface_type {surface_type!r}")


def bev_fname_from_img_fpath(pair_idx: int, pair_uuid: str, surface_type: str, img_fpath: str) -> str:
    """Output file name of a rendered texture map (:582-590)."""
    return f"pair_{pair_idx}___{pair_uuid}_{surface_type}_rgb_{Path(img_fpath).stem}.jpg"


def generate_texture_maps_for_pair(
    img_fpaths_dict: Dict[int, str],
    surface_type: str,
    pair_fpath: str,
    pair_idx: int,
    label_type: str,
    bev_save_root,
    building_id: str,
    floor_id: str,
    depth_save_root: str,
    render_modalities: List[str],
    layout_save_root: str,
    floor_pose_graph=None,
) -> None:
    """Render and save the two texture maps of one alignment hypothesis (:525-663): reads the {R,t,s} JSON,
    names outputs pair_{idx}___{uuid}_{surface}_rgb_{pano stem}.jpg under {bev_save_root}/{label_type}/{building_id},
    skips work if both files exist.  Depth maps must already exist (HoHoNet inference is out of scope)."""
    i2Ti1 = Sim2.from_json(json_fpath=pair_fpath)
    i1, i2 = (int(v) for v in Path(pair_fpath).stem.split("_")[:2])
    img1_fpath, img2_fpath = img_fpaths_dict[i1], img_fpaths_dict[i2]
    pair_uuid = Path(pair_fpath).stem.split("__")[-1]
    save_dir = f"{bev_save_root}/{label_type}/{building_id}"
    os.makedirs(save_dir, exist_ok=True)
    bev_fpath1 = f"{save_dir}/{bev_fname_from_img_fpath(pair_idx, pair_uuid, surface_type, img1_fpath)}"
    bev_fpath2 = f"{save_dir}/{bev_fname_from_img_fpath(pair_idx, pair_uuid, surface_type, img2_fpath)}"
    if "rgb_texture" in render_modalities:
        print(f"On {i1},{i2}")
        args = SimpleNamespace(
            img_i1=img1_fpath, img_i2=img2_fpath,
            depth_i1=f"{depth_save_root}/{building_id}/{Path(img1_fpath).stem}.depth.png",
            depth_i2=f"{depth_save_root}/{building_id}/{Path(img2_fpath).stem}.depth.png",
            scale=0.001, crop_ratio=80 / 512, crop_z_range=_surface_band(surface_type),
        )
        if Path(bev_fpath1).exists() and Path(bev_fpath2).exists():
            print("Both BEV images already exist, skipping...")
            return
        bev_img1, bev_img2 = render_bev_pair(args, building_id, floor_id, i1, i2, i2Ti1, is_semantics=False)
        if bev_img1 is None or bev_img2 is None:
            return
        _imwrite(bev_fpath1, bev_img1)
        _imwrite(bev_fpath2, bev_img2)
    if "layout" in render_modalities:
        raise NotImplementedError("the layout modality (rasterize_room_layout_pair) is outside this build's scope")


def rasterize_room_layout_pair(*args, **kwargs):
    raise NotImplementedError("the layout modality is outside this build's scope (SURVEY.md section 8f, row 4)")
