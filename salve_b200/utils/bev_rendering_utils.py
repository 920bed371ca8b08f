"""BEV rendering entry points (mirror of the reference's salve/utils/bev_rendering_utils.py: the texture
functions :38-45 and :254-663 and the layout modality :48-251), backed by libsalve_bev.so.  Signatures, return
conventions (None for an empty cloud, (None, None) for a pair) and error behaviour follow the reference; the
arithmetic runs on the GPU.

Out of scope here: the dead `is_semantics=True` branches.
"""

from __future__ import annotations

import os
from argparse import Namespace
from pathlib import Path
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from .. import _ctx
from ..common import bevparams as bevparams_mod
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
    if "layout" not in render_modalities:
        return
    # Rasterised layout (:632-661): only for `floor` (the ceiling's would be identical).
    if surface_type != "floor":
        return
    building_layout_save_dir = f"{layout_save_root}/{label_type}/{building_id}"
    os.makedirs(building_layout_save_dir, exist_ok=True)
    layout_fpath1 = f"{building_layout_save_dir}/{bev_fname_from_img_fpath(pair_idx, pair_uuid, surface_type, img1_fpath)}"
    layout_fpath2 = f"{building_layout_save_dir}/{bev_fname_from_img_fpath(pair_idx, pair_uuid, surface_type, img2_fpath)}"
    if Path(layout_fpath1).exists() and Path(layout_fpath2).exists():
        print("Both layout images already exist, skipping...")
        return
    layoutimg1, layoutimg2 = rasterize_room_layout_pair(i2Ti1=i2Ti1, floor_pose_graph=floor_pose_graph, building_id=building_id,
                                                        floor_id=floor_id, i1=i1, i2=i2)
    _imwrite(layout_fpath1, layoutimg1)
    _imwrite(layout_fpath2, layoutimg2)


# ---- layout modality (:48-251): room polygon + window / door / opening strokes, rasterised on the GPU ------------------------
RED = [255, 0, 0]
GREEN = [0, 255, 0]
BLUE = [0, 0, 255]
WDO_COLOR_DICT_CV2 = {"windows": RED, "doors": GREEN, "openings": BLUE}  # :28-31


def _to_px(xy: np.ndarray, bevimg_Sim2_world: Sim2) -> np.ndarray:
    """World -> integer pixel coordinates exactly as rasterize_polygon / rasterize_polyline do (:193-194, :214-215)."""
    return np.round(bevimg_Sim2_world.transform_from(np.asarray(xy, np.float64).reshape(-1, 2))).astype(np.int64)


def _wdo_in_frame(wdo, i2Ti1: Optional[Sim2]):
    """(type, (2, 2) vertices) of a W/D/O object, moved into pano 2's frame when a pose is given (pano_data.WDO.transform_from, wdo.py:129-144).
    Any object with `.type` and `.vertices_local_2d` (or `.pt1` / `.pt2`) will do; the reference's own class is used as it is."""
    v = np.asarray(wdo.vertices_local_2d if hasattr(wdo, "vertices_local_2d") else [wdo.pt1, wdo.pt2], np.float64).reshape(-1, 2)
    if i2Ti1 is not None:
        v = i2Ti1.transform_from(v)
    return wdo.type, v


def _layout_desc(bev_params: BEVParams, room_vertices, wdos, render_mask: bool = True, flip: bool = True) -> dict:
    """One image's drawing list for BevRenderer.rasterize_layouts (what rasterize_single_layout draws, :101-156)."""
    S = bev_params.bevimg_Sim2_world
    thick = bevparams_mod.get_line_width_by_resolution(bevparams_mod.DEFAULT_METERS_PER_PX)  # 8 px at 500 x 500 (:125)
    room_px = _to_px(np.asarray(room_vertices, np.float64) * HOHO_S_ZIND_SCALE_FACTOR, S)
    strokes = []
    polygon = None
    if render_mask:
        polygon = room_px
    else:
        t = int(thick / 3)
        strokes += [(*room_px[k], *room_px[k + 1], (255, 255, 255), t) for k in range(len(room_px) - 1)]
    for wtype, v in wdos:
        px = _to_px(v * HOHO_S_ZIND_SCALE_FACTOR, S)
        strokes += [(*px[k], *px[k + 1], WDO_COLOR_DICT_CV2[wtype], thick) for k in range(len(px) - 1)]
    return dict(polygon=polygon, polygon_rgb=(255, 255, 255), strokes=strokes, flip=flip)


def rasterize_single_layout(bev_params: BEVParams, room_vertices: np.ndarray, wdo_objs, render_mask: bool = True) -> np.ndarray:
    """Room boundary in white (filled mask, or a thin contour), windows / doors / openings as 8-px strokes in their colours, then
    np.flipud (:101-156)."""
    r = _renderer_for(bev_params)
    return r.rasterize_layouts([_layout_desc(bev_params, room_vertices, [_wdo_in_frame(w, None) for w in wdo_objs], render_mask)])[0]


def rasterize_room_layout_pair(i2Ti1: Sim2, floor_pose_graph, building_id: str, floor_id: str, i1: int, i2: int):
    """BEV rasterisation of the two panoramas' room layouts, pano 1's moved into pano 2's frame (:48-98).  `floor_pose_graph.nodes[i]`
    needs `room_vertices_local_2d`, `doors`, `windows`, `openings` (the reference's PoseGraph2d / PanoData, or anything shaped like them)."""
    bev_params = BEVParams()
    n1, n2 = floor_pose_graph.nodes[i1], floor_pose_graph.nodes[i2]
    v1 = np.asarray(n1.room_vertices_local_2d, np.float64)
    v2 = np.asarray(n2.room_vertices_local_2d, np.float64)
    # repeat the first vertex as the last one (:76-77), then pano 1's room goes into pano 2's frame (:79)
    v1 = i2Ti1.transform_from(np.vstack([v1, v1[0].reshape(-1, 2)]))
    v2 = np.vstack([v2, v2[0].reshape(-1, 2)])
    w1 = [_wdo_in_frame(w, i2Ti1) for w in list(n1.doors) + list(n1.windows) + list(n1.openings)]
    w2 = [_wdo_in_frame(w, None) for w in list(n2.doors) + list(n2.windows) + list(n2.openings)]
    r = _renderer_for(bev_params)
    imgs = r.rasterize_layouts([_layout_desc(bev_params, v1, w1), _layout_desc(bev_params, v2, w2)])
    return imgs[0], imgs[1]


def _draw_on(image: np.ndarray, layout: dict) -> np.ndarray:
    h, w = image.shape[:2]
    r = _ctx.get(grid_h=h, grid_w=w)
    out = r.rasterize_layouts([layout], init=np.ascontiguousarray(image, np.uint8)[None])[0]
    image[:] = out
    return image


def draw_polygon_cv2(points: np.ndarray, image: np.ndarray, color) -> np.ndarray:
    """cv2.fillPoly of one (possibly non-convex) polygon onto `image` (:159-181)."""
    return _draw_on(image, dict(polygon=np.asarray(points).astype(np.int32), polygon_rgb=tuple(color), strokes=[], flip=False))


def rasterize_polygon(polygon_xy: np.ndarray, bev_img: np.ndarray, bevimg_Sim2_world: Sim2, color) -> np.ndarray:
    """:184-197"""
    return draw_polygon_cv2(points=_to_px(polygon_xy, bevimg_Sim2_world), image=bev_img, color=color)


def draw_polyline_cv2(line_segments_arr: np.ndarray, image: np.ndarray, color, im_h: int, im_w: int, thickness: int = 2) -> None:
    """Anti-aliased strokes between consecutive points, drawn onto `image` in place (:220-251)."""
    p = np.asarray(line_segments_arr, np.int64).reshape(-1, 2)
    _draw_on(image, dict(polygon=None, strokes=[(*p[k], *p[k + 1], tuple(color), thickness) for k in range(len(p) - 1)], flip=False))


def rasterize_polyline(polyline_xy: np.ndarray, bev_img: np.ndarray, bevimg_Sim2_world: Sim2, color, thickness: int) -> np.ndarray:
    """:200-217"""
    img_h, img_w, _ = bev_img.shape
    draw_polyline_cv2(line_segments_arr=_to_px(polyline_xy, bevimg_Sim2_world), image=bev_img, color=color, im_h=img_h, im_w=img_w, thickness=thickness)
    return bev_img
