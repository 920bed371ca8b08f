"""Batched file driver: the B200 replacement of scripts/render_dataset_bev.py's per-floor fan-out.

The reference enumerates, per building floor, every alignment-hypothesis JSON of the label types
`gt_alignment_approx` and `incorrect_alignment` times {floor, ceiling} and maps
`generate_texture_maps_for_pair` over a multiprocessing.Pool (scripts/render_dataset_bev.py:34-117).  Here the
same enumeration feeds ONE batched render per floor: every pano is read and uploaded once (at full resolution when it
is 2048x1024: the 2x2 mean of bev_rendering_utils.py:373-375 is fused into the colour gather), every distinct
(pano 2, surface) is rendered once, and the JPEG tree written is the one the reference writes
(bev_rendering_utils.py:582-595, 619-630), so scripts/test.py and ZindData run unchanged on it.

File naming, skip-if-exists (:619-621) and the (None, None) rule for empty clouds (:457-458, 626-627) follow the
reference.  Depth maps must already exist under `{depth_save_root}/{building_id}/{pano stem}.depth.png`
(HoHoNet inference is out of scope).  No arithmetic of the path happens in this file.
"""

from __future__ import annotations

import glob
import os
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .common.sim2 import Sim2
from .renderer import IMG_COLLINEAR, IMG_EMPTY, BevRenderer
from .utils import bev_rendering_utils as bru

LABEL_TYPES = ("gt_alignment_approx", "incorrect_alignment")  # scripts/render_dataset_bev.py:87
SURFACES = ("floor", "ceiling")  # :92
PANO_H, PANO_W = 512, 1024


def panoid_from_fpath(fpath: str) -> int:
    """Derive a panorama's id from its file name (scripts/render_dataset_bev.py:29-31)."""
    return int(Path(fpath).stem.split("_")[-1])


def enumerate_floor_hypotheses(hypotheses_save_root: str, building_id: str, floor_id: str) -> List[Tuple[str, int, str]]:
    """(label_type, pair_idx, pair_fpath) in the reference's order (:87-91): pair_idx restarts per label type."""
    out = []
    for label_type in LABEL_TYPES:
        pairs = sorted(glob.glob(f"{hypotheses_save_root}/{building_id}/{floor_id}/{label_type}/*.json"))
        out += [(label_type, k, p) for k, p in enumerate(pairs)]
    return out


def _load_pano_any(depth_fpath: str, rgb_fpath: str):
    """(rgb, depth, fullres).  rgb stays at (1024, 2048) when it is exactly twice the depth map (the renderer averages 2x2 blocks
    on the fly); any other size goes through cv2.resize like the reference (:373-375)."""
    depth = np.ascontiguousarray(bru._imread(depth_fpath), np.uint16)
    rgb = bru._imread(rgb_fpath)
    if rgb.ndim == 2:
        rgb = np.repeat(rgb[:, :, None], 3, axis=2)
    rgb = np.ascontiguousarray(rgb[:, :, :3], np.uint8)
    if depth.shape != (PANO_H, PANO_W):
        raise ValueError(f"depth map must be {PANO_H}x{PANO_W}, got {depth.shape}")
    if rgb.shape[:2] == (2 * PANO_H, 2 * PANO_W):
        return rgb, depth, True
    if rgb.shape[:2] != (PANO_H, PANO_W):
        import cv2

        rgb = np.ascontiguousarray(cv2.resize(rgb, (PANO_W, PANO_H), interpolation=cv2.INTER_LINEAR))
    return rgb, depth, False


def render_building_floor_pairs(
    depth_save_root: str,
    bev_save_root: str,
    hypotheses_save_root: str,
    raw_dataset_dir: str,
    building_id: str,
    floor_id: str,
    layout_save_root: Optional[str] = None,
    render_modalities: Sequence[str] = ("rgb_texture",),
    multiprocess_building_panos: bool = False,
    num_processes: int = 1,
    renderer: Optional[BevRenderer] = None,
    write_threads: int = 8,
    batch_hypotheses: int = 256,
) -> Dict[str, int]:
    """Same arguments as the reference's render_building_floor_pairs (scripts/render_dataset_bev.py:34-117);
    `multiprocess_building_panos` / `num_processes` are accepted and ignored (the batch replaces the pool).
    The floor is rendered in batches of `batch_hypotheses`.  Returns counters {hypotheses, rendered, skipped_existing, skipped_empty, files_written}."""
    if "layout" in render_modalities:
        raise NotImplementedError("the batched driver renders texture maps; for the layout modality call generate_texture_maps_for_pair with a "
                                  "floor pose graph (rasterize_room_layout_pair), as scripts/render_dataset_bev.py:64-117 does")
    stats = dict(hypotheses=0, rendered=0, skipped_existing=0, skipped_empty=0, files_written=0)
    if "rgb_texture" not in render_modalities:
        return stats
    img_fpaths = glob.glob(f"{raw_dataset_dir}/{building_id}/panos/*.jpg") + glob.glob(f"{raw_dataset_dir}/{building_id}/panos/*.png")
    img_fpaths_dict = {panoid_from_fpath(f): f for f in img_fpaths}
    hyps = enumerate_floor_hypotheses(hypotheses_save_root, building_id, floor_id)
    stats["hypotheses"] = len(hyps)
    todo = []  # (label_type, pair_idx, uuid, i1, i2, Sim2, {surface: (path1, path2)})
    for label_type, pair_idx, pair_fpath in hyps:
        if pair_idx == 0:
            print(f"On Building {building_id}, {floor_id}, {label_type}")
        stem = Path(pair_fpath).stem
        i1, i2 = (int(v) for v in stem.split("_")[:2])
        uuid = stem.split("__")[-1]
        save_dir = f"{bev_save_root}/{label_type}/{building_id}"
        out = {}
        for s in SURFACES:
            p1 = f"{save_dir}/{bru.bev_fname_from_img_fpath(pair_idx, uuid, s, img_fpaths_dict[i1])}"
            p2 = f"{save_dir}/{bru.bev_fname_from_img_fpath(pair_idx, uuid, s, img_fpaths_dict[i2])}"
            if Path(p1).exists() and Path(p2).exists():  # bev_rendering_utils.py:619-621
                continue
            out[s] = (p1, p2)
        if not out:
            stats["skipped_existing"] += 1
            continue
        os.makedirs(save_dir, exist_ok=True)
        todo.append((label_type, pair_idx, uuid, i1, i2, Sim2.from_json(pair_fpath), out))
    if not todo:
        return stats
    pano_ids = sorted({t[3] for t in todo} | {t[4] for t in todo})
    slot = {pid: k for k, pid in enumerate(pano_ids)}
    own = renderer is None
    r = renderer or BevRenderer(pano_h=PANO_H, pano_w=PANO_W, max_panos=max(len(pano_ids), 1), max_images=592)
    try:
        for pid in pano_ids:
            f = img_fpaths_dict[pid]
            rgb, depth, full = _load_pano_any(f"{depth_save_root}/{building_id}/{Path(f).stem}.depth.png", f)
            (r.upload_pano_fullres if full else r.upload_pano)(slot[pid], rgb, depth)
        # Hypotheses go through in batches with re-used output buffers: a ZInD floor can have ~10 k hypotheses (15 GB of renders), and a
        # batch's files are on disk before the next one is rendered, so that nothing is lost if a later batch fails.
        img_shape = (len(SURFACES),) + r.img_shape
        posed_buf = np.empty((min(batch_hypotheses, len(todo)),) + img_shape, np.uint8)
        unposed_buf = np.empty((min(batch_hypotheses, len(todo), int(r.cfg.max_panos)),) + img_shape, np.uint8)
        collinear = []
        with ThreadPoolExecutor(max(write_threads, 1)) as pool:
            for b0 in range(0, len(todo), batch_hypotheses):
                batch = todo[b0:b0 + batch_hypotheses]
                p1 = [slot[t[3]] for t in batch]
                p2 = [slot[t[4]] for t in batch]
                R = np.stack([t[5].rotation for t in batch]).astype(np.float32)
                tt = np.stack([t[5].translation for t in batch]).astype(np.float32)
                posed, unposed, idx, cp, cu, sp, su = r.render_hypotheses_compact(
                    p1, p2, R, tt, surfaces=SURFACES, posed_out=posed_buf[: len(batch)], unposed_out=unposed_buf)
                futs = []
                for h, t in enumerate(batch):
                    for si, s in enumerate(SURFACES):
                        if s not in t[6]:
                            continue
                        if sp[h, si] == IMG_EMPTY or su[idx[h], si] == IMG_EMPTY:  # (None, None): nothing is written (:626-627)
                            stats["skipped_empty"] += 1
                            continue
                        if sp[h, si] == IMG_COLLINEAR or su[idx[h], si] == IMG_COLLINEAR:
                            collinear.append((t[0], t[1], s))  # the reference's Qhull call raises for this pair only
                            continue
                        f1, f2 = t[6][s]
                        futs.append(pool.submit(bru._imwrite, f1, posed[h, si]))
                        futs.append(pool.submit(bru._imwrite, f2, unposed[idx[h], si]))
                        stats["rendered"] += 1
                for f in futs:  # the buffers are re-used by the next batch
                    f.result()
                stats["files_written"] += len(futs)
        if collinear:
            from .utils.interpolation_utils import QhullError

            raise QhullError(f"initial simplex is flat (all sites collinear) for {len(collinear)} render(s), first: {collinear[0]}; "
                             "every other pair of the floor has been written")
    finally:
        if own:
            r.close()
    return stats


def render_pairs(depth_save_root: str, bev_save_root: str, raw_dataset_dir: str, hypotheses_save_root: str, building_ids: Sequence[str],
                 render_modalities: Sequence[str] = ("rgb_texture",), rank: int = 0, world_size: int = 1) -> Dict[str, int]:
    """All floors of the given buildings (scripts/render_dataset_bev.py:120-191).  Floors are the sub-directories of
    `{hypotheses_save_root}/{building_id}`.  With world_size > 1 buildings are dealt to ranks by hypothesis count
    (salve_b200.sharding.assign_buildings): one process per GPU, no collective on the data path."""
    from .sharding import assign_buildings

    ids = sorted(b for b in building_ids if b != "1348")  # duplicate pano id in ZInD (:166-168)
    floors = {b: sorted(os.path.basename(d) for d in glob.glob(f"{hypotheses_save_root}/{b}/*") if os.path.isdir(d)) for b in ids}
    counts = [sum(len(enumerate_floor_hypotheses(hypotheses_save_root, b, f)) for f in floors[b]) for b in ids]
    mine = assign_buildings(counts, world_size)[rank]
    total = dict(hypotheses=0, rendered=0, skipped_existing=0, skipped_empty=0, files_written=0)
    for bi in mine:
        for f in floors[ids[bi]]:
            st = render_building_floor_pairs(depth_save_root, bev_save_root, hypotheses_save_root, raw_dataset_dir, ids[bi], f,
                                             render_modalities=render_modalities)
            for k in total:
                total[k] += st[k]
    return total
