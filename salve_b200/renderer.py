"""BevRenderer: thin object wrapper over the C ABI (one context = one GPU, one stream of work).

All arithmetic runs in libsalve_bev.so on the GPU.  This file only marshals numpy arrays /
raw device pointers across ctypes.
"""

from __future__ import annotations

import ctypes
import threading
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _native as nat

SURF_FLOOR = 1
SURF_CEILING = 2
SURFACE_BITS = {"floor": SURF_FLOOR, "ceiling": SURF_CEILING}
NCOUNTS = 8
IMG_OK, IMG_EMPTY, IMG_DEGENERATE, IMG_COLLINEAR = 0, 1, 2, 3
TAP = dict(keygrid=0, color=1, occ=2, nonempty=3, keep=4, tris=5, interp=6, hull=7, qtri=8)


def numpy_sphere_tables(H: int, W: int):
    """Separable factors of the unit sphere with numpy's own trig, so that the GPU reproduces
    get_uni_sphere_xyz (reference salve/utils/hohonet_pano_utils.py:27-43) bit-for-bit."""
    u = np.arange(W)
    v = np.arange(H)
    theta = -(u + 0.5) / W
    theta *= 2 * np.pi
    phi = (v + 0.5) / H
    phi -= 0.5
    phi *= np.pi
    return (
        np.ascontiguousarray(np.cos(phi)),
        np.ascontiguousarray(-np.sin(phi)),
        np.ascontiguousarray(np.cos(theta)),
        np.ascontiguousarray(np.sin(theta)),
    )


def _ptr(a: Optional[np.ndarray], ctype):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.POINTER(ctype))


def _vp(x) -> Optional[int]:
    """numpy array / torch tensor / int -> raw address."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(type(x))


def _locked(fn):
    """Serialise calls on one renderer (a context is not thread-safe); the lock is re-entrant, so a caller may hold it around a sequence."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **kw):
        with self.lock:
            return fn(self, *a, **kw)

    return wrapper


class BevRenderer:
    def __init__(
        self,
        pano_h: int = 512,
        pano_w: int = 1024,
        max_panos: int = 64,
        max_images: int = 592,
        device: int = 0,
        grid_h: int = 501,
        grid_w: int = 501,
        xlims: Tuple[float, float] = (-5.0, 5.0),
        ylims: Tuple[float, float] = (-5.0, 5.0),
        px_per_m: float = 1 / 0.02,
        kernel_sz: int = 11,
        crop_ratio: float = 80 / 512,
        depth_scale: float = 0.001,
        numpy_tables: bool = True,
    ) -> None:
        self._lib = nat.load()
        self.lock = threading.RLock()  # a context is not thread-safe: callers that share a renderer hold this around a sequence of calls
        cfg = nat.Config()
        self._lib.salve_bev_default_config(ctypes.byref(cfg), pano_h, pano_w)
        cfg.device = device
        cfg.max_panos = max_panos
        cfg.max_images = max_images
        cfg.grid_h, cfg.grid_w = grid_h, grid_w
        cfg.kernel_sz = kernel_sz
        cfg.xmin, cfg.xmax = float(xlims[0]), float(xlims[1])
        cfg.ymin, cfg.ymax = float(ylims[0]), float(ylims[1])
        cfg.px_per_m = float(px_per_m)
        cfg.depth_scale = float(np.float32(depth_scale))
        cfg.crop_rows = int(pano_h * crop_ratio)
        self.cfg = cfg
        self.pano_h, self.pano_w = pano_h, pano_w
        self.grid_h, self.grid_w = grid_h, grid_w
        self.img_shape = (grid_h, grid_w, 3)
        self.max_images = max_images
        h = nat.c_vp()
        nat.check(self._lib.salve_bev_ctx_create(ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h
        if numpy_tables:
            t = numpy_sphere_tables(pano_h, pano_w)
            nat.check(self._lib.salve_bev_set_sphere_tables(self._h, *[_ptr(a, ctypes.c_double) for a in t]))

    # -- lifetime -----------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.salve_bev_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- panos -----------------------------------------------------------------------------------
    @_locked
    def upload_pano(self, slot: int, rgb: np.ndarray, depth: np.ndarray, stream: int = 0) -> None:
        rgb = np.ascontiguousarray(rgb, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        if rgb.shape != (self.pano_h, self.pano_w, 3) or depth.shape != (self.pano_h, self.pano_w):
            raise ValueError(f"pano must be ({self.pano_h},{self.pano_w},3) uint8 + ({self.pano_h},{self.pano_w}) uint16")
        nat.check(self._lib.salve_bev_upload_pano(self._h, slot, rgb.ctypes.data, depth.ctypes.data, stream or None))

    @_locked
    def upload_pano_ptr(self, slot: int, rgb_ptr: int, depth_ptr: int, stream: int = 0) -> None:
        """Host pointers (e.g. pinned torch tensors): asynchronous H2D on `stream`."""
        nat.check(self._lib.salve_bev_upload_pano(self._h, slot, rgb_ptr, depth_ptr, stream or None))

    @_locked
    def upload_pano_fullres(self, slot: int, rgb_2x: np.ndarray, depth: np.ndarray, stream: int = 0) -> None:
        """Full-resolution colour: rgb_2x (2H, 2W, 3) uint8 with the (H, W) depth map.  Equivalent to uploading
        cv2.resize(rgb_2x, (W, H), INTER_LINEAR) (reference bev_rendering_utils.py:373-375); the 2x2 mean is fused into the colour gather."""
        rgb_2x = np.ascontiguousarray(rgb_2x, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        if rgb_2x.shape != (2 * self.pano_h, 2 * self.pano_w, 3) or depth.shape != (self.pano_h, self.pano_w):
            raise ValueError(f"need rgb ({2 * self.pano_h},{2 * self.pano_w},3) uint8 + depth ({self.pano_h},{self.pano_w}) uint16")
        nat.check(self._lib.salve_bev_upload_pano_fullres(self._h, slot, rgb_2x.ctypes.data, depth.ctypes.data, stream or None))

    @_locked
    def bind_pano_fullres(self, slot: int, dev_rgb_2x, dev_depth) -> None:
        nat.check(self._lib.salve_bev_bind_pano_fullres(self._h, slot, _vp(dev_rgb_2x), _vp(dev_depth)))

    @_locked
    def bind_pano(self, slot: int, dev_rgb, dev_depth) -> None:
        nat.check(self._lib.salve_bev_bind_pano(self._h, slot, _vp(dev_rgb), _vp(dev_depth)))

    # -- rendering ---------------------------------------------------------------------------------
    @staticmethod
    def _surf_mask(surfaces: Sequence[str]) -> int:
        m = 0
        for s in surfaces:
            m |= SURFACE_BITS[s]
        return m

    def _hyp_args(self, pano1, pano2, R, t):
        p1 = np.ascontiguousarray(pano1, np.int32).reshape(-1)
        p2 = np.ascontiguousarray(pano2, np.int32).reshape(-1)
        n = p1.shape[0]
        R = np.ascontiguousarray(R, np.float32).reshape(n, 4)
        t = np.ascontiguousarray(t, np.float32).reshape(n, 2)
        return n, p1, p2, R, t

    @_locked
    def render_hypotheses(self, pano1, pano2, R, t, surfaces=("floor", "ceiling"), out: Optional[np.ndarray] = None, stream: int = 0):
        """Host-output render.  Returns (images (n, nsurf, 2, gh, gw, 3) u8, counts (n,nsurf,2,8), status (n,nsurf,2))."""
        n, p1, p2, R, t = self._hyp_args(pano1, pano2, R, t)
        mask = self._surf_mask(surfaces)
        nsurf = bin(mask).count("1")
        shape = (n, nsurf, 2) + self.img_shape
        if out is None:
            out = np.empty(shape, np.uint8)
        assert out.dtype == np.uint8 and out.size == int(np.prod(shape)) and out.flags.c_contiguous
        counts = np.zeros((n, nsurf, 2, NCOUNTS), np.int32)
        status = np.zeros((n, nsurf, 2), np.int32)
        nat.check(
            self._lib.salve_bev_render_hypotheses_host(
                self._h, n, _ptr(p1, ctypes.c_int32), _ptr(p2, ctypes.c_int32), _ptr(R, ctypes.c_float), _ptr(t, ctypes.c_float),
                mask, out.ctypes.data, counts.ctypes.data, status.ctypes.data, stream or None,
            )
        )
        return out.reshape(shape), counts, status

    @_locked
    def render_hypotheses_device(self, pano1, pano2, R, t, dev_out, dev_counts=None, dev_status=None, surfaces=("floor", "ceiling"), stream: int = 0) -> int:
        """Device-output render (asynchronous on `stream`).  dev_* are device pointers / torch CUDA tensors.
        Returns the number of images written."""
        n, p1, p2, R, t = self._hyp_args(pano1, pano2, R, t)
        mask = self._surf_mask(surfaces)
        nat.check(
            self._lib.salve_bev_render_hypotheses(
                self._h, n, _ptr(p1, ctypes.c_int32), _ptr(p2, ctypes.c_int32), _ptr(R, ctypes.c_float), _ptr(t, ctypes.c_float),
                mask, _vp(dev_out), _vp(dev_counts), _vp(dev_status), stream or None,
            )
        )
        return n * bin(mask).count("1") * 2

    @_locked
    def render_hypotheses_compact(self, pano1, pano2, R, t, surfaces=("floor", "ceiling"), posed_out: Optional[np.ndarray] = None,
                                  unposed_out: Optional[np.ndarray] = None, stream: int = 0):
        """Host-output render without duplicates: img2 of a pair does not depend on the hypothesis
        (reference bev_rendering_utils.py:451-455), so every distinct (pano 2, surface) is rendered and copied out once.
        Returns (posed (n, nsurf, gh, gw, 3), unposed (n_unique, nsurf, gh, gw, 3), unposed_of_hyp (n,),
        counts_posed (n, nsurf, 8), counts_unposed (n_unique, nsurf, 8), status_posed (n, nsurf), status_unposed (n_unique, nsurf)).
        Hypothesis h, surface s: img1 = posed[h, s], img2 = unposed[unposed_of_hyp[h], s]."""
        n, p1, p2, R, t = self._hyp_args(pano1, pano2, R, t)
        mask = self._surf_mask(surfaces)
        nsurf = bin(mask).count("1")
        cap = min(n, int(self.cfg.max_panos))
        if posed_out is None:
            posed_out = np.empty((n, nsurf) + self.img_shape, np.uint8)
        if unposed_out is None:
            unposed_out = np.empty((cap, nsurf) + self.img_shape, np.uint8)
        assert posed_out.dtype == np.uint8 and posed_out.size == n * nsurf * int(np.prod(self.img_shape)) and posed_out.flags.c_contiguous
        assert unposed_out.dtype == np.uint8 and unposed_out.size >= cap * nsurf * int(np.prod(self.img_shape)) and unposed_out.flags.c_contiguous
        idx = np.zeros(n, np.int32)
        nu = ctypes.c_int32(0)
        cp = np.zeros((n, nsurf, NCOUNTS), np.int32)
        cu = np.zeros((cap, nsurf, NCOUNTS), np.int32)
        sp = np.zeros((n, nsurf), np.int32)
        su = np.zeros((cap, nsurf), np.int32)
        nat.check(
            self._lib.salve_bev_render_hypotheses_compact_host(
                self._h, n, _ptr(p1, ctypes.c_int32), _ptr(p2, ctypes.c_int32), _ptr(R, ctypes.c_float), _ptr(t, ctypes.c_float), mask,
                posed_out.ctypes.data, unposed_out.ctypes.data, _ptr(idx, ctypes.c_int32), ctypes.byref(nu),
                cp.ctypes.data, cu.ctypes.data, sp.ctypes.data, su.ctypes.data, stream or None,
            )
        )
        k = nu.value
        posed = posed_out.reshape((n, nsurf) + self.img_shape)
        unposed = unposed_out.reshape(-1)[: k * nsurf * int(np.prod(self.img_shape))].reshape((k, nsurf) + self.img_shape)
        return posed, unposed, idx, cp, cu[:k], sp, su[:k]

    @_locked
    def render_hypotheses_compact_device(self, pano1, pano2, R, t, dev_posed, dev_unposed, dev_counts_posed=None, dev_counts_unposed=None,
                                         dev_status_posed=None, dev_status_unposed=None, surfaces=("floor", "ceiling"), stream: int = 0):
        """Device-output variant (asynchronous on `stream`).  Returns (unposed_of_hyp (n,) int32, n_unique)."""
        n, p1, p2, R, t = self._hyp_args(pano1, pano2, R, t)
        mask = self._surf_mask(surfaces)
        idx = np.zeros(n, np.int32)
        nu = ctypes.c_int32(0)
        nat.check(
            self._lib.salve_bev_render_hypotheses_compact(
                self._h, n, _ptr(p1, ctypes.c_int32), _ptr(p2, ctypes.c_int32), _ptr(R, ctypes.c_float), _ptr(t, ctypes.c_float), mask,
                _vp(dev_posed), _vp(dev_unposed), _ptr(idx, ctypes.c_int32), ctypes.byref(nu),
                _vp(dev_counts_posed), _vp(dev_counts_unposed), _vp(dev_status_posed), _vp(dev_status_unposed), stream or None,
            )
        )
        return idx, nu.value

    # -- verifier pre-processing ---------------------------------------------------------------------
    def quadruplet_pointers_full(self, dev_out, n: int) -> np.ndarray:
        """(n, 4) device addresses x1c, x2c, x1f, x2f into a full-layout device buffer of render_hypotheses_device
        (floor + ceiling): per hypothesis the images are [floor img1, floor img2, ceiling img1, ceiling img2]."""
        ib = int(np.prod(self.img_shape))
        base = np.uint64(_vp(dev_out)) + np.arange(n, dtype=np.uint64)[:, None] * np.uint64(4 * ib)
        return base + np.array([2, 3, 0, 1], np.uint64)[None, :] * np.uint64(ib)

    def quadruplet_pointers_compact(self, dev_posed, dev_unposed, unposed_of_hyp: np.ndarray) -> np.ndarray:
        """Same for the compact layout of render_hypotheses_compact_device (floor + ceiling)."""
        ib = int(np.prod(self.img_shape))
        n = len(unposed_of_hyp)
        h = np.arange(n, dtype=np.uint64)
        u = np.asarray(unposed_of_hyp, np.uint64)
        p, q = np.uint64(_vp(dev_posed)), np.uint64(_vp(dev_unposed))
        two = np.uint64(2)
        return np.stack([p + (h * two + 1) * np.uint64(ib), q + (u * two + 1) * np.uint64(ib), p + (h * two) * np.uint64(ib), q + (u * two) * np.uint64(ib)], 1)

    @_locked
    def verifier_preprocess(self, src_ptrs: np.ndarray, dev_out, resize_hw: int = 234, crop_hw: int = 224, stream: int = 0) -> None:
        """Fused val/test transform of the reference (resize 234 -> centre crop 224 -> CHW float32 -> ImageNet normalise -> channel
        concatenation; salve/train_utils.py:126-159).  src_ptrs: (n, 4) uint64 device addresses of 501x501x3 uint8 renders in the
        model's order x1c, x2c, x1f, x2f; dev_out: device float32 (n, 12, crop, crop).  Asynchronous on `stream`."""
        src = np.ascontiguousarray(src_ptrs, np.uint64).reshape(-1, 4)
        nat.check(self._lib.salve_bev_verifier_preprocess(self._h, src.shape[0], src.ctypes.data, int(resize_hw), int(crop_hw), _vp(dev_out), stream or None))

    @_locked
    def rasterize_layouts(self, layouts, init: Optional[np.ndarray] = None, stream: int = 0) -> np.ndarray:
        """Layout modality.  layouts: one dict per image with `polygon` ((n, 2) integer pixel vertices or None), `polygon_rgb`,
        `strokes` (list of (x0, y0, x1, y1, (r, g, b), thickness)) and `flip`.  Returns (n, gh, gw, 3) uint8."""
        words, offs = [], [0]
        for L in layouts:
            poly = np.zeros((0, 2), np.int64) if L.get("polygon") is None else np.asarray(L["polygon"], np.int64).reshape(-1, 2)
            r, g, b = (int(v) & 0xFF for v in L.get("polygon_rgb", (255, 255, 255)))
            strokes = L.get("strokes", [])
            w = [poly.shape[0], len(strokes), r | (g << 8) | (b << 16), int(bool(L.get("flip", False)))]
            w += [int(v) for v in np.clip(poly, -(1 << 20), 1 << 20).reshape(-1)]
            for x0, y0, x1, y1, col, th in strokes:
                cr, cg, cb = (int(v) & 0xFF for v in col)
                w += [int(np.clip(v, -(1 << 20), 1 << 20)) for v in (x0, y0, x1, y1)] + [cr | (cg << 8) | (cb << 16), int(th)]
            words += w
            offs.append(len(words))
        desc = np.ascontiguousarray(words, np.int32)
        off = np.ascontiguousarray(offs, np.int64)
        n = len(layouts)
        out = np.empty((n,) + self.img_shape, np.uint8)
        if init is not None:
            init = np.ascontiguousarray(init, np.uint8)
            assert init.shape == out.shape
        nat.check(self._lib.salve_bev_rasterize_layouts_host(self._h, n, _ptr(desc, ctypes.c_int32), _ptr(off, ctypes.c_int64),
                                                              None if init is None else init.ctypes.data, out.ctypes.data, stream or None))
        return out

    @_locked
    def set_dedup_unposed(self, on: bool) -> None:
        nat.check(self._lib.salve_bev_set_dedup_unposed(self._h, int(on)))

    @_locked
    def render_images(self, slots, surfaces: Sequence[str], posed, R, t, stream: int = 0):
        """Individual images.  Returns (images (n, gh, gw, 3), counts (n,8), status (n,))."""
        slots = np.ascontiguousarray(slots, np.int32).reshape(-1)
        n = slots.shape[0]
        surf = np.ascontiguousarray([SURFACE_BITS[s] for s in surfaces], np.int32)
        posed = np.ascontiguousarray(posed, np.int32).reshape(n)
        R = np.ascontiguousarray(R, np.float32).reshape(n, 4)
        t = np.ascontiguousarray(t, np.float32).reshape(n, 2)
        out = np.empty((n,) + self.img_shape, np.uint8)
        counts = np.zeros((n, NCOUNTS), np.int32)
        status = np.zeros(n, np.int32)
        nat.check(
            self._lib.salve_bev_render_images_host(
                self._h, n, _ptr(slots, ctypes.c_int32), _ptr(surf, ctypes.c_int32), _ptr(posed, ctypes.c_int32),
                _ptr(R, ctypes.c_float), _ptr(t, ctypes.c_float), out.ctypes.data, counts.ctypes.data, status.ctypes.data, stream or None,
            )
        )
        return out, counts, status

    @_locked
    def set_bands(self, a=(-float("inf"), -1.0), b=(0.5, float("inf"))) -> None:
        nat.check(self._lib.salve_bev_set_bands(self._h, float(a[0]), float(a[1]), float(b[0]), float(b[1])))

    @_locked
    def backproject(self, slot: int, z_lo: float, z_hi: float, frame: int = 0, R=None, t=None, stream: int = 0) -> np.ndarray:
        """frame 0: HoHoNet frame; 1: ZInD frame; 2: posed by (R, t) into pano 2's frame."""
        n = ctypes.c_int64(0)
        R = None if R is None else np.ascontiguousarray(R, np.float32).reshape(4)
        t = None if t is None else np.ascontiguousarray(t, np.float32).reshape(2)
        a = (self._h, slot, float(z_lo), float(z_hi), int(frame), _ptr(R, ctypes.c_float), _ptr(t, ctypes.c_float))
        nat.check(self._lib.salve_bev_backproject(*a, None, ctypes.byref(n), stream or None))
        out = np.empty((n.value, 6), np.float64)
        if n.value:
            nat.check(self._lib.salve_bev_backproject(*a, _ptr(out, ctypes.c_double), ctypes.byref(n), stream or None))
        return out

    @_locked
    def render_cloud(self, xyzrgb: np.ndarray, stream: int = 0):
        xyzrgb = np.ascontiguousarray(xyzrgb, np.float64)
        if xyzrgb.ndim != 2 or xyzrgb.shape[1] != 6:
            raise ValueError("xyzrgb must have shape (N,6)")
        out = np.empty(self.img_shape, np.uint8)
        counts = np.zeros(NCOUNTS, np.int32)
        status = np.zeros(1, np.int32)
        nat.check(
            self._lib.salve_bev_render_cloud_host(
                self._h, _ptr(xyzrgb, ctypes.c_double), xyzrgb.shape[0], _ptr(out, ctypes.c_uint8), _ptr(counts, ctypes.c_int32),
                _ptr(status, ctypes.c_int32), stream or None,
            )
        )
        return out, counts, int(status[0])

    @_locked
    def choose_elevated(self, x, y, z, zmin: float, zmax: float, num_slices: int) -> np.ndarray:
        x = np.ascontiguousarray(x, np.int64).reshape(-1)
        y = np.ascontiguousarray(y, np.int64).reshape(-1)
        z = np.ascontiguousarray(z, np.float64).reshape(-1)
        valid = np.zeros(x.shape[0], np.uint8)
        nat.check(
            self._lib.salve_bev_choose_elevated(
                self._h, _ptr(x, ctypes.c_int64), _ptr(y, ctypes.c_int64), _ptr(z, ctypes.c_double), x.shape[0], float(zmin), float(zmax),
                int(num_slices), _ptr(valid, ctypes.c_uint8), None,
            )
        )
        return valid.astype(bool)

    @_locked
    def interp_dense(self, points_xy: np.ndarray, values: np.ndarray, grid_h: int, grid_w: int, want_hull: bool = False):
        """Returns (img or None if degenerate, hull or None, status)."""
        pts = np.ascontiguousarray(points_xy, np.int64).reshape(-1, 2)
        vals = np.ascontiguousarray(values, np.float64).reshape(-1, 3)
        img = np.zeros((grid_h, grid_w, 3), np.uint8)
        hull = np.zeros((grid_h, grid_w), np.uint8) if want_hull else None
        status = ctypes.c_int32(0)
        nat.check(
            self._lib.salve_bev_interp_dense(
                self._h, _ptr(pts, ctypes.c_int64), _ptr(vals, ctypes.c_double), pts.shape[0], grid_h, grid_w, _ptr(img, ctypes.c_uint8),
                _ptr(hull, ctypes.c_uint8), ctypes.byref(status), None,
            )
        )
        if status.value == IMG_DEGENERATE:
            return None, None, status.value
        return img, (hull.astype(bool) if want_hull else None), status.value

    @_locked
    def remove_hallucinated(self, sparse: np.ndarray, interp: np.ndarray, K: int) -> np.ndarray:
        sparse = np.ascontiguousarray(sparse, np.uint8)
        interp = np.ascontiguousarray(interp, np.uint8)
        h, w, _ = interp.shape
        out = np.empty((h, w, 3), np.uint8)
        nat.check(self._lib.salve_bev_remove_hallucinated(self._h, _ptr(sparse, ctypes.c_uint8), _ptr(interp, ctypes.c_uint8), h, w, int(K), _ptr(out, ctypes.c_uint8), None))
        return out

    @_locked
    def uni_sphere_xyz(self) -> np.ndarray:
        out = np.empty((self.pano_h, self.pano_w, 3), np.float64)
        nat.check(self._lib.salve_bev_get_uni_sphere_xyz(self._h, _ptr(out, ctypes.c_double)))
        return out

    # -- taps / diagnostics ------------------------------------------------------------------------------
    @_locked
    def tap(self, image: int, what: str) -> np.ndarray:
        g = self.grid_h * self.grid_w
        wpr = (self.grid_w + 31) // 32
        if what in ("keygrid", "color"):
            buf = np.zeros((self.grid_h, self.grid_w), np.uint32)
        elif what in ("occ", "nonempty", "keep"):
            buf = np.zeros((self.grid_h, wpr), np.uint32)
        elif what == "tris":
            buf = np.full((2 * g, 3), -2, np.int32)
        elif what == "interp":
            buf = np.zeros((self.grid_h, self.grid_w, 3), np.uint8)
        elif what == "hull":
            buf = np.zeros((self.grid_h, self.grid_w), np.uint8)
        elif what == "qtri":
            buf = np.full((self.grid_h, self.grid_w, 3), -1, np.int32)
        else:
            raise ValueError(what)
        nat.check(self._lib.salve_bev_tap(self._h, image, TAP[what], buf.ctypes.data, buf.nbytes, None))
        if what in ("occ", "nonempty", "keep"):
            bits = np.unpackbits(buf.view(np.uint8).reshape(self.grid_h, wpr * 4), axis=1, bitorder="little")
            return bits[:, : self.grid_w].astype(bool)
        if what == "tris":
            return buf[(buf != -2).all(1)]
        if what == "hull":
            return buf.astype(bool)
        return buf

    @_locked
    def enable_timing(self, on: bool = True) -> None:
        nat.check(self._lib.salve_bev_enable_timing(self._h, int(on)))

    @_locked
    def last_timings(self) -> dict:
        """Per-stage device time (ms) of the last pano render call: splat, the six stages of the image pipeline, `image` (their sum)
        and `total`."""
        ms = np.zeros(8, np.float32)
        nat.check(self._lib.salve_bev_last_timings(self._h, _ptr(ms, ctypes.c_float)))
        d = dict(splat=float(ms[0]), sites=float(ms[1]), prep=float(ms[2]), local=float(ms[6]), window=float(ms[3]), shade=float(ms[4]),
                 finish=float(ms[5]), total=float(ms[7]))
        d["image"] = d["sites"] + d["prep"] + d["local"] + d["window"] + d["shade"] + d["finish"]
        return d

    def launch_count(self) -> int:
        return int(self._lib.salve_bev_launch_count(self._h))
