"""ctypes wrapper of oracle/canonical_dt.c (CPU checker for the densification stage).

TEST INFRASTRUCTURE ONLY -- see the header of canonical_dt.c.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcanonical_dt.so")
_SRC = os.path.join(_HERE, "canonical_dt.c")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, _SRC])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        _lib.cdt_triangulate.restype = ctypes.c_int
        _lib.cdt_triangulate.argtypes = [ctypes.c_int, ctypes.c_int, i32p, i32p, i32p, i32p, i64p]
        _lib.cdt_check_delaunay.restype = ctypes.c_int64
        _lib.cdt_check_delaunay.argtypes = [ctypes.c_int, i32p, i32p, ctypes.c_int, i32p, i64p, u8p, i64p]
        _lib.cdt_rasterize.restype = None
        _lib.cdt_rasterize.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, i32p, i32p, u8p, ctypes.c_int, i32p, u8p, u8p, i32p]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def sort_sites(site_rc: np.ndarray, site_rgb: np.ndarray):
    """Row-major order (row, then col) -- the order the CUDA path compacts sites in."""
    order = np.lexsort((site_rc[:, 1], site_rc[:, 0]))
    return np.ascontiguousarray(site_rc[order]).astype(np.int32), np.ascontiguousarray(site_rgb[order]).astype(np.uint8)


def triangulate(row: np.ndarray, col: np.ndarray, img_w: int):
    """Canonical Delaunay of row-major-sorted sites.  Returns (tri_v (2S-2,3) i32 with -1 = ghost, stats dict)."""
    row = np.ascontiguousarray(row, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    S = row.shape[0]
    tri_v = np.empty((2 * S, 3), np.int32)
    tri_n = np.empty((2 * S, 3), np.int32)
    stats = np.zeros(4, np.int64)
    nt = lib().cdt_triangulate(S, img_w, _p(row, ctypes.c_int32), _p(col, ctypes.c_int32), _p(tri_v, ctypes.c_int32),
                               _p(tri_n, ctypes.c_int32), _p(stats, ctypes.c_int64))
    if nt < 0:
        raise ValueError(f"cdt_triangulate failed: {nt}")
    return tri_v[:nt].copy(), dict(flips=int(stats[0]), init_check=int(stats[1]), final_check=int(stats[2]), residual_ties=int(stats[3]))


def check_delaunay(row, col, tri_v):
    """Exact validity check of any triangulation.  Returns dict(violations, n_strict, strict_flag, hull_edges)."""
    row = np.ascontiguousarray(row, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    tri_v = np.ascontiguousarray(tri_v, np.int32)
    nt = tri_v.shape[0]
    flag = np.zeros(nt, np.uint8)
    ns = ctypes.c_int64(0)
    un = ctypes.c_int64(0)
    viol = lib().cdt_check_delaunay(row.shape[0], _p(row, ctypes.c_int32), _p(col, ctypes.c_int32), nt, _p(tri_v, ctypes.c_int32),
                                    ctypes.byref(ns), _p(flag, ctypes.c_uint8), ctypes.byref(un))
    return dict(violations=int(viol), n_strict=int(ns.value), strict_flag=flag.astype(bool), hull_edges=int(un.value))


def rasterize(row, col, rgb, tri_v, img_h: int, img_w: int):
    """Exact integer barycentric interpolation.  Returns (interp u8 (h,w,3), hull bool (h,w), tri_id i32 (h,w))."""
    row = np.ascontiguousarray(row, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    rgb = np.ascontiguousarray(rgb, np.uint8)
    tri_v = np.ascontiguousarray(tri_v, np.int32)
    interp = np.zeros((img_h, img_w, 3), np.uint8)
    hull = np.zeros((img_h, img_w), np.uint8)
    tri_id = np.empty((img_h, img_w), np.int32)
    lib().cdt_rasterize(row.shape[0], img_h, img_w, _p(row, ctypes.c_int32), _p(col, ctypes.c_int32), _p(rgb, ctypes.c_uint8),
                        tri_v.shape[0], _p(tri_v, ctypes.c_int32), _p(interp, ctypes.c_uint8), _p(hull, ctypes.c_uint8),
                        _p(tri_id, ctypes.c_int32))
    return interp, hull.astype(bool), tri_id


def canonical_triangle_set(row, col, tri_v) -> np.ndarray:
    """Real triangles as sorted vertex triples, sorted lexicographically -- for set comparison."""
    real = tri_v[(tri_v >= 0).all(1)]
    t = np.sort(real, axis=1)
    return t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))]
