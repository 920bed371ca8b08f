"""ctypes binding of libsalve_bev.so (include/salve_bev.h).

The library is the product; there is no Python or CPU fallback.  Loading fails loudly if the
shared object is missing, and every call fails loudly if there is no CUDA device.
"""

from __future__ import annotations

import ctypes
import os

from . import build as _build

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u16p = ctypes.POINTER(ctypes.c_uint16)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_vp = ctypes.c_void_p


class Config(ctypes.Structure):
    """struct salve_bev_config (include/salve_bev.h)."""

    _fields_ = [
        ("device", ctypes.c_int32),
        ("pano_h", ctypes.c_int32),
        ("pano_w", ctypes.c_int32),
        ("max_panos", ctypes.c_int32),
        ("max_images", ctypes.c_int32),
        ("grid_h", ctypes.c_int32),
        ("grid_w", ctypes.c_int32),
        ("kernel_sz", ctypes.c_int32),
        ("xmin", ctypes.c_double),
        ("ymin", ctypes.c_double),
        ("xmax", ctypes.c_double),
        ("ymax", ctypes.c_double),
        ("px_per_m", ctypes.c_double),
        ("depth_scale", ctypes.c_float),
        ("crop_rows", ctypes.c_int32),
    ]


# every symbol include/salve_bev.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "salve_bev_default_config": (None, [ctypes.POINTER(Config), ctypes.c_int32, ctypes.c_int32]),
    "salve_bev_last_error": (ctypes.c_char_p, []),
    "salve_bev_ctx_create": (ctypes.c_int, [ctypes.POINTER(Config), ctypes.POINTER(c_vp)]),
    "salve_bev_ctx_destroy": (None, [c_vp]),
    "salve_bev_set_sphere_tables": (ctypes.c_int, [c_vp, c_f64p, c_f64p, c_f64p, c_f64p]),
    "salve_bev_get_uni_sphere_xyz": (ctypes.c_int, [c_vp, c_f64p]),
    "salve_bev_upload_pano": (ctypes.c_int, [c_vp, ctypes.c_int32, c_vp, c_vp, c_vp]),
    "salve_bev_bind_pano": (ctypes.c_int, [c_vp, ctypes.c_int32, c_vp, c_vp]),
    "salve_bev_upload_pano_fullres": (ctypes.c_int, [c_vp, ctypes.c_int32, c_vp, c_vp, c_vp]),
    "salve_bev_bind_pano_fullres": (ctypes.c_int, [c_vp, ctypes.c_int32, c_vp, c_vp]),
    "salve_bev_render_hypotheses": (ctypes.c_int, [c_vp, ctypes.c_int32, c_i32p, c_i32p, c_f32p, c_f32p, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp]),
    "salve_bev_render_hypotheses_host": (ctypes.c_int, [c_vp, ctypes.c_int32, c_i32p, c_i32p, c_f32p, c_f32p, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp]),
    "salve_bev_render_hypotheses_compact": (ctypes.c_int, [c_vp, ctypes.c_int32, c_i32p, c_i32p, c_f32p, c_f32p, ctypes.c_uint32, c_vp, c_vp, c_i32p, c_i32p, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "salve_bev_render_hypotheses_compact_host": (ctypes.c_int, [c_vp, ctypes.c_int32, c_i32p, c_i32p, c_f32p, c_f32p, ctypes.c_uint32, c_vp, c_vp, c_i32p, c_i32p, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "salve_bev_verifier_preprocess": (ctypes.c_int, [c_vp, ctypes.c_int32, c_vp, ctypes.c_int32, ctypes.c_int32, c_vp, c_vp]),
    "salve_bev_rasterize_layouts_host": (ctypes.c_int, [c_vp, ctypes.c_int32, c_i32p, c_i64p, c_vp, c_vp, c_vp]),
    "salve_bev_set_dedup_unposed": (ctypes.c_int, [c_vp, ctypes.c_int32]),
    "salve_bev_render_images_host": (ctypes.c_int, [c_vp, ctypes.c_int32, c_i32p, c_i32p, c_i32p, c_f32p, c_f32p, c_vp, c_vp, c_vp, c_vp]),
    "salve_bev_set_bands": (ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double]),
    "salve_bev_backproject": (ctypes.c_int, [c_vp, ctypes.c_int32, ctypes.c_double, ctypes.c_double, ctypes.c_int32, c_f32p, c_f32p, c_f64p, c_i64p, c_vp]),
    "salve_bev_render_cloud_host": (ctypes.c_int, [c_vp, c_f64p, ctypes.c_int64, c_u8p, c_i32p, c_i32p, c_vp]),
    "salve_bev_choose_elevated": (ctypes.c_int, [c_vp, c_i64p, c_i64p, c_f64p, ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_int32, c_u8p, c_vp]),
    "salve_bev_interp_dense": (ctypes.c_int, [c_vp, c_i64p, c_f64p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, c_u8p, c_u8p, c_i32p, c_vp]),
    "salve_bev_remove_hallucinated": (ctypes.c_int, [c_vp, c_u8p, c_u8p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_u8p, c_vp]),
    "salve_bev_tap": (ctypes.c_int, [c_vp, ctypes.c_int32, ctypes.c_int32, c_vp, ctypes.c_int64, c_vp]),
    "salve_bev_last_timings": (ctypes.c_int, [c_vp, c_f32p]),
    "salve_bev_enable_timing": (ctypes.c_int, [c_vp, ctypes.c_int32]),
    "salve_bev_local_rule_tables": (ctypes.c_int64, [c_vp, ctypes.c_int64]),
    "salve_bev_launch_count": (ctypes.c_int64, [c_vp]),
}

_lib = None


class SalveBevError(RuntimeError):
    pass


def lib_path() -> str:
    # SALVE_BEV_LIB: developer override used by scripts/variant_bench.py to time differently compiled builds of the same sources
    return os.environ.get("SALVE_BEV_LIB") or _build.LIB_PATH


def load():
    """dlopen libsalve_bev.so and bind every declared symbol.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise SalveBevError(
            f"{path} is missing: build it with `python -m salve_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback for this path."
        )
    lib = ctypes.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().salve_bev_last_error().decode(errors="replace")
        raise SalveBevError(f"libsalve_bev error {rc}: {msg}")
