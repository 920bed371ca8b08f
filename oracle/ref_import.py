"""Import the *unmodified* reference (zillow/salve) with stubbed I/O, to pin the oracle and to time it.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is looked up at $SALVE_REFERENCE_ROOT, else /root/reference
(the build container), else oracle/_ref (unmodified copies of the few modules the path loads, made by
scripts/make_oracle_ref.py; git-ignored, they travel to the GPU box like the built .so files).  It is used by
scripts/make_golden.py to freeze golden vectors under tests/golden/, by tests marked `needs_reference`, and by
bench.py's CPU legs (`--impl reference`, `cpu_baseline`).  Nothing under salve_b200/ imports it.

The reference's hot path imports packages that are absent here (imageio, matplotlib, gtsam,
gtsfm, colour, ...) but none of them carries hot-path arithmetic; they are replaced by stub
modules, and imageio.imread/imwrite by an in-memory store, so that
salve/utils/bev_rendering_utils.py:417-480 (render_bev_pair) runs as written.
"""

from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types
from types import SimpleNamespace
from unittest.mock import MagicMock

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    env = os.environ.get("SALVE_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isdir(os.path.join(cand, "salve")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()

_MISSING = {
    "imageio", "matplotlib", "gtsam", "gtsfm", "colour", "shapely", "rdp", "hydra",
    "omegaconf", "open3d", "seaborn", "mpl_toolkits",
}


class _Stub(types.ModuleType):
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in _MISSING:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, module):
        pass


STORE: dict = {}
_loaded = None


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "salve"))


def load():
    """Return the reference modules as a namespace (bru, interp, zorder, sphere, bevparams, sim2)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    # our own repo may contain a drop-in `salve` shim; make sure the reference wins here
    for k in [k for k in sys.modules if k == "salve" or k.startswith("salve.")]:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    sys.meta_path.insert(0, _Finder())
    import imageio  # stub

    imageio.imread = lambda p: STORE[p].copy()
    imageio.imwrite = lambda p, a: STORE.__setitem__(p, a)
    hi = types.ModuleType("salve.utils.hohonet_inference")
    hi.infer_depth_if_nonexistent = lambda **k: None
    sys.modules["salve.utils.hohonet_inference"] = hi

    import torch

    torch.set_num_threads(1)
    import salve.common.bevparams as bevparams
    import salve.common.sim2 as sim2
    import salve.utils.bev_rendering_utils as bru
    import salve.utils.hohonet_pano_utils as sphere
    import salve.utils.interpolation_utils as interp
    import salve.utils.zorder_utils as zorder

    _loaded = SimpleNamespace(bru=bru, interp=interp, zorder=zorder, sphere=sphere, bevparams=bevparams, sim2=sim2)
    return _loaded


def render_bev_pair(rgb1, depth1, rgb2, depth2, R, t, surface: str):
    """Run the reference's render_bev_pair on in-memory arrays (512x1024 only)."""
    import contextlib
    import io
    import warnings

    ref = load()
    STORE["rgb1"], STORE["d1"], STORE["rgb2"], STORE["d2"] = rgb1, depth1, rgb2, depth2
    crop = [-float("inf"), -1.0] if surface == "floor" else [0.5, float("inf")]
    args = SimpleNamespace(
        img_i1="rgb1", img_i2="rgb2", depth_i1="d1", depth_i2="d2", scale=0.001, crop_ratio=80 / 512, crop_z_range=crop
    )
    pose = ref.sim2.Sim2(R, t, 1.0)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        return ref.bru.render_bev_pair(args, "b", "f", 0, 1, pose, False)
