"""Register this package's modules under the reference's module names, so that
`import salve.utils.bev_rendering_utils` (and friends) resolve to the GPU implementation.

    import salve_b200.dropin; salve_b200.dropin.install()

Only the modules on the BEV rendering path are aliased; anything else under `salve.` keeps
resolving to whatever is on sys.path (the reference, if installed).
"""

import importlib
import sys
import types

_ALIASES = {
    "salve.utils.bev_rendering_utils": "salve_b200.utils.bev_rendering_utils",
    "salve.utils.interpolation_utils": "salve_b200.utils.interpolation_utils",
    "salve.utils.zorder_utils": "salve_b200.utils.zorder_utils",
    "salve.utils.mesh_grid": "salve_b200.utils.mesh_grid",
    "salve.utils.hohonet_pano_utils": "salve_b200.utils.hohonet_pano_utils",
    "salve.common.bevparams": "salve_b200.common.bevparams",
    "salve.common.sim2": "salve_b200.common.sim2",
}


def install() -> None:
    for pkg in ("salve", "salve.utils", "salve.common"):
        if pkg not in sys.modules:
            try:
                importlib.import_module(pkg)
            except ImportError:
                m = types.ModuleType(pkg)
                m.__path__ = []  # namespace-like
                sys.modules[pkg] = m
    for ref_name, ours in _ALIASES.items():
        mod = importlib.import_module(ours)
        sys.modules[ref_name] = mod
        parent, _, leaf = ref_name.rpartition(".")
        setattr(sys.modules[parent], leaf, mod)
