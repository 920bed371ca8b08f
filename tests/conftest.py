import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    from oracle import ref_import

    if not ref_import.available():
        skip = pytest.mark.skip(reason="reference not mounted on this box")
        for it in items:
            if "needs_reference" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import json

    import numpy as np

    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "pair_c1.npz")))
    with open(os.path.join(ROOT, "tests", "golden", "pair_c1.json")) as f:
        meta = json.load(f)
    return g, meta


@pytest.fixture(scope="session")
def golden_inputs(golden):
    from oracle import synth

    _, meta = golden
    c = meta["case"]
    rgb1, d1 = synth.synth_pano(meta["H"], meta["W"], c["pano1_seed"], c["tex1"])
    rgb2, d2 = synth.synth_pano(meta["H"], meta["W"], c["pano2_seed"], c["tex2"])
    R, t = synth.synth_pose(c["pose_seed"])
    return rgb1, d1, rgb2, d2, R, t
