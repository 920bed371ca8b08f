"""Freeze golden vectors of the layout modality from the UNMODIFIED reference (its rasterize_room_layout_pair, run with the real cv2).

Run in the build container only:   python scripts/make_golden_layout.py
Writes tests/golden/layout_c1.npz: for a few seeded synthetic layouts (oracle/layout_synth.py) and poses the reference's two images.
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import layout_synth, ref_import, synth  # noqa: E402

CASES = [(0, 1, 3), (2, 3, 4), (4, 5, 6)]  # (room seed 1, room seed 2, pose seed)


def pose(seed):
    R, t = synth.synth_pose(seed)
    return R, (t * 0.25).astype(np.float32)  # keep pano 1's room inside the 10 m box


def main():
    ref = ref_import.load()
    import salve.common.pano_data as pano_data  # the reference's WDO class

    out = {}
    for k, (s1, s2, ps) in enumerate(CASES):
        g = layout_synth.nodes([s1, s2], wdo_cls=pano_data.WDO, sim2_cls=ref.sim2.Sim2)
        R, t = pose(ps)
        i1, i2 = ref.bru.rasterize_room_layout_pair(ref.sim2.Sim2(R, t, 1.0), g, "b", "f", 0, 1)
        out[f"case{k}_img1"], out[f"case{k}_img2"] = i1, i2
        print(k, i1.shape, int((i1 > 0).any(2).sum()), int((i2 > 0).any(2).sum()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "layout_c1.npz"), **out)
    print("wrote tests/golden/layout_c1.npz", os.path.getsize(os.path.join(ROOT, "tests", "golden", "layout_c1.npz")), "bytes")


if __name__ == "__main__":
    main()
