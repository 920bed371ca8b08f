"""Quick end-to-end parity check on a GPU box (developer tool; the real checks live in tests/).

    python scripts/gpu_check.py [H W]
"""

import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import numpy as np

from oracle import bev_oracle as bo
from oracle import synth
from salve_b200.renderer import BevRenderer
import parity_utils as pu


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (512, 1024)
    rgb1, d1 = synth.synth_pano(H, W, 0, "iid")
    rgb2, d2 = synth.synth_pano(H, W, 1, "smooth")
    R, t = synth.synth_pose(0)
    r = BevRenderer(pano_h=H, pano_w=W, max_panos=4, max_images=8)
    r.upload_pano(0, rgb1, d1)
    r.upload_pano(1, rgb2, d2)
    r.enable_timing(True)
    t0 = time.time()
    imgs, counts, status = r.render_hypotheses([0], [1], R[None], t[None])
    print("render wall %.3fs" % (time.time() - t0), "timings", r.last_timings())
    print("counts", counts.reshape(-1, 8).tolist(), "status", status.reshape(-1).tolist())
    ok_all = True
    for si, surf in enumerate(["floor", "ceiling"]):
        s1, s2 = bo.render_pair(rgb1, d1, rgb2, d2, R, t, surf)
        for pi, st in enumerate((s1, s2)):
            img_idx = si * 2 + pi
            can = pu.oracle_canonical(st)
            res = {}
            res["crop"] = counts[0, si, pi, 0] == st.count_crop
            res["bbox"] = counts[0, si, pi, 1] == st.count_bbox
            res["sites"] = counts[0, si, pi, 2] == len(st.site_rc)
            kg = r.tap(img_idx, "keygrid").astype(np.int64)
            res["keygrid"] = np.array_equal(kg, st.key_grid)
            res["nonempty"] = np.array_equal(r.tap(img_idx, "nonempty"), st.nonempty)
            res["keep"] = np.array_equal(r.tap(img_idx, "keep"), st.keep)
            tri = r.tap(img_idx, "tris")
            res["tris"] = np.array_equal(pu.tri_pixel_set(tri), pu.oracle_tri_pixel_set(can))
            res["hull"] = np.array_equal(r.tap(img_idx, "hull"), can["hull"]) and pu.hull_check(can["hull"], st.hull) >= 0
            res["interp"] = np.array_equal(r.tap(img_idx, "interp"), can["interp"])
            res["final_canonical"] = np.array_equal(imgs[0, si, pi], pu.canonical_final(st, can))
            rep = pu.rgb_report(imgs[0, si, pi], st, can)
            print(surf, "pano", pi + 1, res, "rounds/flips", counts[0, si, pi, 6:8].tolist(), "oracle flips", can["stats"]["flips"])
            print("   rgb vs SciPy:", rep)
            ok_all &= all(res.values()) and rep["safe_gt1"] == 0.0 and rep["outside_kept_diff"] == 0
    print("ALL OK" if ok_all else "MISMATCH")
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
