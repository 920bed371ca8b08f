"""Freeze golden vectors from the UNMODIFIED reference (imported from /root/reference with I/O stubs,
oracle/ref_import.py) and check the oracle restatement against it bit-for-bit.

Run in the build container only (the reference does not travel to the GPU box):
    python scripts/make_golden.py

Writes tests/golden/pair_c1.npz:
  inputs are re-generated from seeds at test time (oracle/synth.py); their sha256 is stored to catch drift.
  per image (floor/ceiling x pano1-posed/pano2): the reference's final image, plus the stage values the
  oracle exposes (counts, winner keys digest, bit-packed non-empty / keep / hull masks).
Also re-runs the reference's own known-answer tests for the path and stores their expected values
(tests/golden/kat.json) so that tests/ can replay them against the oracle and the CUDA entry points.
"""

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import bev_oracle as bo  # noqa: E402
from oracle import synth  # noqa: E402
from oracle import ref_import  # noqa: E402

H, W = 512, 1024
CASE = dict(pano1_seed=0, pano2_seed=1, pose_seed=0, tex1="iid", tex2="smooth")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    assert ref_import.available(), "reference not mounted"
    rgb1, d1 = synth.synth_pano(H, W, CASE["pano1_seed"], CASE["tex1"])
    rgb2, d2 = synth.synth_pano(H, W, CASE["pano2_seed"], CASE["tex2"])
    R, t = synth.synth_pose(CASE["pose_seed"])
    out = dict(R=R, t=t)
    meta = dict(case=CASE, H=H, W=W, inputs_sha=dict(rgb1=sha(rgb1), d1=sha(d1), rgb2=sha(rgb2), d2=sha(d2)), images={})
    ref = ref_import.load()
    for surf in ("floor", "ceiling"):
        r1, r2 = ref_import.render_bev_pair(rgb1, d1, rgb2, d2, R, t, surf)
        s1, s2 = bo.render_pair(rgb1, d1, rgb2, d2, R, t, surf)
        for name, ref_img, st in ((f"{surf}_1", r1, s1), (f"{surf}_2", r2, s2)):
            assert np.array_equal(ref_img, st.final), f"oracle != reference on {name}"
            out[f"{name}_final"] = ref_img
            out[f"{name}_nonempty"] = np.packbits(st.nonempty)
            out[f"{name}_keep"] = np.packbits(st.keep)
            out[f"{name}_hull"] = np.packbits(st.hull)
            meta["images"][name] = dict(count_crop=int(st.count_crop), count_bbox=int(st.count_bbox), n_sites=int(len(st.site_rc)),
                                        key_grid_sha=sha(st.key_grid.astype(np.int64)), final_sha=sha(ref_img))
            print(name, "reference == oracle; sites", len(st.site_rc))
    # stage-level pins against the reference's own functions
    assert np.array_equal(ref.sphere.get_uni_sphere_xyz(H, W), bo.uni_sphere_xyz(H, W))
    meta["sphere_sha"] = sha(bo.uni_sphere_xyz(H, W))
    rng = np.random.default_rng(5)
    x = rng.integers(0, 40, 3000); y = rng.integers(0, 30, 3000); z = rng.uniform(-2.5, 2.5, 3000)
    assert np.array_equal(ref.zorder.choose_elevated_repeated_vals(x, y, z), bo.choose_elevated(x, y, z))
    sp = rng.integers(0, 256, (60, 70, 3)).astype(np.uint8); sp[rng.random((60, 70)) < 0.9] = 0
    it = rng.integers(0, 256, (60, 70, 3)).astype(np.uint8)
    ref_h = ref.interp.remove_hallucinated_content(sp, it, 11)
    mine = (np.repeat(bo.keep_mask(bo.nonempty_mask(sp), 11)[:, :, None], 3, 2) * it).astype(np.uint8)
    assert np.array_equal(ref_h, mine)
    print("stage pins ok (sphere table, z-order rule, hallucination mask incl. uint8 wrap)")
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pair_c1.npz"), **out)
    import scipy

    meta["generated_with"] = dict(numpy=np.__version__, scipy=scipy.__version__)
    with open(os.path.join(ROOT, "tests", "golden", "pair_c1.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote tests/golden/pair_c1.{npz,json}", os.path.getsize(os.path.join(ROOT, "tests", "golden", "pair_c1.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
