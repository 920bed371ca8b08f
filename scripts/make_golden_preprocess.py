"""Freeze a golden vector of the verifier pre-processing from the UNMODIFIED reference transform classes
(salve/utils/transform.py imported from /root/reference; cv2 + torch do the arithmetic) and check the oracle restatement
(oracle/preprocess_oracle.py) against it bit-for-bit.

Run in the build container only:   python scripts/make_golden_preprocess.py
Writes tests/golden/preprocess_c1.npz: the inputs are re-generated from the seed at test time; stored are the sha256 of the
(12, 224, 224) float32 result and a few of its rows.
"""

import collections
import collections.abc
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import preprocess_oracle as po  # noqa: E402
from oracle import ref_import  # noqa: E402

SEED = 1234
ROWS = (0, 1, 111, 222, 223)


def golden_inputs(seed: int = SEED):
    """Four 501x501x3 uint8 images: two noise-like, two smooth with zero borders (like real renders)."""
    rng = np.random.default_rng(seed)
    imgs = []
    for k in range(4):
        if k % 2 == 0:
            a = rng.integers(0, 256, size=(501, 501, 3), dtype=np.int64).astype(np.uint8)
        else:
            yy, xx = np.mgrid[0:501, 0:501]
            a = np.stack([(yy * 255 // 500), (xx * 255 // 500), ((yy * 3 + xx * 5 + 17 * k) % 256)], -1).astype(np.uint8)
        a[:40] = 0
        a[:, -55:] = 0
        imgs.append(a)
    return imgs


def main():
    assert ref_import.available(), "reference not mounted"
    collections.Iterable = collections.abc.Iterable  # transform.py:244 predates Python 3.10
    ref_import.load()
    import salve.utils.normalization_utils as nu
    import salve.utils.transform as T

    mean, std = nu.get_imagenet_mean_std()
    chain = T.ComposeQuadruplet(
        [T.ResizeQuadruplet((234, 234)), T.CropQuadruplet(size=(224, 224), crop_type="center", padding=mean), T.ToTensorQuadruplet(),
         T.NormalizeQuadruplet(mean=mean, std=std)]
    )
    x1c, x2c, x1f, x2f = golden_inputs()
    out = chain(x1c.copy(), x2c.copy(), x1f.copy(), x2f.copy())
    import torch

    ref = torch.cat([o[None] for o in out], dim=1)[0].numpy()
    mine = po.preprocess_quadruplet(x1c, x2c, x1f, x2f)
    assert ref.shape == (12, 224, 224) and ref.dtype == np.float32
    assert np.array_equal(ref, mine), "oracle restatement != reference transform chain"
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "preprocess_c1.npz"),
        seed=np.int64(SEED), rows=np.array(ROWS), values=ref[:, ROWS, :],
        sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(ref).tobytes()).digest(), np.uint8),
    )
    print("reference chain == oracle restatement; golden written")


if __name__ == "__main__":
    main()
