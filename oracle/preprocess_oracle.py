"""CPU restatement of the verifier's val/test pre-processing chain -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path
(salve_b200/) never does.

Reference chain (salve/train_utils.py:126-159, get_val_test_transform, two modalities):
    ResizeQuadruplet((234, 234))      salve/utils/transform.py:256-272   cv2.resize(img, (w, h), INTER_LINEAR)
    CropQuadruplet((224, 224), "center")  transform.py:386-420           h_off = w_off = int((234 - 224) / 2) = 5
    ToTensorQuadruplet()              transform.py:79-85, 105-123        HWC uint8 -> CHW float32 (no /255)
    NormalizeQuadruplet(mean, std)    transform.py:177-202               t.sub_(m).div_(s), m/s = ImageNet * 255
                                      salve/utils/normalization_utils.py:13-26
and the model concatenates x1c, x2c, x1f, x2f along channels (salve/models/early_fusion.py:60-61,
salve/dataset/zind_data.py:306-315): ceiling pano 1, ceiling pano 2, floor pano 1, floor pano 2.

The resize arithmetic lives in a third-party dependency absent from /root/reference: OpenCV (cv2 4.13.0 here; the
reference does not pin it).  For 8-bit images cv2.resize(INTER_LINEAR) is fixed point (imgproc/src/resize.cpp,
HResizeLinear / VResizeLinear<uchar>): coefficients are rounded to 11 bits (x 2048), the horizontal pass keeps
S*a0 + S'*a1 as int, the vertical pass computes
    ((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2.
`resize_linear_u8` restates that; tests/test_oracle_cpu.py pins it bit-for-bit against cv2.resize itself.
"""

from __future__ import annotations

import numpy as np

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def get_imagenet_mean_std():
    """normalization_utils.py:13-26: python floats, value * 255."""
    mean = [item * 255 for item in IMAGENET_MEAN]
    std = [item * 255 for item in IMAGENET_STD]
    return mean, std


def linear_coefficients(src: int, dst: int):
    """Per destination index: (source index s0, s1, int16 weights w0, w1) as cv2 computes them
    (resize.cpp: scale in double, f = (float)((d + 0.5) * scale - 0.5), s = floor(f), f -= s, border clamps,
    weights saturate_cast<short>(float * 2048) = round half to even)."""
    scale = np.float64(src) / np.float64(dst)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    f[lo] = 0.0
    s[lo] = 0
    hi = s >= src - 1
    f[hi] = 0.0
    s[hi] = src - 1
    w0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int32)
    w1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int32)
    s1 = np.minimum(s + 1, src - 1)
    return s, s1, w0, w1


def resize_linear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """cv2.resize(img, (out_w, out_h), interpolation=cv2.INTER_LINEAR) for uint8 HWC images."""
    assert img.dtype == np.uint8 and img.ndim == 3
    h, w, _ = img.shape
    sx0, sx1, ax0, ax1 = linear_coefficients(w, out_w)
    sy0, sy1, by0, by1 = linear_coefficients(h, out_h)
    src = img.astype(np.int32)
    hor = src[:, sx0, :] * ax0[None, :, None] + src[:, sx1, :] * ax1[None, :, None]  # (h, out_w, c) int
    r0 = hor[sy0]
    r1 = hor[sy1]
    v = ((by0[:, None, None] * (r0 >> 4)) >> 16) + ((by1[:, None, None] * (r1 >> 4)) >> 16)
    return ((v + 2) >> 2).astype(np.uint8)


def preprocess_image(img: np.ndarray, resize_hw: int = 234, crop_hw: int = 224) -> np.ndarray:
    """One HWC uint8 image -> (3, crop, crop) float32, normalised."""
    r = resize_linear_u8(img, resize_hw, resize_hw)
    off = int((resize_hw - crop_hw) / 2)
    c = r[off : off + crop_hw, off : off + crop_hw]
    t = c.transpose(2, 0, 1).astype(np.float32)
    mean, std = get_imagenet_mean_std()
    for ch in range(3):
        # torch: t.sub_(m).div_(s) with python-float scalars on a float32 tensor = float32 arithmetic
        t[ch] = (t[ch] - np.float32(mean[ch])) / np.float32(std[ch])
    return t


def preprocess_quadruplet(x1c, x2c, x1f, x2f, resize_hw: int = 234, crop_hw: int = 224) -> np.ndarray:
    """(12, crop, crop) float32 = cat(x1c, x2c, x1f, x2f) after the val/test transform."""
    return np.concatenate([preprocess_image(x, resize_hw, crop_hw) for x in (x1c, x2c, x1f, x2f)], axis=0)
