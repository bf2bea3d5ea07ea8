"""CPU restatement of the reference's input transform — TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu legs).

quick_start/keep_inference.py:88-93 (and WSI_evaluation/zeroshot_*_WSI.py:38-43):
    Resize(224, BICUBIC) -> CenterCrop(224) -> ToTensor -> Normalize(mean, std)
torchvision applies Resize/CenterCrop to the PIL image, so the arithmetic is Pillow's (un-vendored dependency, pinned
Pillow==10.0.0, training/requirements.txt:10): libImaging/Resample.c precompute_coeffs, normalize_coeffs_8bpc,
ImagingResampleHorizontal_8bpc, ImagingResampleVertical_8bpc; torchvision.transforms.functional.resize /
center_crop for the output extent and the crop offsets. Pinned against the real torchvision + Pillow pipeline in
tests/golden/transform.npz (oracle/make_golden.py) and, when Pillow is importable, live in tests/test_oracle.py.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:  # Resample.c bicubic_filter, a = -0.5
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coefficients(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc: per output coordinate (xmin, n, int weights)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    out = []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        k = [w / ww if ww != 0.0 else w for w in k]
        ki = [int(-0.5 + w * (1 << PRECISION_BITS)) if w < 0 else int(0.5 + w * (1 << PRECISION_BITS)) for w in k]
        out.append((xmin, xmax, np.asarray(ki, dtype=np.int64)))
    return out


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_resize_bicubic(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Image.resize((out_w, out_h), BICUBIC) on an 8-bit RGB image [H,W,3]: horizontal pass, then vertical pass."""
    x = img.astype(np.int64)
    H, W, _ = x.shape
    if out_w != W:
        tmp = np.zeros((H, out_w, 3), np.uint8)
        for xo, (lo, n, k) in enumerate(coefficients(W, out_w)):
            tmp[:, xo] = _clip8((1 << (PRECISION_BITS - 1)) + np.tensordot(x[:, lo:lo + n], k, axes=([1], [0])))
        x = tmp.astype(np.int64)
    if out_h != H:
        tmp = np.zeros((out_h, x.shape[1], 3), np.uint8)
        for yo, (lo, n, k) in enumerate(coefficients(H, out_h)):
            tmp[yo] = _clip8((1 << (PRECISION_BITS - 1)) + np.tensordot(k, x[lo:lo + n], axes=([0], [0])))
        x = tmp
    return x.astype(np.uint8)


def resize_center_crop(img: np.ndarray, size: int = 224) -> np.ndarray:
    """transforms.Resize(size, BICUBIC) + CenterCrop(size) on uint8 [H,W,3] (keep_inference.py:89-90)."""
    H, W, _ = img.shape
    short, long_ = (W, H) if W <= H else (H, W)
    new_short, new_long = size, int(size * long_ / short)  # torchvision _compute_resized_output_size
    out_w, out_h = (new_short, new_long) if W <= H else (new_long, new_short)
    r = img if (out_h, out_w) == (H, W) else pil_resize_bicubic(img, out_h, out_w)
    top, left = int(round((out_h - size) / 2.0)), int(round((out_w - size) / 2.0))  # center_crop: Python round()
    return r[top:top + size, left:left + size]


def to_tensor_normalize(img_u8: np.ndarray) -> np.ndarray:
    """ToTensor + Normalize (keep_inference.py:91-92): uint8 [H,W,3] -> float32 [3,H,W]."""
    mean = np.asarray([0.485, 0.456, 0.406], np.float32).reshape(3, 1, 1)
    std = np.asarray([0.229, 0.224, 0.225], np.float32).reshape(3, 1, 1)
    x = img_u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    return (x - mean) / std
