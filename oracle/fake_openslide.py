"""TEST INFRASTRUCTURE: an in-memory stand-in for the `openslide` module, enough for the reference's mask readers
(WSI_evaluation/segment_utils.py:91-152: `open_slide`, `read_region(location, level, size)` returning an RGBA PIL image,
`level_downsamples`, `level_dimensions`). openslide and the ground-truth mask TIFFs are not available offline; the
segmentation metric code is exercised on a seeded synthetic mask pyramid instead.

    with fake_openslide.installed(mask):   # mask: uint8 [H, W] at level 0
        ...                                 # `import openslide; openslide.open_slide(anything)` serves the mask
"""
from __future__ import annotations

import contextlib
import sys
import types

import numpy as np
from PIL import Image


class FakeSlide:
    def __init__(self, mask: np.ndarray, downsamples=(1.0, 4.0, 16.0, 32.0)):
        self.levels = [np.ascontiguousarray(mask[::int(d), ::int(d)]) for d in downsamples]
        self.level_downsamples = tuple(float(d) for d in downsamples)
        self.level_dimensions = tuple((lv.shape[1], lv.shape[0]) for lv in self.levels)  # (width, height)

    def read_region(self, location, level, size):
        """location: (x, y) in level-0 pixels; size: (w, h) in `level` pixels; outside the slide reads as 0."""
        d = int(self.level_downsamples[level])
        x0, y0 = int(location[0]) // d, int(location[1]) // d
        w, h = int(size[0]), int(size[1])
        lv = self.levels[level]
        out = np.zeros((h, w), dtype=np.uint8)
        ys, xs = slice(max(y0, 0), min(y0 + h, lv.shape[0])), slice(max(x0, 0), min(x0 + w, lv.shape[1]))
        if ys.stop > ys.start and xs.stop > xs.start:
            out[ys.start - y0:ys.stop - y0, xs.start - x0:xs.stop - x0] = lv[ys, xs]
        return Image.fromarray(np.stack([out, out, out, np.full_like(out, 255)], -1), mode="RGBA")


@contextlib.contextmanager
def installed(mask: np.ndarray):
    mod = types.ModuleType("openslide")
    mod.open_slide = lambda path: FakeSlide(mask)
    mod.OpenSlide = lambda path: FakeSlide(mask)
    old = sys.modules.get("openslide")
    sys.modules["openslide"] = mod
    try:
        yield mod
    finally:
        if old is None:
            sys.modules.pop("openslide", None)
        else:
            sys.modules["openslide"] = old


def synthetic_case(seed: int = 5, grid: int = 24, patch: int = 224):
    """A seeded tumour mask (two blobs) on a grid x grid tile slide, tile keys 'x_y' and probabilities correlated with it."""
    rng = np.random.default_rng(seed)
    H = W = grid * patch
    yy, xx = np.mgrid[0:H, 0:W]
    mask = (((xx - 0.33 * W) ** 2 + (yy - 0.4 * H) ** 2 < (0.18 * W) ** 2) |
            ((xx - 0.72 * W) ** 2 / 2.0 + (yy - 0.7 * H) ** 2 < (0.12 * W) ** 2)).astype(np.uint8) * 255
    probs = {}
    for gy in range(grid):
        for gx in range(grid):
            if rng.random() < 0.15:
                continue  # holes in the tissue
            frac = mask[gy * patch:(gy + 1) * patch, gx * patch:(gx + 1) * patch].mean() / 255.0
            probs[f"{gx * patch}_{gy * patch}"] = float(np.clip(0.15 + 0.7 * frac + rng.normal(0, 0.18), 0.0, 1.0))
    return mask, probs
