"""Slide-level segmentation metrics (CPU, needs `openslide`): restated from
WSI_evaluation/segment_utils.py:91-119 (`eval_seg_auc`) and :122-152 (`eval_seg_coarse`).

This is metric code outside the accelerated path (it needs openslide and a ground-truth mask
TIFF that are not in the tree); it exists so that `keep_b200.wsi.zero_shot_segment` keeps the reference's
return value `(auc, dice)` when openslide is available.
"""
from __future__ import annotations

import numpy as np


def _coords(key: str):
    x, y = key.split("_")
    return int(x), int(y)


def eval_seg_auc(probs_all_refined: dict, mask_path: str, patch_size: int = 224, save_path: str = "./"):
    import openslide
    from sklearn import metrics

    mask = openslide.open_slide(mask_path)
    gt, pr = [], []
    for key, prob in probs_all_refined.items():
        x, y = _coords(key)
        region = np.array(mask.read_region((x, y), 0, (patch_size, patch_size)).convert("L"))
        gt.append(1 if np.count_nonzero(region) > patch_size * patch_size / 2 else 0)  # majority-tumour tile
        pr.append(prob)
    gt, pr = np.array(gt), np.array(pr)
    auc = metrics.roc_auc_score(gt, pr)
    fpr, tpr, thresholds = metrics.roc_curve(gt, pr)
    return auc, thresholds[np.argmax(tpr - fpr)]  # Youden's J


def eval_seg_coarse(probs_all_refined: dict, mask_path: str, patch_size: int = 224, thd: float = 0.5):
    import openslide

    mask = openslide.open_slide(mask_path)
    downs = mask.level_downsamples
    level = min(range(len(downs)), key=lambda i: abs(downs[i] - 16))
    mask_img = np.array(mask.read_region([0, 0], level, mask.level_dimensions[level]).convert("L"))
    mag = int(downs[level])
    pred = np.zeros_like(mask_img)
    for key, prob in probs_all_refined.items():
        if prob > thd:
            x, y = _coords(key)
            pred[int(y / mag):int(y / mag + patch_size / mag), int(x / mag):int(x / mag + patch_size / mag)] = 255
    mask_sum = np.count_nonzero(mask_img) * 256
    pred_sum = np.count_nonzero(pred) * 256
    inter = np.count_nonzero(mask_img * pred) * 256
    if mask_sum + pred_sum == 0:
        return 1
    return 2 * inter / (mask_sum + pred_sum)
