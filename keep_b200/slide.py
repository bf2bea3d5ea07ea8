"""Slide-level entry point: what the three reference scripts do between "prompts JSON + slide features" and the answer.

    reference script                                   lines     here
    WSI_evaluation/zeroshot_detection_WSI.py           26-71     zero_shot_slide("detection", ...)
    WSI_evaluation/zeroshot_subtyping_WSI.py           27-74     zero_shot_slide("subtyping", ...)
    WSI_evaluation/zeroshot_segmentation_WSI.py        22-68     zero_shot_slide("segmentation", ...)

The scripts share one flow: for every entry of the prompt file build a [hidden, C] classifier (K = 660-1782 entries, one
batch-1 `encode_text` call per class: 2.8k-7.1k BERT calls and as many host round trips) -> rank the K classifiers on the
slide (`zero_shot_prompt_select`, one small GEMM + topk + `.item()` per classifier) and merge the top-n -> run the task head.
Here the bank is ONE batched `encode_text` over the distinct class texts, the ranking ONE pass of the screening kernel, and
the task heads the device kernels of keep_b200.wsi. Defaults (label maps, add_normal, patch_size, overlap, topn) are the
scripts'. `timings`, when given, receives the wall time of every stage in seconds (device-synchronised).
"""
from __future__ import annotations

import random
import time
from typing import Mapping, MutableMapping, Optional

import torch

from . import wsi

# per task: label map, add_normal, patch_size, overlap  — zeroshot_detection_WSI.py:22-24,70; zeroshot_subtyping_WSI.py:25-27,
# 58,72; zeroshot_segmentation_WSI.py:21,66
TASK_DEFAULTS = {
    "detection": dict(label_map={"Normal": 0, "Tumor": 1}, add_normal=False, patch_size=256, overlap=False),
    "subtyping": dict(label_map={"CHRCC": 0, "CCRCC": 1, "PRCC": 2}, add_normal=True, patch_size=256, overlap=True),
    "segmentation": dict(label_map={"Normal": 0, "Tumor": 1}, add_normal=False, patch_size=224, overlap=True),
}


def _stamp(timings: Optional[MutableMapping], key: str, t0: float, device) -> float:
    if timings is not None:
        if torch.device(device).type == "cuda":
            torch.cuda.synchronize(device)
        timings[key] = time.perf_counter() - t0
    return time.perf_counter()


def zero_shot_slide(task: str, KEEP_model: Mapping, prompts: Mapping, tile_features: torch.Tensor, tile_coords, device,
                    label_map: Optional[Mapping] = None, topn: int = 50, prompt_screening: bool = True,
                    add_normal: Optional[bool] = None, patch_size: Optional[int] = None, overlap: Optional[bool] = None,
                    mask_path: Optional[str] = None, timings: Optional[MutableMapping] = None):
    """Run one slide through one of the reference's three zero-shot tasks.

    `prompts` is the parsed prompt file ({"0": {"classnames": {...}, "templates": ...}, ...}); `tile_features` are the
    slide's [N, D] features (the reference's h5 `features`) or raw tiles ([N,3,H,W] float / [N,H,W,3] uint8, encoded here);
    `tile_coords` its [N, 2] coordinates. Returns what the script prints: detection -> tumour-tile fraction (float);
    subtyping -> 0-dim LongTensor (index into the label map); segmentation -> {"x_y": refined tumour probability}, or
    (auc, dice) when `mask_path` is given (needs openslide, as the reference does)."""
    if task not in TASK_DEFAULTS:
        raise ValueError(f"task must be one of {sorted(TASK_DEFAULTS)}, got {task!r}")
    d = TASK_DEFAULTS[task]
    label_map = d["label_map"] if label_map is None else label_map
    add_normal = d["add_normal"] if add_normal is None else add_normal
    patch_size = d["patch_size"] if patch_size is None else patch_size
    overlap = d["overlap"] if overlap is None else overlap

    t0 = time.perf_counter()
    feats = tile_features.to(device)
    if feats.dim() == 4:  # raw tiles: the stage the reference leaves to an upstream feature extractor
        feats = KEEP_model["model"].encode_image(feats)
    t0 = _stamp(timings, "features_s", t0, device)

    # ---- the classifier bank: one classifier per prompt-file entry (scripts: "generate prompt classifier") ----
    bank = wsi.build_classifier_bank(KEEP_model, label_map, prompts, device, add_normal=add_normal)
    t0 = _stamp(timings, "classifier_bank_s", t0, device)

    # ---- screening / ensembling (scripts: "select prompt classifier") ----
    if prompt_screening:
        ensemble = wsi.zero_shot_prompt_select(bank, feats, topn=topn, device=device)
    else:  # the scripts' fallback: topn classifiers drawn with random.seed(cter)
        merged = torch.zeros_like(bank[0])
        for cter in range(topn):
            random.seed(cter)
            merged += bank[random.randint(0, len(bank) - 1)]
        ensemble = torch.nn.functional.normalize(merged, p=2, dim=0)
    t0 = _stamp(timings, "screening_s", t0, device)

    # ---- task head ----
    if task == "detection":
        out = wsi.zero_shot_detection(ensemble, feats, tile_coords, patch_size=patch_size, overlap=overlap)
    elif task == "subtyping":
        out = wsi.zero_shot_subtyping(ensemble, feats, tile_coords, patch_size=patch_size, overlap=overlap)
    elif mask_path is not None:
        out = wsi.zero_shot_segment(ensemble, feats, tile_coords, mask_path, patch_size=patch_size, overlap=overlap)
    else:
        out = wsi.zero_shot_segment_probs(ensemble, feats, tile_coords, patch_size=patch_size, overlap=overlap)
    _stamp(timings, "task_head_s", t0, device)
    if timings is not None:
        timings["classifiers"] = len(bank)
        timings["tiles"] = int(feats.shape[0])
    return out
