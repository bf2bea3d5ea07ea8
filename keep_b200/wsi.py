"""WSI_evaluation entry points on the B200 path — same names, arguments and results as the reference.

    reference function                         file:line                                   here
    zero_shot_classifier                       WSI_evaluation/utils.py:64-84               zero_shot_classifier
    get_zeroshot_classifier                    WSI_evaluation/utils.py:86-104              get_zeroshot_classifier
    rank_cls_score                             WSI_evaluation/utils.py:107-117             rank_cls_score
    zero_shot_prompt_select                    WSI_evaluation/utils.py:119-146             zero_shot_prompt_select
    zero_shot_detection / refine_seg           WSI_evaluation/detection_utils.py:88-100, 39-74
    zero_shot_subtyping / refine_seg           WSI_evaluation/subtyping_utils.py:67-83, 38-65
    zero_shot_segment   / refine_seg           WSI_evaluation/segment_utils.py:44-60, 63-89

Differences that are additions, not changes: `tile_features` may also be raw tiles ([N,3,224,224] float or
[N,224,224,3] uint8) when `model=` is given (the reference scripts only ever see pre-extracted features,
SURVEY.md D4); `build_classifier_bank` encodes a whole prompt file in one batched `encode_text` call instead
of K*C batch-1 calls; `*_tensors` variants return device tensors instead of "x_y"-keyed dicts.
Quirks kept on purpose: only the first template's embedding is used (utils.py:74), the softmax temperature is
the literal 10 (detection_utils.py:93), the first tile at a coordinate wins (detection_utils.py:45), subtyping
ignores the last ('Normal') column when picking the slide label (subtyping_utils.py:82).

All arithmetic happens in libkeep_b200.so (similarity, screening and refine kernels); this file only orders
calls and converts results to the reference's Python containers.
"""
from __future__ import annotations

from typing import Mapping, Sequence

import numpy as np
import torch

from . import ops

SOFTMAX_TEMPERATURE = 10.0  # hard-coded in the reference, not logit_scale.exp()


def cood2str(cood):  # utils.py:148-149
    return str(cood[0]) + "_" + str(cood[1])


def str2cood(s):  # utils.py:150-151
    return [int(item) for item in s.split("_")]


# ---------------------------------------------------------------------------------------------------------
# prompt classifiers
# ---------------------------------------------------------------------------------------------------------
def _texts_for(classname: str, templates) -> list:
    if isinstance(templates, list):
        return [t.replace("CLASSNAME", classname) for t in templates]
    if isinstance(templates, str):
        return [templates.replace("CLASSNAME", classname)]
    raise TypeError("templates must be a str or a list of str")


def zero_shot_classifier(KEEP_model: Mapping, classnames: Sequence[str], templates, device):
    """[hidden, C] stack of unit-norm class embeddings (utils.py:64-84)."""
    cols = []
    for classname in classnames:
        texts = _texts_for(classname, templates)
        enc = KEEP_model["tokenizer"](texts, max_length=256, padding="max_length", truncation=True,
                                      return_tensors="pt").to(device)
        emb = KEEP_model["model"].encode_text(enc)[0]  # first template only — reference behaviour (utils.py:74)
        cols.append(emb)
    w = torch.stack(cols, dim=1).to(device)
    # per column: normalize -> mean over one row -> renormalize (utils.py:79-80); on unit vectors this is a
    # renormalisation, done here for all columns at once
    return w / w.norm(dim=0, keepdim=True)


def get_zeroshot_classifier(model: Mapping, label_map: Mapping, prompts: Mapping, device, add_normal: bool = False):
    """utils.py:86-104: columns ordered by label index, optional trailing 'Normal' column."""
    classnames = prompts["classnames"]
    idx_to_class = {v: k for k, v in label_map.items()}
    if add_normal:
        idx_to_class[len(idx_to_class)] = "Normal"
    ordered = [classnames[idx_to_class[i]] for i in range(len(idx_to_class))]
    return zero_shot_classifier(model, ordered, prompts["templates"], device)


def build_classifier_bank(model: Mapping, label_map: Mapping, prompts_by_index: Mapping, device,
                          add_normal: bool = False) -> list:
    """All K prompt-classifiers of a prompt file (the loop at zeroshot_detection_WSI.py:50-53) with ONE batched
    encode_text over the distinct class texts instead of K*C batch-1 calls."""
    idx_to_class = {v: k for k, v in label_map.items()}
    if add_normal:
        idx_to_class[len(idx_to_class)] = "Normal"
    order = [idx_to_class[i] for i in range(len(idx_to_class))]
    texts, per_prompt = {}, []
    for k in range(len(prompts_by_index)):
        pr = prompts_by_index[str(k)] if str(k) in prompts_by_index else prompts_by_index[k]
        row = []
        for cls in order:
            t = _texts_for(pr["classnames"][cls], pr["templates"])[0]  # first template only (utils.py:74)
            row.append(texts.setdefault(t, len(texts)))
        per_prompt.append(row)
    uniq = list(texts.keys())
    enc = model["tokenizer"](uniq, max_length=256, padding="max_length", truncation=True, return_tensors="pt").to(device)
    emb = model["model"].encode_text(enc)  # [U, hidden], unit norm
    emb = emb / emb.norm(dim=1, keepdim=True)
    return [emb[row].t().contiguous() for row in per_prompt]


# ---------------------------------------------------------------------------------------------------------
# prompt screening
# ---------------------------------------------------------------------------------------------------------
def rank_cls_score(logits: torch.Tensor) -> float:
    """utils.py:107-117 for one classifier's logits [N,C] (device reduction, one host read)."""
    top = torch.topk(logits, k=2, dim=1).values
    return ((top[:, 0] - top[:, 1]) - (top[:, 0] + top[:, 1] - 1).abs()).mean().item()


def prompt_scores(classifiers: Sequence[torch.Tensor], tile_features: torch.Tensor) -> torch.Tensor:
    """cls_score of every classifier in one pass: one [N,D]x[D,K*C] product with the top-2 margin reduced on the
    device, instead of K small GEMMs with a host sync each (utils.py:127-130)."""
    feats = _features(tile_features, None)
    K, C = len(classifiers), classifiers[0].shape[1]
    stacked = torch.cat([c.to(feats.device, torch.float32) for c in classifiers], dim=1).contiguous()
    return ops.prompt_scores(feats, stacked, K, C)


def zero_shot_prompt_select(classifiers: Sequence[torch.Tensor], tile_features: torch.Tensor, topn: int, device):
    """utils.py:119-146: rank classifiers by cls_score, sum the top-n, L2-normalise the columns."""
    feats = tile_features.to(device, non_blocking=True).squeeze(0)
    scores = prompt_scores(classifiers, feats).cpu()
    _, index = torch.sort(scores, descending=True, stable=True)
    merged = torch.zeros_like(classifiers[0])
    for i in index[0:topn].tolist():
        merged += classifiers[i]
    return torch.nn.functional.normalize(merged, p=2, dim=0)


# ---------------------------------------------------------------------------------------------------------
# tile x prompt similarity
# ---------------------------------------------------------------------------------------------------------
def _features(tile_features: torch.Tensor, model) -> torch.Tensor:
    if tile_features.dim() == 4:
        if model is None:
            raise ValueError("raw tiles were passed; give model=<KEEPModel> so they can be encoded")
        m = model["model"] if isinstance(model, Mapping) else model
        return m.encode_image(tile_features)
    if tile_features.dim() != 2:
        raise ValueError(f"tile_features must be [N,D] features or [N,3,H,W] tiles, got {tuple(tile_features.shape)}")
    return tile_features


def tile_probabilities(classifier: torch.Tensor, tile_features: torch.Tensor, model=None):
    """(logits, probs) with probs = softmax(10 * normalize(features) @ classifier, dim=1)
    (detection_utils.py:90-93, subtyping_utils.py:69-72, segment_utils.py:46-49)."""
    feats = _features(tile_features, model)
    return ops.similarity(feats, classifier.to(feats.device), group=0, temp=SOFTMAX_TEMPERATURE)


# ---------------------------------------------------------------------------------------------------------
# refine_seg and the three task heads
# ---------------------------------------------------------------------------------------------------------
def refine_tensors(probs: torch.Tensor, tile_coords, patch_size: int, overlap: bool):
    """(kept index [M], coords [M,2] int64, refined probs [M,C]) in first-occurrence order, on the device."""
    if isinstance(tile_coords, torch.Tensor):  # the reference passes the h5 `coords` array; device tensors are taken as they are
        coords = tile_coords.to(device=probs.device, dtype=torch.long).reshape(-1, 2)
    else:
        coords = torch.as_tensor(np.asarray(tile_coords)).to(device=probs.device, dtype=torch.long).reshape(-1, 2)
    keep, refined = ops.refine(coords, probs, patch_size, overlap)
    idx = keep.nonzero().flatten()
    return idx, coords[idx], refined[idx]


def _keys(coords: torch.Tensor) -> list:
    return [f"{x}_{y}" for x, y in coords.cpu().tolist()]


def refine_seg_detection(logits_slide, coords_slide, patch_size=224, threshold=0.5, overlap=True):
    """detection_utils.py:39-74 -> ({"x_y": 0|1}, {"x_y": tumour prob})."""
    _, coords, ref = refine_tensors(logits_slide, coords_slide, patch_size, overlap)
    keys = _keys(coords)
    p1 = ref[:, 1].cpu()
    return dict(zip(keys, (p1 > threshold).long().tolist())), dict(zip(keys, p1.tolist()))


def refine_seg_subtyping(logits_slide, coords_slide, patch_size=224, overlap=True):
    """subtyping_utils.py:38-65 -> {"x_y": argmax class}."""
    _, coords, ref = refine_tensors(logits_slide, coords_slide, patch_size, overlap)
    return dict(zip(_keys(coords), ref.argmax(dim=1).cpu().tolist()))


def refine_seg_segment(logits_slide, coords_slide, patch_size=224, overlap=True):
    """segment_utils.py:63-89 -> {"x_y": tumour prob}."""
    _, coords, ref = refine_tensors(logits_slide, coords_slide, patch_size, overlap)
    return dict(zip(_keys(coords), ref[:, 1].cpu().tolist()))


def zero_shot_detection(classifier, tile_features, tile_coords, patch_size=256, overlap=False, model=None):
    """Tumour-tile fraction of the slide (detection_utils.py:88-100)."""
    _, probs = tile_probabilities(classifier, tile_features, model)
    _, _, ref = refine_tensors(probs, tile_coords, patch_size, overlap)
    return float((ref[:, 1] > 0.5).sum().item()) / ref.shape[0]


def zero_shot_subtyping(classifier, tile_features, tile_coords, patch_size=256, overlap=True, model=None):
    """Slide subtype = argmax over the per-class tile fractions, last ('Normal') column excluded
    (subtyping_utils.py:67-83). Returns a 0-dim LongTensor like the reference."""
    _, probs = tile_probabilities(classifier, tile_features, model)
    _, _, ref = refine_tensors(probs, tile_coords, patch_size, overlap)
    C = classifier.shape[1]
    counts = torch.bincount(ref.argmax(dim=1), minlength=C).cpu().numpy()
    frac = [counts[ix] / ref.shape[0] for ix in range(C)]
    _, max_label = torch.tensor(frac[0:-1]).max(0)
    return max_label


def zero_shot_segment_probs(classifier, tile_features, tile_coords, patch_size=224, overlap=True, model=None):
    """The device part of zero_shot_segment (segment_utils.py:44-52): refined per-tile tumour probabilities."""
    _, probs = tile_probabilities(classifier, tile_features, model)
    return refine_seg_segment(probs, tile_coords, patch_size=patch_size, overlap=overlap)


def zero_shot_segment(classifier, tile_features, tile_coords, mask_path, patch_size=224, overlap=True, model=None):
    """segment_utils.py:44-60. The AUROC/Dice evaluation reads the ground-truth mask with openslide on the CPU
    (segment_utils.py:91-152); that metric code is outside the accelerated path and needs openslide."""
    probs_all_refined = zero_shot_segment_probs(classifier, tile_features, tile_coords, patch_size, overlap, model)
    try:
        import openslide  # noqa: F401
    except ImportError as e:
        raise ImportError("zero_shot_segment needs `openslide` to read the ground-truth mask; "
                          "use zero_shot_segment_probs for the refined tile probabilities") from e
    from .seg_eval import eval_seg_auc, eval_seg_coarse

    auc, best_thd = eval_seg_auc(probs_all_refined, mask_path, patch_size=patch_size)
    dice = eval_seg_coarse(probs_all_refined, mask_path, patch_size=patch_size, thd=best_thd)
    return auc, dice
