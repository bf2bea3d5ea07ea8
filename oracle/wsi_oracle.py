"""CPU oracle for the WSI task logic around the hot path — TEST INFRASTRUCTURE ONLY (see keep_oracle.py).

Restates, in plain torch/numpy on the CPU, the reference functions (paths relative to /root/reference):

    zero_shot_classifier      WSI_evaluation/utils.py:64-84     (incl. the `[0]` first-template quirk, :74)
    get_zeroshot_classifier   WSI_evaluation/utils.py:86-104
    rank_cls_score            WSI_evaluation/utils.py:107-117
    zero_shot_prompt_select   WSI_evaluation/utils.py:119-146
    tile softmax              detection_utils.py:90-93, subtyping_utils.py:69-72, segment_utils.py:46-49
    refine_seg (3 variants)   detection_utils.py:39-74, subtyping_utils.py:38-65, segment_utils.py:63-89
    zero_shot_detection       detection_utils.py:88-100
    zero_shot_subtyping       subtyping_utils.py:67-83

Pinned by tests/test_oracle.py against golden vectors produced by RUNNING the reference's own functions in the
build container (oracle/make_golden.py imports them from /root/reference with stub modules for the absent
h5py/openslide).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def tile_probs(classifier: torch.Tensor, tile_features: torch.Tensor, temp: float = 10.0):
    """logits = normalize(features) @ classifier ; probs = softmax(temp * logits, dim=1)."""
    feats = F.normalize(tile_features, dim=-1)
    logits = feats @ classifier
    return logits, torch.softmax(logits * temp, dim=1)


def zero_shot_classifier(model, tokenizer, classnames, templates, device="cpu"):
    cols = []
    with torch.no_grad():
        for name in classnames:
            if isinstance(templates, list):
                texts = [tpl.replace("CLASSNAME", name) for tpl in templates]
            else:
                texts = [templates.replace("CLASSNAME", name)]
            enc = tokenizer(texts, max_length=256, padding="max_length", truncation=True, return_tensors="pt").to(device)
            emb = model.encode_text(enc)[0]  # utils.py:74 — only the first template's embedding survives
            if emb.dim() == 1:
                emb = emb.unsqueeze(0)
            e = F.normalize(emb, dim=-1).mean(dim=0)
            cols.append(e / e.norm())
    return torch.stack(cols, dim=1).to(device)


def get_zeroshot_classifier(model, tokenizer, label_map, prompts, device="cpu", add_normal=False):
    names = prompts["classnames"]
    by_index = {v: k for k, v in label_map.items()}
    if add_normal:
        by_index[len(by_index)] = "Normal"
    ordered = [names[by_index[i]] for i in range(len(by_index))]
    return zero_shot_classifier(model, tokenizer, ordered, prompts["templates"], device)


def rank_cls_score(logits: torch.Tensor) -> float:
    top = torch.topk(logits, k=logits.shape[1], dim=1).values
    first, second = top[:, 0], top[:, 1]
    return ((first - second) - (first + second - 1).abs()).mean().item()


def zero_shot_prompt_select(classifiers, tile_features, topn):
    feats = F.normalize(tile_features.squeeze(0), dim=-1)
    scores = [rank_cls_score(feats @ c) for c in classifiers]
    order = torch.sort(torch.tensor(scores), descending=True).indices
    merged = torch.zeros_like(classifiers[0])
    for i in order[:topn]:
        merged += classifiers[i]
    return F.normalize(merged, p=2, dim=0), scores


def _first_occurrence(coords):
    first = {}
    for i, c in enumerate(np.asarray(coords).tolist()):
        first.setdefault((int(c[0]), int(c[1])), i)
    return first


def refine_mean(probs: np.ndarray, coords, patch_size: int, overlap: bool):
    """Common core of the three refine_seg variants: returns {(x,y): refined float32 prob vector} in insertion
    order. First tile at a coordinate wins; with overlap the vector is the float32 mean over the kept tiles
    present among (x-ps,y-ps), (x,y-ps), (x-ps,y), (x,y) in that order."""
    probs = np.asarray(probs, dtype=np.float32)
    first = _first_occurrence(coords)
    out = {}
    for (x, y), i in first.items():
        if not overlap:
            out[(x, y)] = probs[i]
            continue
        stack = [probs[first[c]] for c in ((x - patch_size, y - patch_size), (x, y - patch_size),
                                          (x - patch_size, y), (x, y)) if c in first]
        out[(x, y)] = np.array(stack).mean(0)
    return out


def refine_seg_detection(probs, coords, patch_size=224, threshold=0.5, overlap=True):
    ref = refine_mean(probs, coords, patch_size, overlap)
    preds = {f"{x}_{y}": int(v[1] > threshold) for (x, y), v in ref.items()}
    pr = {f"{x}_{y}": float(v[1]) for (x, y), v in ref.items()}
    return preds, pr


def refine_seg_subtyping(probs, coords, patch_size=224, overlap=True):
    ref = refine_mean(probs, coords, patch_size, overlap)
    return {f"{x}_{y}": int(np.argmax(v)) for (x, y), v in ref.items()}


def refine_seg_segment(probs, coords, patch_size=224, overlap=True):
    ref = refine_mean(probs, coords, patch_size, overlap)
    return {f"{x}_{y}": float(v[1]) for (x, y), v in ref.items()}


def zero_shot_detection(classifier, tile_features, tile_coords, patch_size=256, overlap=False):
    _, probs = tile_probs(classifier, tile_features)
    preds, _ = refine_seg_detection(probs.numpy(), tile_coords, patch_size=patch_size, overlap=overlap)
    return np.array(list(preds.values())).sum() / len(preds)


def zero_shot_subtyping(classifier, tile_features, tile_coords, patch_size=256, overlap=True):
    _, probs = tile_probs(classifier, tile_features)
    preds = refine_seg_subtyping(probs.numpy(), tile_coords, patch_size=patch_size, overlap=overlap)
    vals = np.array(list(preds.values()))
    frac = [(vals == c).sum() / len(preds) for c in range(classifier.shape[1])]
    return torch.tensor(frac[0:-1]).max(0).indices  # the appended 'Normal' column is excluded (subtyping_utils.py:82)
