#!/usr/bin/env python
"""Slide-level harness: one synthetic slide through the reference's three zero-shot scripts, stage by stage.

    python tools/run_slide.py [--task detection|subtyping|segmentation|all] [--tiles N] [--prompts FILE.json]
                              [--oracle-classifiers K] [--out profiles/r02_slide_<task>.json]

What runs (keep_b200.slide.zero_shot_slide = zeroshot_{detection,subtyping,segmentation}_WSI.py:26-71 on the B200 path):
prompt file -> K classifiers (ONE batched encode_text) -> screening (ONE pass of the screening kernel) -> top-50 ensemble
-> task head (similarity + softmax(x10) + refine_seg + slide reduction). Inputs are synthetic: a `.npz` slide of unit-norm
768-d features on a regular grid, and - when no prompt file is given (the reference's own files do not travel to the GPU box)
- a generated prompt file with the shape of the reference's three: 1386 x 2 (cptac_cm), 660 x 2 (camelyon), 1782 x 4 (tcga_rcc)
entries of "CLASSNAME." templates. There is no tokenizer vocabulary offline, so texts are tokenised by a word-hash tokenizer
with the PubMedBERT conventions ([CLS] = 2, [SEP] = 3, [PAD] = 0, max_length 256); weights are random-init.
`--oracle-classifiers K` also times the REFERENCE flow (batch-1 encode_text per class, one small GEMM + topk + .item() per
classifier, Python dict walk) on the CPU oracle for the first K classifiers and extrapolates to the bank.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import sys
import tempfile
import time
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BANK_SHAPES = {"detection": (1386, ["Normal", "Tumor"]), "segmentation": (660, ["Tumor", "Normal"]),
               "subtyping": (1782, ["CCRCC", "PRCC", "CHRCC", "Normal"])}
ADJ = ["malignant", "benign", "invasive", "necrotic", "fibrotic", "dense", "atypical", "pleomorphic", "clear", "papillary",
       "chromophobe", "pigmented", "metastatic", "inflamed", "normal", "healthy", "stromal", "epithelial", "lymphoid", "renal"]
NOUN = ["tissue", "carcinoma", "melanoma", "tumor", "neoplasm", "parenchyma", "stroma", "cells", "lesion", "cortex", "nodule", "region"]


class HashTokenizer:
    """Word-hash stand-in for the HF tokenizer (same call signature as WSI_evaluation/utils.py:73)."""

    def __init__(self, vocab_size=30522):
        self.vocab_size = vocab_size

    def __call__(self, texts, max_length=256, padding="max_length", truncation=True, return_tensors="pt"):
        from transformers import BatchEncoding

        texts = [texts] if isinstance(texts, str) else texts
        ids = torch.zeros(len(texts), max_length, dtype=torch.long)
        mask = torch.zeros_like(ids)
        for i, t in enumerate(texts):
            w = [5 + zlib.crc32(x.encode()) % (self.vocab_size - 5) for x in re.findall(r"[a-z0-9]+|[^\sa-z0-9]", t.lower())]
            row = [2] + w[: max_length - 2] + [3]
            ids[i, : len(row)] = torch.tensor(row)
            mask[i, : len(row)] = 1
        return BatchEncoding({"input_ids": ids, "token_type_ids": torch.zeros_like(ids), "attention_mask": mask})


def synthetic_prompt_file(task: str, seed: int = 0) -> dict:
    k, classes = BANK_SHAPES[task]
    rng = np.random.default_rng(seed)
    out = {}
    for i in range(k):
        names = {}
        for c in classes:
            n_adj = int(rng.integers(1, 4))
            words = [ADJ[j] for j in rng.integers(0, len(ADJ), n_adj)] + [NOUN[int(rng.integers(0, len(NOUN)))]]
            names[c] = ("normal " if c == "Normal" else "") + " ".join(words)
        out[str(i)] = {"classnames": names, "templates": "CLASSNAME."}
    return out


def synthetic_slide(n_tiles: int, dim: int, seed: int, path: str) -> str:
    """A CLAM-style slide file: features [N, D] float32, coords [N, 2] int64 (stride 256 grid with a few duplicates)."""
    g = torch.Generator().manual_seed(seed)
    side = int(np.ceil(np.sqrt(n_tiles)))
    gy, gx = np.divmod(np.arange(n_tiles), side)
    coords = np.stack([gx * 256, gy * 256], 1).astype(np.int64)
    coords[-min(16, n_tiles):] = coords[: min(16, n_tiles)]  # re-visited coordinates: the first tile wins
    np.savez(path, features=torch.randn(n_tiles, dim, generator=g).numpy(), coords=coords)
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="all", choices=["detection", "subtyping", "segmentation", "all"])
    ap.add_argument("--tiles", type=int, default=10_000)
    ap.add_argument("--prompts", default=None, help="a reference prompt JSON (WSI_evaluation/prompts/*.json); default: generated")
    ap.add_argument("--topn", type=int, default=50)
    ap.add_argument("--oracle-classifiers", type=int, default=0)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    from keep_b200 import KEEPConfig, KEEPModel, io
    from keep_b200.slide import TASK_DEFAULTS, zero_shot_slide
    from keep_b200.weights import random_state_dict

    dev = torch.device("cuda", 0)
    cfg = KEEPConfig()
    with torch.device(dev):
        model = KEEPModel(cfg)
    model.load_state_dict(random_state_dict(cfg, seed=0, device=dev), strict=True)
    model.eval()
    KEEP = {"model": model, "tokenizer": HashTokenizer(), "transform": None}
    results = []
    for task in (["detection", "subtyping", "segmentation"] if args.task == "all" else [args.task]):
        prompts = json.load(open(args.prompts)) if args.prompts else synthetic_prompt_file(task)
        with tempfile.TemporaryDirectory() as td:
            feats, coords = io.load_slide(synthetic_slide(args.tiles, 768, 7, os.path.join(td, "slide.npz")))
        label_map = TASK_DEFAULTS[task]["label_map"]
        rec = {"task": task, "prompt_file": args.prompts or f"generated ({len(prompts)} entries, shape of the reference's file)",
               "tiles": args.tiles, "topn": args.topn}
        for rep in range(2):  # the second pass is the warm one that is reported
            timings = {}
            t0 = time.perf_counter()
            dfeats = io.to_device(feats, dev)
            torch.cuda.synchronize(dev)
            timings["upload_s"] = time.perf_counter() - t0
            out = zero_shot_slide(task, KEEP, prompts, dfeats, coords, dev, label_map=label_map, topn=args.topn, timings=timings)
            timings["total_s"] = time.perf_counter() - t0
        rec["gpu"] = timings
        rec["answer"] = (float(out) if task == "detection" else int(out) if task == "subtyping"
                         else {"tiles_kept": len(out), "mean_tumour_prob": float(np.mean(list(out.values())))})
        if args.oracle_classifiers > 0:
            rec["cpu_reference_flow"] = time_reference_flow(task, prompts, feats, coords, label_map, args.oracle_classifiers, args.topn)
        results.append(rec)
        print(json.dumps(rec))
    if args.out:
        json.dump(results, open(args.out, "w"), indent=1)


def time_reference_flow(task, prompts, feats, coords, label_map, k, topn):
    """The reference's own flow on the CPU (oracle classes) for the first k classifiers; per-stage seconds and the
    extrapolation to the whole prompt file (both stages are linear in the number of classifiers)."""
    from keep_b200.slide import TASK_DEFAULTS
    from oracle import keep_oracle as ko, wsi_oracle as wo  # CPU baseline leg only

    torch.set_num_threads(os.cpu_count() or 1)
    m = ko.KEEPModel(ko.DEFAULT_TEXT_CONFIG, 768, ko.DEFAULT_VISION_CONFIG).eval()
    m.load_state_dict(ko.synthetic_state_dict(m, seed=0))
    tok = HashTokenizer()
    d = TASK_DEFAULTS[task]
    t0 = time.perf_counter()
    bank = [wo.get_zeroshot_classifier(m, tok, label_map, prompts[str(i)], "cpu", add_normal=d["add_normal"]) for i in range(k)]
    t_bank = time.perf_counter() - t0
    t0 = time.perf_counter()
    ens, _ = wo.zero_shot_prompt_select(bank, feats, topn=min(topn, k))
    t_screen = time.perf_counter() - t0
    t0 = time.perf_counter()
    if task == "detection":
        wo.zero_shot_detection(ens, feats, coords.numpy(), patch_size=d["patch_size"], overlap=d["overlap"])
    elif task == "subtyping":
        wo.zero_shot_subtyping(ens, feats, coords.numpy(), patch_size=d["patch_size"], overlap=d["overlap"])
    else:
        _, probs = wo.tile_probs(ens, feats)
        wo.refine_seg_segment(probs.numpy(), coords.numpy(), patch_size=d["patch_size"], overlap=d["overlap"])
    t_head = time.perf_counter() - t0
    scale = len(prompts) / k
    return {"classifiers_timed": k, "cores": torch.get_num_threads(), "classifier_bank_s": t_bank, "screening_s": t_screen,
            "task_head_s": t_head, "extrapolated_total_s": t_bank * scale + t_screen * scale + t_head}


if __name__ == "__main__":
    main()
