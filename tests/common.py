"""Shared builders for the tests: seeded inputs identical to oracle/make_golden.py, model pairs, metrics."""
from __future__ import annotations

import numpy as np
import torch

from oracle import keep_oracle as ko
from oracle.make_golden import full_inputs, tiny_inputs, weight_checksum, wsi_inputs  # noqa: F401  (same seeds)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


def row_metrics(got: torch.Tensor, ref: torch.Tensor):
    """max over rows of rel-L2 and min cosine (the parity gate of SURVEY.md §8d)."""
    got, ref = got.double().cpu(), ref.double().cpu()
    rl = ((got - ref).norm(dim=1) / ref.norm(dim=1)).max().item()
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=1).min().item()
    return rl, cos


def tiny_oracle(seed: int = 1, max_pos: int = 64):
    text = dict(ko.TINY_TEXT_CONFIG, max_position_embeddings=max_pos)
    m = ko.KEEPModel(text, 128, ko.TINY_VISION_CONFIG).eval()
    sd = ko.synthetic_state_dict(m, seed=seed)
    m.load_state_dict(sd)
    return m, sd, text


def tiny_product(sd, text_cfg, operand_dtype="float16", device="cuda:0"):
    from keep_b200 import KEEPConfig, KEEPModel

    cfg = KEEPConfig(text_config=text_cfg, vision_config=ko.TINY_VISION_CONFIG, projection_dim=128,
                     operand_dtype=operand_dtype)
    m = KEEPModel(cfg)
    m.load_state_dict(sd, strict=True)
    return m.to(device).eval()


def full_oracle(seed: int = 0):
    m = ko.KEEPModel(ko.DEFAULT_TEXT_CONFIG, 768, ko.DEFAULT_VISION_CONFIG).eval()
    sd = ko.synthetic_state_dict(m, seed=seed)
    m.load_state_dict(sd)
    return m, sd


def full_product(sd, operand_dtype="float16", device="cuda:0"):
    from keep_b200 import KEEPConfig, KEEPModel

    m = KEEPModel(KEEPConfig(text_config=ko.DEFAULT_TEXT_CONFIG, projection_dim=768, operand_dtype=operand_dtype))
    m.load_state_dict(sd, strict=True)
    return m.to(device).eval()


def to_device(text: dict, device):
    return {k: v.to(device) for k, v in text.items()}


def load_golden(golden_dir, name):
    return np.load(f"{golden_dir}/{name}")
