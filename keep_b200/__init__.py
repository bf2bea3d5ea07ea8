"""keep_b200 — B200-native implementation of KEEP's zero-shot WSI inference hot path.

Public surface (mirrors the reference, MAGIC-AI4Med/KEEP):
    KEEPConfig, KEEPModel                       quick_start/keep_inference.py:9-76
    keep_b200.wsi.*                             WSI_evaluation/{utils,detection_utils,subtyping_utils,segment_utils}.py
    keep_b200.distributed.*                     tile sharding + all-gather (new; the reference is single-GPU)
All arithmetic runs in libkeep_b200.so (include/keep_b200.h); see DESIGN.md.
"""
from .configuration_keep import KEEPConfig  # noqa: F401
from .modeling_keep import KEEPModel  # noqa: F401
from ._lib import KeepB200Error  # noqa: F401

__all__ = ["KEEPConfig", "KEEPModel", "KeepB200Error"]
