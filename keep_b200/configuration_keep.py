"""KEEPConfig — same constructor and `model_type` as the reference (quick_start/keep_inference.py:9-22)."""
from __future__ import annotations

from transformers import PretrainedConfig

# PubMedBERT-base geometry (training/README.md:27); the released text_config lives in the HF config.json
# (keep_inference.py:49,80) and overrides these defaults key by key.
DEFAULT_TEXT_CONFIG = dict(
    vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
    max_position_embeddings=512, type_vocab_size=2, hidden_act="gelu", layer_norm_eps=1e-12,
)
# What timm.create_model("vit_large_patch16_224", img_size=224, patch_size=16, init_values=1e-5, num_classes=0)
# builds (keep_inference.py:32-40). The reference ignores config.vision_config; so do we unless a dict with these
# keys is given (used by the small configurations in the tests).
DEFAULT_VISION_CONFIG = dict(img_size=224, patch_size=16, width=1024, depth=24, heads=16, mlp=4096, ln_eps=1e-6)


class KEEPConfig(PretrainedConfig):
    model_type = "keep"

    def __init__(self, vision_config=None, text_config=None, projection_dim=768, operand_dtype="float16",
                 text_precision="auto", image_precision="auto", **kwargs):
        super().__init__(**kwargs)
        self.vision_config = vision_config
        self.text_config = text_config
        self.projection_dim = projection_dim
        # keep_b200 extensions.  operand_dtype: 16-bit type of the tensor-core operands ("float16" | "bfloat16").
        # text_precision / image_precision (include/keep_b200.h KEEPB200_PRECISION_*): "high" = split-operand GEMMs
        # through the whole tower (hi + lo 16-bit pairs, three MMA passes: ~3e-4 rel-L2 against the fp32 reference),
        # "fast" = one pass (~1.0-1.4e-3: the fp16 operand rounding; the throughput path), "balanced" = two passes (hi + lo
        # weights, activations rounded once: ~7e-4, inside the 1e-3 target at twice the MMA work), "auto" = high for calls of up
        # to 8192 prompts (every WSI classifier bank) / 16 tiles (quick-start use), fast for anything larger.
        self.operand_dtype = operand_dtype
        self.text_precision = text_precision
        self.image_precision = image_precision

    # resolved geometry ---------------------------------------------------------------------------
    def vision(self) -> dict:
        vc = dict(DEFAULT_VISION_CONFIG)
        if isinstance(self.vision_config, dict):
            vc.update({k: v for k, v in self.vision_config.items() if k in vc})
        return vc

    def text(self) -> dict:
        tc = dict(DEFAULT_TEXT_CONFIG)
        if isinstance(self.text_config, dict):
            tc.update(self.text_config)
        act = tc.get("hidden_act", "gelu")
        if act != "gelu":
            raise ValueError(f"keep_b200 implements BERT with exact-erf GELU only (hidden_act={act!r})")
        if tc.get("position_embedding_type", "absolute") != "absolute":
            raise ValueError("keep_b200 implements absolute position embeddings only")
        return tc
