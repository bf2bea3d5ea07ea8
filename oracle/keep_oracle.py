"""CPU oracle for the KEEP zero-shot inference hot path — TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product path
(`keep_b200/`) never does and fails loudly when its CUDA library is missing.

What is restated, and from where (paths relative to /root/reference):

* `KEEPConfig`, `KEEPModel.{__init__, encode_image, encode_text, forward}`
      quick_start/keep_inference.py:9-22, 25-73  (fp32, no autocast: SURVEY.md D5)
* the vision tower the reference obtains from the UN-VENDORED dependency `timm==1.0.15`
  (training/requirements.txt:13) through
      timm.create_model("vit_large_patch16_224", img_size=224, patch_size=16, init_values=1e-5,
                        num_classes=0, dynamic_img_size=True)      quick_start/keep_inference.py:32-40
  timm's source is not on disk, so `VisionTransformer` below restates its published algorithm for exactly
  those kwargs: Conv2d(3,D,16,16) patch embed -> [cls | patches] + learned pos_embed ->
  depth x pre-LN blocks with LayerScale (x + g1*attn(LN(x)); x + g2*mlp(LN(x))), qkv bias, SDPA scale
  dh^-0.5, exact-erf GELU, LayerNorm eps 1e-6 -> final LayerNorm -> CLS token. State-dict keys are timm's.
* the text tower is the real `transformers.BertModel` (installed 5.5.0; reference pins 4.34.0 — same math).

Pinning (tests/test_oracle.py, oracle/make_golden.py):
  - the reference file's own class body (lines 1-76, exec'd verbatim with this module's timm shim injected)
    must agree bit-for-bit with `KEEPModel` here on the same state-dict;
  - the ViT restatement must agree with the independent implementation `torchvision.models.vit_l_16`
    (weights remapped, LayerScale = 1) to fp32 round-off;
  - parameter count 303,350,784 (ViT-L/16) and the 546-key state-dict.
The reference ships NO golden vectors or known-answer tests for this path, so numerical
parity is pinned to the reference's own code run here, not to published outputs.
"""
from __future__ import annotations

import math
import sys
import types
from typing import Mapping

import torch
import torch.nn as nn
import torch.nn.functional as F

# BERT geometry of PubMedBERT-base, which training/README.md:27 names as the text encoder. The real
# text_config lives in the external HF config.json (keep_inference.py:49,80), so everything is config-driven.
DEFAULT_TEXT_CONFIG = dict(
    vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
    max_position_embeddings=512, type_vocab_size=2, hidden_act="gelu", layer_norm_eps=1e-12,
    hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
)
DEFAULT_VISION_CONFIG = dict(img_size=224, patch_size=16, width=1024, depth=24, heads=16, mlp=4096, ln_eps=1e-6)

TINY_TEXT_CONFIG = dict(DEFAULT_TEXT_CONFIG, vocab_size=1000, hidden_size=128, num_hidden_layers=2,
                        num_attention_heads=2, intermediate_size=256, max_position_embeddings=64)
TINY_VISION_CONFIG = dict(img_size=224, patch_size=16, width=128, depth=2, heads=2, mlp=256, ln_eps=1e-6)


# ------------------------------------------------------------------------------------------------
# vision tower (restated timm VisionTransformer for the kwargs at keep_inference.py:32-40)
# ------------------------------------------------------------------------------------------------
class _PatchEmbed(nn.Module):
    def __init__(self, patch, width):
        super().__init__()
        self.proj = nn.Conv2d(3, width, kernel_size=patch, stride=patch, bias=True)

    def forward(self, x):  # [B,3,H,W] -> [B,gh*gw,D] (row-major over the patch grid)
        return self.proj(x).flatten(2).transpose(1, 2)


class _Attention(nn.Module):
    def __init__(self, width, heads):
        super().__init__()
        self.heads = heads
        self.qkv = nn.Linear(width, 3 * width, bias=True)
        self.proj = nn.Linear(width, width)

    def forward(self, x):
        B, N, D = x.shape
        q, k, v = self.qkv(x).reshape(B, N, 3, self.heads, D // self.heads).permute(2, 0, 3, 1, 4).unbind(0)
        y = F.scaled_dot_product_attention(q, k, v)  # scale = dh ** -0.5, no mask
        return self.proj(y.transpose(1, 2).reshape(B, N, D))


class _Mlp(nn.Module):
    def __init__(self, width, hidden):
        super().__init__()
        self.fc1 = nn.Linear(width, hidden)
        self.act = nn.GELU()  # exact erf
        self.fc2 = nn.Linear(hidden, width)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _LayerScale(nn.Module):
    def __init__(self, width, init):
        super().__init__()
        self.gamma = nn.Parameter(init * torch.ones(width))

    def forward(self, x):
        return x * self.gamma


class _Block(nn.Module):
    def __init__(self, width, heads, hidden, init_values, eps):
        super().__init__()
        self.norm1 = nn.LayerNorm(width, eps=eps)
        self.attn = _Attention(width, heads)
        self.ls1 = _LayerScale(width, init_values)
        self.norm2 = nn.LayerNorm(width, eps=eps)
        self.mlp = _Mlp(width, hidden)
        self.ls2 = _LayerScale(width, init_values)

    def forward(self, x):
        x = x + self.ls1(self.attn(self.norm1(x)))
        return x + self.ls2(self.mlp(self.norm2(x)))


class VisionTransformer(nn.Module):
    """timm `vit_large_patch16_224(..., init_values=1e-5, num_classes=0, dynamic_img_size=True)` restated."""

    def __init__(self, img_size=224, patch_size=16, width=1024, depth=24, heads=16, mlp=4096, ln_eps=1e-6,
                 init_values=1e-5):
        super().__init__()
        self.num_features = self.embed_dim = width
        self.grid = img_size // patch_size
        self.patch_size = patch_size
        self.patch_embed = _PatchEmbed(patch_size, width)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, width))
        self.pos_embed = nn.Parameter(torch.randn(1, self.grid * self.grid + 1, width) * 0.02)
        self.blocks = nn.Sequential(*[_Block(width, heads, mlp, init_values, ln_eps) for _ in range(depth)])
        self.norm = nn.LayerNorm(width, eps=ln_eps)

    def _pos(self, gh, gw):
        if (gh, gw) == (self.grid, self.grid):
            return self.pos_embed
        # dynamic_img_size: bicubic, antialiased resample of the grid part of pos_embed, prefix token kept
        prefix, grid = self.pos_embed[:, :1], self.pos_embed[:, 1:]
        grid = grid.reshape(1, self.grid, self.grid, -1).permute(0, 3, 1, 2).float()
        grid = F.interpolate(grid, size=(gh, gw), mode="bicubic", antialias=True)
        return torch.cat([prefix, grid.permute(0, 2, 3, 1).reshape(1, gh * gw, -1)], dim=1)

    def forward(self, x):
        B, _, H, W = x.shape
        x = self.patch_embed(x)
        x = torch.cat([self.cls_token.expand(B, -1, -1), x], dim=1) + self._pos(H // self.patch_size, W // self.patch_size)
        x = self.norm(self.blocks(x))
        return x[:, 0]  # global_pool='token'; fc_norm and head are Identity for num_classes=0


def _create_model(name, pretrained=False, img_size=224, patch_size=16, init_values=None, num_classes=1000,
                  dynamic_img_size=False, **kw):
    if name != "vit_large_patch16_224" or pretrained or num_classes != 0:
        raise NotImplementedError(f"timm shim only serves the KEEP call (got {name!r}, num_classes={num_classes})")
    return VisionTransformer(img_size=img_size, patch_size=patch_size, init_values=init_values, **kw)


def install_timm_shim():
    """Make `import timm; timm.create_model(...)` resolve to the restatement (timm is not installed here)."""
    if "timm" in sys.modules:  # a real timm, or the shim installed earlier (module identity must be stable)
        return sys.modules["timm"]
    mod = types.ModuleType("timm")
    mod.create_model = _create_model
    mod._keep_oracle_shim = True
    sys.modules["timm"] = mod
    return mod


# ------------------------------------------------------------------------------------------------
# KEEPModel (quick_start/keep_inference.py:25-73), as a plain nn.Module
# ------------------------------------------------------------------------------------------------
class KEEPModel(nn.Module):
    def __init__(self, text_config: Mapping | None = None, projection_dim: int = 768,
                 vision_config: Mapping | None = None):
        super().__init__()
        from transformers import BertConfig, BertModel

        vc = dict(DEFAULT_VISION_CONFIG, **(vision_config or {}))
        self.visual = VisionTransformer(init_values=1e-5, **vc)                      # :32-40
        self.visual_head = nn.Sequential(                                            # :42-46
            nn.Linear(self.visual.num_features, projection_dim), nn.GELU(), nn.Linear(projection_dim, projection_dim))
        self.text = BertModel(BertConfig(**dict(text_config or DEFAULT_TEXT_CONFIG)))  # :49-50
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.04))         # :52

    def encode_image(self, image_inputs):                                            # :54-58
        return F.normalize(self.visual_head(self.visual(image_inputs)), dim=-1)

    def encode_text(self, text_inputs):                                              # :60-62
        return F.normalize(self.text(**text_inputs).pooler_output, dim=-1)

    def forward(self, image_inputs, text_inputs):                                    # :65-73
        return {"vision_features": self.encode_image(image_inputs), "text_features": self.encode_text(text_inputs)}


# ------------------------------------------------------------------------------------------------
# seeded synthetic weights (no checkpoint is available offline)
# ------------------------------------------------------------------------------------------------
def synthetic_state_dict(model: nn.Module, seed: int = 0) -> dict:
    """Deterministic, non-degenerate weights for every tensor of `model.state_dict()`.

    LayerScale gamma ~ U(0.05, 0.5) instead of the 1e-5 init: with 1e-5 every block contributes ~1e-5 and a
    broken attention/MLP kernel would still pass parity. LayerNorm weights 1 +- 0.1, biases small, linear
    weights N(0, 1/sqrt(fan_in)) so activations stay O(1) through 24 layers.
    """
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, ref in model.state_dict().items():
        shape = tuple(ref.shape)
        if not ref.dtype.is_floating_point:  # e.g. position_ids buffers
            out[name] = ref.clone()
            continue
        leaf = name.rsplit(".", 1)[-1]
        if name.endswith("logit_scale"):
            t = torch.tensor(math.log(1 / 0.04))
        elif "gamma" in leaf:
            t = torch.rand(shape, generator=g) * 0.45 + 0.05
        elif "norm" in name.lower() and leaf == "weight":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = 0.02 * torch.randn(shape, generator=g)
        elif leaf in ("cls_token", "pos_embed"):
            t = 0.02 * torch.randn(shape, generator=g) if leaf == "cls_token" else 0.1 * torch.randn(shape, generator=g)
        elif "embeddings" in name:  # BERT embedding tables
            t = 0.05 * torch.randn(shape, generator=g)
        elif len(shape) >= 2:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        out[name] = t.to(torch.float32)
    return out


def synthetic_text_inputs(n: int, seq_len: int = 256, vocab: int = 30522, seed: int = 0, min_len: int = 4,
                          max_len: int = 32) -> dict:
    """Tokenizer-shaped inputs without a vocabulary file: [CLS]=2 ... [SEP]=3, [PAD]=0 (PubMedBERT convention),
    lengths U{min_len..max_len}, padding='max_length' (keep_inference.py:99)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros(n, seq_len, dtype=torch.long)
    mask = torch.zeros(n, seq_len, dtype=torch.long)
    lens = torch.randint(min_len, min(max_len, seq_len) + 1, (n,), generator=g)
    for i, L in enumerate(lens.tolist()):
        ids[i, 0] = 2
        ids[i, 1:L - 1] = torch.randint(5, vocab, (L - 2,), generator=g)
        ids[i, L - 1] = 3
        mask[i, :L] = 1
    return {"input_ids": ids, "token_type_ids": torch.zeros_like(ids), "attention_mask": mask}
