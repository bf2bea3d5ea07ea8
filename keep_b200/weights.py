"""State-dict contract of KEEPModel: names and shapes of the 546 reference tensors.

The names are those of `KEEPModel.state_dict()` in the reference (quick_start/keep_inference.py:28-52):
`visual.*` from timm's VisionTransformer, `visual_head.{0,2}.*`, `logit_scale`, `text.*` from
transformers' BertModel.  The C library holds the same table
(keepb200_num_weights / keepb200_weight_name); a GPU test checks that the two agree.
"""
from __future__ import annotations

from collections import OrderedDict


def state_dict_spec(config) -> "OrderedDict[str, tuple]":
    v, t = config.vision(), config.text()
    D, F, depth, ps = v["width"], v["mlp"], v["depth"], v["patch_size"]
    T = (v["img_size"] // ps) ** 2 + 1
    proj = config.projection_dim
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    spec["logit_scale"] = ()
    spec["visual.cls_token"] = (1, 1, D)
    spec["visual.pos_embed"] = (1, T, D)
    spec["visual.patch_embed.proj.weight"] = (D, 3, ps, ps)
    spec["visual.patch_embed.proj.bias"] = (D,)
    for i in range(depth):
        p = f"visual.blocks.{i}."
        spec[p + "norm1.weight"] = (D,)
        spec[p + "norm1.bias"] = (D,)
        spec[p + "attn.qkv.weight"] = (3 * D, D)
        spec[p + "attn.qkv.bias"] = (3 * D,)
        spec[p + "attn.proj.weight"] = (D, D)
        spec[p + "attn.proj.bias"] = (D,)
        spec[p + "ls1.gamma"] = (D,)
        spec[p + "norm2.weight"] = (D,)
        spec[p + "norm2.bias"] = (D,)
        spec[p + "mlp.fc1.weight"] = (F, D)
        spec[p + "mlp.fc1.bias"] = (F,)
        spec[p + "mlp.fc2.weight"] = (D, F)
        spec[p + "mlp.fc2.bias"] = (D,)
        spec[p + "ls2.gamma"] = (D,)
    spec["visual.norm.weight"] = (D,)
    spec["visual.norm.bias"] = (D,)
    spec["visual_head.0.weight"] = (proj, D)
    spec["visual_head.0.bias"] = (proj,)
    spec["visual_head.2.weight"] = (proj, proj)
    spec["visual_head.2.bias"] = (proj,)
    d, I = t["hidden_size"], t["intermediate_size"]
    spec["text.embeddings.word_embeddings.weight"] = (t["vocab_size"], d)
    spec["text.embeddings.position_embeddings.weight"] = (t["max_position_embeddings"], d)
    spec["text.embeddings.token_type_embeddings.weight"] = (t["type_vocab_size"], d)
    spec["text.embeddings.LayerNorm.weight"] = (d,)
    spec["text.embeddings.LayerNorm.bias"] = (d,)
    for i in range(t["num_hidden_layers"]):
        p = f"text.encoder.layer.{i}."
        for nm in ("query", "key", "value"):
            spec[p + f"attention.self.{nm}.weight"] = (d, d)
            spec[p + f"attention.self.{nm}.bias"] = (d,)
        spec[p + "attention.output.dense.weight"] = (d, d)
        spec[p + "attention.output.dense.bias"] = (d,)
        spec[p + "attention.output.LayerNorm.weight"] = (d,)
        spec[p + "attention.output.LayerNorm.bias"] = (d,)
        spec[p + "intermediate.dense.weight"] = (I, d)
        spec[p + "intermediate.dense.bias"] = (I,)
        spec[p + "output.dense.weight"] = (d, I)
        spec[p + "output.dense.bias"] = (d,)
        spec[p + "output.LayerNorm.weight"] = (d,)
        spec[p + "output.LayerNorm.bias"] = (d,)
    spec["text.pooler.dense.weight"] = (d, d)
    spec["text.pooler.dense.bias"] = (d,)
    return spec


# buffers that some transformers versions serialise with BertModel; accepted and ignored on load
IGNORED_BUFFERS = ("text.embeddings.position_ids", "text.embeddings.token_type_ids")


def random_state_dict(config, seed: int = 0, device="cpu"):
    """Random-init weights of the right architecture for benchmarks and smoke runs (no checkpoint is available
    offline): linear weights N(0, 1/fan_in), LayerNorm weights 1 +- 0.1, LayerScale U(0.05, 0.5) so that every
    block contributes, small biases/embeddings. Generated directly on `device`."""
    import math

    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    out = {}
    for name, shape in state_dict_spec(config).items():
        leaf = name.rsplit(".", 1)[-1]
        if name == "logit_scale":
            t = torch.tensor(math.log(1 / 0.04), device=device)
        elif leaf == "gamma":
            t = torch.rand(shape, generator=g, device=device) * 0.45 + 0.05
        elif "norm" in name.lower() and leaf == "weight":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=device)
        elif leaf == "bias":
            t = 0.02 * torch.randn(shape, generator=g, device=device)
        elif leaf == "cls_token":
            t = 0.02 * torch.randn(shape, generator=g, device=device)
        elif leaf == "pos_embed":
            t = 0.1 * torch.randn(shape, generator=g, device=device)
        elif "embeddings" in name:
            t = 0.05 * torch.randn(shape, generator=g, device=device)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g, device=device) / math.sqrt(fan_in)
        out[name] = t.float()
    return out
