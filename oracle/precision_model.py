"""Where does the 16-bit operand error of the product path come from? — TEST INFRASTRUCTURE ONLY (CPU, no product code).

    python oracle/precision_model.py [--bf16]

Emulates the product's arithmetic on the CPU oracle: every Linear sees operands rounded to the 16-bit operand type
(optionally "split": hi + lo parts, i.e. ~22 mantissa bits), attention sees 16-bit q/k/v and 16-bit probabilities,
everything else (accumulation, residual stream, LayerNorm, softmax, GELU) stays fp32 — then reports the rel-L2 of the
embeddings against the plain fp32 oracle for BASELINE config 1, for several choices of WHICH sites are rounded.
This is the model behind the precision policy in DESIGN.md section 2 (which GEMMs run split-operand).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import keep_oracle as ko  # noqa: E402
from oracle.make_golden import full_inputs  # noqa: E402


class Policy:
    """Which sites are rounded. `lin(name)` -> 'r' (rounded operands), 's' (split: exact to ~2^-22) ; attention flag."""

    def __init__(self, dt, lin_rule, attn_round=True, attn_out_split=False):
        self.dt, self.lin_rule, self.attn_round, self.attn_out_split = dt, lin_rule, attn_round, attn_out_split

    def r(self, x):
        return x.to(self.dt).float()

    def op(self, x, how):
        if how == "s":
            hi = x.to(self.dt).float()
            return hi + (x - hi).to(self.dt).float()
        if how == "r":
            return self.r(x)
        return x


def linear(pol, name, x, w, b):
    how = pol.lin_rule(name)
    if how == "w":  # weights split (exact to ~2^-22), activations rounded: two MMA passes, Ah*Wh + Ah*Wl
        return F.linear(pol.op(x, "r"), pol.op(w, "s"), b)
    if how == "a":  # activations split, weights rounded
        return F.linear(pol.op(x, "s"), pol.op(w, "r"), b)
    return F.linear(pol.op(x, how), pol.op(w, how), b)


def vit_forward(model, pol, tiles):
    v = model.visual
    B = tiles.shape[0]
    w = v.patch_embed.proj.weight
    patches = F.unfold(tiles, kernel_size=16, stride=16).transpose(1, 2)  # [B,196,768]
    x = linear(pol, "patch", patches, w.reshape(w.shape[0], -1), v.patch_embed.proj.bias)
    x = torch.cat([v.cls_token.expand(B, -1, -1), x], dim=1) + v.pos_embed
    depth = len(v.blocks)
    for i, blk in enumerate(v.blocks):
        tag = f"b{i}."
        h = F.layer_norm(x, (x.shape[-1],), blk.norm1.weight, blk.norm1.bias, blk.norm1.eps)
        qkv = linear(pol, tag + "qkv", h, blk.attn.qkv.weight, blk.attn.qkv.bias)
        if pol.attn_round:
            qkv = pol.r(qkv)
        Bq, N, _ = qkv.shape
        H = blk.attn.heads
        q, k, vv = qkv.reshape(Bq, N, 3, H, -1).permute(2, 0, 3, 1, 4).unbind(0)
        s = (q @ k.transpose(-1, -2)) * 0.125
        p = torch.softmax(s, dim=-1)
        if pol.attn_round:
            m = p.max(dim=-1, keepdim=True).values
            p = pol.r(p / m) * m  # the kernel rounds exp(s - ref), not the normalised probability
            p = p / p.sum(-1, keepdim=True) * 1.0
        ctx = (p @ vv).transpose(1, 2).reshape(Bq, N, -1)
        if pol.attn_round and not pol.attn_out_split:
            ctx = pol.r(ctx)
        if i == depth - 1:  # CLS rows only from here on
            x, ctx = x[:, :1], ctx[:, :1]
        x = x + blk.ls1.gamma * linear(pol, tag + "proj", ctx, blk.attn.proj.weight, blk.attn.proj.bias)
        h = F.layer_norm(x, (x.shape[-1],), blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
        h = F.gelu(linear(pol, tag + "fc1", h, blk.mlp.fc1.weight, blk.mlp.fc1.bias))
        x = x + blk.ls2.gamma * linear(pol, tag + "fc2", h, blk.mlp.fc2.weight, blk.mlp.fc2.bias)
    x = F.layer_norm(x, (x.shape[-1],), v.norm.weight, v.norm.bias, v.norm.eps)[:, 0]
    h0, h2 = model.visual_head[0], model.visual_head[2]
    y = linear(pol, "head0", x, h0.weight, h0.bias)
    y = linear(pol, "head2", F.gelu(y), h2.weight, h2.bias)
    return F.normalize(y, dim=-1)


def bert_forward(model, pol, text):
    t = model.text
    ids, mask = text["input_ids"], text["attention_mask"]
    S = int(mask.any(0).nonzero().max()) + 1
    ids, mask = ids[:, :S], mask[:, :S]
    e = t.embeddings
    x = e.word_embeddings(ids) + e.token_type_embeddings(torch.zeros_like(ids)) + e.position_embeddings.weight[:S]
    x = F.layer_norm(x, (x.shape[-1],), e.LayerNorm.weight, e.LayerNorm.bias, e.LayerNorm.eps)
    bias = torch.zeros(mask.shape, dtype=torch.float32).masked_fill(mask == 0, float("-inf"))[:, None, None, :]
    for i, L in enumerate(t.encoder.layer):
        a = L.attention.self
        tag = f"l{i}."
        q = linear(pol, tag + "q", x, a.query.weight, a.query.bias)
        k = linear(pol, tag + "k", x, a.key.weight, a.key.bias)
        v = linear(pol, tag + "v", x, a.value.weight, a.value.bias)
        if pol.attn_round:
            q, k, v = pol.r(q), pol.r(k), pol.r(v)
        P, _, D = q.shape
        H = a.num_attention_heads
        sp = lambda z: z.reshape(P, S, H, D // H).transpose(1, 2)
        s = (sp(q) @ sp(k).transpose(-1, -2)) * 0.125 + bias
        p = torch.softmax(s, dim=-1)
        if pol.attn_round:
            m = p.max(dim=-1, keepdim=True).values
            p = pol.r(p / m) * m
            p = p / p.sum(-1, keepdim=True)
        ctx = (p @ sp(v)).transpose(1, 2).reshape(P, S, D)
        if pol.attn_round and not pol.attn_out_split:
            ctx = pol.r(ctx)
        o = L.attention.output
        x = F.layer_norm(x + linear(pol, tag + "ao", ctx, o.dense.weight, o.dense.bias), (D,), o.LayerNorm.weight,
                         o.LayerNorm.bias, o.LayerNorm.eps)
        h = F.gelu(linear(pol, tag + "in", x, L.intermediate.dense.weight, L.intermediate.dense.bias))
        o2 = L.output
        x = F.layer_norm(x + linear(pol, tag + "out", h, o2.dense.weight, o2.dense.bias), (D,), o2.LayerNorm.weight,
                         o2.LayerNorm.bias, o2.LayerNorm.eps)
    pooled = torch.tanh(linear(pol, "pool", x[:, 0], t.pooler.dense.weight, t.pooler.dense.bias))
    return F.normalize(pooled, dim=-1)


def row_rel(a, b):
    return float(((a - b).double().norm(dim=1) / b.double().norm(dim=1)).max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bf16", action="store_true")
    ap.add_argument("--only", default="", help="run only the policies whose name contains this string")
    args = ap.parse_args()
    dt = torch.bfloat16 if args.bf16 else torch.float16
    torch.set_num_threads(os.cpu_count() or 1)
    model = ko.KEEPModel(ko.DEFAULT_TEXT_CONFIG, 768, ko.DEFAULT_VISION_CONFIG).eval()
    model.load_state_dict(ko.synthetic_state_dict(model, seed=0))
    g = np.load(os.path.join(ROOT, "tests", "golden", "keep_full.npz"))
    tiles, text = full_inputs(torch.from_numpy(g["example_tile_f16"]))
    with torch.no_grad():
        ref_i, ref_t = model.encode_image(tiles), model.encode_text(text)
        last = "b23."
        policies = {
            "all GEMMs + attention rounded (round-1 product)": Policy(dt, lambda n: "r"),
            "head + CLS tail of the last block split": Policy(dt, lambda n: "s" if n.startswith(("head", last + "proj", last + "fc")) else "r"),
            "... and fp32 head": Policy(dt, lambda n: "x" if n.startswith("head") else ("s" if n.startswith((last + "proj", last + "fc")) else "r")),
            "all GEMMs split, attention rounded": Policy(dt, lambda n: "s"),
            "all GEMMs split, attention rounded, context hi+lo": Policy(dt, lambda n: "s", attn_out_split=True),
            "only attention rounded, exact GEMMs": Policy(dt, lambda n: "x"),
            # two-pass candidates for a level between FAST and HIGH (the fp32 head and the split CLS tail are kept)
            "W split everywhere, A rounded (2 passes)": Policy(dt, lambda n: "x" if n.startswith("head") else "w"),
            "A split everywhere, W rounded (2 passes)": Policy(dt, lambda n: "x" if n.startswith("head") else "a"),
            "W split in fc1/fc2 only": Policy(dt, lambda n: "x" if n.startswith("head") else ("w" if ".fc" in n else "r")),
            "W split in fc2 only": Policy(dt, lambda n: "x" if n.startswith("head") else ("w" if ".fc2" in n else "r")),
            "fc1/fc2 fully split, rest rounded": Policy(dt, lambda n: "x" if n.startswith("head") else ("s" if ".fc" in n else "r")),
        }
        if args.only:
            policies = {k: v for k, v in policies.items() if args.only in k}
        for name, pol in policies.items():
            ri = row_rel(vit_forward(model, pol, tiles), ref_i)
            rt = row_rel(bert_forward(model, pol, text), ref_t)
            print(f"{name:58s} image rel-L2 {ri:.2e}   text rel-L2 {rt:.2e}")


if __name__ == "__main__":
    main()
