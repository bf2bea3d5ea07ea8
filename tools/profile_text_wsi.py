"""One pass of the kernels that the tile benchmark does not exercise, for `ncu` (tools/gpu_round.sh): the text tower on a
WSI-sized prompt set (high precision: packed tcgen05 attention, split-operand GEMMs, fused pooler) and on a padded S = 256
batch (single-tile attention with O inside the unit's TMEM region), the similarity / screening / refine kernels at BASELINE sizes, and the uint8 resize."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keep_b200 import KEEPConfig, KEEPModel, ops
from keep_b200.transform import preprocess
from keep_b200.weights import random_state_dict

dev = torch.device("cuda", 0)
cfg = KEEPConfig()
with torch.device(dev):
    model = KEEPModel(cfg)
model.load_state_dict(random_state_dict(cfg, seed=0, device=dev), strict=True)
# WSI kernels first, then the text tower: ncu captures the first N matching launches
gd = torch.Generator(device=dev).manual_seed(1)
feats = torch.randn(50_000, 768, device=dev, generator=gd)
cls = torch.nn.functional.normalize(torch.randn(768, 256, device=dev, generator=gd), dim=0)
ops.similarity(feats, cls, group=4, temp=10.0)
ops.similarity(feats, cls, group=4, temp=10.0, want_logits=False)
ops.similarity(feats[:10_000], cls[:, :32].contiguous(), group=2, temp=10.0, want_logits=False)
f200 = torch.randn(200_000, 768, device=dev, generator=gd)
_, p2 = ops.similarity(f200, cls[:, :2].contiguous(), group=2, temp=10.0, want_logits=False)
ops.prompt_scores(feats, torch.nn.functional.normalize(torch.randn(768, 1782 * 4, device=dev, generator=gd), dim=0), 1782, 4)
side = 448
coords = torch.stack(torch.meshgrid(torch.arange(side), torch.arange(side), indexing="ij"), -1).reshape(-1, 2)[:200_000].to(dev) * 112
ops.refine(coords, p2, 224, True)
preprocess(torch.randint(0, 256, (256, 256, 256, 3), device=dev, dtype=torch.uint8, generator=gd))
qkv256 = torch.randn(512 * 256, 3 * 12 * 64, device=dev, generator=gd).half()   # one padded-BERT attention layer (512 prompts, S = 256)
ops.attention(qkv256, 512, 256, 12)
del qkv256
g = torch.Generator().manual_seed(0)
P = 512
lens = torch.randint(4, 33, (P,), generator=g)
ids = torch.randint(5, 30522, (P, 256), generator=g)
mask = (torch.arange(256)[None, :] < lens[:, None]).long()
text = {"input_ids": (ids * mask).to(dev), "token_type_ids": torch.zeros_like(ids).to(dev), "attention_mask": mask.to(dev)}
emb = model.encode_text(text)                      # trimmed, high precision
model.trim_text = False
model.config.text_precision = "fast"
emb_padded = model.encode_text(text)               # padded S = 256, one-pass GEMMs
torch.cuda.synchronize()
print("ok", float(emb.sum()), float(emb_padded.sum()))
