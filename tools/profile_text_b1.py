"""One batch-1 encode_text call (the reference's per-class call pattern, WSI_evaluation/utils.py:67-74) for an ncu launch list."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keep_b200 import KEEPConfig, KEEPModel  # noqa: E402
from keep_b200.weights import random_state_dict  # noqa: E402

dev = torch.device("cuda:0")
cfg = KEEPConfig()
with torch.device(dev):
    model = KEEPModel(cfg)
model.load_state_dict(random_state_dict(cfg, seed=0, device=dev))
model.eval()
n_tok = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ids = torch.zeros(1, 256, dtype=torch.long, device=dev)
ids[0, :n_tok] = torch.randint(5, 30000, (n_tok,), device=dev)
mask = (torch.arange(256, device=dev)[None, :] < n_tok).long()
text = {"input_ids": ids, "token_type_ids": torch.zeros_like(ids), "attention_mask": mask}
for _ in range(3):
    out = model.encode_text(text)
torch.cuda.synchronize()
