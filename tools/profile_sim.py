"""One launch of sim_tc_kernel per BASELINE shape (after an L2 flush) for `ncu -k regex:sim_tc`."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keep_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (N, P, G, logits) in [(10_000, 32, 2, True), (50_000, 256, 4, True), (50_000, 256, 4, False), (200_000, 2, 0, True)]:
    feats = torch.randn(N, 768, device=dev)
    cls = F.normalize(torch.randn(768, P, device=dev), dim=0)
    for _ in range(2):
        flush.zero_()
        ops.similarity(feats, cls, group=G, temp=10.0, want_logits=logits)
    torch.cuda.synchronize()
