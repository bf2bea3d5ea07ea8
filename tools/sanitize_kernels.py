"""One small launch of every tcgen05 / TMA kernel of the library, for compute-sanitizer:

    compute-sanitizer --tool racecheck  python tools/sanitize_kernels.py      (shared-memory hazards between the warp roles)
    compute-sanitizer --tool synccheck  python tools/sanitize_kernels.py      (barrier misuse)
    compute-sanitizer --tool memcheck   python tools/sanitize_kernels.py

Shapes are small (the tools slow kernels down 10-100x) but cover every role pipeline: the CTA-pair GEMM (two tiles per pair:
both accumulator stages), the single-CTA GEMM in split-operand mode with the hi|lo GELU epilogue, both attention kernels
(single tile with the shared O accumulator and with O inside the unit's region; packed sequences; three key blocks), the
TF32 similarity kernel (single CTA with direct stores and small row tiles; CTA pairs with staging tiles + TMA stores), the
fused fp32 head."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keep_b200 import ops

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)

a, w = (rn(256 * 150, 128) * 0.5).half(), (rn(256, 128) * 0.1).half()          # 150 pair tiles on 74 pairs
bias = rn(256)
out = ops.gemm(a, w, ops.EPI_BIAS_HALF, bias=bias)
assert (out.float() - (a.float() @ w.float().T + bias)).abs().max() < 0.05
resid = rn(256 * 150, 256)
ops.gemm(a, w, ops.EPI_RESID_F32, bias=bias, gamma=torch.ones(256, device=dev), resid=resid.clone())
ops.gemm_resid_stats(a, w, resid.clone(), bias=bias)
x32, w32 = rn(300, 128), rn(256, 128) * 0.1
h = ops.gemm_split(ops.cast_hilo(x32), ops.cast_hilo(w32), 128, ops.EPI_BIAS_GELU_HILO, ops.SPLIT_AW, bias=bias)
assert (h[:, :256].float() + h[:, 256:].float() - F.gelu(x32 @ w32.T + bias)).abs().max() < 1e-3
# BALANCED level: hi|lo weights against plain 16-bit activations on the CTA-pair kernel (two passes over K), GELU epilogue
whl = ops.cast_hilo(w32)
a16 = (rn(256 * 150, 128) * 0.5).half()
h2 = ops.gemm_split(a16, whl, 128, ops.EPI_BIAS_GELU_HALF, ops.SPLIT_W, bias=bias)
assert (h2.float() - F.gelu(a16.float() @ w32.T + bias)).abs().max() < 5e-3
ab, wb = a[: 256 * 80 + 37].bfloat16(), w.bfloat16()                           # ragged row tiles, bf16 epilogue variants
ops.gemm(ab, wb, ops.EPI_BIAS_GELU_HALF, bias=bias)
for B, S, H, masked in ((4, 197, 2, False), (40, 12, 2, True), (2, 300, 2, True), (3, 100, 2, True), (3, 256, 2, True), (2, 240, 2, False)):
    qkv = rn(B * S, 3 * H * 64).half()
    mask = None
    if masked:
        lens = torch.randint(1, S + 1, (B,), device=dev, generator=g)
        mask = (torch.arange(S, device=dev)[None, :] < lens[:, None]).long()
    o = ops.attention(qkv, B, S, H, key_mask=mask, hilo=masked)
    assert torch.isfinite(o).all()
feats, cls = rn(2000, 768), F.normalize(rn(768, 32), dim=0)
lg, pr = ops.similarity(feats, cls, group=2, temp=10.0)
assert (lg - F.normalize(feats, dim=-1) @ cls).abs().max() < 1e-3
big, cls128 = rn(19_200, 768), F.normalize(rn(768, 128), dim=0)                 # 75 pair tiles of 256 rows on 74 pairs
lg, pr = ops.similarity(big, cls128, group=4, temp=10.0)
assert (lg - F.normalize(big, dim=-1) @ cls128).abs().max() < 1e-3 and (pr.view(-1, 32, 4).sum(-1) - 1).abs().max() < 1e-4
ops.prompt_scores(feats, F.normalize(rn(768, 64), dim=0), 16, 4, fused=True)
ops.visual_head(rn(20, 1024), torch.ones(1024, device=dev), torch.zeros(1024, device=dev), 1e-6, rn(768, 1024) / 32, rn(768), rn(768, 768) / 27, rn(768))
ops.pooler(rn(5, 768), rn(768, 768) / 27, rn(768))
xy = torch.tensor([[-1, -1], [223, 223], [-1, 223], [-225, -225], [-1, -1], [2 ** 31 - 2, 5], [-2 ** 31, 5]], device=dev)
keep, _ = ops.refine(xy, torch.rand(7, 2, device=dev, generator=g), 224, True)   # biased 32-bit keys, negative coordinates
assert keep.tolist() == [1, 1, 1, 1, 0, 1, 1]
torch.cuda.synchronize()
print("sanitize_kernels: all launches completed")
