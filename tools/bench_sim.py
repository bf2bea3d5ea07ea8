"""Tile x prompt similarity (sim_tc_kernel) at the BASELINE shapes against the measured HBM peak, outputs and workspace
preallocated (the slide loop's use), inputs evicted from L2 between launches when they would fit in it.

    python tools/bench_sim.py            (KEEPB200_LIB=<other build> for an A/B inside one gpurun call)
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from keep_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
HBM = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6551.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def gpu_ms(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print(os.environ.get("KEEPB200_LIB", "in-tree library"))
L = _lib.lib()
for (N, P, G) in [(10_000, 32, 2), (50_000, 256, 4), (200_000, 2, 0), (50_000, 256, 2), (100_000, 64, 2)]:
    feats = torch.randn(N, 768, device=dev)
    cls = F.normalize(torch.randn(768, P, device=dev), dim=0)
    lg = torch.empty(N, P, device=dev)
    pr = torch.empty(N, P, device=dev)
    ws = torch.empty(max(L.keepb200_similarity_workspace_bytes(768, P), 16), dtype=torch.uint8, device=dev)
    need_flush = N * 768 * 4 < (200 << 20)
    t_flush = gpu_ms(lambda: flush.zero_()) if need_flush else 0.0
    for name, kw, out_bytes in (("logits+probs", dict(out_logits=lg, out_probs=pr), 2 * N * P * 4),
                                ("probs only", dict(want_logits=False, out_probs=pr), N * P * 4)):
        def run():
            if need_flush:
                flush.zero_()
            ops.similarity(feats, cls, group=G, temp=10.0, workspace=ws, **kw)

        ms = gpu_ms(run) - t_flush
        byts = (N * 768 + 768 * P) * 4 + out_bytes
        print(f"  {N:7d} x {P:3d} group {G} {name:13s}: {ms * 1e3:7.1f} us  {byts / ms / 1e6:6.0f} GB/s = {byts / ms / 1e6 / HBM:.2f} of {HBM:.0f}", flush=True)
