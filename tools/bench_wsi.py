"""The rows around the encoders at BASELINE sizes on one B200: tile x prompt similarity,
prompt screening and refine_seg, each timed with CUDA events around the C-ABI call, with its algorithmic bytes against
the measured HBM peak and the CPU oracle (the reference's own algorithm: oracle/wsi_oracle.py) timed beside it.
    python tools/bench_wsi.py > profiles/<round>_bench_wsi_rows.json
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from keep_b200 import ops  # noqa: E402
from oracle import wsi_oracle as wo  # noqa: E402  (CPU baseline leg only)

dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = peaks["hbm_gbs"]


def gpu_ms(fn, iters=10, warm=3):
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


def cpu_s(fn, reps=1):
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


rows = []
torch.set_num_threads(os.cpu_count() or 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > L2: written between timed calls of small inputs

# ---- a12: similarity + grouped softmax ------------------------------------------------------------------------------
for (N, P, G) in [(10_000, 32, 2), (50_000, 256, 4), (200_000, 2, 0)]:
    feats = torch.randn(N, 768, device=dev)
    cls = F.normalize(torch.randn(768, P, device=dev), dim=0)

    def run():
        flush.zero_() if N * 768 * 4 < (200 << 20) else None
        ops.similarity(feats, cls, group=G, temp=10.0)

    def run_flush_only():
        flush.zero_() if N * 768 * 4 < (200 << 20) else None

    ms = gpu_ms(run) - gpu_ms(run_flush_only)
    byts = (N * 768 + 768 * P) * 4 + 2 * N * P * 4
    fc, cc = feats.cpu(), cls.cpu()
    t_cpu = cpu_s(lambda: torch.softmax((F.normalize(fc, dim=-1) @ cc).view(N, -1, G or P) * 10, -1), reps=3)
    rows.append({"row": "a12 similarity + softmax(x10)", "N": N, "P": P, "group": G, "kernel": "sim_tc_kernel (TF32 tcgen05)",
                 "us": ms * 1e3, "algorithmic_bytes": byts, "achieved_gbs": byts / ms / 1e6, "hbm_peak_gbs": HBM,
                 "frac": byts / ms / 1e6 / HBM, "cpu_oracle_ms": t_cpu * 1e3, "cpu_threads": torch.get_num_threads()})

# ---- f1: prompt screening (utils.py:107-146): K classifiers x C classes over N tiles -----------------------------------
for (N, K, C) in [(10_000, 1386, 2), (50_000, 1782, 4)]:
    feats = torch.randn(N, 768, device=dev)
    cls = F.normalize(torch.randn(768, K * C, device=dev), dim=0)
    ms = gpu_ms(lambda: ops.prompt_scores(feats, cls, K, C), iters=5)
    fc = F.normalize(feats.cpu(), dim=-1)
    cl = [cls[:, k * C:(k + 1) * C].cpu() for k in range(K)]
    kk = min(K, 200)  # bounded sample of the reference's per-classifier loop, scaled up

    def cpu_loop():
        for k in range(kk):
            wo.rank_cls_score(fc @ cl[k])

    t_cpu = cpu_s(cpu_loop) * K / kk
    flops = 2.0 * N * 768 * K * C
    rows.append({"row": "f1 prompt screening", "N": N, "K": K, "C": C, "kernel": "sim_tc_kernel + prompt_score_kernel",
                 "ms": ms, "tflops_tf32": flops / ms / 1e9, "cpu_oracle_ms": t_cpu * 1e3,
                 "cpu_note": f"reference loop (one matmul + topk per classifier), {kk} of {K} classifiers timed and scaled"})

# ---- f2: refine_seg over a slide ---------------------------------------------------------------------------------------------
for (N, C) in [(10_000, 2), (200_000, 2), (200_000, 4)]:
    side = int(np.ceil(np.sqrt(N)))
    ys, xs = torch.meshgrid(torch.arange(side), torch.arange(side), indexing="ij")
    coords = (torch.stack([xs.reshape(-1), ys.reshape(-1)], 1)[:N] * 112)[torch.randperm(N)]
    probs = torch.softmax(torch.randn(N, C), 1)
    cd, pd = coords.to(dev), probs.to(dev)
    ms = gpu_ms(lambda: ops.refine(cd, pd, 224, True), iters=5)
    t_cpu = cpu_s(lambda: wo.refine_mean(probs.numpy(), coords.numpy(), 224, True))
    byts = N * (16 + 2 * C * 4 + 1)
    rows.append({"row": "f2 refine_seg (dedupe + 4-neighbour mean)", "N": N, "C": C, "kernel": "table_* + refine_kernel",
                 "us": ms * 1e3, "algorithmic_bytes": byts, "cpu_oracle_ms": t_cpu * 1e3,
                 "cpu_note": "reference dict walk (oracle/wsi_oracle.py::refine_mean), 1 thread"})

print(json.dumps({"device": torch.cuda.get_device_name(0), "rows": rows}, indent=1))
