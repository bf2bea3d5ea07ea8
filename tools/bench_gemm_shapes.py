"""Sustained throughput of the ViT layer GEMMs by (shape, epilogue): each case runs long enough (default 300 launches,
~0.2 s) for the power cap to settle, so the numbers compare with the in-step `roofline.per_shape` of bench.py.

    python tools/bench_gemm_shapes.py [iters]            (KEEPB200_LIB=<other build> for an A/B in the same gpurun call)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keep_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
try:  # SM clock and board power while a case runs (the step is power-capped: tells energy-bound from time-bound)
    import threading
    import time

    import pynvml

    pynvml.nvmlInit()
    _h = pynvml.nvmlDeviceGetHandleByIndex(0)

    class Sampler:
        def __enter__(self):
            self.clk, self.pw, self.stop = [], [], False

            def loop():
                while not self.stop:
                    self.clk.append(pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM))
                    self.pw.append(pynvml.nvmlDeviceGetPowerUsage(_h) / 1000.0)
                    time.sleep(0.01)

            self.t = threading.Thread(target=loop)
            self.t.start()
            return self

        def __exit__(self, *a):
            self.stop = True
            self.t.join()

        def text(self):
            c, w = sorted(self.clk), sorted(self.pw)
            return f"clk {c[len(c) // 2]} MHz  {w[len(w) // 2]:.0f} W ({len(c)} samples)" if c else ""
except Exception:  # noqa: BLE001
    class Sampler:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            pass

        def text(self):
            return ""
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
M = 197 * 512
CASES = [  # (name, N, K, epi)
    ("fc1-shape bias->16bit", 4096, 1024, 0),
    ("fc1 bias+gelu->16bit", 4096, 1024, 1),
    ("qkv-shape bias->16bit", 3072, 1024, 0),
    ("proj resid fp32", 1024, 1024, 2),
    ("fc2 resid fp32", 1024, 4096, 2),
    ("fc2 resid+stats", 1024, 4096, 5),
]
print(os.environ.get("KEEPB200_LIB", "in-tree library"))
for name, N, K, epi in CASES:
    a = (torch.randn(M, K, device=dev) * 0.5).half()
    w = (torch.randn(N, K, device=dev) * 0.05).half()
    bias = torch.zeros(N, device=dev)
    x = torch.zeros(M, N, device=dev) if epi in (2, 5) else None

    def run():
        if epi == 5:
            ops.gemm_resid_stats(a, w, x, bias=bias)
        elif epi == 2:
            ops.gemm(a, w, 2, bias=bias, resid=x, out=x)
        else:
            run.out = ops.gemm(a, w, epi, bias=bias, out=getattr(run, "out", None))

    for _ in range(20):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with Sampler() as sm:
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"  {name:26s} M={M} N={N} K={K}: {ms * 1e3:7.1f} us  {2.0 * M * N * K / ms / 1e9:6.0f} TFLOP/s  {sm.text()}", flush=True)
    del a, w, x
    if hasattr(run, "out"):
        del run.out
