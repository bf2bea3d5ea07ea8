"""Event timeline (clock64 deltas) of CTA 0 of the tcgen05 attention kernel, and the kernel's device time.

    python tools/trace_attention.py [B S H]        (default: the ViT layer shape 512 x 197 x 16)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keep_b200 import _lib, ops

dev = torch.device("cuda:0")
B, S, H = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (512, 197, 16)
qkv = torch.randn(B * S, 3 * H * 64, device=dev).half()
for _ in range(3):
    ops.attention(qkv, B, S, H)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attention(qkv, B, S, H)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
flops = 4.0 * B * H * S * S * 64
print(f"attention B={B} S={S} H={H}: {us:.1f} us per launch, {flops / us / 1e6:.0f} TFLOP/s (4*B*H*S*S*64 FLOP)")
buf = torch.zeros(64 * 16, dtype=torch.int64, device=dev)
L = _lib.lib()
L.keepb200_debug_attention_trace(buf.data_ptr())
ops.attention(qkv, B, S, H)
torch.cuda.synchronize()
L.keepb200_debug_attention_trace(None)
t = buf.view(64, 16).cpu()
t0 = int(t[0, 0])
names = ["S_issue", "PV_waited", "PV_issued", "sm_start", "sm_p1", "sm_baton", "sm_end", "out_start", "out_done"]
print("unit " + " ".join(n.rjust(10) for n in names))
for u in range(24):
    print(f"{u:4d} " + " ".join(str(int(t[u, e]) - t0).rjust(10) for e in range(len(names))))
d = (t[8:40] - t0).double()
print("steady state per unit (cycles): S_issue -> S_issue of the next unit", float((d[1:, 0] - d[:-1, 0]).mean()),
      " S_issue -> softmax start", float((d[:, 3] - d[:, 0]).mean()), " softmax (start->end)", float((d[:, 6] - d[:, 3]).mean()),
      " softmax end -> PV issued", float((d[:, 2] - d[:, 6]).mean()), " PV issued -> O drained", float((d[:, 8] - d[:, 2]).mean()))
