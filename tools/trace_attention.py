"""Print the event timeline (clock64 deltas) of CTA 0 of the tcgen05 attention kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keep_b200 import _lib, ops

dev = torch.device("cuda:0")
B, S, H = 512, 197, 16
qkv = torch.randn(B * S, 3 * H * 64, device=dev).half()
for _ in range(2):
    ops.attention(qkv, B, S, H)
buf = torch.zeros(64 * 16, dtype=torch.int64, device=dev)
L = _lib.lib()
L.keepb200_debug_attention_trace(buf.data_ptr())
ops.attention(qkv, B, S, H)
torch.cuda.synchronize()
L.keepb200_debug_attention_trace(None)
t = buf.view(64, 16).cpu()
t0 = int(t[0, 0])
names = ["S_issue", "PV_waited", "PV_issued", "sm_start", "sm_p1", "sm_baton", "sm_p2", "out_start", "out_done"]
print("unit " + " ".join(n.rjust(10) for n in names))
for u in range(24):
    print(f"{u:4d} " + " ".join(str(int(t[u, e]) - t0).rjust(10) for e in range(9)))
d = t[8:40]
print("steady state per unit (cycles): S_issue->next S_issue(same region)", float((d[2:, 0] - d[:-2, 0]).float().mean()) / 2,
      " pass1", float((d[:, 4] - d[:, 3]).float().mean()), " baton wait", float((d[:, 5] - d[:, 4]).float().mean()),
      " pass2", float((d[:, 6] - d[:, 5]).float().mean()), " out", float((d[:, 8] - d[:, 7]).float().mean()))
