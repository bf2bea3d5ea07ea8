import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keep_b200 import ops
dev = torch.device("cuda:0")
B, S, H = 512, 256, 12
qkv = torch.randn(B * S, 3 * H * 64, device=dev).half()
g = torch.Generator().manual_seed(0)
lens = torch.randint(4, 33, (B,), generator=g)
mask = (torch.arange(S)[None, :] < lens[:, None]).long().to(dev)
for name, m in (("no mask", None), ("prefix masks 4..32 of 256", mask)):
    for _ in range(3): ops.attention(qkv, B, S, H, key_mask=m)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.attention(qkv, B, S, H, key_mask=m)
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) * 50, "us")
