"""Experiment: two half-size persistent grids from two streams side by side (attention / LayerNorm phases of one
stream under the GEMM phases of the other) against the normal one-stream, all-SM schedule.
    KEEPB200_SMS=74 python tools/exp_two_streams.py 2      # two streams, 74-SM grids
    python tools/exp_two_streams.py 1                      # baseline
"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keep_b200 import KEEPConfig, KEEPModel
from keep_b200.weights import random_state_dict

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
cfg = KEEPConfig()
sd = random_state_dict(cfg, seed=0, device=dev)
models = []
for _ in range(ns):
    with torch.device(dev):
        m = KEEPModel(cfg)
    m.load_state_dict(sd, strict=True)
    m.eval()
    m.image_chunk = chunk
    models.append(m)
N = 10240
tiles = torch.randn(N, 3, 224, 224, device=dev)
streams = [torch.cuda.Stream(dev) for _ in range(ns)]
outs = [None] * (N // chunk)

def step():
    for i, b0 in enumerate(range(0, N, chunk)):
        k = i % ns
        with torch.cuda.stream(streams[k]):
            outs[i] = models[k].encode_image(tiles[b0:b0 + chunk])

for _ in range(2):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print(f"streams={ns} chunk={chunk} KEEPB200_SMS={os.environ.get('KEEPB200_SMS')}: {N / dt:.0f} tiles/s ({dt * 1e3:.1f} ms/step)")
