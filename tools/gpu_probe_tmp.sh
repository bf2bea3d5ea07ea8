timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -x -q -s -k "preprocess or raw_uint8 or uint8" 2>&1 | grep -v "^$" | tail -12
python - <<'PY'
import torch, time
from keep_b200.transform import preprocess
x = torch.randint(0, 256, (1024, 256, 256, 3), dtype=torch.uint8, device="cuda")
for _ in range(3): preprocess(x)
torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): preprocess(x)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print(f"preprocess 1024 x 256x256 -> 224x224: {ms*1e3:.0f} us = {1024/ms*1e3:.0f} tiles/s, {(1024*(256*256*3+224*224*3))/ms/1e6:.0f} GB/s algorithmic")
PY
