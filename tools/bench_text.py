"""BASELINE config 5: encode_text over the prompt bank (11,454 disease names x 8 templates = 91,632 prompts,
S = 256 padded, lengths U{4..32}, ids U{5..30521}, seed 3000) on one B200. Reports prompts/s for the reference's
padded computation (every position computed) and for the exact trimmed computation (positions masked in every
row skipped), plus the fraction of the BERT tensor roofline (45.904 GFLOP/prompt at S=256)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keep_b200 import KEEPConfig, KEEPModel  # noqa: E402
from keep_b200.weights import random_state_dict  # noqa: E402

dev = torch.device("cuda:0")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 91632
cfg = KEEPConfig()
with torch.device(dev):
    model = KEEPModel(cfg)
model.load_state_dict(random_state_dict(cfg, seed=0, device=dev))
model.eval()
g = torch.Generator().manual_seed(3000)
lens = torch.randint(4, 33, (P,), generator=g)
ids = torch.zeros(P, 256, dtype=torch.long)
rnd = torch.randint(5, 30522, (P, 32), generator=g)
pos = torch.arange(32)[None, :]
ids[:, :32] = torch.where(pos < lens[:, None], rnd, torch.zeros_like(rnd))
ids[:, 0] = 2
ids[torch.arange(P), lens - 1] = 3
mask = (torch.arange(256)[None, :] < lens[:, None]).long()
text = {"input_ids": ids.to(dev), "token_type_ids": torch.zeros_like(ids).to(dev), "attention_mask": mask.to(dev)}
padded = dict(text)
padded["attention_mask"] = text["attention_mask"].clone()
padded["attention_mask"][0, 255] = 1  # one attended key at the last position forces the full S=256 computation


def timed(inputs, chunk):
    out = None
    for _ in range(2):
        out = model.encode_text({k: v[:chunk] for k, v in inputs.items()})
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = model.encode_text(inputs)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out


def latency_b1(n_calls=200):
    """The reference's own call pattern (WSI_evaluation/utils.py:67-74): one encode_text per class prompt, batch 1."""
    one = {k: v[:1] for k, v in text.items()}
    for _ in range(10):
        model.encode_text(one)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n_calls):
        model.encode_text(one)
    e1.record()
    torch.cuda.synchronize()
    return {"tokens": int(lens[0]), "calls": n_calls, "device_us_per_call": e0.elapsed_time(e1) * 1e3 / n_calls,
            "wall_us_per_call": (time.perf_counter() - t0) * 1e6 / n_calls}


b1 = latency_b1()
t_trim, o_trim = timed(text, 4096)
n_pad = min(P, 16384)
t_pad, o_pad = timed({k: v[:n_pad] for k, v in padded.items()}, 1024)
diff = (o_trim[1:n_pad] - o_pad[1:]).abs().max().item()
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 1400.0
line = {
    "workload": f"prompt bank: {P} prompts, S=256 padded, lengths U{{4..32}}",
    "trimmed": {"prompts_per_s": P / t_trim, "seconds": t_trim, "s_eff": int(lens.max())},
    "padded_S256": {"prompts": n_pad, "prompts_per_s": n_pad / t_pad, "seconds": t_pad,
                    "frac_of_tensor_roofline": (n_pad / t_pad) * 45.904e9 / (peak * 1e12)},
    "max_abs_diff_trimmed_vs_padded": diff,
    "batch_1_latency": b1,
}
print(json.dumps(line))
