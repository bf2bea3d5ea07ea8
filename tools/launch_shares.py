"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of one encode_image
chunk (from the last im2col launch to the end)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [x["Kernel Name"] for x in rows]
starts = [i for i, n in enumerate(names) if "im2col" in n]
start = starts[-2] if len(starts) > 1 else starts[-1]
end = starts[-1] if len(starts) > 1 else len(rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for x in rows[start:end]:
    v = float(x["Metric Value"])
    u = x["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    n = x["Kernel Name"]
    m = re.search(r"(\w+)<([^>]*)>\(", n) or re.search(r"(\w+)\(", n)
    key = m.group(1) + ("<" + m.group(2) + ">" if m.lastindex and m.lastindex > 1 else "")
    agg[key][0] += 1
    agg[key][1] += v
tot = sum(v[1] for v in agg.values())
print(f"one chunk: launches {end - start}, total {tot / 1e3:.2f} ms (serialised, cold-cache ncu times: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / 1e3:9.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:4d} avg={v[1] / v[0]:9.1f} us  {k}")
