"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit, HMMA = legacy mma.sync, FFMA2 = packed fp32.
    python tools/sass_evidence.py > profiles/<round>_sass_evidence.txt        (needs cuobjdump; no GPU)
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "keep_b200", "libkeep_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
keys = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "FFMA2", "MUFU.EX2", "SYNCS", "total"]
funcs, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line) if cur else None
    if m:
        for k in keys[:-1]:
            if m.group(1).startswith(k):
                funcs[cur][k] += 1
        funcs[cur]["total"] += 1
names = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
print("SASS mnemonic counts per kernel of keep_b200/libkeep_b200.so (cuobjdump -sass, sm_100a), built from the committed sources.")
print("UTC*MMA = tcgen05.mma (UTCQMMA: kind::tf32), LDTM/STTM = tcgen05.ld/st, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit, HMMA = legacy")
print("mma.sync, FFMA2 = packed fp32 FMA, SYNCS = mbarrier operations.\n")
print(f"{'kernel':58s} " + " ".join(f"{k:>8s}" for k in keys))
for name, d in zip(names, funcs.values()):
    short = re.sub(r"\(CUtensorMap_st.*|\(.*", "", name.replace("kb::(anonymous namespace)::", ""))
    if d["total"] >= 20:
        print(f"{short[:58]:58s} " + " ".join(f"{d[k]:8d}" for k in keys))
