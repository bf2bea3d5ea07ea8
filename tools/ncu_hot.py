"""Top stall locations from `ncu -i rep --page source --csv --kernel-id :::N > file.csv`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
print("kernel:", rows[0][1][:100])
print("total samples", tot, "instructions", len(data))
agg = {hdr[i]: sum(int(r[i]) for r in data) for i in stalls}
print("stall mix:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for r in sorted(data, key=lambda r: -int(r[isamp]))[:n]:
    why = sorted(((int(r[i]), hdr[i][6:]) for i in stalls), reverse=True)[:2]
    print(r[isamp].rjust(7), r[iex].rjust(9), r[ia].strip()[:70].ljust(70), why)
