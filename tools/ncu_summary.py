"""Compact per-launch table from an `ncu --page raw --csv` export (tools/gpu_round.sh): duration, DRAM bytes and
throughput, tensor-pipe activity, SM clock, registers - the numbers DESIGN.md and bench.py's roofline.traffic quote."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
units = rows[1]


def num(r, name):
    i = col.get(name)
    if i is None:
        return float("nan")
    try:
        return float(r[i].replace(",", ""))
    except ValueError:  # "n/a", "no data"
        return float("nan")


def scaled(r, name, to):
    """value converted to the unit `to` using the unit row of the CSV"""
    v, u = num(r, name), units[col[name]] if name in col else ""
    f = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    base = v * f.get(u, 1.0)
    return base / f[to]


print("kernel,grid,block,regs,duration_us,dram_read_MB,dram_write_MB,dram_GBps,dram_pct,tensor_pipe_active_pct,sm_GHz")
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[col["Kernel Name"]]
    m = re.search(r"(\w+)<([^>]*)>", name) or re.search(r"(\w+)\(", name) or re.search(r"(\w+)", name)
    short = m.group(0).rstrip("(").replace(", ", " ")  # template arguments: "gemm2_kernel<1 0 1>" = <epilogue, bf16, full row tiles>
    dur = scaled(r, "gpu__time_duration.sum", "us")
    rd, wr = scaled(r, "dram__bytes_read.sum", "Mbyte"), scaled(r, "dram__bytes_write.sum", "Mbyte")
    print(f"{short},{r[col['Grid Size']].replace(',', ' ')},{r[col['Block Size']].replace(',', ' ')},{num(r, 'launch__registers_per_thread'):.0f},"
          f"{dur:.1f},{rd:.1f},{wr:.1f},{(rd + wr) / dur * 1e3:.0f},{num(r, 'FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed'):.1f},"
          f"{num(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f},{num(r, 'sm__cycles_elapsed.avg.per_second') / 1e9 if units[col['sm__cycles_elapsed.avg.per_second']] == 'hz' else num(r, 'sm__cycles_elapsed.avg.per_second'):.2f}")
