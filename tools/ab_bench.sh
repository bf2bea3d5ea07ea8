#!/bin/bash
# A/B of two builds of the same ABI inside ONE gpurun call (box-to-box spread on the pool is +-3%):
#   bash tools/ab_bench.sh <tag> <other libkeep_b200.so> [rounds]
# alternates <other> and the in-tree library, prints value / GEMM TFLOP/s per shape for each run.
TAG=${1:-ab}; OTHER=${2:-_ab/libkeep_b200_base.so}; ROUNDS=${3:-2}
OUT=gpurun_out; mkdir -p $OUT
for r in $(seq 1 $ROUNDS); do
  for which in other tree; do
    if [ $which = other ]; then export KEEPB200_LIB=$OTHER; else unset KEEPB200_LIB; fi
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $OUT/${TAG}_${which}_$r.json 2> $OUT/${TAG}_${which}_$r.err
    python - "$OUT/${TAG}_${which}_$r.json" $which <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print(sys.argv[2].ljust(6), "tiles/s %.0f" % d["value"], "gemm %.0f TF/s" % r["achieved"],
      " ".join("epi%d:%.0f" % (s["epi"], s["tflops"]) for s in r["per_shape"]), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
  done
done
