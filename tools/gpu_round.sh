#!/bin/bash
# One GPU-box pass: parity tests, the bench line, the ncu launch list and one full capture of the hot kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag> [tests|notests]'
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/.
set -u
TAG=${1:-r01}
TESTS=${2:-tests}
OUT=gpurun_out
mkdir -p $OUT
python -m keep_b200.build > $OUT/${TAG}_build.log 2>&1 || { cat $OUT/${TAG}_build.log; exit 1; }
if [ "$TESTS" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
  tail -5 $OUT/${TAG}_pytest_gpu.log
fi
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench exit $?"; cat $OUT/${TAG}_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k 'regex:^(?!.*(elementwise|distribution|at::|cub|Fill)).*$' -c 1200 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --tiles 512 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_launch_bench.log 2>&1
python tools/launch_shares.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launch_shares.txt 2>&1; cat $OUT/${TAG}_launch_shares.txt
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gemm2|attention_tc|layernorm' -s 60 -c 7 -f -o $OUT/${TAG}_full \
  python bench.py --tiles 512 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_full_bench.log 2>&1
echo "ncu full exit $?"
ls -la $OUT
