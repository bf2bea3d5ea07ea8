#!/bin/bash
# One GPU-box pass: parity tests, the bench line, the ncu launch list and full captures of the hot kernels.
#   gpurun --timeout 1800 -- 'bash tools/gpu_round.sh <tag> [tests|notests]'
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/.
set -u
TAG=${1:-r02}
TESTS=${2:-tests}
OUT=gpurun_out
mkdir -p $OUT
python -m keep_b200.build > $OUT/${TAG}_build.log 2>&1 || { cat $OUT/${TAG}_build.log; exit 1; }
if [ "$TESTS" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
  grep -E "passed|failed|FAILED|pytest exit" $OUT/${TAG}_pytest_gpu.log | tail -5
fi
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench exit $?"; cut -c1-1500 $OUT/${TAG}_bench_n1.json
# every launch of one 512-tile chunk with its device time (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k 'regex:^(?!.*(elementwise|distribution|at::|cub|Fill)).*$' -c 1200 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --tiles 512 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $OUT/${TAG}_ncu_launch_bench.log 2>&1
python tools/launch_shares.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launch_shares.txt 2>&1; cat $OUT/${TAG}_launch_shares.txt
# --set full of one ViT block's kernels (+ the similarity kernel and the fused head at the end of the chunk)
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gemm2|attention_tc1|layernorm|sim_tc|head_kernel' -s 60 -c 9 -f -o $OUT/${TAG}_full \
  python bench.py --tiles 512 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $OUT/${TAG}_ncu_full_bench.log 2>&1
echo "ncu full exit $?"
# the text tower (packed tcgen05 attention, split-operand GEMMs, embedding gather, fused pooler) and the WSI kernels
timeout 600 ncu --set full --clock-control none -k 'regex:attention_tc|bert_embed|gemm_kernel|head_kernel|refine|prompt_score|score_reduce|sim_tc|resample|crop_kernel|table_' -c 44 -f -o $OUT/${TAG}_full_text_wsi \
  python tools/profile_text_wsi.py > $OUT/${TAG}_ncu_text_wsi.log 2>&1
echo "ncu text/wsi exit $?"
# gpurun brings back at most 64 MiB: keep the raw pages (every metric of every captured launch) as CSV, drop the reports
for rep in ${TAG}_full ${TAG}_full_text_wsi; do
  ncu -i $OUT/$rep.ncu-rep --page raw --csv > $OUT/$rep.raw.csv 2>/dev/null
  rm -f $OUT/$rep.ncu-rep
done
ls -la $OUT | tail -20
