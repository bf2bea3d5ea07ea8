mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "attention" 2>&1 | tail -3
python tools/trace_attention.py > gpurun_out/attention_trace_new.txt 2>&1; tail -1 gpurun_out/attention_trace_new.txt
python tools/probe_kernels.py attperf 2>&1 | grep -i "perf" | head -3
KEEPB200_LIB=keep_b200/libkeep_b200_old.so python tools/probe_kernels.py attperf 2>&1 | grep -i "perf" | head -3
for v in old new old new; do
if [ $v = old ]; then export KEEPB200_LIB=keep_b200/libkeep_b200_old.so; else unset KEEPB200_LIB; fi
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$v.json 2> gpurun_out/bench_tmp.err; python -c "
import json;d=json.load(open('gpurun_out/bench_$v.json'));print('$v', d['value'], d['roofline']['achieved'], d['clocks']['sm_mhz'], d['roofline']['gemm_share_of_step'])"
done
