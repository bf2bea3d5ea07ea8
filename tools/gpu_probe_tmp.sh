mkdir -p gpurun_out
echo "== new, prefetch auto"; python tools/probe_kernels.py perf 2>&1 | grep perf
echo "== new, prefetch on"; KEEPB200_RESID_PREFETCH=1 python tools/probe_kernels.py perf 2>&1 | grep perf
echo "== new, prefetch off"; KEEPB200_RESID_PREFETCH=0 python tools/probe_kernels.py perf 2>&1 | grep perf
echo "== old"; KEEPB200_LIB=keep_b200/libkeep_b200_old.so python tools/probe_kernels.py perf 2>&1 | grep perf
echo "== new again"; python tools/probe_kernels.py perf 2>&1 | grep perf
for v in old new old new; do
if [ $v = old ]; then export KEEPB200_LIB=keep_b200/libkeep_b200_old.so; else unset KEEPB200_LIB; fi
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$v.json 2> gpurun_out/bench_tmp.err; python -c "
import json;d=json.load(open('gpurun_out/bench_$v.json'));print('$v', d['value'], d['roofline']['achieved'], d['clocks']['sm_mhz'], [(r['N'],r['K'],r['tflops']) for r in d['roofline']['per_shape'][:4]])"
done
