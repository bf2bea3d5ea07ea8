mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -12
for v in 0 1 2 0 1 2; do
KEEPB200_LN_FUSE=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_fuse$v.json 2> gpurun_out/bench_tmp.err; python -c "
import json;d=json.load(open('gpurun_out/bench_fuse$v.json'));print('fuse=$v', d['value'], d['roofline']['achieved'], d['clocks']['sm_mhz'], [(r['N'],r['K'],r['epi'],r['tflops']) for r in d['roofline']['per_shape'][:5]])"
done
