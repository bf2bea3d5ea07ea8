mkdir -p gpurun_out
timeout 900 python tools/bench_wsi.py > gpurun_out/bench_wsi_rows.json 2> gpurun_out/bench_wsi_rows.err; tail -3 gpurun_out/bench_wsi_rows.err; python -c "
import json;d=json.load(open('gpurun_out/bench_wsi_rows.json'))
for r in d['rows']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k not in ('kernel','cpu_note')})"
