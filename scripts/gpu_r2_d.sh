#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
( time timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.txt 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
tail -30 gpurun_out/pytest_gpu.txt | cut -c1-220
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c4.json'))
print('c4 ms/step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity_check']['ok'], 'cpu', d.get('cpu_baseline',{}).get('value'))
for k,v in d.get('secondary',{}).items(): print(k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk!='workload'})
PY
tail -3 gpurun_out/bench_c4.err
