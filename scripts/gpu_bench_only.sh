#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c4_reference.json 2> gpurun_out/bench_c4_reference.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c4.json'))
print('c4 ms/step %.4f'%d['ms_per_step'], 'frac %.4f'%d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity_check']['ok'], 'cpu', d['cpu_baseline']['value'])
r=json.load(open('gpurun_out/bench_c4_reference.json'))
print('reference arm', r['value'], r['unit'], r['cpu_baseline'])
PY
