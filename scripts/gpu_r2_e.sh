#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
( time timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 ) > gpurun_out/pytest_gpu.txt 2>&1
python scripts/phase_timing2.py c4 > gpurun_out/phase2_c4.txt 2>&1
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/bench_c4_quick.json 2> gpurun_out/bench_c4_quick.err
tail -12 gpurun_out/pytest_gpu.txt | cut -c1-220; head -14 gpurun_out/phase2_c4.txt
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c4_quick.json'))
print('c4 ms/step %.4f'%d['ms_per_step'], 'frac %.4f'%d['roofline']['frac'], 'parity', d['parity_check']['ok'], d['parity_check']['per_iteration_max_rad'])
PY
