#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) 2>&1 | cut -c1-200
for i in 1 2; do
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-secondary --no-parity-check 2>gpurun_out/bench_e2e.err | python -c "import json,sys; d=json.load(sys.stdin); print('value ms/step %.4f'%d['ms_per_step'], 'e2e ms/step %.4f'%d['e2e']['ms_per_step'], 'bytes', d['e2e']['h2d_bytes_per_step'])"
done
