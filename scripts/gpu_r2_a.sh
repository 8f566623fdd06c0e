#!/bin/bash
# round 2, call A: baseline of the round-1 kernel + the new parity tests + bench with parity_check
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
nvidia-smi -L | head -2
( time timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 ) > gpurun_out/pytest_gpu.txt 2>&1
python scripts/phase_timing.py c4 > gpurun_out/phase_c4.txt 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
tail -32 gpurun_out/pytest_gpu.txt; cat gpurun_out/parity_metrics.jsonl; head -12 gpurun_out/phase_c4.txt
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c4.json'))
print({k:d[k] for k in ('value','ms_per_step','parity_check','clocks')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic_note'])
PY
tail -3 gpurun_out/bench_c4.err
