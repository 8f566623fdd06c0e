#!/bin/bash
# round 2: k_track2 correctness + timing vs k_track, TMA on / off
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
( time timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.txt 2>&1
python scripts/phase_timing2.py c4 > gpurun_out/phase2_c4.txt 2>&1
M3TB_TMA=0 python scripts/phase_timing2.py c4 > gpurun_out/phase2_c4_legacy.txt 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_k0.json 2> gpurun_out/bench_c4_k0.err
M3TB_TMA=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c4_k0_legacy.json 2> gpurun_out/bench_c4_k0_legacy.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --workload c2 > gpurun_out/bench_c2_k0.json 2> gpurun_out/bench_c2_k0.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --workload c3 > gpurun_out/bench_c3_k0.json 2> gpurun_out/bench_c3_k0.err
tail -42 gpurun_out/pytest_gpu.txt | cut -c1-250; cat gpurun_out/parity_metrics.jsonl | cut -c1-330; head -12 gpurun_out/phase2_c4.txt; sed -n 12,24p gpurun_out/phase2_c4.txt;  head -3 gpurun_out/phase2_c4_legacy.txt
python - <<PY
import json
for w in ('c4_k0','c4_k0_legacy','c2_k0','c3_k0'):
    try:
        d=json.load(open(f'gpurun_out/bench_{w}.json'))
        pc=d.get('parity_check') or {}
        print(w, 'ms/step %.4f'%d['ms_per_step'], 'frac %.4f'%d['roofline']['frac'], 'e2e', d.get('e2e',{}).get('ms_per_step'), 'parity', pc.get('ok'), pc.get('per_iteration_max_rad'))
    except Exception as e:
        print(w,'failed',e); print(open(f'gpurun_out/bench_{w}.err').read()[-800:])
PY
