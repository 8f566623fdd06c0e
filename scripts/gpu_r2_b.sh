#!/bin/bash
# round 2, call B: k_track2 correctness + timing vs k_track
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
( time timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.txt 2>&1
cp gpurun_out/parity_metrics.jsonl gpurun_out/parity_metrics_k2.jsonl 2>/dev/null
python scripts/phase_timing2.py c4 > gpurun_out/phase2_c4.txt 2>&1
for k in 0 1; do
M3TB_KERNEL=$k python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_k$k.json 2> gpurun_out/bench_c4_k$k.err
M3TB_KERNEL=$k python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --workload c2 > gpurun_out/bench_c2_k$k.json 2> gpurun_out/bench_c2_k$k.err
M3TB_KERNEL=$k python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --workload c3 > gpurun_out/bench_c3_k$k.json 2> gpurun_out/bench_c3_k$k.err
done
tail -42 gpurun_out/pytest_gpu.txt | cut -c1-250; cat gpurun_out/parity_metrics.jsonl; cat gpurun_out/phase2_c4.txt | head -40
python - <<PY
import json
for w in ('c4','c2','c3'):
  for k in (0,1):
    try:
        d=json.load(open(f'gpurun_out/bench_{w}_k{k}.json'))
        print(w, 'kernel', 'k_track2' if k==0 else 'k_track', 'ms/step %.4f'%d['ms_per_step'], 'frac %.4f'%d['roofline']['frac'], 'e2e', d.get('e2e',{}).get('ms_per_step'), d.get('parity_check'))
    except Exception as e:
        print(w,k,'failed',e); print(open(f'gpurun_out/bench_{w}_k{k}.err').read()[-800:])
PY
