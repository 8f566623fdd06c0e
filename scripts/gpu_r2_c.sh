#!/bin/bash
# round 2: tests + timing + ncu evidence of k_track2
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
( time timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 ) > gpurun_out/pytest_gpu.txt 2>&1
python scripts/phase_timing2.py c4 > gpurun_out/phase2_c4.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check --no-secondary > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_track2 -s 4 -c 1 -f -o gpurun_out/prof_k_track2 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity-check --no-secondary > gpurun_out/ncu_full.log 2>&1
# the capture's DRAM traffic goes into profiles/traffic.json (tied to the kernel sources), THEN the bench line is taken
python scripts/ncu_summarise.py gpurun_out/prof_k_track2.ncu-rep r02_k_track2 > /dev/null 2>&1; cp profiles/traffic.json gpurun_out/traffic.json
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
tail -12 gpurun_out/pytest_gpu.txt | cut -c1-200; head -16 gpurun_out/phase2_c4.txt
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c4.json'))
print('c4 ms/step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], 'parity', d['parity_check'], 'cpu', d.get('cpu_baseline',{}).get('value'))
PY
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/prof_k_track2.ncu-rep
