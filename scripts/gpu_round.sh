#!/bin/bash
# One gpurun call: parity tests, smoke, bench lines (+ optional ncu captures with NCU=1).
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python bench.py --steps 20 --warmup 3 --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python bench.py --steps 20 --warmup 3 --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
if [ "$NCU" = "1" ]; then
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_c4.json 2> gpurun_out/bench_ref_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_track -s 4 -c 2 -f -o gpurun_out/prof_k_track \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
fi
tail -8 gpurun_out/pytest_gpu.txt; tail -2 gpurun_out/smoke.txt
for w in c4 c2 c3; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$w.json'))
    print('$w', 'value %.4g it/s'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'e2e %.4g'%d.get('e2e',{}).get('value',0), 'cpu %.4g (%s thr) 1thr %.4g'%(d.get('cpu_baseline',{}).get('value',0), d.get('cpu_baseline',{}).get('cores'), d.get('cpu_baseline',{}).get('single_thread_value',0)), 'frac %.4f'%d['roofline']['frac'], d['clocks'])
except Exception as e:
    print('$w failed', e); print(open('gpurun_out/bench_$w.err').read()[-1500:])
PY
done
