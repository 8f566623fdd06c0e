#!/bin/bash
# One gpurun call: parity tests, smoke, bench lines, ncu launch list + full capture of k_track.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python bench.py --steps 20 --warmup 3 --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python bench.py --steps 20 --warmup 3 --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_c4.json 2> gpurun_out/bench_ref_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_track -s 4 -c 2 -f -o gpurun_out/prof_k_track \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu.txt; cat gpurun_out/smoke.txt | tail -3; cat gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
