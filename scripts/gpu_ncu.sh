#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_track -s 4 -c 1 -f -o gpurun_out/prof_c4 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_track -s 4 -c 1 -f -o gpurun_out/prof_c2 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workload c2 > gpurun_out/ncu_c2.log 2>&1
M3TB_NO_TILES=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c4_notiles.json 2>&1
tail -2 gpurun_out/ncu_c4.log; tail -2 gpurun_out/ncu_c2.log; cat gpurun_out/bench_c4_notiles.json | cut -c1-200
