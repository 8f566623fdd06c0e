set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/bench_c5_projected.json 2> gpurun_out/bench_c5_projected.err; tail -c 3000 gpurun_out/bench_c5_projected.json; tail -5 gpurun_out/bench_c5_projected.err
python bench.py --workload c5 --variant constrained --steps 10 --warmup 3 > gpurun_out/bench_c5_constrained.json 2> gpurun_out/bench_c5_constrained.err; tail -c 1500 gpurun_out/bench_c5_constrained.json; tail -5 gpurun_out/bench_c5_constrained.err
