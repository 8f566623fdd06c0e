#!/bin/bash
# end-to-end step against the grid of the prefetch ingest (M3TB_INGEST_CTAS; 0 = the default, a quarter of the SMs)
mkdir -p gpurun_out
for ctas in 0 24 32 40 64 128 0; do
M3TB_INGEST_CTAS=$ctas BENCH_DEBUG_E2E=1 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-secondary --no-parity-check 2>gpurun_out/bench_e2e_$ctas.err | python -c "import json,sys; d=json.load(sys.stdin); print('ingest CTAs $ctas: value ms/step %.4f'%d['ms_per_step'], 'e2e ms/step %.4f'%d['e2e']['ms_per_step'])"
done
