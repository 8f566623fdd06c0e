#!/bin/bash
# e2e with the ingest-plan switches (M3TB_INGEST_PLAN bits: 1 early pose snapshot, 2 third stream, 4 pinned staging) and grid sizes
for cfg in "7 32" "7 24" "7 40" "7 48" "7 64" "7 37" "0 128" "7 32" "0 32" "7 80"; do
set -- $cfg
M3TB_INGEST_PLAN=$1 M3TB_INGEST_CTAS=$2 BENCH_DEBUG_E2E=1 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-secondary --no-parity-check 2>gpurun_out/bench_e2e_$1_$2.err | python -c "import json,sys; d=json.load(sys.stdin); print('plan $1 ctas $2 value ms/step %.4f'%d['ms_per_step'], 'e2e ms/step %.4f'%d['e2e']['ms_per_step'])"
done
