#!/bin/bash
mkdir -p gpurun_out
for v in A B A B; do
  if [ $v = A ]; then unset M3TB_LIB; else export M3TB_LIB=$PWD/3dobjecttracking_b200/csrc/libm3t_b200_B.so; fi
  python scripts/phase_timing2.py c4 2>&1 | head -1 | sed "s/^/$v: /"
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary --no-parity-check 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$v bench ms/step %.4f'%d['ms_per_step'])"
done
