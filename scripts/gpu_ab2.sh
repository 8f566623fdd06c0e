#!/bin/bash
# A/B of library variants: usage gpu_ab2.sh <tag> <tag> ... ; tag "cur" = the in-tree build, else csrc/libm3t_b200_<tag>.so
mkdir -p gpurun_out/ab
for round in 1 2; do
for v in "$@"; do
  if [ $v = cur ]; then unset M3TB_LIB; else export M3TB_LIB=$PWD/3dobjecttracking_b200/csrc/libm3t_b200_$v.so; fi
  if [ $round = 1 ]; then
    python scripts/ab_pose_dump.py gpurun_out/ab/state_$v.npz
    python scripts/phase_timing2.py c4 > gpurun_out/ab/phase_$v.txt 2>&1
    head -1 gpurun_out/ab/phase_$v.txt | sed "s/^/$v: /"
  fi
  python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary --no-parity-check 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$v bench ms/step %.4f'%d['ms_per_step'])"
done
done
python - "$@" <<PY
import sys, numpy as np
tags = sys.argv[1:]
ref = np.load(f"gpurun_out/ab/state_{tags[0]}.npz")
for t in tags[1:]:
    z = np.load(f"gpurun_out/ab/state_{t}.npz")
    print(t, "vs", tags[0], "poses identical:", bool((ref["poses"].view(np.uint32) == z["poses"].view(np.uint32)).all()),
          "line state identical:", bool(np.array_equal(ref["lines"], z["lines"])))
PY
