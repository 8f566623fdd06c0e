#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
( time timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 ) > gpurun_out/pytest_gpu.txt 2>&1
tail -14 gpurun_out/pytest_gpu.txt | cut -c1-220
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | cut -c1-200
echo "== racecheck"; timeout 1200 compute-sanitizer --tool racecheck --print-limit 8 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 | cut -c1-220
echo "== synccheck"; timeout 900 compute-sanitizer --tool synccheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | cut -c1-200
