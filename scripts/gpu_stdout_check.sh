#!/bin/bash
# the bench's stdout must be exactly one JSON line, also under torchrun (NCCL banner and friends go to stderr)
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-secondary > gpurun_out/stdout_n$N.txt 2> gpurun_out/stderr_n$N.txt
echo "rc $? stdout lines: $(wc -l < gpurun_out/stdout_n$N.txt)"; head -c 200 gpurun_out/stdout_n$N.txt; echo
grep -c "NCCL version" gpurun_out/stderr_n$N.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/stdout_n1.txt 2> gpurun_out/stderr_n1.txt
echo "rc $? stdout lines: $(wc -l < gpurun_out/stdout_n1.txt)"; python -c "import json; d=json.load(open('gpurun_out/stdout_n1.txt')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['parity_check']['ok'])"
