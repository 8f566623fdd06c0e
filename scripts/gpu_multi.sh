#!/bin/bash
# multi-GPU: NCCL pose equality test (2 ranks) and the bench at N GPUs (weak scaling, e2e)
mkdir -p gpurun_out
N=${1:-2}
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -5
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c4_n$N.json').read().strip().split('\n')[-1])
    print('N=$N value %.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'e2e %.4g it/s  %.4f ms'%(d['e2e']['value'], d['e2e']['ms_per_step']), d['e2e'].get('host_binding'))
except Exception as e:
    print('failed', e); print(open('gpurun_out/bench_c4_n$N.err').read()[-1500:])
PY
