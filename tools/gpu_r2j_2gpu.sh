#!/bin/bash
TAG=r2j
mkdir -p gpurun_out
for path in module functional module functional; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --step-path $path 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$path', d['ms_per_step'], d['ms_per_step_fastest_rank'], d['kernels']['forward_ms'], d['kernels']['backward_ms'])"
done
