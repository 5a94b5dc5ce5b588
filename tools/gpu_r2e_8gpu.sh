#!/bin/bash
# round-2 call E (8 GPUs): host-copy bandwidth alone vs 8 ranks at once, with / without NUMA binding; bench e2e at N = 8
TAG=r2e
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo8.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/hostcopy_probe.py > gpurun_out/${TAG}_hostcopy_8gpu.json 2> gpurun_out/${TAG}_hostcopy_8gpu.err; echo "hostcopy8 rc=$?"; cat gpurun_out/${TAG}_hostcopy_8gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > gpurun_out/${TAG}_bench8_numa.json 2> gpurun_out/${TAG}_bench8_numa.err; echo "bench8 rc=$?"
FN2B200_NUMA=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > gpurun_out/${TAG}_bench8_nonuma.json 2> gpurun_out/${TAG}_bench8_nonuma.err; echo "bench8 (no numa) rc=$?"
python - <<PY
import json
for f in ("bench8_numa","bench8_nonuma"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s.json"%f)); print(f, {k:d.get(k) for k in ("value","ms_per_step","e2e","numa")})
    except Exception as e: print(f, "ERR", e)
PY
