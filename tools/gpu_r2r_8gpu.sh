#!/bin/bash
# final sources, 8 GPUs: the driver's launch of the ours arm, extras included
TAG=r2r
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench8_ours.json 2> gpurun_out/${TAG}_bench8_ours.err; echo "bench8 ours rc=$?"
python - <<PY
import json
for line in open("gpurun_out/${TAG}_bench8_ours.json").read().splitlines():
    if line.startswith("{"):
        d=json.loads(line); print({k:d.get(k) for k in ("value","ms_per_step","ms_per_step_fastest_rank","n_gpus")}, d["e2e"]["ms_per_step"], d.get("flownet2"))
PY
tail -2 gpurun_out/${TAG}_bench8_ours.err
