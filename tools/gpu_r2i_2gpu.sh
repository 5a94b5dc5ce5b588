#!/bin/bash
# round-2 call I (2 GPUs): the driver's multi-GPU launch of bench.py, both arms, extras included
TAG=r2i
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench2_ours.json 2> gpurun_out/${TAG}_bench2_ours.err; echo "bench2 ours rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench2_ref.json 2> gpurun_out/${TAG}_bench2_ref.err; echo "bench2 ref rc=$?"
python - <<PY
import json
for f in ("bench2_ours","bench2_ref"):
    try:
        for line in open("gpurun_out/${TAG}_%s.json"%f).read().splitlines():
            if line.startswith("{"):
                d=json.loads(line); print(f, {k:d.get(k) for k in ("value","ms_per_step","n_gpus","numa")}, d["e2e"]["ms_per_step"], d.get("flownet2"))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/${TAG}_bench2_ours.err
