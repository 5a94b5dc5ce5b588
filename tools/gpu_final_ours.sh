#!/bin/bash
# ours-only variant of gpu_final.sh (the reference arm does not change between kernel revisions)
TAG=${1:-r1n}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ours.json 2> gpurun_out/${TAG}_bench_ours.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_ours.json"))
print({k:d[k] for k in ("value","ms_per_step","kernels","e2e","gpu_launches","clocks")}); print(d["roofline"]["frac"], d.get("flownet2"))
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/${TAG}_bench_under_ncu.log 2>&1; echo "ncu-list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"corr_|resample2d_|channelnorm_" -c 12 -o gpurun_out/${TAG}_prof python tools/prof_ops.py all 1 > gpurun_out/${TAG}_prof.log 2>&1; echo "ncu-full rc=$?"
