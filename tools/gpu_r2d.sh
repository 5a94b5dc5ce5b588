#!/bin/bash
# round-2 call D: full GPU test suite + smoke + bench (both arms)
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ours.json 2> gpurun_out/${TAG}_bench_ours.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench_ours.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_ours.json"))
print({k:d.get(k) for k in ("value","ms_per_step","kernels","roofline","e2e","gpu_launches","numa","cpu_baseline")})
print(json.dumps(d.get("flownet2"), indent=1)); print(json.dumps(d.get("ops"), indent=0))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "benchref rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_ref.json"))
print({k:d.get(k) for k in ("value","ms_per_step","e2e","steps_cap","numa")}); print(d.get("flownet2"))
PY
