#!/bin/bash
# round-2 call G: full GPU test suite + smoke + bench (both arms) + ncu launch list of the bench command + ncu full capture
TAG=${1:-r2g}
mkdir -p gpurun_out
python -c "
import importlib.util,os
s=importlib.util.spec_from_file_location('b','flownet2-pytorch_b200/build.py'); m=importlib.util.module_from_spec(s); s.loader.exec_module(m); print(m.product_digest())" > gpurun_out/${TAG}_digest.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ours.json 2> gpurun_out/${TAG}_bench_ours.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_ours.json"))
print({k:d.get(k) for k in ("value","ms_per_step","kernels","roofline","gpu_launches")}); print(d["e2e"]["ms_per_step"], {k:(v.get("pairs_per_sec_per_gpu"), v.get("fused",{}).get("pairs_per_sec_per_gpu"), v.get("agreement")) for k,v in d["flownet2"].items()})
print({k:(v["ms"],v["frac_hbm"]) for k,v in d["ops"].items()})
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "benchref rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/${TAG}_bench_under_ncu.log 2>&1; echo "ncu-list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"corr_|resample2d_|channelnorm_|warp_concat" -c 14 -o gpurun_out/${TAG}_prof python tools/prof_ops.py all 1 > gpurun_out/${TAG}_prof.log 2>&1; echo "ncu-full rc=$?"
