#!/bin/bash
# One gpurun call worth of work: smoke -> tests -> golden -> bench (both arms) -> kernel comparison -> ncu.
# Every stage is wrapped in its own timeout so a hung kernel cannot eat the box.
mkdir -p gpurun_out
TAG=${1:-r1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${TAG}_pytest.log
timeout 200 python tests/golden/make_golden.py gpurun_out/golden > gpurun_out/${TAG}_golden.log 2>&1; echo "golden rc=$?"
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_ours.json 2> gpurun_out/${TAG}_bench_ours.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench_ours.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "benchref rc=$?"
cat gpurun_out/${TAG}_bench_ref.json
timeout 600 python tools/compare_ref.py > gpurun_out/${TAG}_compare.json 2> gpurun_out/${TAG}_compare.err; echo "compare rc=$?"
