#!/bin/bash
# bench both arms + ncu launch list (per-launch durations of the same bench command)
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ours.json 2> gpurun_out/${TAG}_bench_ours.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench_ours.json; tail -3 gpurun_out/${TAG}_bench_ours.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "benchref rc=$?"
cat gpurun_out/${TAG}_bench_ref.json; tail -3 gpurun_out/${TAG}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/${TAG}_bench_under_ncu.log 2>&1; echo "ncu rc=$?"
