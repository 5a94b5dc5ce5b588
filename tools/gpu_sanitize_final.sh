#!/bin/bash
# compute-sanitizer over the final round-2 sources: every kernel family (memcheck), the non-tensor-core kernels (racecheck)
TAG=${1:-r2q}
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() { local tool=$1 group=$2 to=$3
  timeout $to $S --tool $tool --error-exitcode 9 --log-file gpurun_out/${TAG}_sanitize_${tool}_${group}.log python tools/sanitize_ops.py $group > gpurun_out/${TAG}_sanitize_${tool}_${group}.out 2>&1
  echo "sanitize $tool $group rc=$? : $(tail -1 gpurun_out/${TAG}_sanitize_${tool}_${group}.out) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_sanitize_${tool}_${group}.log | tail -1)"; }
run memcheck all 500
run racecheck rs 200
run racecheck cn 200
run racecheck fused 200
run initcheck tc 300
