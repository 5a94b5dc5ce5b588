#!/bin/bash
# round-2 final validation: sanitizer over the tensor-core kernels (TS-mode forward), full GPU tests, smoke, bench both arms,
# ncu launch list + full capture
TAG=${1:-r2m}
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 400 $S --tool $tool --error-exitcode 9 --log-file gpurun_out/${TAG}_sanitize_${tool}_tc.log python tools/sanitize_ops.py tc > gpurun_out/${TAG}_sanitize_${tool}_tc.out 2>&1
  echo "sanitize $tool tc rc=$? : $(tail -1 gpurun_out/${TAG}_sanitize_${tool}_tc.out) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_sanitize_${tool}_tc.log | tail -1)"
done
timeout 400 $S --tool memcheck --error-exitcode 9 --log-file gpurun_out/${TAG}_sanitize_memcheck_fused.log python tools/sanitize_ops.py fused > gpurun_out/${TAG}_sanitize_memcheck_fused.out 2>&1; echo "sanitize memcheck fused rc=$? $(tail -1 gpurun_out/${TAG}_sanitize_memcheck_fused.out)"
bash tools/gpu_r2g.sh ${TAG}
