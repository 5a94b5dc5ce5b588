#!/bin/bash
# compute-sanitizer passes over every product kernel (SURVEY section 5 "race detection").
# Logs land in gpurun_out/<tag>_sanitize_<tool>.log; copy the summaries to profiles/.
TAG=${1:-r2a}
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool group timeout extra-args...
    local tool=$1 group=$2 to=$3; shift 3
    local log=gpurun_out/${TAG}_sanitize_${tool}_${group}.log
    timeout $to $S --tool $tool --error-exitcode 9 "$@" --log-file $log python tools/sanitize_ops.py $group > gpurun_out/${TAG}_sanitize_${tool}_${group}.out 2>&1
    echo "sanitize $tool $group rc=$? : $(tail -1 gpurun_out/${TAG}_sanitize_${tool}_${group}.out) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
}
run memcheck all 420
run synccheck tc 300
run racecheck tc 480
run racecheck rs 200
run initcheck tc 300
