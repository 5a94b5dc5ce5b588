#!/bin/bash
# round-2 call C: Resample2d kernel families (tests, sweep, ncu of the tile kernels) + fused model forwards
TAG=r2c
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "resample or warp or fused or cat or full_size_output" > gpurun_out/${TAG}_pytest_rs.log 2>&1; echo "pytest rs rc=$?"; tail -12 gpurun_out/${TAG}_pytest_rs.log | cut -c1-400
timeout 900 python tools/rs_sweep.py > gpurun_out/${TAG}_rs_sweep.txt 2>&1; echo "sweep rc=$?"; cat gpurun_out/${TAG}_rs_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"resample2d_" -c 6 -o gpurun_out/${TAG}_prof_rs python tools/prof_ops.py small 1 > gpurun_out/${TAG}_prof.log 2>&1; echo "ncu rc=$?"
