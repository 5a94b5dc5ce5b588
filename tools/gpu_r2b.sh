#!/bin/bash
# round-2 call B: new Resample2d kernels (tests, sweep, ncu) + reduction-flavour micro-benchmark
TAG=r2b
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "resample or warp or channelnorm or umma" > gpurun_out/${TAG}_pytest_rs.log 2>&1; echo "pytest rs rc=$?"; tail -15 gpurun_out/${TAG}_pytest_rs.log
timeout 200 python -m pytest tests/test_gpu_umma.py -m gpu -x -q > gpurun_out/${TAG}_pytest_umma.log 2>&1; echo "pytest umma rc=$?"; tail -3 gpurun_out/${TAG}_pytest_umma.log
timeout 300 python tools/atomics_bench.py > gpurun_out/${TAG}_atomics.txt 2>&1; echo "atomics rc=$?"; cat gpurun_out/${TAG}_atomics.txt
timeout 600 python tools/rs_sweep.py > gpurun_out/${TAG}_rs_sweep.txt 2>&1; echo "sweep rc=$?"; cat gpurun_out/${TAG}_rs_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"resample2d_" -c 6 -o gpurun_out/${TAG}_prof_rs python tools/prof_ops.py small 1 > gpurun_out/${TAG}_prof.log 2>&1; echo "ncu rc=$?"
