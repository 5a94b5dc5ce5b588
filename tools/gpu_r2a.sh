#!/bin/bash
# round-2 call A: topology + TS-mode operand check + sanitizer passes + host-copy probe (1 GPU)
TAG=r2a
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,pci.bus_id --format=csv; nproc; lscpu | grep -i -E "model name|socket|numa|thread|core"; nvidia-smi topo -m; \
  for d in /sys/bus/pci/devices/*; do c=$(cat $d/class 2>/dev/null); if [[ $c == 0x0302* || $c == 0x0300* ]]; then echo "$d numa_node=$(cat $d/numa_node)"; fi; done; free -g; } > gpurun_out/${TAG}_topo.txt 2>&1
timeout 200 python tools/umma_ts_check.py > gpurun_out/${TAG}_ts_check.log 2>&1; echo "ts_check rc=$?"; cat gpurun_out/${TAG}_ts_check.log | tail -4
timeout 200 python tools/hostcopy_probe.py > gpurun_out/${TAG}_hostcopy_1gpu.json 2> gpurun_out/${TAG}_hostcopy_1gpu.err; echo "hostcopy rc=$?"; cat gpurun_out/${TAG}_hostcopy_1gpu.json
bash tools/gpu_sanitize.sh ${TAG}
