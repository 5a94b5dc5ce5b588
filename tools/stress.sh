#!/bin/bash
# repeat the correlation parity tests + a bench to shake out rare races / hangs (each run under timeout)
for i in 1 2 3 4 5 6; do
  timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "correlation" 2>&1 | tail -1
done
for i in 1 2 3; do
  timeout 200 python bench.py --steps 30 --warmup 3 --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done
