#!/bin/bash
# bench (no extras) for each producer-warp count: prints ms/step and the fwd / bwd kernel times
for np in ${@:-1 2 3 4}; do
  FN2B200_TC_NP=$np timeout 100 python bench.py --steps 10 --warmup 3 --no-extras > /tmp/np_$np.json 2> /tmp/np_$np.err
  python - "$np" <<'PY'
import json, sys
np_ = sys.argv[1]
try:
    d = json.loads(open('/tmp/np_%s.json' % np_).read())
    print("NP=%s ms_per_step=%.4f %s" % (np_, d["ms_per_step"], d["kernels"]))
except Exception as e:
    print("NP=%s failed: %s" % (np_, e)); print(open('/tmp/np_%s.err' % np_).read()[-600:])
PY
done
