#!/bin/bash
# EXPERIMENT helper: time the C2 kernel for each launch-shape variant (DEB_DP_VARIANT) at a reduced ensemble size.
N=${1:-2000000}
for v in 0 1 2 3 4 5 6; do
  echo -n "variant $v: "
  DEB_DP_VARIANT=$v python bench.py --n-traj $N --steps 2 --warmup 1 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.3f G/s  kernel %.1f ms  frac %.3f' % (d['value']/1e9, d['roofline']['kernel_ms'], d['roofline']['frac']))
    elif 'Error' in l or 'error' in l: print(l.strip())
"
done
