#!/bin/bash
# EXPERIMENT helper: A/B alternative builds of libdeb200.so (build/alt/*.so) on the C2 kernel.
N=${1:-2000000}
for lib in differential-equations_b200/libdeb200.so build/alt/*.so; do
  echo -n "$lib: "
  DEB200_LIB=$PWD/$lib python bench.py --n-traj $N --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.3f G/s  kernel %.1f ms  frac %.3f' % (d['value']/1e9, d['roofline']['kernel_ms'], d['roofline']['frac']))
"
done
