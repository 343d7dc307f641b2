#!/bin/bash
# throughput of the decode kernel at 1..6 resident CTAs per SM (run on the GPU box)
# usage: scripts/ab_occupancy.sh "<bench args>" [nvcc flags]
cfg="$1"; flags="$2"
RADIAN_NVCC_EXTRA="$flags" python radian_b200/build.py > /dev/null 2>&1 || { echo "build failed"; exit 1; }
for w in ${WS:-1 2 3 4 5 6}; do
  RADIAN_CTAS_PER_SM=$w python bench.py $cfg --no-cpu --no-e2e 2> /tmp/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$flags] ctas/sm=$w', 'frames/s %.4g' % d['roofline']['frames_per_s'], 'ms %.2f' % d['roofline']['kernel_ms'], d['parity_check'])"
done
python radian_b200/build.py > /dev/null 2>&1
