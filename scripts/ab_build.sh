#!/bin/bash
# A/B on the GPU box: build variants of the library (extra nvcc flags) and run one bench config each.
# usage: scripts/ab_build.sh "<bench args>" "<flags A>" "<flags B>" ...
cfg="$1"; shift
for v in "$@"; do
  RADIAN_NVCC_EXTRA="$v" python radian_b200/build.py > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  python bench.py $cfg --no-cpu --no-e2e 2> /tmp/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$v]', 'frames/s %.4g' % d['roofline']['frames_per_s'], 'ms %.2f' % d['roofline']['kernel_ms'], d['parity_check'])"
  grep stage2 /tmp/ab.err
done
python radian_b200/build.py > /dev/null 2>&1
