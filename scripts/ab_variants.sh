#!/bin/bash
# A/B of compile-time variants of decode_kernel on the GPU box: rebuilds the library per variant.
# usage: scripts/ab_variants.sh "-DX=0" "-DX=1" ...
for v in "$@"; do
  RADIAN_NVCC_EXTRA="$v" python radian_b200/build.py > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  python bench.py --reads 35520 --fixed-len 500 --steps 3 --warmup 3 --no-cpu --no-e2e --check-reads 4 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'frames/s %.4g' % d['roofline']['frames_per_s'], d['parity_check'])"
done
