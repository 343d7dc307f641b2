#!/bin/bash
# kernel-only numbers for a list of bench workloads (run on the GPU box): one line per workload
# usage: scripts/run_cells.sh "<extra bench args>" workload [workload ...]
extra="$1"; shift
for wl in "$@"; do
  python bench.py --workload "$wl" --steps 3 --warmup 3 --no-cpu --no-e2e --check-reads 8 $extra 2> /tmp/cell.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$wl', 'reads', d['config']['reads_per_gpu_per_step'], 'frames/s %.4g' % d['roofline']['frames_per_s'], 'bases/s %.4g' % d['value'], 'ms %.2f' % d['roofline']['kernel_ms'], d['parity_check'])
except Exception as e:
    print('$wl', 'FAILED', e)"
  grep -v "^$" /tmp/cell.err | tail -3
done
