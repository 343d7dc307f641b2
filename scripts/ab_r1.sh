#!/bin/bash
# the round-1 kernel on a bench configuration, for A/B against the current one (run on the GPU box;
# scripts/probes/r1_tree.tgz = `git archive` of the round-1 commit, not committed)
# usage: scripts/ab_r1.sh "<round-1 bench.py args>"
rm -rf /tmp/r1tree && tar xzf scripts/probes/r1_tree.tgz -C /tmp && cd /tmp/r1tree || exit 1
python -c "from radian_b200 import build; build.build(force=True)" > /dev/null 2>&1 || { echo "r1 build failed"; exit 1; }
python bench.py $1 --no-cpu --no-e2e 2> /tmp/r1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[r1] $1', 'frames/s %.4g' % d['roofline']['frames_per_s'], 'ms %.2f' % d['roofline']['kernel_ms'], d['parity_check'])"
