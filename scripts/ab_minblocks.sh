#!/bin/bash
# A/B of resident CTAs per SM for the decode kernel (run on the GPU box).
for mb in ${MBS:-4 5 6}; do
  RADIAN_NVCC_EXTRA="-DRADIAN_MIN_BLOCKS=$mb" python -m radian_b200.build > /dev/null 2>&1
  for cfg in "--reads 35520 --fixed-len 500" "--reads 47360 --fixed-len 500 --no-lm --beam-width 6"; do
    python bench.py $cfg --steps 2 --warmup 1 --no-e2e --no-cpu --check-reads 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('minb=$mb', '$cfg', 'bases/s %.3g frames/s %.4g ms %.1f frac %.3f' % (d['value'], d['roofline']['frames_per_s'], d['ms_per_step'], d['roofline']['frac']), d['parity_check'])"
  done
done
python -m radian_b200.build > /dev/null 2>&1
