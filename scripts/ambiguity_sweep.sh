#!/bin/bash
# how the headline moves with the ambiguity of the posteriors (SURVEY.md 8d knob): probability of a second,
# competing spike per base; kernel-only, 20 000 reads per point (run on the GPU box)
for p in 0 0.15 0.3 0.45 0.6; do
  python bench.py --reads 20000 --ambiguity $p --steps 3 --warmup 3 --no-cpu --no-e2e --check-reads 16 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('second_spike_prob $p', 'bases/s %.4g' % d['value'], 'frames/s %.4g' % r['frames_per_s'], 'gate_open_frac %.3f' % r['gate_open_frac'], 'quiet_frame_frac %.3f' % r['quiet_frame_frac'], 'n_lookup/frame %.1f' % r['n_lookup_per_frame'], 'frac_alg %.2f frac_min %.2f' % (r['frac_alg'], r['frac_min']), d['parity_check'])"
done
