#!/usr/bin/env python
"""Per-read decode times of one launch (library built with -DRADIAN_READ_TIMES): which reads are slow?
usage: RADIAN_NVCC_EXTRA=-DRADIAN_READ_TIMES python radian_b200/build.py; python scripts/read_times.py [bench args]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from radian_b200 import decode, synth  # noqa: E402

sys.argv = [sys.argv[0]] + sys.argv[1:]
a = bench.parse()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
table = None if a.no_lm else decode.RnaTable(synth.make_table(a.context_len, 5), 0)
post, fo, nb = bench.make_batch(a, 0, dev)
T = fo[1:] - fo[:-1]
order = torch.argsort(T, descending=True).to(torch.int32)
so = torch.zeros(a.reads + 1, dtype=torch.int64, device=dev)
so[1:] = torch.cumsum(T // 4 + 64, 0)
for rep in range(2):
    res = decode.decode_batch_device(post, fo, a.beam_width, table, 0.5, 0.5, max_frames=int(T.max()), order=order,
                                     seq_offsets=so, counters=True)
    torch.cuda.synchronize()
c = res.counters.cpu().numpy().astype(np.uint64)
ns = (c[:, 3] >> np.uint64(32)).astype(np.float64)
slow = (c[:, 3] & np.uint64(0xffffffff)).astype(np.float64)
Tn = T.cpu().numpy().astype(np.float64)
per = ns / np.maximum(Tn, 1)
q = np.percentile(per, [0, 10, 50, 90, 99, 100])
print("ns per frame: min %.0f p10 %.0f median %.0f p90 %.0f p99 %.0f max %.0f" % tuple(q))
print("slow-frame share: median %.3f max %.3f" % (np.median(slow / Tn), (slow / Tn).max()))
worst = np.argsort(-per)[:8]
for i in worst:
    print("read %d: T %d, %.0f ns/frame, slow share %.3f, near-tie frames %d, bases %d" %
          (i, Tn[i], per[i], slow[i] / Tn[i], int(c[i, 2]), int(res.lengths[i])))
print("corr(ns/frame, slow share) = %.2f" % np.corrcoef(per, slow / Tn)[0, 1])
