#!/usr/bin/env python
"""Where the cycles of the wide decode kernel (beam widths 33..128) go (library built with -DRADIAN_WIDE_PROBE).
usage: RADIAN_NVCC_EXTRA=-DRADIAN_WIDE_PROBE python radian_b200/build.py; python scripts/wide_probe.py --workload c5:64:12:1"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from radian_b200 import decode, synth  # noqa: E402

a = bench.parse()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
table = None if a.no_lm else decode.RnaTable(synth.make_table(a.context_len, 5), 0)
post, fo, nb = bench.make_batch(a, 0, dev)
T = fo[1:] - fo[:-1]
order = torch.argsort(T, descending=True).to(torch.int32)
so = torch.zeros(a.reads + 1, dtype=torch.int64, device=dev)
so[1:] = torch.cumsum(T // 4 + 64, 0)
import ctypes  # noqa: E402
from radian_b200 import _native  # noqa: E402

lib = ctypes.CDLL(_native.LIB_PATH) if hasattr(_native, "LIB_PATH") else ctypes.CDLL(os.path.join(ROOT, "radian_b200", "libradian_b200.so"))
laps = (ctypes.c_ulonglong * 24)()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(2):
    lib.radian_debug_wide_laps(laps, 1)
    e0.record()
    res = decode.decode_batch_device(post, fo, a.beam_width, table, 0.5, 0.5, max_frames=int(T.max()), order=order,
                                     seq_offsets=so, counters=True)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
assert lib.radian_debug_wide_laps(laps, 0) == 0
v = [float(x) for x in laps]
frames = float(T.sum())
names = ["quiet loop", "wait for table rows in flight", "nursery + rescale", "copy, merge, order check", "extensions, who competes",
         "candidate list", "incremental ranks", "exact ranks", "survivors, list of new beams", "staging + creation of new beams",
         "KeyError check, restaging", "orphans find their parents", "merge masks", "records, prefetch, loop", "successors, tie flags", "refresh"]
slow = v[16] + v[17]
tot = sum(v[:16])
print("bw %d: %.4g frames/s (launch %.1f ms), reads %d; quiet %.3f of the frames, ranked %.3f; %.0f cycles per frame of a read" %
      (a.beam_width, frames / ms * 1e3, ms, a.reads, 1 - slow / frames, v[17] / frames, tot / frames))
for i, n in enumerate(names):
    per = {0: frames - slow, 1: slow, 2: slow, 3: slow, 4: slow, 13: frames, 15: slow}.get(i, v[17])
    print("  %-36s %5.1f %% of the cycles, %7.0f cycles per %s" %
          (n, 100 * v[i] / tot, v[i] / max(per, 1), {0: "quiet frame", 13: "frame"}.get(i, "frame the long way" if per == slow else "ranked frame")))
