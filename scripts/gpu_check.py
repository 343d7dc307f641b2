"""First-contact diagnostic for the GPU box: runs every golden case and prints a mismatch
summary instead of stopping at the first failure."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_io  # noqa: E402
from radian_b200 import decode  # noqa: E402

files = ["decode_kat.npz", "decode_random.npz", "decode_synth.npz", "decode_long.npz", "decode_wide.npz"]
cases = [c for f in files for c in golden_io.decode_cases(f)]
bad = 0
tabs = {}
t0 = time.time()
for c in cases:
    lm = None
    if c.L:
        key = (c.L, c.tseed)
        if key not in tabs:
            if len(tabs) > 4:
                tabs.clear()
            tabs[key] = decode.RnaTable(golden_io.table(*key))
        lm = tabs[key]
    try:
        seqs, sc, cnt = decode.beam_search_batch([c.mat], c.bw, lm, c.s_thr, c.r_thr, c.L, return_details=True)
    except Exception as e:
        print("EXC", c, type(e).__name__, e)
        bad += 1
        continue
    want = "".join("ACGT"[s] for s in c.seq)
    w0 = c.scores[0]
    ok_seq = seqs[0] == want
    ok_sc = (sc[0, 0] == w0) or abs(sc[0, 0] - w0) <= 1e-9 * max(1, abs(w0))
    ok_cnt = int(cnt[0, 0]) == c.n_lookup and int(cnt[0, 1]) == c.n_combine
    if not (ok_seq and ok_sc and ok_cnt):
        bad += 1
        if bad <= 25:
            print("MISMATCH", c, "seq", ok_seq, "score", sc[0], "want", c.scores[:2], "cnt", cnt[0],
                  (c.n_lookup, c.n_combine))
            if not ok_seq:
                print("   got ", seqs[0][:80], len(seqs[0]))
                print("   want", want[:80], len(want))
print(f"{len(cases)} cases, {bad} bad, {time.time() - t0:.1f}s")
