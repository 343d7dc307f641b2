#!/usr/bin/env python
"""BASELINE.json configs[0] shape through the drop-in, the way the reference calls it (one read per
call, basecall.py:70-123): five reads with the bundled file's lengths, windows of 1024 every 128,
assemble_matrices -> beam_search with the default flags (bw 6, context 11, thresholds 0.5/0.5).
Synthetic posteriors and table (the model blobs are missing).  Reports the latency of
every call and the whole loop, next to the oracle on one host thread.
usage: python scripts/bench_c1.py > profiles/r2_c1_latency.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402  (checker / CPU baseline)
from radian_b200 import decode, matrix_assembly, synth  # noqa: E402

T_REAL = [12833, 14799, 4863, 9905, 11388]  # signal lengths of radian/data/reads.fast5
L, BW = 11, 6
tab = synth.make_table(L, 5)
lm = decode.RnaTable(tab)
post, off = synth.make_reads(np.array([t // 43 + 40 for t in T_REAL]), seed=1)
post, off = post.numpy(), off.numpy()
reads = [post[off[i]:off[i] + T_REAL[i]] for i in range(5)]
chunk_lists = [synth.split_windows(r, 1024, 128) for r in reads]
decode.beam_search(matrix_assembly.assemble_matrices(chunk_lists[2], 128), "ACGT", BW, lm, 0.5, 0.5, L, {})  # warm-up
out = {"config": {"frames": T_REAL, "beam_width": BW, "context_len": L, "chunk_len": 1024, "step_size": 128},
       "per_read": []}
t_loop = time.perf_counter()
seqs = []
for i, mats in enumerate(chunk_lists):
    t0 = time.perf_counter()
    m = matrix_assembly.assemble_matrices(mats, 128)
    t1 = time.perf_counter()
    s = decode.beam_search(m, "ACGT", BW, lm, 0.5, 0.5, L, {})
    t2 = time.perf_counter()
    seqs.append(s)
    out["per_read"].append({"frames": T_REAL[i], "windows": len(mats), "assemble_ms": 1e3 * (t1 - t0),
                            "beam_search_ms": 1e3 * (t2 - t1), "bases": len(s)})
out["loop_s"] = time.perf_counter() - t_loop
out["bases_per_s"] = sum(len(s) for s in seqs) / out["loop_s"]
t0 = time.perf_counter()
want = ["".join("ACGT"[c] for c in oracle.beam_search(oracle.assemble(mats, 128), BW, tab, L, 0.5, 0.5, topk=1)[0])
        for mats in chunk_lists]
out["oracle_one_thread_s"] = time.perf_counter() - t0
out["identical_to_oracle"] = seqs == want
print(json.dumps(out, indent=1))
