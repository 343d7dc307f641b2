#!/usr/bin/env python
"""Throughput sweep of BASELINE.json configs[4] at reduced read counts (one GPU):
beam width x context length x read length, kernel-only (resident inputs, CUDA events).
Writes one JSON line per cell.  Usage: python scripts/sweep.py [--quick] > profiles/rN_sweep.jsonl"""
import argparse
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from radian_b200 import decode, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--frames", type=float, default=1.0e9, help="target frames per cell")
ap.add_argument("--bw", type=int, nargs="*", default=[6, 8, 16, 32, 64])
ap.add_argument("--ctx", type=int, nargs="*", default=[0, 6, 8, 10, 11, 12])
ap.add_argument("--kb", type=float, nargs="*", default=[0.5, 1, 2, 5, 10])
args = ap.parse_args()
if args.quick:
    args.bw, args.ctx, args.kb, args.frames = [6, 16, 64], [0, 12], [1, 5], 3e8

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
peak = 6546.6
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
for kb in args.kb:
    nbases = int(kb * 1000)
    reads = int(min(40000, max(2048, args.frames / (43 * nbases))))
    a = types.SimpleNamespace(reads=reads, seed=11, fixed_len=nbases, f64=False)
    post, fo, nb = bench.make_batch(a, 0, dev)
    T = fo[1:] - fo[:-1]
    order = torch.argsort(T, descending=True).to(torch.int32)
    so = torch.zeros(reads + 1, dtype=torch.int64, device=dev)
    so[1:] = torch.cumsum(T // 4 + 64, 0)
    frames = int(post.shape[0])
    for L in args.ctx:
        table = decode.RnaTable(synth.make_table(L, 5), 0) if L else None
        for bw in args.bw:
            def run(counters=False, out=None):
                return decode.decode_batch_device(post, fo, bw, table, 0.5, 0.5, max_frames=int(T.max()), order=order,
                                                  seq_offsets=so, counters=counters, out=out)
            res = run(counters=True)
            torch.cuda.synchronize()
            st = res.status.cpu().numpy()
            n_lookup = int(res.counters[:, 0].sum())
            bases = int(res.lengths.sum())
            out = run()
            run(out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                run(out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 2
            b_alg = 20 * frames + 16 * n_lookup
            print(json.dumps({"read_kb": kb, "reads": reads, "frames": frames, "context_len": L, "beam_width": bw,
                              "kernel_ms": ms, "bases_per_s": bases / (ms * 1e-3), "frames_per_s": frames / (ms * 1e-3),
                              "algorithmic_GBps": b_alg / (ms * 1e-3) / 1e9, "roofline_frac": b_alg / (ms * 1e-3) / 1e9 / peak,
                              "failed_reads": int((st != 0).sum())}), flush=True)
        del table
    del post, fo
    torch.cuda.empty_cache()
