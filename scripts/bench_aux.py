#!/usr/bin/env python
"""Kernel measurements for the widened rows of the scope table on one B200:
preprocessing (normalise, windows) timed alone with CUDA events on resident data, stitching through
the host entry point with its kernel times from RADIAN_TRACE, each beside the CPU oracle on a
bounded sample.  Usage: RADIAN_TRACE=1 python scripts/bench_aux.py > profiles/rN_aux_bench.json"""
import ctypes
import json
import os
import re
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402  (CPU baseline only)
from radian_b200 import _native, sequence_assembly, synth  # noqa: E402

lib = _native.lib
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6546.6)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
out = {}
threads = os.cpu_count() or 1

# ---------------------------------------------------------------- N2: mad_normalise + get_windows
n_reads = 20000
nb = synth.read_lengths(n_reads, 5)
T = (nb * 43).astype(np.int64)
off = np.zeros(n_reads + 1, np.int64)
off[1:] = np.cumsum(T)
g = torch.Generator(device=dev)
g.manual_seed(1)
sig = (torch.randn(int(off[-1]), device=dev, generator=g) * 70 + 680).round().clamp(-32768, 32767).to(torch.int16)
d_off = torch.from_numpy(off).to(dev)
norm = torch.empty(int(off[-1]), dtype=torch.float64, device=dev)
is_int = torch.zeros(n_reads, dtype=torch.int32, device=dev)
status = torch.zeros(n_reads, dtype=torch.int32, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def normalise():
    _native.check(lib.radian_normalise_batch_dev(sig.data_ptr(), d_off.data_ptr(), n_reads, 4.0, 1, norm.data_ptr(),
                                                 is_int.data_ptr(), status.data_ptr(), st))


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(normalise)
assert int(status.sum()) == 0
samples = int(off[-1])
k = 64
t0 = time.perf_counter()
sig_h = [sig[off[i]:off[i + 1]].cpu().numpy() for i in range(k)]
with ThreadPoolExecutor(threads) as ex:
    t0 = time.perf_counter()
    ref = list(ex.map(lambda s: oracle.mad_normalise(s, 4), sig_h))
    cpu_dt = time.perf_counter() - t0
for i in range(k):
    assert np.array_equal(ref[i].view(np.int64), norm[off[i]:off[i + 1]].cpu().numpy().view(np.int64))
out["normalise"] = {"kernel": "normalise_kernel", "reads": n_reads, "samples": samples, "kernel_ms": ms,
                    "samples_per_s": samples / (ms * 1e-3), "algorithmic_bytes_per_sample": 10,
                    "algorithmic_GBps": 10 * samples / (ms * 1e-3) / 1e9, "peak_GBps": peak,
                    "frac": 10 * samples / (ms * 1e-3) / 1e9 / peak,
                    "cpu_baseline": {"samples_per_s": sum(len(s) for s in sig_h) / cpu_dt, "cores": threads,
                                     "kind": "port", "sample": f"{k} reads"},
                    "bit_exact_vs_oracle_reads": k}

W, S = 1024, 128
nw = np.zeros(n_reads, np.int64)
pad = np.zeros(n_reads, np.int32)
_native.check(lib.radian_windows_plan(_native.np_ptr(off), n_reads, W, S, _native.np_ptr(nw), _native.np_ptr(pad)))
sub = 4000  # windows are 8x the signal: a fifth of the reads keeps the output at ~17 GB
woff = np.zeros(sub + 1, np.int64)
woff[1:] = np.cumsum(nw[:sub])
d_woff = torch.from_numpy(woff).to(dev)
wins = torch.empty((int(woff[-1]), W), dtype=torch.float64, device=dev)


def windows():
    _native.check(lib.radian_windows_batch_dev(norm.data_ptr(), d_off.data_ptr(), d_woff.data_ptr(), sub, W, S,
                                               wins.data_ptr(), st))


ms = timed(windows)
w0, p0 = oracle.get_windows(norm[:off[1]].cpu().numpy(), W, S)
assert np.array_equal(w0, wins[:woff[1]].cpu().numpy()) and p0 == pad[0]
wbytes = 8 * int(off[sub]) + 8 * W * int(woff[-1])
out["windows"] = {"kernel": "windows_kernel", "reads": sub, "windows": int(woff[-1]), "kernel_ms": ms,
                  "algorithmic_GBps": wbytes / (ms * 1e-3) / 1e9, "peak_GBps": peak,
                  "frac": wbytes / (ms * 1e-3) / 1e9 / peak}
del wins, norm, sig
torch.cuda.empty_cache()

# ---------------------------------------------------------------- N1: stitching
rng = np.random.default_rng(3)
n_reads = 4000
frag_lists = []
for r in range(n_reads):
    n = int(nb[r])
    truth = rng.integers(0, 4, n).astype(np.uint8)
    frags, start = [], 0
    while start < n:  # a 1024-frame window every 128 frames: ~24 bases every ~3
        f = truth[start:start + int(rng.integers(18, 31))].copy()
        flip = rng.random(f.size) < 0.03
        f[flip] = rng.integers(0, 4, int(flip.sum()))
        frags.append(f)
        start += int(rng.integers(2, 5))
    frag_lists.append(frags)
n_frags = sum(len(f) for f in frag_lists)
import io
import contextlib

sequence_assembly.stitch_batch(frag_lists[:8])
rfd, wfd = os.pipe()
saved = os.dup(2)
os.dup2(wfd, 2)
t0 = time.perf_counter()
res = sequence_assembly.stitch_batch(frag_lists)
dt = time.perf_counter() - t0
os.dup2(saved, 2)
os.close(wfd)
trace = os.read(rfd, 65536).decode()
m = re.search(r"pair\+place ([0-9.]+) ms, vote\+argmax ([0-9.]+) ms", trace)
k = 64
t0 = time.perf_counter()
with ThreadPoolExecutor(threads) as ex:
    ref = list(ex.map(lambda fl: oracle.stitch(["".join("ACGT"[s] for s in f) for f in fl])[0], frag_lists[:k]))
cpu_dt = time.perf_counter() - t0
assert ref == res[:k]
bases = sum(len(s) for s in res)
out["stitch"] = {"kernels": "pair_kernel + place_kernel + vote_kernel + argmax_kernel", "reads": n_reads,
                 "fragments": n_frags, "consensus_bases": bases,
                 "pair_place_ms": float(m.group(1)) if m else None, "vote_argmax_ms": float(m.group(2)) if m else None,
                 "fragments_per_s_kernels": n_frags / ((float(m.group(1)) + float(m.group(2))) * 1e-3) if m else None,
                 "host_call_s": dt, "bases_per_s_host_call": bases / dt,
                 "cpu_baseline": {"bases_per_s": sum(len(s) for s in ref) / cpu_dt, "cores": threads, "kind": "port",
                                  "sample": f"{k} reads"},
                 "identical_to_oracle_reads": k,
                 "note": "set RADIAN_TRACE=1 for the kernel times; the host call includes Python list handling"}
print(json.dumps(out, indent=1))
