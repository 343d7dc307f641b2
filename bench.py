#!/usr/bin/env python
"""Benchmark of the RADIAN decode hot path on B200 (contract: see DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, all host threads)

Default workload (BASELINE.json configs[2], the one `metric` is quoted on): RNA-LM global decode,
beam width 16, 12-mer context, sig/rna thresholds 0.5/0.5, synthetic reads of the ~1.5 kb
LogNormal length distribution, float32 posteriors, synthetic dense table.
A "step" is one decode of one resident batch of `--reads` reads per GPU (default: all 100k reads
of the config, 132 GB of float32 posteriors).  Reads are independent: each rank owns its own batch
and table replica, there is no collective on the data path (weak scaling); ranks only exchange
step times and counts.  Other workloads of BASELINE.json (`--workload`):

    c2            configs[1]: pure CTC prefix beam search, model off, bw 6, 10k reads
    c3            configs[2]: the default
    c4-global     configs[3]: windows of 1024 frames every 128 -> overlap merge -> RNA-LM decode, bw 16
    c4-chunk      configs[3] as the reference's chunk mode: every window decoded with the model off
                  (8x the frames), fragments stitched
    c5:BW:L:KB    one cell of configs[4]: beam width BW, context L (0 = model off), reads of KB kilobases

Under ncu: the end-to-end leg streams its input while the kernel runs, which a kernel-replay
profiler cannot do; the library then copies first (it detects the injection, or set
RADIAN_HOST_COPY_FIRST=1), and `--no-e2e` skips the leg altogether.

Exit status: 3 if the parity spot check against the oracle found a mismatch (the JSON line is
still printed, with `parity_mismatches` > 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decoded bases/sec (RNA-LM beam search, bw=16)"
UNIT = "bases/s"
W_LEN, W_STEP = 1024, 128  # --chunk-len / --step-size of configs[3]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", default="c3", help="c2 | c3 | c4-global | c4-chunk | c5:BW:L:KB")
    ap.add_argument("--reads", type=int, default=None, help="reads per GPU per step (default: the workload's)")
    ap.add_argument("--beam-width", type=int, default=None)
    ap.add_argument("--context-len", type=int, default=None)
    ap.add_argument("--no-lm", action="store_true", help="model off (same as --workload c2 with --beam-width 6)")
    ap.add_argument("--fixed-len", type=int, default=None, help="bases per read instead of the LogNormal")
    ap.add_argument("--f64", action="store_true", help="float64 posteriors (assembled global matrices)")
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--ambiguity", type=float, default=0.3,
                    help="probability of a second, competing spike per base (SURVEY.md 8d knob)")
    ap.add_argument("--e2e-reads", type=int, default=16384, help="reads in the host-buffer end-to-end leg")
    ap.add_argument("--cpu-reads", type=int, default=512, help="reads in the CPU baseline sample (also parity-checked)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-python", action="store_true")
    ap.add_argument("--e2e-pageable", action="store_true", help="end-to-end leg from pageable host memory")
    ap.add_argument("--check-reads", type=int, default=4, help="reads re-decoded by the oracle after timing")
    ap.add_argument("--rotate", type=int, default=0,
                    help="rank r takes the batch of rank (r + rotate) %% world: tells a slow device from a slow batch")
    a = ap.parse_args()
    wl = a.workload.lower()
    a.kind = "decode"
    defaults = {"reads": 100000, "bw": 16, "L": 12}
    if wl == "c2":
        a.no_lm = True
        defaults = {"reads": 10000, "bw": 6, "L": 0}
    elif wl.startswith("c5:"):
        bw, L, kb = wl.split(":")[1:4]
        nbases = int(float(kb) * 1000)
        a.fixed_len = a.fixed_len or nbases
        defaults = {"reads": int(min(40000, max(2048, 1.0e9 / (43 * nbases)))), "bw": int(bw), "L": int(L)}
        a.no_lm = a.no_lm or int(L) == 0
    elif wl in ("c4-global", "c4-chunk"):
        a.kind = wl
        defaults = {"reads": 4096, "bw": 16, "L": 12 if wl == "c4-global" else 0}
        a.no_lm = wl == "c4-chunk"
    elif wl != "c3":
        ap.error(f"unknown workload {a.workload}")
    a.reads = a.reads or defaults["reads"]
    a.beam_width = a.beam_width or defaults["bw"]
    a.context_len = defaults["L"] if a.context_len is None else a.context_len
    if a.no_lm:
        a.context_len = 0
    return a


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def workload_name(a):
    if a.kind == "c4-global":
        return (f"C4 chunk-len {W_LEN} / step-size {W_STEP}: overlap merge + RNA-LM global decode, bw={a.beam_width}, "
                f"context_len={a.context_len}")
    if a.kind == "c4-chunk":
        return (f"C4 chunk mode: every {W_LEN}-frame window (step {W_STEP}) decoded with the model off, "
                f"bw={a.beam_width}, then stitched")
    if a.no_lm:
        return f"C2 pure CTC prefix beam search, RNA model off, bw={a.beam_width}"
    return (f"C3 RNA-LM global decode, bw={a.beam_width}, context_len={a.context_len}, "
            f"sig/rna thresholds 0.5/0.5")


def make_batch(a, rank, device):
    """Resident synthetic batch, generated on the device in pieces."""
    import torch

    from radian_b200 import synth

    nb = synth.read_lengths(a.reads, a.seed * 1000 + rank, fixed=a.fixed_len)
    n = len(nb)
    est_frames = int(nb.sum() * 43.5) + 4096
    dt = torch.float64 if a.f64 else torch.float32
    post = torch.empty((est_frames, 5), dtype=dt, device=device)
    offs = [torch.zeros(1, dtype=torch.int64, device=device)]
    pos = 0
    i = 0
    piece = 0
    amb = getattr(a, "ambiguity", 0.3)
    while i < n:
        j = i
        frames = 0
        while j < n and (frames == 0 or frames + nb[j] * 43 < 16_000_000):
            frames += nb[j] * 43
            j += 1
        p, o = synth.make_reads(nb[i:j], seed=a.seed * 7919 + rank * 104729 + piece, device=device, dtype=dt,
                                second_spike=amb)
        T = p.shape[0]
        if pos + T > post.shape[0]:
            grown = torch.empty((int((pos + T) * 1.05), 5), dtype=dt, device=device)
            grown[:pos] = post[:pos]
            post = grown
        post[pos:pos + T] = p
        offs.append(o[1:] + pos)
        pos += T
        i = j
        piece += 1
        del p, o
    frame_offsets = torch.cat(offs)
    return post[:pos], frame_offsets, nb


def cpu_sample(a, post_np, fo_np, table_np, threads, n_reads):
    """Decode a bounded sample with the oracle port on `threads` host threads -> (bases/s, desc, seqs)."""
    from oracle import oracle

    n = min(n_reads, len(fo_np) - 1)
    sub = post_np[fo_np[0]:fo_np[n]]
    fo = fo_np[:n + 1] - fo_np[0]
    t0 = time.perf_counter()
    seqs, _, _ = oracle.beam_search_batch(sub, fo, a.beam_width, table_np, a.context_len, 0.5, 0.5, threads=threads)
    dt = time.perf_counter() - t0
    bases = int(sum(len(s) for s in seqs))
    cpu_sample.last_sec = dt
    return bases / dt, f"{n} reads / {int(fo[-1])} frames / {bases} bases in {dt:.2f}s on {threads} threads", seqs


def cpu_sample_c4(a, reads_np, table_np, threads):
    """configs[3] on the host: the oracle's assemble + decode, or per-window decode + stitch."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle
    from radian_b200 import synth

    t0 = time.perf_counter()

    def one(m):
        mats = synth.split_windows(m, W_LEN, W_STEP)
        if a.kind == "c4-global":
            seq = oracle.beam_search(oracle.assemble(mats, W_STEP), a.beam_width, table_np, a.context_len, 0.5, 0.5,
                                     topk=1)[0]
            return "".join("ACGT"[s] for s in seq)
        frags = ["".join("ACGT"[s] for s in oracle.beam_search(w, a.beam_width, topk=1)[0]) for w in mats]
        return oracle.stitch(frags)[0]

    with ThreadPoolExecutor(threads) as ex:
        seqs = list(ex.map(one, reads_np))
    dt = time.perf_counter() - t0
    bases = sum(len(s) for s in seqs)
    cpu_sample.last_sec = dt
    frames = sum(len(m) for m in reads_np)
    return bases / dt, f"{len(reads_np)} reads / {frames} frames / {bases} bases in {dt:.2f}s on {threads} threads", seqs


def run_reference(a):
    """CPU arm: the oracle port (kind 'port': the reference itself is pure Python and does not
    exist on the GPU box) on all host threads, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from radian_b200 import synth

    threads = os.cpu_count() or 1
    table = None if a.no_lm else synth.make_table(a.context_len, 5)
    c4 = a.kind != "decode"
    n = max(threads, a.cpu_reads if not c4 else min(a.cpu_reads, 64))
    nb = synth.read_lengths(n, a.seed * 1000, fixed=a.fixed_len)
    post, off = synth.make_reads(nb, seed=a.seed * 7919, device="cpu", second_spike=a.ambiguity)
    post_np = post.numpy()
    if a.f64:
        post_np = post_np.astype(np.float64)
    fo_np = off.numpy()
    vals, secs = [], []
    desc = ""
    for s in range(a.warmup + a.steps):
        if c4:
            v, desc, _ = cpu_sample_c4(a, [post_np[fo_np[i]:fo_np[i + 1]] for i in range(n)], table, threads)
        else:
            v, desc, _ = cpu_sample(a, post_np, fo_np, table, threads, n)
        if s >= a.warmup:
            vals.append(v)
            secs.append(cpu_sample.last_sec)
    value = float(np.mean(vals))
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": desc, "reads_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback 6650"


def committed_traffic(a, frames):
    """DRAM bytes / issue-slot use of one launch from the committed ncu capture of this very workload
    (profiles/r2_traffic.json, regenerated whenever the kernel changes: it names the build it was
    taken from); None for any other configuration."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            tr = json.load(f)
        c = tr["config"]
        if (c["reads_per_gpu_per_step"] == a.reads and c["frames_per_gpu_per_step"] == frames
                and c["beam_width"] == a.beam_width and c["context_len"] == a.context_len and c["seed"] == a.seed
                and a.kind == "decode" and not a.f64 and a.fixed_len is None and a.ambiguity == 0.3):
            return tr
    except Exception:
        pass
    return None


def gather_per_rank(world, rank, device, vals):
    """all_gather of a few floats per rank -> list of lists (rank order)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(vals, dtype=torch.float64, device=device)
    if world == 1:
        return [t.tolist()]
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


def timed_steps(step, a, world, rank, local):
    """W warm-up calls, then K calls timed with CUDA events on the current stream, bracketed by a
    barrier + synchronize on both sides -> (per-step ms, total ms, clocks of this rank)."""
    import torch
    import torch.distributed as dist

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for s in range(a.steps):
        step()
        ev[s + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    step_ms = [ev[s].elapsed_time(ev[s + 1]) for s in range(a.steps)]
    return step_ms, ev[0].elapsed_time(ev[-1]), clocks


def e2e_leg(a, world, rank, local, post, fo, table):
    """The same metric through the reference-facing entry points with HOST buffers: page-locked
    posteriors -> radian_decode_batch_host (H2D + kernel + D2H inside) -> FASTA records
    (`>{id}\\n{seq[::-1]}\\n`, basecall.py:129) as one text buffer.  All ranks start together."""
    import torch
    import torch.distributed as dist

    from radian_b200 import _native, decode, fasta

    lib = decode.lib
    # with many ranks on one host the calls are halved (8 x 22 GB of locked memory is not a given)
    ne = min(a.e2e_reads if world <= 2 else a.e2e_reads // 2, a.reads)
    while True:
        fe = int(fo[ne].item())
        try:
            h_post = torch.empty((fe, 5), dtype=post.dtype)
            if not a.e2e_pageable:
                h_post = h_post.pin_memory()
            break
        except RuntimeError:
            if ne <= 256:
                raise
            ne //= 2
    h_post.copy_(post[:fe])
    post_np = h_post.numpy()
    fo_np = fo[:ne + 1].cpu().numpy()
    so = np.zeros(ne + 1, dtype=np.int64)
    so[1:] = np.cumsum((fo_np[1:] - fo_np[:-1]) // 4 + 64)
    seq = np.zeros(int(so[-1]), dtype=np.uint8)
    ln = np.zeros(ne, dtype=np.int64)
    sc = np.zeros((ne, 2))
    st = np.zeros(ne, dtype=np.int32)
    ids = fasta.pack_ids([f"read{rank}_{i}" for i in range(ne)])
    L = a.context_len

    def call():
        rc = lib.radian_decode_batch_host(
            _native.np_ptr(post_np), int(a.f64), _native.np_ptr(fo_np), ne, a.beam_width,
            table._h if table else None, L, 0.5, 0.5, _native.np_ptr(seq), _native.np_ptr(so), _native.np_ptr(ln),
            _native.np_ptr(sc), _native.np_ptr(st), None, local)
        _native.check(rc)
        return fasta.format_records(ids, seq, so, ln)

    txt = call()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        txt = call()
    dt = (time.perf_counter() - t0) / reps
    bases = int(ln.sum())
    out = {"bases": bases, "sec": dt, "reads": ne,
           "bytes_in": int(post_np.nbytes + fo_np.nbytes + so.nbytes),
           "bytes_out": int(seq.nbytes + ln.nbytes + sc.nbytes + st.nbytes), "fasta_bytes": len(txt)}
    # the Python drop-in on the same reads as a list of numpy arrays -> list of str
    if not a.no_e2e_python:
        npy = max(256, ne // 4)
        mats = [post_np[fo_np[i]:fo_np[i + 1]] for i in range(npy)]
        decode.beam_search_batch(mats[:64], a.beam_width, table, 0.5, 0.5, L)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        strs = decode.beam_search_batch(mats, a.beam_width, table, 0.5, 0.5, L)
        dtp = time.perf_counter() - t0
        out.update({"py_bases": int(sum(len(s) for s in strs)), "py_sec": dtp, "py_reads": npy})
        k = int(np.argmax(ln[:npy]))
        rec = bytes(txt).split(b"\n")
        assert strs[k] == rec[2 * k + 1][::-1].decode(), "drop-in and C ABI disagree"
    return out


def run_decode(a, world, rank, local, device):
    import torch

    from radian_b200 import decode, synth

    L = a.context_len
    table_np = None if a.no_lm else synth.make_table(L, 5)
    table = decode.RnaTable(table_np, local) if table_np is not None else None
    post, fo, nb = make_batch(a, (rank + a.rotate) % world, device)
    T = fo[1:] - fo[:-1]
    max_frames = int(T.max().item())
    frames = int(post.shape[0])
    order = torch.argsort(T, descending=True).to(torch.int32)
    # output slots: a decoded read is far shorter than its frame count; T/4+64 symbols per read
    seq_offsets = torch.zeros(a.reads + 1, dtype=torch.int64, device=device)
    seq_offsets[1:] = torch.cumsum(T // 4 + 64, 0)

    def step(counters=False, out=None):
        return decode.decode_batch_device(post, fo, a.beam_width, table, 0.5, 0.5, max_frames=max_frames,
                                          order=order, seq_offsets=seq_offsets, counters=counters, out=out)

    # counting pass (untimed): N_lookup for the algorithmic byte count, and status check
    res = step(counters=True)
    torch.cuda.synchronize()
    status = res.status.cpu().numpy()
    if status.any():
        raise SystemExit(f"bench.py: {int((status != 0).sum())} reads failed, status codes {np.unique(status)}")
    n_lookup = int(res.counters[:, 0].sum().item())
    n_combine = int(res.counters[:, 1].sum().item())
    gate_open = int((res.counters[:, 3] >> 32).sum().item())
    slow = int((res.counters[:, 3] & 0xffffffff).sum().item())
    bases = int(res.lengths.sum().item())
    del res
    out = step()
    step_ms, total_ms, clocks = timed_steps(lambda: step(out=out), a, world, rank, local)
    assert int(out.lengths.sum().item()) == bases

    torch.cuda.empty_cache()  # (blocks torch cached during the counting pass: the e2e leg allocates outside torch)
    e2e = None if a.no_e2e else e2e_leg(a, world, rank, local, post, fo, table)

    # ---- parity spot check against the oracle (outside every timed region)
    check = None
    cpu = None
    if rank == 0 and (a.check_reads > 0 or not a.no_cpu):
        nc = min(max(a.check_reads, 0 if a.no_cpu else a.cpu_reads), a.reads)
        idx = np.arange(nc)  # the first reads of the batch: an unbiased sample of the length distribution
        sub = [post[int(fo[i]):int(fo[i + 1])].cpu().numpy() for i in idx]
        fo_s = np.zeros(nc + 1, dtype=np.int64)
        fo_s[1:] = np.cumsum([m.shape[0] for m in sub])
        threads = os.cpu_count() or 1
        cpu_v, cpu_desc, seqs = cpu_sample(a, np.concatenate(sub), fo_s, table_np, threads, nc)
        mine = out.seq.cpu().numpy()
        so_np = seq_offsets.cpu().numpy()
        ln_np = out.lengths.cpu().numpy()
        bad = sum(not np.array_equal(mine[so_np[i]:so_np[i] + ln_np[i]], seqs[k]) for k, i in enumerate(idx))
        check = {"reads": int(nc), "mismatches": int(bad)}
        cpu = {"value": cpu_v, "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu_desc}

    per_rank = gather_per_rank(world, rank, device, [
        float(rank), float(local), float(np.mean(step_ms)), float(total_ms), float(frames), float(bases),
        float(n_lookup), float(n_combine), float(gate_open), float(slow), float(clocks["sm_mhz"] or 0),
        e2e["bases"] if e2e else 0.0, e2e["sec"] if e2e else 0.0,
        e2e.get("py_bases", 0.0) if e2e else 0.0, e2e.get("py_sec", 0.0) if e2e else 0.0])
    if rank != 0:
        return 0
    pr = np.array(per_rank)
    hbm, peak_src = peaks()
    ms_per_step = float(pr[:, 3].max()) / a.steps
    tot_bases = float(pr[:, 5].sum())
    value = tot_bases / (ms_per_step * 1e-3)
    # roofline of the dominant (only) kernel, per launch on rank 0:
    # B_alg = 20 B x frames + 16 B x lm[context] reads the reference performs
    kernel_ms = float(np.mean(step_ms))
    row_b = 40 if a.f64 else 20
    b_alg = row_b * frames + 16 * n_lookup
    b_min = row_b * frames + 16 * n_combine
    sec = kernel_ms * 1e-3
    tr = committed_traffic(a, frames)
    roof = {"bound": "hbm", "achieved": b_alg / sec / 1e9, "peak": hbm, "unit": "GB/s",
            "frac": b_alg / sec / 1e9 / hbm, "traffic": tr["traffic_bytes_per_launch"] if tr else None,
            "traffic_source": tr["source"] if tr else None, "peak_source": peak_src,
            "kernel": "decode_kernel",
            # three fractions of the same measured HBM peak: algorithmic bytes of the reference's
            # access pattern (frac = frac_alg), only the table rows the reference actually mixes in
            # (frac_min), and what the kernel really moves through DRAM (frac_dram: rows ride in registers)
            "frac_alg": b_alg / sec / 1e9 / hbm, "frac_min": b_min / sec / 1e9 / hbm,
            "frac_dram": (tr["traffic_bytes_per_launch"] / sec / 1e9 / hbm) if tr else None,
            "issue_active_pct": tr.get("issue_active_pct") if tr else None,
            "note": "the kernel is bound by instruction issue and dependent latency, not by HBM: frac_alg is the "
                    "SURVEY 8(d) label, frac_dram the physical one",
            "algorithmic_bytes_per_launch": b_alg, "bytes_min_per_launch": b_min,
            "frames_per_s": frames / sec, "kernel_ms": kernel_ms,
            "n_lookup_per_frame": n_lookup / max(frames, 1),
            "gate_open_frac": gate_open / max(frames, 1), "quiet_frame_frac": 1.0 - slow / max(frames, 1)}
    e2e_line = None
    if e2e:
        # all ranks started together: whole-job bases over the slowest rank's time
        e2e_line = {"value": float(pr[:, 11].sum() / pr[:, 12].max()), "unit": UNIT,
                    "h2d_bytes_per_step": e2e["bytes_in"] * world, "d2h_bytes_per_step": e2e["bytes_out"] * world,
                    "h2d_gbps_per_rank": e2e["bytes_in"] / float(pr[:, 12].max()) / 1e9,
                    "reads_per_call": e2e["reads"], "sec_per_call": e2e["sec"],
                    "what": "radian_decode_batch_host on page-locked buffers + FASTA records formed, per rank",
                    "fasta_bytes_per_call": e2e["fasta_bytes"]}
        if pr[:, 14].max() > 0:
            e2e_line["python_dropin"] = {
                "value": float(pr[:, 13].sum() / pr[:, 14].max()), "unit": UNIT, "reads_per_call": e2e["py_reads"],
                "what": "decode.beam_search_batch(list of numpy arrays) -> list of str (pageable memory)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "reads_per_gpu_per_step": a.reads,
                   "frames_per_gpu_per_step": frames, "bases_per_gpu_per_step": bases,
                   "posterior_dtype": "f64" if a.f64 else "f32", "table": "synthetic dense 4^L x 4 f64",
                   "read_lengths": f"fixed {a.fixed_len}" if a.fixed_len else "LogNormal(1300,0.6) in [200,10000] bases",
                   "second_spike_prob": a.ambiguity,
                   "l2": f"inputs {post.element_size() * post.numel() / 1e9:.1f} GB per step, far larger than L2",
                   "parallelism": f"reads sharded over {world} GPU(s), table replicated, no collective"},
        "roofline": roof,
        "cpu_baseline": cpu,
        "e2e": e2e_line,
        "gpu_launches": a.steps,
        "clocks": clocks,
        "parity_check": check,
        "parity_mismatches": check["mismatches"] if check else None,
        "per_rank": [{"rank": int(r[0]), "device": int(r[1]), "kernel_ms": r[2], "frames": int(r[4]),
                      "bases": int(r[5]), "sm_mhz": r[10], "batch_of_rank": int((r[0] + a.rotate) % world)} for r in pr],
    }
    print(json.dumps(line))
    return 3 if check and check["mismatches"] else 0


def window_layout(T, fo_h):
    """Chunk row offsets / read chunk ranges / source (row, length) of every window of
    preprocess.get_windows + the trim of basecall.py:96."""
    cro, rcr, src = [0], [0], []
    for r, t in enumerate(T):
        start = 0
        while start + W_LEN <= t:
            src.append((fo_h[r] + start, W_LEN))
            cro.append(cro[-1] + W_LEN)
            start += W_STEP
        src.append((fo_h[r] + start, int(t - start)))
        cro.append(cro[-1] + int(t - start))
        rcr.append(len(cro) - 1)
    return np.asarray(cro, np.int64), np.asarray(rcr, np.int64), src


def run_c4(a, world, rank, local, device):
    """configs[3] with everything resident: window matrices as the signal model would leave them."""
    import torch

    from radian_b200 import decode, matrix_assembly, sequence_assembly, synth

    post, fo, nb = make_batch(a, (rank + a.rotate) % world, device)
    T = (fo[1:] - fo[:-1]).cpu().numpy()
    fo_h = fo.cpu().numpy()
    cro, rcr, src = window_layout(T, fo_h)
    n_chunks = len(src)
    starts = torch.tensor([s for s, _ in src], dtype=torch.int64, device=device)
    lens = torch.tensor([ln for _, ln in src], dtype=torch.int64, device=device)
    d_cro = torch.from_numpy(cro).to(device)
    rep = torch.repeat_interleave(torch.arange(n_chunks, device=device), lens)
    idx = starts[rep] + (torch.arange(int(cro[-1]), device=device) - d_cro[:-1][rep])
    chunks = post[idx].contiguous()
    del idx, rep
    frames = int(T.sum())
    table_np = None
    table = None
    if a.kind == "c4-global":
        table_np = synth.make_table(a.context_len, 5)
        table = decode.RnaTable(table_np, local)
        plan = matrix_assembly.AssemblePlan(cro, rcr, W_STEP, device)
        mat, oro = matrix_assembly.assemble_batch_device(chunks, plan=plan)
        Tt = fo[1:] - fo[:-1]
        order = torch.argsort(Tt, descending=True).to(torch.int32)
        mf = int(Tt.max())
        res = decode.decode_batch_device(mat, oro, a.beam_width, table, 0.5, 0.5, max_frames=mf, order=order)

        def step():
            matrix_assembly.assemble_batch_device(chunks, plan=plan, out=mat)
            decode.decode_batch_device(mat, oro, a.beam_width, table, 0.5, 0.5, max_frames=mf, order=order, out=res)

        launches = 2
        decoded_frames = frames
    else:
        clen = torch.from_numpy(np.diff(cro)).to(device)
        corder = torch.argsort(clen, descending=True).to(torch.int32)
        so = torch.zeros(n_chunks + 1, dtype=torch.int64, device=device)
        so[1:] = torch.cumsum(clen // 2 + 8, 0)
        res = decode.decode_batch_device(chunks, d_cro, a.beam_width, None, max_frames=W_LEN, order=corder,
                                         seq_offsets=so)
        d_rcr = torch.from_numpy(rcr).to(device)
        fstart = so[:-1].contiguous()
        stitched = [None]

        def step():
            decode.decode_batch_device(chunks, d_cro, a.beam_width, None, max_frames=W_LEN, order=corder,
                                       seq_offsets=so, out=res)
            stitched[0] = sequence_assembly.stitch_device(res.seq, fstart, res.lengths, d_rcr)

        launches = 5
        decoded_frames = int(cro[-1])
    step()
    torch.cuda.synchronize()
    assert int(res.status.abs().sum()) == 0
    step_ms, total_ms, clocks = timed_steps(step, a, world, rank, local)
    if a.kind == "c4-global":
        got = res.strings()
    else:
        dseq, doff, dlen, dst = stitched[0]
        assert int(dst.abs().sum()) == 0
        hs, ho, hl = dseq.cpu().numpy(), doff.cpu().numpy(), dlen.cpu().numpy()
        got = ["".join("ACGT"[s] for s in hs[ho[r]:ho[r] + hl[r]]) for r in range(a.reads)]
    bases = sum(len(g) for g in got)
    check = cpu = None
    if rank == 0 and (a.check_reads > 0 or not a.no_cpu):
        nc = min(max(a.check_reads, 0 if a.no_cpu else min(a.cpu_reads, 64)), a.reads)
        sub = [post[int(fo[i]):int(fo[i + 1])].cpu().numpy() for i in range(nc)]
        threads = os.cpu_count() or 1
        cpu_v, cpu_desc, seqs = cpu_sample_c4(a, sub, table_np, threads)
        bad = sum(got[i] != seqs[i] for i in range(nc))
        check = {"reads": nc, "mismatches": int(bad)}
        cpu = {"value": cpu_v, "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu_desc}
    # end to end through the drop-in: lists of numpy window matrices -> strings (basecall.basecall_batch)
    e2e = None
    if not a.no_e2e:
        import types

        import torch.distributed as dist

        from radian_b200 import basecall

        ne = min(512 if world <= 2 else 256, a.reads)
        ch = chunks[:int(cro[rcr[ne]])].cpu().numpy()
        chunk_lists = [[ch[cro[k]:cro[k + 1]] for k in range(int(rcr[r]), int(rcr[r + 1]))] for r in range(ne)]
        args = types.SimpleNamespace(decode_type="global" if a.kind == "c4-global" else "chunk",
                                     beam_width=a.beam_width, step_size=W_STEP, sig_threshold=0.5, rna_threshold=0.5,
                                     context_len=a.context_len)
        basecall.basecall_batch(None, chunk_lists[:8], args, table)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        seqs = basecall.basecall_batch(None, chunk_lists, args, table)
        dt = time.perf_counter() - t0
        assert seqs[:4] == got[:4]
        e2e = {"bases": sum(len(s) for s in seqs), "sec": dt, "reads": ne, "bytes_in": int(ch.nbytes)}
    per_rank = gather_per_rank(world, rank, device, [
        float(rank), float(local), float(np.mean(step_ms)), float(total_ms), float(frames), float(bases),
        float(clocks["sm_mhz"] or 0), e2e["bases"] if e2e else 0.0, e2e["sec"] if e2e else 0.0])
    if rank != 0:
        return 0
    pr = np.array(per_rank)
    hbm, peak_src = peaks()
    ms_per_step = float(pr[:, 3].max()) / a.steps
    value = float(pr[:, 5].sum()) / (ms_per_step * 1e-3)
    kernel_ms = float(np.mean(step_ms))
    sec = kernel_ms * 1e-3
    # algorithmic bytes: the decoder reads every decoded frame once (20 B float32 windows, 40 B float64
    # assembled rows); the merge reads 20 B per window row and writes 40 B per assembled row
    b_alg = 20 * decoded_frames if a.kind == "c4-chunk" else 20 * int(cro[-1]) + 40 * frames + 40 * frames
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "reads_per_gpu_per_step": a.reads, "frames_per_gpu_per_step": frames,
                   "windows_per_gpu_per_step": n_chunks, "decoded_frames_per_gpu_per_step": decoded_frames,
                   "bases_per_gpu_per_step": bases, "read_lengths": "LogNormal(1300,0.6) in [200,10000] bases",
                   "l2": f"window matrices {chunks.element_size() * chunks.numel() / 1e9:.1f} GB per step, far larger than L2",
                   "parallelism": f"reads sharded over {world} GPU(s), table replicated, no collective"},
        "roofline": {"bound": "hbm", "achieved": b_alg / sec / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": b_alg / sec / 1e9 / hbm, "traffic": None, "peak_source": peak_src,
                     "kernel": ("decode_kernel (+ assemble_kernel)" if a.kind == "c4-global"
                                else "decode_kernel (+ stitch kernels)"),
                     "algorithmic_bytes_per_launch": b_alg, "kernel_ms": kernel_ms,
                     "frames_per_s": decoded_frames / sec},
        "cpu_baseline": cpu,
        "e2e": ({"value": float(pr[:, 7].sum() / pr[:, 8].max()), "unit": UNIT, "h2d_bytes_per_step": e2e["bytes_in"],
                 "d2h_bytes_per_step": e2e["bases"], "reads_per_call": e2e["reads"], "sec_per_call": e2e["sec"],
                 "what": "basecall.basecall_batch on lists of numpy window matrices -> strings, per rank"} if e2e else None),
        "gpu_launches": a.steps * launches,
        "clocks": clocks,
        "parity_check": check,
        "parity_mismatches": check["mismatches"] if check else None,
        "per_rank": [{"rank": int(r[0]), "device": int(r[1]), "kernel_ms": r[2], "frames": int(r[4]),
                      "bases": int(r[5]), "sm_mhz": r[6]} for r in pr],
    }
    print(json.dumps(line))
    return 3 if check and check["mismatches"] else 0


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return 0

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    try:
        rc = run_decode(a, world, rank, local, device) if a.kind == "decode" else run_c4(a, world, rank, local, device)
    finally:
        if world > 1:
            dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
