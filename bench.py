#!/usr/bin/env python
"""Benchmark of the RADIAN decode hot path on B200 (contract: see DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, all host threads)

Workload (BASELINE.json configs[2], the one `metric` is quoted on): RNA-LM global decode,
beam width 16, 12-mer context, sig/rna thresholds 0.5/0.5, synthetic reads of the ~1.5 kb
LogNormal length distribution, float32 posteriors, synthetic dense table.
A "step" is one decode of one resident batch of `--reads` reads per GPU (default: all 100k reads
of the config, 132 GB of float32 posteriors).  Reads are independent:
each rank owns its own batch and table replica, there is no collective on the data path
(weak scaling); ranks only exchange the step time (max) and the decoded base count (sum).

Under ncu: the end-to-end leg streams its input while the kernel runs, which a kernel-replay
profiler cannot do; the library then copies first (it detects the injection, or set
RADIAN_HOST_COPY_FIRST=1), and `--no-e2e` skips the leg altogether.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--reads", type=int, default=100000, help="reads per GPU per step (config 3: 100k)")
    ap.add_argument("--beam-width", type=int, default=16)
    ap.add_argument("--context-len", type=int, default=12)
    ap.add_argument("--no-lm", action="store_true", help="config 2: pure CTC (use with --beam-width 6)")
    ap.add_argument("--fixed-len", type=int, default=None, help="bases per read instead of the LogNormal")
    ap.add_argument("--f64", action="store_true", help="float64 posteriors (assembled global matrices)")
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--e2e-reads", type=int, default=16384, help="reads in the host-buffer end-to-end leg")
    ap.add_argument("--cpu-reads", type=int, default=512, help="reads in the CPU baseline sample (also parity-checked)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-pageable", action="store_true", help="end-to-end leg from pageable host memory")
    ap.add_argument("--check-reads", type=int, default=4, help="reads re-decoded by the oracle after timing")
    return ap.parse_args()


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
    while i < n:
        j = i
        frames = 0
        while j < n and (frames == 0 or frames + nb[j] * 43 < 16_000_000):
            frames += nb[j] * 43
            j += 1
        p, o = synth.make_reads(nb[i:j], seed=a.seed * 7919 + rank * 104729 + piece, device=device, dtype=dt)
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
    """Decode a bounded sample with the oracle port on `threads` host threads -> (bases/s, desc)."""
    from oracle import oracle

    n = min(n_reads, len(fo_np) - 1)
    sub = post_np[fo_np[0]:fo_np[n]]
    fo = fo_np[:n + 1] - fo_np[0]
    L = 0 if a.no_lm else a.context_len
    t0 = time.perf_counter()
    seqs, _, _ = oracle.beam_search_batch(sub, fo, a.beam_width, table_np, L, 0.5, 0.5, threads=threads)
    dt = time.perf_counter() - t0
    bases = int(sum(len(s) for s in seqs))
    cpu_sample.last_sec = dt
    return bases / dt, f"{n} reads / {int(fo[-1])} frames / {bases} bases in {dt:.2f}s on {threads} threads", seqs


def run_reference(a):
    """CPU arm: the oracle port (kind 'port': the reference itself is pure Python and does not
    exist on the GPU box) on all host threads, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from radian_b200 import synth

    threads = os.cpu_count() or 1
    L = 0 if a.no_lm else a.context_len
    table = None if a.no_lm else synth.make_table(L, 5)
    n = max(threads, a.cpu_reads)
    nb = synth.read_lengths(n, a.seed * 1000, fixed=a.fixed_len)
    post, off = synth.make_reads(nb, seed=a.seed * 7919, device="cpu")
    post_np = post.numpy()
    if a.f64:
        post_np = post_np.astype(np.float64)
    fo_np = off.numpy()
    vals, secs = [], []
    desc = ""
    for s in range(a.warmup + a.steps):
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


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return

    import torch
    import torch.distributed as dist

    from radian_b200 import decode, synth

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

    L = 0 if a.no_lm else a.context_len
    table_np = None if a.no_lm else synth.make_table(L, 5)
    table = decode.RnaTable(table_np, local) if table_np is not None else None

    post, fo, nb = make_batch(a, rank, device)
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
    if os.environ.get("RADIAN_STAGE_STATS"):  # library built with -DRADIAN_STAGE_STATS
        c3 = res.counters[:, 3]
        print(json.dumps({"stage2_frames": int((c3 >> 32).sum().item()), "slow_frames": int((c3 & 0xffffffff).sum().item()),
                          "frames": int(post.shape[0])}), file=sys.stderr)
    bases = int(res.lengths.sum().item())
    out = decode.decode_batch_device(post, fo, a.beam_width, table, 0.5, 0.5, max_frames=max_frames, order=order,
                                     seq_offsets=seq_offsets)
    for _ in range(a.warmup):
        step(out=out)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for s in range(a.steps):
        step(out=out)
        ev[s + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [ev[s].elapsed_time(ev[s + 1]) for s in range(a.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    assert int(out.lengths.sum().item()) == bases

    # ---- end to end through the C ABI with host buffers (H2D + kernel + D2H in the timed region)
    e2e = None
    if not a.no_e2e:
        # page-locked host copy of the first reads of the batch; with many ranks on one host the
        # calls are halved (8 x 22 GB of locked memory is not a given), as they are if locking fails
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
        lib = decode.lib
        from radian_b200 import _native

        so = np.zeros(ne + 1, dtype=np.int64)
        so[1:] = np.cumsum((fo_np[1:] - fo_np[:-1]) // 4 + 64)
        seq = np.zeros(int(so[-1]), dtype=np.uint8)
        ln = np.zeros(ne, dtype=np.int64)
        sc = np.zeros((ne, 2))
        st = np.zeros(ne, dtype=np.int32)

        def e2e_call():
            rc = lib.radian_decode_batch_host(
                _native.np_ptr(post_np), int(a.f64), _native.np_ptr(fo_np), ne, a.beam_width,
                table._h if table else None, L, 0.5, 0.5, _native.np_ptr(seq), _native.np_ptr(so), _native.np_ptr(ln),
                _native.np_ptr(sc), _native.np_ptr(st), None, local)
            _native.check(rc)

        e2e_call()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            e2e_call()
        dt = (time.perf_counter() - t0) / reps
        e2e_bases = int(ln.sum())
        e2e = {"value": e2e_bases / dt, "bytes_in": int(post_np.nbytes + fo_np.nbytes + so.nbytes),
               "bytes_out": int(seq.nbytes + ln.nbytes + sc.nbytes + st.nbytes), "reads": ne, "sec": dt}

    # ---- parity spot check against the oracle (outside every timed region)
    check = None
    cpu = None
    if rank == 0 and (a.check_reads > 0 or not a.no_cpu):
        from oracle import oracle

        nc = max(a.check_reads, 0 if a.no_cpu else a.cpu_reads)
        nc = min(nc, a.reads)
        # the first reads of the batch: an unbiased sample of the length distribution
        idx = np.arange(nc)
        sub = [post[int(fo[i]):int(fo[i + 1])].cpu().numpy() for i in idx]
        fo_s = np.zeros(nc + 1, dtype=np.int64)
        fo_s[1:] = np.cumsum([m.shape[0] for m in sub])
        sub_np = np.concatenate(sub)
        threads = os.cpu_count() or 1
        cpu_v, cpu_desc, seqs = cpu_sample(a, sub_np, fo_s, table_np, threads, nc)
        mine = out.seq.cpu().numpy()
        so_np = seq_offsets.cpu().numpy()
        ln_np = out.lengths.cpu().numpy()
        bad = 0
        for k, i in enumerate(idx):
            got = mine[so_np[i]:so_np[i] + ln_np[i]]
            bad += not np.array_equal(got, seqs[k])
        check = {"reads": int(nc), "mismatches": int(bad)}
        cpu = {"value": cpu_v, "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu_desc}

    # ---- aggregate over ranks: max time, sum of bases
    tt = torch.tensor([total_ms], dtype=torch.float64, device=device)
    bb = torch.tensor([float(bases), float(frames), float(n_lookup), float(n_combine),
                       e2e["value"] if e2e else 0.0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(bb, op=dist.ReduceOp.SUM)
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650"
        ms_per_step = float(tt.item()) / a.steps
        tot_bases, tot_frames, tot_lookup, tot_comb = (float(x) for x in bb[:4].tolist())
        value = tot_bases / (ms_per_step * 1e-3)
        # roofline of the dominant (only) kernel, per launch on rank 0:
        # B_alg = 20 B x frames + 16 B x lm[context] reads the reference performs
        kernel_ms = float(np.mean(step_ms))
        b_alg = (8 * 5 if a.f64 else 20) * frames + 16 * n_lookup
        b_min = (8 * 5 if a.f64 else 20) * frames + 16 * n_combine
        achieved = b_alg / (kernel_ms * 1e-3) / 1e9
        # DRAM bytes of one launch from the committed ncu capture of this very workload
        # (profiles/r1_traffic.json); null for any other configuration
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                tr = json.load(f)
            c = tr["config"]
            if (c["reads_per_gpu_per_step"] == a.reads and c["frames_per_gpu_per_step"] == frames
                    and c["beam_width"] == a.beam_width and c["context_len"] == a.context_len and c["seed"] == a.seed
                    and not a.no_lm and not a.f64 and a.fixed_len is None):
                traffic, traffic_src = tr["traffic_bytes_per_launch"], tr["source"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "reads_per_gpu_per_step": a.reads,
                       "frames_per_gpu_per_step": frames, "bases_per_gpu_per_step": bases,
                       "posterior_dtype": "f64" if a.f64 else "f32", "table": "synthetic dense 4^L x 4 f64",
                       "read_lengths": f"fixed {a.fixed_len}" if a.fixed_len else "LogNormal(1300,0.6) in [200,10000] bases",
                       "l2": f"inputs {post.element_size() * post.numel() / 1e9:.1f} GB per step, far larger than L2",
                       "parallelism": f"reads sharded over {world} GPU(s), table replicated, no collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": "decode_kernel",
                         "algorithmic_bytes_per_launch": b_alg, "bytes_min_per_launch": b_min,
                         "frames_per_s": frames / (kernel_ms * 1e-3), "kernel_ms": kernel_ms,
                         "n_lookup_per_frame": n_lookup / max(frames, 1)},
            "cpu_baseline": cpu,
            "e2e": ({"value": float(bb[4].item()), "unit": UNIT, "h2d_bytes_per_step": e2e["bytes_in"],
                     "d2h_bytes_per_step": e2e["bytes_out"], "reads_per_call": e2e["reads"],
                     "sec_per_call": e2e["sec"]} if e2e else None),
            "gpu_launches": a.steps,
            "clocks": clocks,
            "parity_check": check,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
