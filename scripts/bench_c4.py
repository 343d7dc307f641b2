#!/usr/bin/env python
"""BASELINE.json configs[3] on one GPU (chunk-len 1024, step-size 128, beam width 16), kernel times
with CUDA events on resident window matrices as the signal model would leave them:
  (ii) assemble_matrices merge -> RNA-LM global decode of the float64 matrix (basecall.py:99-109)
  (i)  the reference's chunk mode: every window decoded with the model off, fragments stitched
       (basecall.py:110-123)
Each checked against the oracle on the first reads.  Usage: python scripts/bench_c4.py [n_reads]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402  (checker)
from radian_b200 import decode, matrix_assembly, sequence_assembly, synth  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
W, S, BW, L = 1024, 128, 16, 12
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
nb = synth.read_lengths(n_reads, 9)
post, fo = synth.make_reads(nb, seed=4, device=dev)
T = (fo[1:] - fo[:-1]).cpu().numpy()
fo_h = fo.cpu().numpy()
# window layout of preprocess.get_windows + the trim of basecall.py:96
cro, rcr, src = [0], [0], []
for r, t in enumerate(T):
    start = 0
    while start + W <= t:
        src.append((fo_h[r] + start, W))
        cro.append(cro[-1] + W)
        start += S
    src.append((fo_h[r] + start, int(t - start)))
    cro.append(cro[-1] + int(t - start))
    rcr.append(len(cro) - 1)
cro = np.asarray(cro, np.int64)
rcr = np.asarray(rcr, np.int64)
n_chunks = len(src)
idx = torch.empty(int(cro[-1]), dtype=torch.int64, device=dev)
starts = torch.tensor([s for s, _ in src], dtype=torch.int64, device=dev)
lens = torch.tensor([l for _, l in src], dtype=torch.int64, device=dev)
d_cro = torch.from_numpy(cro).to(dev)
rep = torch.repeat_interleave(torch.arange(n_chunks, device=dev), lens)
idx = starts[rep] + (torch.arange(int(cro[-1]), device=dev) - d_cro[:-1][rep])
chunks = post[idx].contiguous()
del idx, rep


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"config": {"reads": n_reads, "frames": int(T.sum()), "chunks": n_chunks, "chunk_rows": int(cro[-1]),
                  "chunk_len": W, "step_size": S, "beam_width": BW}}
# ---- (ii) assemble + global decode with the RNA model
plan = matrix_assembly.AssemblePlan(cro, rcr, S, dev)
mat, oro = matrix_assembly.assemble_batch_device(chunks, plan=plan)
ms_asm = timed(lambda: matrix_assembly.assemble_batch_device(chunks, plan=plan, out=mat))
tab_np = synth.make_table(L, 5)
table = decode.RnaTable(tab_np, 0)
Tt = fo[1:] - fo[:-1]
order = torch.argsort(Tt, descending=True).to(torch.int32)
res = decode.decode_batch_device(mat, oro, BW, table, 0.5, 0.5, max_frames=int(Tt.max()), order=order)
ms_dec = timed(lambda: decode.decode_batch_device(mat, oro, BW, table, 0.5, 0.5, max_frames=int(Tt.max()),
                                                  order=order, out=res))
assert int(res.status.abs().sum()) == 0
bases = int(res.lengths.sum())
got = res.strings()
for r in range(2):
    c = chunks[int(cro[rcr[r]]):int(cro[rcr[r + 1]])].cpu().numpy()
    mats = [c[cro[k] - cro[rcr[r]]:cro[k + 1] - cro[rcr[r]]] for k in range(int(rcr[r]), int(rcr[r + 1]))]
    want = "".join("ACGT"[s] for s in oracle.beam_search(oracle.assemble(mats, S), BW, tab_np, L, 0.5, 0.5)[0])
    assert got[r] == want, f"global mode differs from the oracle on read {r}"
out["global_assemble_then_decode"] = {
    "assemble_ms": ms_asm, "decode_ms": ms_dec, "bases": bases, "bases_per_s": bases / ((ms_asm + ms_dec) * 1e-3),
    "decode_frames_per_s": int(T.sum()) / (ms_dec * 1e-3), "posterior_dtype": "f64 (assembled)", "context_len": L,
    "identical_to_oracle_reads": 2}
del mat, res
# ---- (i) chunk mode: every window decoded on its own, model off, then stitched
clen = torch.from_numpy(np.diff(cro)).to(dev)
corder = torch.argsort(clen, descending=True).to(torch.int32)
so = torch.zeros(n_chunks + 1, dtype=torch.int64, device=dev)
so[1:] = torch.cumsum(clen // 2 + 8, 0)
cres = decode.decode_batch_device(chunks, d_cro, BW, None, max_frames=W, order=corder, seq_offsets=so)
ms_cdec = timed(lambda: decode.decode_batch_device(chunks, d_cro, BW, None, max_frames=W, order=corder,
                                                   seq_offsets=so, out=cres))
assert int(cres.status.abs().sum()) == 0
seq = cres.seq.cpu().numpy()
so_h = so.cpu().numpy()
ln = cres.lengths.cpu().numpy()
t0 = time.perf_counter()
foff = np.zeros(n_chunks + 1, dtype=np.int64)
foff[1:] = np.cumsum(ln)
gidx = np.repeat(so_h[:-1] - foff[:-1], ln) + np.arange(int(foff[-1]), dtype=np.int64)
cseq, coff = sequence_assembly.stitch_flat(seq[gidx], foff, rcr)
dt_st = time.perf_counter() - t0
cons = ["".join("ACGT"[s] for s in cseq[coff[r]:coff[r + 1]]) for r in range(n_reads)]
frag_lists = [[seq[so_h[k]:so_h[k] + ln[k]] for k in range(int(rcr[r]), int(rcr[r + 1]))] for r in range(2)]
for r in range(2):
    frags = ["".join("ACGT"[s] for s in f) for f in frag_lists[r]]
    c = chunks[int(cro[rcr[r]]):int(cro[rcr[r + 1]])].cpu().numpy()
    mats = [c[cro[k] - cro[rcr[r]]:cro[k + 1] - cro[rcr[r]]] for k in range(int(rcr[r]), int(rcr[r + 1]))]
    ofr = ["".join("ACGT"[s] for s in oracle.beam_search(m, BW)[0]) for m in mats]
    assert frags == ofr and cons[r] == oracle.stitch(ofr)[0], f"chunk mode differs from the oracle on read {r}"
# the same with everything resident: decoder output -> stitch, no host in between
d_rcr = torch.from_numpy(rcr).to(dev)
fstart = so[:-1].contiguous()
dseq, doff, dlen, dst = sequence_assembly.stitch_device(cres.seq, fstart, cres.lengths, d_rcr)
ms_st = timed(lambda: sequence_assembly.stitch_device(cres.seq, fstart, cres.lengths, d_rcr))
assert int(dst.abs().sum()) == 0
hs, ho, hl = dseq.cpu().numpy(), doff.cpu().numpy(), dlen.cpu().numpy()
assert ["".join("ACGT"[s] for s in hs[ho[r]:ho[r] + hl[r]]) for r in range(n_reads)] == cons
cbases = sum(len(c) for c in cons)
out["chunk_decode_then_stitch"] = {
    "decode_ms": ms_cdec, "decode_frames_per_s": int(cro[-1]) / (ms_cdec * 1e-3), "stitch_host_call_s": dt_st,
    "stitch_resident_ms": ms_st, "bases_per_s_resident": cbases / ((ms_cdec + ms_st) * 1e-3),
    "bases": cbases, "bases_per_s_decode_only": cbases / (ms_cdec * 1e-3),
    "bases_per_s_with_stitch_call": cbases / (ms_cdec * 1e-3 + dt_st), "identical_to_oracle_reads": 2,
    "note": "8x the frames of global mode (every frame is decoded in 8 windows); the stitch call is "
            "stitch_flat on host arrays (compaction, H2D, four kernels, D2H)"}
print(json.dumps(out, indent=1))
