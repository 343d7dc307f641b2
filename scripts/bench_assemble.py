"""Roofline measurement of the chunk-assembly kernel (run on the GPU box).
Workload: reads of the config-4 shape (chunk-len 1024, step 128): every read's window matrices are
resident in HBM as the sig model would leave them; one launch assembles all reads.
Algorithmic bytes per output row: 20 B read (float32 row of the winning chunk) + 40 B written
(float64 row) = 60 B."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radian_b200 import matrix_assembly, synth  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
W, S = 1024, 128
dev = torch.device("cuda", 0)
nb = synth.read_lengths(n_reads, 9)
T = (nb * 43).astype(np.int64)
# window layout of preprocess.get_windows + basecall.py:96 trim
cro = [0]
rcr = [0]
for t in T:
    start = 0
    while start + W <= t:
        cro.append(cro[-1] + W)
        start += S
    cro.append(cro[-1] + (t - start))
    rcr.append(len(cro) - 1)
cro = np.asarray(cro, np.int64)
rcr = np.asarray(rcr, np.int64)
rows_in = int(cro[-1])
chunks = torch.rand((rows_in, 5), device=dev, dtype=torch.float32)
chunks /= chunks.sum(1, keepdim=True)
out, oro = matrix_assembly.assemble_batch_device(chunks, cro, rcr, S)
torch.cuda.synchronize()
rows_out = out.shape[0]
assert rows_out == int(T.sum())
# spot check against the oracle on the first read
from oracle import oracle  # noqa: E402

c = chunks[: int(cro[rcr[1]])].cpu().numpy()
mats = [c[cro[k]:cro[k + 1]] for k in range(int(rcr[1]))]
want = oracle.assemble(mats, S)
got = out[: want.shape[0]].cpu().numpy()
ok = bool(np.array_equal(got, want))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
plan = matrix_assembly.AssemblePlan(cro, rcr, S, dev)
for _ in range(3):
    matrix_assembly.assemble_batch_device(chunks, plan=plan, out=out)
torch.cuda.synchronize()
ev0.record()
for _ in range(reps):
    matrix_assembly.assemble_batch_device(chunks, plan=plan, out=out)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / reps
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
gbs = 60.0 * rows_out / (ms * 1e-3) / 1e9
print(json.dumps({"kernel": "assemble_kernel<f64>", "reads": n_reads, "rows_in": rows_in, "rows_out": rows_out,
                  "input_GB": rows_in * 20 / 1e9, "kernel_ms": ms, "algorithmic_GBps": gbs,
                  "peak_GBps": peak, "frac": gbs / peak, "bit_exact_vs_oracle_read0": ok}))
