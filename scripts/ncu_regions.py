"""Summarise an ncu report by source line / region of decode.cu (run in the build container)."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
frames_per_warp = float(sys.argv[2]) if len(sys.argv) > 2 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]
iInst = hdr.index("Instructions Executed")
iSamp = hdr.index("# Samples")
by = collections.Counter()
sm = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= iInst or r[0] == "":
        continue
    try:
        n = int(r[iInst]); s = int(r[iSamp]); ln = int(r[0])
    except ValueError:
        continue
    by[ln] += n
    sm[ln] += s
tot = sum(by.values())
ts = sum(sm.values())
src = open("radian_b200/csrc/decode.cu").read().split("\n")
print("total warp instructions", tot)
marks = []
for i, l in enumerate(src, 1):
    if l.strip().startswith("// ----") or l.strip().startswith("// COPY") or l.strip().startswith("// EXTEND") \
            or l.strip().startswith("// MERGE") or l.strip().startswith("// SELECT") or l.strip().startswith("// RESCALE") \
            or "exact path:" in l or "rank = number of candidates" in l or "const bool survive" in l or "if (n_new > 0)" in l:
        marks.append((i, l.strip()[:60]))
marks.append((len(src) + 1, "end"))
prev = (1, "prologue")
for mk in marks:
    a, b = prev[0], mk[0] - 1
    n = sum(v for k, v in by.items() if a <= k <= b)
    s = sum(v for k, v in sm.items() if a <= k <= b)
    if n:
        extra = f" {n / frames_per_warp:7.1f}/warp-frame" if frames_per_warp else ""
        print(f"{a:4d}-{b:4d} {100 * n / tot:5.1f}% inst {100 * s / max(ts, 1):5.1f}% samp{extra}  {prev[1]}")
    prev = mk
print("--- top lines")
for ln, n in by.most_common(25):
    print(f"{ln:5d} {100 * n / tot:5.1f}% inst {100 * sm[ln] / max(ts, 1):5.1f}% samp  {src[ln - 1].strip()[:90]}")
