"""Summarise an ncu report by source file and by region of the kernel source (run in the build
container):  python scripts/ncu_regions.py REPORT.ncu-rep [warp_frames] [source.cu]

Regions are delimited by the marker comments of the kernel source ("// ----", "// COPY", ...).
`warp_frames` = frames x reads / reads-per-warp of the profiled launch turns instruction counts
into instructions per warp-frame."""
import collections
import csv
import os
import subprocess
import sys

rep = sys.argv[1]
warp_frames = float(sys.argv[2]) if len(sys.argv) > 2 else None
main_src = sys.argv[3] if len(sys.argv) > 3 else "radian_b200/csrc/decode.cu"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
files = collections.OrderedDict()  # path -> {line: [inst, samples, stalls Counter]}
cur = None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = files.setdefault(r[1], {})
        continue
    if r[0] == "Line No":
        hdr = r
        iInst = hdr.index("Instructions Executed")
        iSamp = hdr.index("# Samples")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if cur is None or hdr is None or len(r) <= iInst:
        continue
    try:
        ln = int(r[0]); n = int(r[iInst]); s = int(r[iSamp])
    except ValueError:
        continue
    e = cur.setdefault(ln, [0, 0, collections.Counter()])
    e[0] += n
    e[1] += s
    for i, h in stall_cols:
        try:
            e[2][h] += int(r[i])
        except (ValueError, IndexError):
            pass

tot = sum(e[0] for f in files.values() for e in f.values())
ts = sum(e[1] for f in files.values() for e in f.values())
print(f"total warp instructions {tot}, samples {ts}" + (f", {tot / warp_frames:.1f} per warp-frame" if warp_frames else ""))
print("--- by file")
for path, f in files.items():
    n = sum(e[0] for e in f.values())
    s = sum(e[1] for e in f.values())
    if n:
        extra = f" {n / warp_frames:7.1f}/warp-frame" if warp_frames else ""
        print(f"{100 * n / tot:5.1f}% inst {100 * s / max(ts, 1):5.1f}% samp{extra}  {os.path.basename(path)}")

key = [p for p in files if p.endswith(os.path.basename(main_src))]
if key:
    by = files[key[0]]
    src = open(main_src).read().split("\n")
    marks = []
    for i, l in enumerate(src, 1):
        t = l.strip()
        if t.startswith("// ----") or t.startswith("// COPY") or t.startswith("// EXTEND") or t.startswith("// MERGE") \
                or t.startswith("// SELECT") or t.startswith("// RESCALE") or "exact path:" in t \
                or "rank = number of candidates" in t or "const bool survive" in t or "if (n_new > 0)" in t \
                or t.startswith("// ---- exact ranks") or t.startswith("// ---- candidate list"):
            marks.append((i, t[:60]))
    marks.append((len(src) + 1, "end"))
    print(f"--- regions of {main_src}")
    prev = (1, "prologue")
    for mk in marks:
        a, b = prev[0], mk[0] - 1
        n = sum(v[0] for k, v in by.items() if a <= k <= b)
        s = sum(v[1] for k, v in by.items() if a <= k <= b)
        st = collections.Counter()
        for k, v in by.items():
            if a <= k <= b:
                st.update(v[2])
        if n:
            extra = f" {n / warp_frames:7.1f}/warp-frame" if warp_frames else ""
            top = " ".join(f"{h[6:]}:{100 * c / max(s, 1):.0f}%" for h, c in st.most_common(3))
            print(f"{a:4d}-{b:4d} {100 * n / tot:5.1f}% inst {100 * s / max(ts, 1):5.1f}% samp{extra}  {prev[1]}  [{top}]")
        prev = mk
    print("--- top lines (all files)")
    allv = [(e[0], e[1], os.path.basename(p), ln) for p, f in files.items() for ln, e in f.items()]
    allv.sort(reverse=True)
    cache = {}
    for n, s, fn, ln in allv[:30]:
        full = [p for p in files if os.path.basename(p) == fn][0]
        if full not in cache:
            try:
                cache[full] = open(full.replace("/root/repo/", "")).read().split("\n")
            except OSError:
                cache[full] = []
        text = cache[full][ln - 1].strip()[:80] if ln - 1 < len(cache[full]) else ""
        print(f"{fn}:{ln:<5d} {100 * n / tot:5.1f}% inst {100 * s / max(ts, 1):5.1f}% samp  {text}")
