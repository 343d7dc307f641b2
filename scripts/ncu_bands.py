"""Instruction volume of an ncu report by execution-frequency band and opcode, per warp-frame:
python scripts/ncu_bands.py REPORT.ncu-rep WARP_FRAMES [--stream]   (--stream lists the hot path)"""
import collections
import csv
import re
import subprocess
import sys

rep, wf = sys.argv[1], float(sys.argv[2])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None
ln = None
seen = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        continue
    if r[0] != "":
        try:
            ln = int(r[0])
        except ValueError:
            pass
        continue
    if len(r) < 8 or not r[2].startswith("0x"):
        continue
    try:
        n = int(r[7])
    except ValueError:
        continue
    a = int(r[2], 16)
    if a not in seen:
        seen[a] = (n / wf, cur, ln, r[3].strip())
bands = [(0.85, ">=0.85"), (0.3, "0.3-0.85"), (0.1, "0.1-0.3"), (0.03, "0.03-0.1"), (-1, "<0.03")]
vol = collections.Counter()
cnt = collections.Counter()
ops = collections.Counter()
for a, (f, c, l, s) in seen.items():
    b = next(name for lo, name in bands if f >= lo)
    vol[b] += f
    cnt[b] += 1
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", s)
    ops[m.group(2).split(".")[0] if m else s[:8]] += f
print("total %.1f instr per warp-frame" % sum(vol.values()))
for _, name in bands:
    print(f"  {name:9s} {vol[name]:6.1f} from {cnt[name]} static instructions")
print("  " + " ".join(f"{o}:{n:.1f}" for o, n in ops.most_common(22)))
if "--stream" in sys.argv:
    lo = float(sys.argv[sys.argv.index("--stream") + 1]) if len(sys.argv) > sys.argv.index("--stream") + 1 else 0.85
    for a in sorted(seen):
        f, c, l, s = seen[a]
        if f >= lo:
            print(f"{a & 0xfffff:05x} {c[:14]}:{l:<4d} {f:4.2f} {s[:72]}")
