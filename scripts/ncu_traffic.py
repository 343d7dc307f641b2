#!/usr/bin/env python
"""profiles/rN_traffic.json from an `ncu --set full` capture of the bench's decode launch (build
container): DRAM bytes, duration, issue-slot use, stall mix, L2/L1 hit rates.
usage: python scripts/ncu_traffic.py REPORT.ncu-rep "source command" reads frames bw L seed > profiles/r2_traffic.json"""
import csv
import json
import subprocess
import sys

rep, source, reads, frames, bw, L, seed = sys.argv[1], sys.argv[2], *(int(x) for x in sys.argv[3:8])
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, val = rows[0], rows[2] if len(rows) > 2 else rows[1]
m = {}
for h, v in zip(hdr, val):
    try:
        m[h] = float(v.replace(",", ""))
    except ValueError:
        m[h] = v
unit = dict(zip(hdr, rows[1]))


def scaled(name):
    v, u = m[name], unit.get(name, "")
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1)


rd, wr = scaled("dram__bytes_read.sum"), scaled("dram__bytes_write.sum")
stalls = {k.split("issue_stalled_")[1].split("_per_issue")[0]: round(v, 3) for k, v in m.items()
          if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and isinstance(v, float) and v >= 0.05
          and "selected_per" not in k.replace("not_selected", "")}
out = {
    "source": source,
    "kernel": m.get("Kernel Name", "decode_kernel"),
    "config": {"reads_per_gpu_per_step": reads, "frames_per_gpu_per_step": frames, "beam_width": bw, "context_len": L, "seed": seed},
    "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes_per_launch": rd + wr, "bytes_per_frame": (rd + wr) / frames,
    "gpu_time_ms_under_ncu": scaled("gpu__time_duration.sum"),
    "issue_active_pct": m["smsp__issue_active.avg.pct_of_peak_sustained_active"],
    "inst_executed_warp": m["smsp__inst_executed.sum"], "inst_per_warp_frame": m["smsp__inst_executed.sum"] / (frames / 2),
    "l2_sector_hit_pct": m["lts__t_sector_hit_rate.pct"], "l1tex_sector_hit_pct": m["l1tex__t_sector_hit_rate.pct"],
    "registers_per_thread": m["launch__registers_per_thread"], "grid": m["launch__grid_size"],
    "warps_active_pct": m["sm__warps_active.avg.pct_of_peak_sustained_active"],
    "stalls_per_issue": dict(sorted(stalls.items(), key=lambda kv: -kv[1])),
}
print(json.dumps(out, indent=1))
