#!/usr/bin/env python3
"""Turn the raw ncu exports a GPU-box visit left in gpurun_out/ into the small tracked summaries under profiles/.
usage: summarize_ncu.py <tag>     (reads gpurun_out/prof_add_raw.csv, prof_add_source.csv, launches.csv)"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G, P = ROOT / "gpurun_out", ROOT / "profiles"
tag = sys.argv[1]
KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__occupancy_limit", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "sm__inst_executed.avg.per_cycle", "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "smsp__issue_active.avg.pct",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum")
rows = list(csv.reader(open(G / "prof_add_raw.csv")))
hdr, units, vals = rows[0], rows[1], rows[2]
kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "add_kernel"
out = [f"# ncu --set full --clock-control none, {kname} (tools/gpu_prof.sh, tools/prof_add.py)"]
m = {}
for h, u, v in zip(hdr, units, vals):
    m[h] = v
    if h.startswith(KEEP) or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")) or \
       (h.startswith("sm__inst_executed_pipe_") and h.endswith(".avg.pct_of_peak_sustained_active")):
        out.append(f"{h:90s} {u:16s} {v}")
(P / f"{tag}_add_kernel_ncu_summary.txt").write_text("\n".join(out) + "\n")
stalls = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_stalls.py"), str(G / "prof_add_source.csv"), "512"],
                        capture_output=True, text=True).stdout
(P / f"{tag}_add_kernel_stalls.txt").write_text(stalls)
if (G / "launches.csv").exists():
    (P / f"{tag}_launches_bench.csv").write_text((G / "launches.csv").read_text())
# per-key DRAM traffic of the kernel, for bench.py's roofline.traffic
log = (G / "prof_add.log").read_text() if (G / "prof_add.log").exists() else ""
gb = float(m["dram__bytes_read.sum"]) + float(m["dram__bytes_write.sum"])
unit = units[hdr.index("dram__bytes_read.sum")]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
info = {"kernel": kname, "dram_bytes_per_launch": gb * scale, "keys_per_launch": None, "source": f"profiles/{tag}_add_kernel_ncu_summary.txt"}
import re
mm = re.search(r"2\^(\d+)", " ".join(sys.argv[2:]))
if mm:
    info["keys_per_launch"] = 1 << int(mm.group(1))
    info["dram_bytes_per_key"] = gb * scale / info["keys_per_launch"]
(P / "add_kernel_traffic.json").write_text(json.dumps(info, indent=1) + "\n")
print("\n".join(out[:12]))
print(stalls.splitlines()[-1])
print(info)
