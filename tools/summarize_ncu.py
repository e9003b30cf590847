#!/usr/bin/env python3
"""Turn the raw ncu exports a GPU-box visit left in gpurun_out/ into the small tracked summaries under profiles/.
usage: summarize_ncu.py <gpurun_out/X_raw.csv> <profiles/out.txt> ["note"] [--traffic KEYS_PER_LAUNCH]
One block per profiled kernel launch: launch shape, time, DRAM bytes, pipe utilisation, stall reasons per issued
instruction, instructions executed. With --traffic also rewrites profiles/add_kernel_traffic.json (bench.py's
roofline.traffic) from the FIRST kernel of the file."""
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__occupancy_limit", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "sm__inst_executed.avg.per_cycle", "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "smsp__issue_active.avg.pct",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "lts__t_sector_hit_rate.pct")
src, dst = Path(sys.argv[1]), Path(sys.argv[2])
note = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else ""
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
out = [f"# ncu --set full --clock-control none; {note}".rstrip("; ")]
first = None
for vals in rows[2:]:
    m = dict(zip(hdr, vals))
    first = first or m
    out.append(f"\n## {m.get('Kernel Name', '?')}  (launch id {m.get('ID', '?')})")
    for h, u, v in zip(hdr, units, vals):
        if h.startswith(KEEP) or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")) or \
           (h.startswith("sm__inst_executed_pipe_") and h.endswith(".avg.pct_of_peak_sustained_active")):
            out.append(f"{h:90s} {u:16s} {v}")
dst.write_text("\n".join(out) + "\n")
print(dst, len(out), "lines")
if "--traffic" in sys.argv:
    keys = int(sys.argv[sys.argv.index("--traffic") + 1])
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    b = sum(float(first[k]) * scale[units[hdr.index(k)]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    inst = float(first["smsp__inst_executed.sum"]) * 32 / keys
    info = {"kernel": first.get("Kernel Name"), "dram_bytes_per_launch": b, "keys_per_launch": keys, "dram_bytes_per_key": b / keys,
            "thread_instructions_per_key": inst, "grid": int(first["launch__grid_size"]), "source": str(dst.resolve().relative_to(ROOT))}
    (ROOT / "profiles" / "add_kernel_traffic.json").write_text(json.dumps(info, indent=1) + "\n")
    print(info)
