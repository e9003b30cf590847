#!/usr/bin/env python3
"""SASS census of the pipelined addr33 kernel's hot loop (one pass-2 step = 2 keys) from the built object:
ALU-pipe instructions, IMAD.WIDE (occupies an issue slot on both integer pipes), FMA-only instructions per key.
Merges the result into profiles/add_kernel_traffic.json (bench.py's roofline.alu_slots). Usage: loop_census.py"""
import collections
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
obj = ROOT / "build" / "obj" / "add_inst_1.o"
sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
sass = sass[sass.index("add_kernel_sp"):]
ins = [(int(m.group(1), 16), re.sub(r"^@!?U?P\d\s+", "", m.group(2).strip()))
       for m in (re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l) for l in sass.splitlines()) if m]
loops = []
for a, t in ins:
    m = re.search(r"BRA(?:\.\S+)*\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
# the pass-2 loop: the largest loop that is nested inside another one (the group loop) and holds two hashes (> 5000 instructions)
cands = sorted(((hi - lo) // 16, lo, hi) for lo, hi in loops if 5000 < (hi - lo) // 16 < 8000)
n, lo, hi = cands[0]
ALU = {"SHF", "LOP3", "IADD3", "LEA", "SEL", "ISETP", "VIADD", "PRMT", "IABS", "IMNMX", "PLOP3", "SGXT", "BMSK", "POPC", "FLO"}
c = collections.Counter()
for a, t in ins:
    if lo <= a <= hi:
        op = t.split()[0]
        base = op.split(".")[0]
        if base == "IMAD":
            c["imad_wide" if "WIDE" in op else "fma_only"] += 1
        elif base in ALU:
            c["alu"] += 1
        else:
            c["other"] += 1
out = {"loop_instructions_per_key": (n + 1) / 2, "alu_pipe_per_key": c["alu"] / 2, "imad_wide_per_key": c["imad_wide"] / 2,
       "fma_only_per_key": c["fma_only"] / 2, "other_per_key": c["other"] / 2,
       "alu_issue_slots_per_key": (c["alu"] + c["imad_wide"]) / 2,
       "note": "SASS census of the pass-2 loop of add_kernel_sp<A33> (tools/loop_census.py); IMAD.WIDE takes an issue slot on both integer pipes"}
p = ROOT / "profiles" / "add_kernel_traffic.json"
d = json.loads(p.read_text()) if p.exists() else {}
d["loop_census"] = out
p.write_text(json.dumps(d, indent=1) + "\n")
print(json.dumps(out, indent=1))
