#!/usr/bin/env python3
"""Build tuning variants of libecloop_b200.so (different -D defines) into build/variants/ so that one GPU-box
visit can time them all (tools/bench_variants.sh). Usage: build_variants.py tag:DEF=V,DEF=V ..."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ecloop_b200 import build as B  # noqa: E402

for spec in sys.argv[1:]:
    tag, _, defs = spec.partition(":")
    defines = tuple(d for d in defs.split(",") if d)
    print(tag, defines, B.build(defines=defines, variant=tag), flush=True)
