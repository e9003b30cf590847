#!/usr/bin/env python3
"""One short add-mode run for ncu: a warm-up submit and a single measured launch of add_kernel over
2^LOG2 keys at 2^70 (addr33, puzzle list filter). Usage: python tools/prof_add.py [log2_keys] [flags]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ecloop_b200 as E  # noqa: E402
import ecloop_b200.host as H  # noqa: E402

log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 28
flags = int(sys.argv[2]) if len(sys.argv) > 2 else E.A33
flt = H.load_filter(ROOT / "tests" / "golden" / "btc-puzzles-hash")
with E.Device(0) as dev:
    dev.set_filter(flt.bits)
    dev.set_stride(1)
    for i in range(3):
        hits = dev.batch_add(2**70 + (i << log2), 1 << log2, flags)
        total, hot, launches = dev.last_elapsed_ms()
        print(f"run {i}: {len(hits)} bloom-positive, total {total:.3f} ms, add_kernel {hot:.3f} ms, "
              f"{(1 << log2) / hot / 1e3:.1f} Mkeys/s (kernel), {launches} launches")
