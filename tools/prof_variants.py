#!/usr/bin/env python3
"""Kernel rate of every add-kernel instance, filter in shared memory and in HBM, for the library selected by
ECLOOP_B200_LIB (tools/build_variants.py). Usage: prof_variants.py [log2_filter_bytes_hbm=32] [flags ...]
Prints one line per (flags, placement): M base keys/s of the add kernel (CUDA events around its launches)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ecloop_b200 as E  # noqa: E402
import ecloop_b200.host as H  # noqa: E402

lb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
flag_list = [int(x) for x in sys.argv[2:]] or [1, 2, 3, 5, 7]
LOG2 = {1: 32, 2: 31, 3: 31, 5: 30, 6: 29, 7: 29}
puzzles = H.load_filter(ROOT / "tests" / "golden" / "btc-puzzles-hash")
with E.Device(0) as dev:
    dev.set_stride(1)
    for place in ("smem", "hbm"):
        if place == "smem":
            dev.set_filter(puzzles.bits)
        else:
            dev.filter_generate((1 << lb) // 8 - 5, 0.37, 4)
        for flags in flag_list:
            lk = LOG2[flags]
            best = 0.0
            for i in range(2):
                dev.batch_add(2**70 + (i << lk), 1 << lk, flags)
                total, hot, launches = dev.last_elapsed_ms()
                best = max(best, (1 << lk) / hot / 1e3)
            per = 6 if flags & E.ENDO else 1
            print(f"flags {flags} {place}: {best:8.1f} M base keys/s ({best * per:8.1f} counted), {launches} launches, 2^{lk} keys", flush=True)
