#!/usr/bin/env python3
"""Device time of the mul path (ecl_mul_submit + ecl_collect) for 2^LOG2 random keys, -a cu. Usage: prof_mul.py [log2]"""
import random
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ecloop_b200 as E  # noqa: E402
import ecloop_b200.host as H  # noqa: E402

log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 20
r = random.Random(5)
keys = E._fe_array([r.getrandbits(256) for _ in range(1 << log2)])
flt = H.load_filter(ROOT / "tests" / "golden" / "btc-puzzles-hash")
t_open = time.perf_counter()
with E.Device(0) as dev:
    print(f"ecl_open (context + window table build): {time.perf_counter() - t_open:.3f} s")
    t_open = time.perf_counter()
    E.Device(0).close()
    print(f"second ecl_open in the same process (window table build + allocations only): {time.perf_counter() - t_open:.3f} s")
    dev.set_filter(flt.bits)
    for i in range(3):
        t0 = time.perf_counter()
        dev.mul_submit(keys, E.A33 | E.A65)
        hits, _ = dev.collect()
        wall = time.perf_counter() - t0
        total, hot, launches = dev.last_elapsed_ms()
        print(f"run {i}: {1 << log2} keys, wall {wall * 1e3:.2f} ms, device total {total:.3f} ms (incl. H2D), mul_kernel {hot:.3f} ms, "
              f"{(1 << log2) / hot / 1e3:.1f} Mkeys/s (kernel)")
