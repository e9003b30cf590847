#!/usr/bin/env python3
"""BASELINE configs[3] shape on one GPU: `add -endo` against a large .blf-style bloom filter resident in HBM.
Usage: prof_bloom.py [log2_filter_bytes=32] [log2_keys=30] [flags=5]   (filter bits i.i.d. with fill 0.37)"""
import sys
import time
from pathlib import Path


ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ecloop_b200 as E  # noqa: E402

lb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lk = int(sys.argv[2]) if len(sys.argv) > 2 else 30
flags = int(sys.argv[3]) if len(sys.argv) > 3 else (E.A33 | E.ENDO)
words = (1 << lb) // 8 - 5  # not a power of two, like a real blf-gen size
with E.Device(0) as dev:
    t0 = time.perf_counter()
    dev.filter_generate(words, 0.37, 4)  # bits i.i.d. with p = 95/256, generated on the device
    print(f"filter: {words * 8 / 2**30:.2f} GiB, fill {dev.filter_fill():.4f}, generated on the device in {time.perf_counter() - t0:.2f} s", flush=True)
    per_key = 6 if flags & E.ENDO else 1
    for i in range(3):
        hits = dev.batch_add(2**70 + (i << lk), 1 << lk, flags)
        total, hot, launches = dev.last_elapsed_ms()
        print(f"run {i}: flags {flags}, {len(hits)} bloom-positive of {per_key * (1 << lk)} hashes (expected ~{per_key * (1 << lk) * 0.371**20:.1f}), "
              f"add_kernel {hot:.2f} ms, {(1 << lk) / hot / 1e3:.1f} M base keys/s = {per_key * (1 << lk) / hot / 1e3:.1f} Mkeys/s counted", flush=True)
