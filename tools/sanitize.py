#!/usr/bin/env python3
"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck): the fused add kernels with
the filter in shared memory and in HBM (probe pipe + candidate queue), the mul kernel, the primitives."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ecloop_b200 as E  # noqa: E402

rng = np.random.default_rng(1)


def bloom(words, fill):
    bits = np.zeros(words, dtype=np.uint64)
    for b in range(64):
        bits |= (rng.random(words) < fill).astype(np.uint64) << np.uint64(b)
    return bits


with E.Device(0) as dev:
    dev.set_tuning(1, 0)
    for name, bits in (("smem filter", bloom(509, 0.8)), ("hbm filter", bloom((1 << 14) + 3, 0.6))):
        dev.set_filter(bits)
        for flags in (E.A33, E.A65, E.A33 | E.A65 | E.ENDO):
            hits = dev.batch_add(2**70 + 77, 2048 * 3, flags)
            print(f"{name}: flags {flags}: {len(hits)} bloom-positive", flush=True)
    hits = dev.mul_batch([int(x) for x in rng.integers(1, 2**62, size=3000)], E.A33 | E.A65)
    print("mul:", len(hits), "bloom-positive")
    print("fp:", dev.fp(E.OP_INV, [5, 7, 2**200 + 1])[:1])
    print("bloom prim:", dev.bloom_has([(1, 2, 3, 4, 5)]))
