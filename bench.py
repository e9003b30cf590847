#!/usr/bin/env python3
"""bench.py — the reference's headline metric on B200: Mkeys/s, `add` mode, addr33, list filter.

Workload (BASELINE.json configs[1]): `ecloop add -r 400000000000000000:40000000ffffffffff` (keys 2^70 ..
2^70+2^40-1, compressed addresses), filter = the 160 puzzle hashes + 64 planted keys. One "step" is one pass of the
hot path (ecl_add_submit + ecl_collect) over one batch of 2^LOG2 consecutive keys of that range per GPU; rank g of
N sweeps its own contiguous shard of the range (no data-path collective; weak scaling: per-GPU work is fixed).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm (torchrun for N > 1)
    python bench.py --impl reference [--gpus N] ...              # the unmodified reference on the host cores

Prints ONE JSON line (rank 0). `value` is device-timed (CUDA events on the launch stream, max over ranks);
`e2e` is wall-clock through the public API with host buffers; `roofline` is the integer-ALU roofline the
metric names (keys/s x 2700 canonical ALU ops) against the LOP3 issue rate measured in the same
process, with the HBM view beside it; `cpu_baseline` times oracle/_ref on a bounded sample of the same range.

After the timed headline the same line gets `secondary`: the other three BASELINE.json configs, each with its own
device-timed value, end-to-end value, roofline and clocks (`--no-secondary` skips them):
  mul_10M_cu     configs[2]: 10 M seeded keys per GPU through ecl_mul_submit from host buffers, -a cu
  add_endo_blf   configs[3]: 6 GiB filter (Bernoulli(0.37), generated on the device, seed 4) in HBM, -endo, the range
                 of Makefile:58 cut into one job-aligned shard per GPU
  rnd_128_32_cu  configs[4]: `ecloop rnd -d 128:32 -a cu` — the C host on all GPUs of the run, consecutive 2^32 windows
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

RANGE_S = 0x400000000000000000
RANGE_E = 0x40000000FFFFFFFFFF
RANGE_KEYS = 1 << 40
ALU_OPS_PER_KEY = 2700  # SURVEY §8d canonical int32 ALU-pipe ops per addr33 key
SCRATCH_BYTES_PER_KEY = 32  # prefix products: 32 B written + 32 B read per 2 keys
METRIC = "Mkeys/s (add mode, addr33)"


def shard_of(rank: int, world: int):
    """contiguous job-aligned shard [start, start + n_keys) of the 2^40-key range for `rank` of `world`"""
    per = RANGE_KEYS // world // (1 << 21) * (1 << 21)
    return RANGE_S + rank * per, per


def planted_offsets(n_steps: int, log2_step: int, per_rank: int = 8):
    """key offsets (relative to a rank's shard start) of the planted keys: inside the swept prefix"""
    import random

    r = random.Random(71)
    span = n_steps << log2_step
    return sorted(r.randrange(span) for _ in range(per_rank))


def py_hash160(k: int):
    """(hash160 of the compressed key, hash160 of the uncompressed key) of k: textbook double-and-add over python
    ints + hashlib"""
    import hashlib

    p = 2**256 - 2**32 - 977
    gx = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
    gy = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8

    def add(a, b):
        if a is None:
            return b
        if b is None:
            return a
        if a[0] == b[0]:
            if (a[1] + b[1]) % p == 0:
                return None
            lam = 3 * a[0] * a[0] * pow(2 * a[1], -1, p) % p
        else:
            lam = (b[1] - a[1]) * pow(b[0] - a[0], -1, p) % p
        x = (lam * lam - a[0] - b[0]) % p
        return x, (lam * (a[0] - x) - a[1]) % p

    acc, base = None, (gx, gy)
    while k:
        if k & 1:
            acc = add(acc, base)
        base = add(base, base)
        k >>= 1
    ser = bytes([2 + (acc[1] & 1)]) + acc[0].to_bytes(32, "big")
    ser65 = b"\x04" + acc[0].to_bytes(32, "big") + acc[1].to_bytes(32, "big")
    return (hashlib.new("ripemd160", hashlib.sha256(ser).digest()).hexdigest(),
            hashlib.new("ripemd160", hashlib.sha256(ser65).digest()).hexdigest())


def py_hash160_33(k: int) -> str:
    return py_hash160(k)[0]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()

    def run(self):
        try:
            p = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                  "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        while not self._stop.is_set():
            line = p.stdout.readline()
            if not line:
                break
            self.rows.append([c.strip() for c in line.split(",")])
        p.terminate()

    def stop(self):
        self._stop.set()

    def summary(self):
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


def measured_traffic():
    """DRAM bytes per key of add_kernel from the committed `ncu --set full` capture (tools/summarize_ncu.py)"""
    p = ROOT / "profiles" / "add_kernel_traffic.json"
    try:
        d = json.loads(p.read_text())
        return float(d["dram_bytes_per_key"]), d.get("source", str(p))
    except Exception:
        return None, None


def kernel_slot_census():
    """the kernel's OWN ALU issue slots per key (SASS census of the hot loop + ncu's executed-instruction count), so that
    the canonical-model fraction and the pipe fraction sit side by side"""
    try:
        d = json.loads((ROOT / "profiles" / "add_kernel_traffic.json").read_text())
        c = d["loop_census"]
        return {"alu_pipe_per_key": c["alu_pipe_per_key"], "imad_wide_per_key": c["imad_wide_per_key"],
                "slots_per_key": c["alu_issue_slots_per_key"], "executed_thread_instructions_per_key_ncu": round(d.get("thread_instructions_per_key", 0), 1)}
    except Exception:
        return None


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text()), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------- the reference arm / cpu baseline


def run_reference_sample(start: int, n_keys: int, threads: int):
    """time the unmodified reference binary (oracle/_ref) on keys [start, start+n_keys): -> (Mkeys/s wall, info)"""
    import oracle as O

    exe = O.ref_binary()
    if exe is None:
        return None, "oracle/_ref not built"
    flt = ROOT / "tests" / "golden" / "btc-puzzles-hash"
    args = [str(exe), "add", "-f", str(flt), "-r", "%x:%x" % (start, start + n_keys - 1), "-t", str(threads), "-q", "-o", "/dev/null"]
    t0 = time.perf_counter()
    r = subprocess.run(args, capture_output=True, env=dict(os.environ, LC_ALL="C"))
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        return None, f"reference exited {r.returncode}"
    status = [l for l in r.stderr.decode(errors="replace").replace("\r", "\n").splitlines() if "Mkeys/s ~" in l]
    rate = n_keys / dt / 1e6
    import re

    m = re.search(r"~ ([\d.]+) Mkeys/s", status[-1]) if status else None
    if m:  # the reference's own metric (k_checked / elapsed since ITS start, main.c:137-139): no process start-up in it
        rate = float(m.group(1))
    return rate, {"binary": exe.name, "status_line": status[-1].strip() if status else "", "wall_mkeys": round(n_keys / dt / 1e6, 3)}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n = 1 << 27  # bounded sample per step of the same range (2^27 keys, job-aligned)
    vals = []
    info = {}
    for i in range(args.warmup + args.steps):
        v, info = run_reference_sample(RANGE_S + i * n, n, cores)
        if v is None:
            print(json.dumps({"impl": "reference", "unavailable": str(info)}))
            return 0
        if i >= args.warmup:
            vals.append(v)
    ms = sum(n / (v * 1e6) for v in vals) / len(vals) * 1e3
    value = n * len(vals) / sum(n / (v * 1e6) for v in vals) / 1e6
    sample = f"2^27 consecutive keys of configs[1] per step, -t {cores}, the reference's own status-line rate ({info.get('binary')}; wall clock incl. process start: {info.get('wall_mkeys')} Mkeys/s)"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "Mkeys/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 (4x64-bit limbs, CPU)", "data": "synthetic",
        "config": {"workload": "add -r 400000000000000000:40000000ffffffffff addr33 (BASELINE configs[1]), bounded sample",
                   "keys_per_step": n, "filter": "list: 160 puzzle hashes"},
        "cpu_baseline": {"value": round(value, 3), "unit": "Mkeys/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "Mkeys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0



# ---------------------------------------------------------------- secondary legs: BASELINE.json configs[2..4]

CU_OPS_PER_KEY = 321 + 3 * 1384 + 2 * 950 + 2 * 12 + 2 * 42  # field + 3 SHA-256 blocks + 2 RIPEMD-160 blocks + serialise + probes
ENDO_EXTRA_OPS = 2400  # each extra endomorphism image (hash160_33 + its share of the beta multiplications)
RANGE71_S, RANGE71_E = 0x400000000000000000, 0x7FFFFFFFFFFFFFFFFF  # Makefile:58 (range_71)


def shard71_of(rank: int, world: int):
    """contiguous job-aligned shard [start, start + n_keys) of 400000000000000000:7fffffffffffffffff (Makefile:58) for `rank`"""
    per = (RANGE71_E - RANGE71_S) // world // (1 << 21) * (1 << 21)
    return RANGE71_S + rank * per, per


def _allreduce(world, local, vals, op):
    import torch
    import torch.distributed as dist

    if world == 1:
        return list(vals)
    t = torch.tensor(list(vals), device=f"cuda:{local}", dtype=torch.float64)
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
    return [float(x) for x in t.tolist()]


def peaks_all(dev, peaks, world, local):
    """rank 0 measured the pipe peaks; the legs only need them there"""
    return peaks if peaks is not None else {"lop3_gops": 1.0}


def leg_mul(E, H, dev, rank, world, local, peaks, n_keys, with_reference):
    """configs[2]: `mul` on n_keys seeded random 256-bit keys per GPU, -a cu, keys in host memory (numpy), batches of
    2^20 with ECL_MUL_DEPTH submits in flight; hits = every (n_keys/100)-th key, alternating encodings."""
    import numpy as np
    import torch
    import torch.distributed as dist

    rng = np.random.default_rng(3 + rank)
    keys = rng.integers(0, 2**64, size=(n_keys, 4), dtype=np.uint64)
    keys[:, 3] &= np.uint64(0x7FFFFFFFFFFFFFFF)  # below n without a reduction step, like parsed hex keys
    every = max(1, n_keys // 100)
    planted = list(range(0, n_keys, every))
    want = []
    hashes = []
    for j, i in enumerate(planted):
        k = sum(int(keys[i, l]) << (64 * l) for l in range(4))
        h33, h65 = py_hash160(k)
        hashes.append(h33 if j % 2 == 0 else h65)
        want.append((i, j % 2))
    flt = H.filter_from_hashes(hashes)
    dev.set_filter(np.ascontiguousarray(flt.bits, dtype=np.uint64))
    lib, h = dev._lib, dev._h
    batch = 1 << 22
    import ctypes as C

    def run_all():
        found, hot, pend = [], 0.0, []
        for b in range(0, n_keys, batch):
            part = keys[b:b + batch]
            dev._ck(lib.ecl_mul_submit(h, C.c_void_p(part.ctypes.data), part.shape[0], E.A33 | E.A65))
            pend.append(b)
            if len(pend) == E.MUL_DEPTH:
                b0 = pend.pop(0)
                hits, _ = dev.collect()
                hot += dev.last_elapsed_ms()[1]
                found += [(b0 + k, kd) for k, _, kd, hh in hits if flt.check_exact(hh)]
        while pend:
            b0 = pend.pop(0)
            hits, _ = dev.collect()
            hot += dev.last_elapsed_ms()[1]
            found += [(b0 + k, kd) for k, _, kd, hh in hits if flt.check_exact(hh)]
        return found, hot

    run_all()  # warm-up (allocations, clocks)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    reps = 30  # the 10 M keys take ~35 ms: repeated so that the clock sampler (200 ms period) sees the leg under load
    t0 = time.perf_counter()
    hot_ms, ok = 0.0, True
    for _ in range(reps):
        found, h_ms = run_all()
        hot_ms += h_ms
        ok = ok and sorted(found) == sorted(want)
    wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.stop()
    wall_ms, hot_ms = _allreduce(world, local, [wall_ms, hot_ms], "MAX")
    ok = _allreduce(world, local, [1.0 if ok else 0.0], "MIN")[0] == 1.0
    if rank != 0:
        return None, ok
    kern = n_keys * reps / (hot_ms * 1e-3) / 1e6  # per GPU (max rank)
    e2e = n_keys * reps * world / (wall_ms * 1e-3) / 1e6
    B = max(1, -(-batch // (148 * 512)))
    field_mults = 10 * 11 + 7 + 270.0 / B  # W = 24: 11 windows, the first is a load
    ops = field_mults * 45 + 10 * 8 * 16 + (CU_OPS_PER_KEY - 321)  # canonical ALU-pipe ops: 45 per multiplication
    alu_peak = peaks["lop3_gops"] * 1e9
    leg = {
        "workload": f"mul: {n_keys} seeded 256-bit keys per GPU from host memory, -a cu (BASELINE configs[2]), batches of 2^22, {E.MUL_DEPTH} submits in flight, "
                    f"the pass over the {n_keys} keys repeated {reps}x inside the timed region",
        "metric": "Mkeys/s (mul mode, -a cu)", "unit": "Mkeys/s", "n_gpus": world,
        "value": round(kern * world, 2), "value_note": "mul_points_kernel + mul_hash_kernel, CUDA events on the launch stream, max over ranks, x n_gpus",
        "e2e": {"value": round(e2e, 2), "unit": "Mkeys/s", "h2d_bytes_per_step": 32 * batch, "d2h_bytes_per_step": 4,
                "note": "wall clock through ecl_mul_submit/ecl_collect from pageable host arrays, staging copy + H2D inside"},
        "roofline": {"bound": "int_alu", "achieved": round(kern * 1e6 * ops / 1e12, 3), "peak": round(alu_peak / 1e12, 3), "unit": "Tops/s",
                     "frac": round(kern * 1e6 * ops / alu_peak, 4),
                     "model": f"{field_mults:.0f} field multiplications per key (10 mixed additions x 11 from the W=24 window table, normalisation 7, inversion 270/{B}) x 45 canonical "
                              f"ALU ops + 3 SHA-256 + 2 RIPEMD-160 blocks = {ops:.0f} ALU-pipe ops per key; IMAD.WIDE also occupies the ALU pipe on "
                              "sm_100 (DESIGN.md), which this canonical count ignores"},
        "parity_gate": "every planted key found, nothing else" if ok else "FAILED", "clocks": sampler.summary(),
    }
    if with_reference:
        leg["reference"] = reference_mul_sample(min(n_keys, 1 << 20))
    return leg, ok


def reference_mul_sample(n):
    """the unmodified reference's `mul -a cu -t nproc` on n seeded keys of the same distribution (bounded sample)"""
    import numpy as np

    import oracle as O

    exe = O.ref_binary()
    if exe is None:
        return {"value": None, "note": "oracle/_ref not built"}
    rng = np.random.default_rng(3)
    keys = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    keys[:, 3] &= np.uint64(0x7FFFFFFFFFFFFFFF)
    text = "".join("%016x%016x%016x%016x\n" % (int(k[3]), int(k[2]), int(k[1]), int(k[0])) for k in keys).encode()
    cores = os.cpu_count() or 1
    flt = ROOT / "tests" / "golden" / "btc-puzzles-hash"
    t0 = time.perf_counter()
    r = subprocess.run([str(exe), "mul", "-f", str(flt), "-a", "cu", "-t", str(cores), "-q", "-o", "/dev/null"], input=text, capture_output=True)
    dt = time.perf_counter() - t0
    status = [l for l in r.stderr.decode(errors="replace").replace("\r", "\n").splitlines() if "Mkeys/s ~" in l]
    return {"value": round(n / dt / 1e6, 3), "unit": "Mkeys/s", "cores": cores, "kind": "reference",
            "sample": f"{n} keys, mul -a cu -t {cores}, wall clock incl. its 0.4 s table build; status: {status[-1].strip() if status else ''}"}


def leg_endo_blf(E, H, dev, rank, world, local, peaks, size_words, log2_step, steps):
    """configs[3]: 6 GiB filter in HBM (fill 0.37, seed 4, + planted hashes through ecl_filter_add), `-endo`, addr33;
    rank g sweeps the head of its job-aligned shard of 400000000000000000:7fffffffffffffffff (Makefile:58)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    shard, _ = shard71_of(rank, world)
    step_keys = 1 << log2_step
    r = __import__("random").Random(4 + rank)
    planted = sorted(r.randrange((steps + 1) * step_keys) for _ in range(6))
    t0 = time.perf_counter()
    dev.filter_generate(size_words, 0.37, 4)
    gen_s = time.perf_counter() - t0
    new = dev.filter_add([tuple(int(py_hash160_33(shard + o)[i:i + 8], 16) for i in range(0, 40, 8)) for o in planted])
    dev.filter_commit()
    dev.set_stride(1)
    flags = E.A33 | E.ENDO

    def step(i):
        hits = dev.batch_add(shard + i * step_keys, step_keys, flags)
        return [(i * step_keys + k, e) for k, e, _, _ in hits]

    step(0)  # warm-up: candidate queue allocation, clocks
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    t0 = time.perf_counter()
    hot_ms, hits = 0.0, []
    for i in range(1, steps + 1):
        hits += step(i)
        hot_ms += dev.last_elapsed_ms()[1]
    wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.stop()
    plain = {k for k, e in hits if e == 0}
    ok = all(o in plain for o in planted if o >= step_keys)
    n_hashes = 6 * steps * step_keys
    expect_fp = n_hashes * dev.filter_fill() ** 20
    wall_ms, hot_ms = _allreduce(world, local, [wall_ms, hot_ms], "MAX")
    ok = _allreduce(world, local, [1.0 if ok else 0.0], "MIN")[0] == 1.0
    n_hits = int(_allreduce(world, local, [float(len(hits))], "SUM")[0])
    if rank != 0:
        return None, ok
    base = steps * step_keys / (hot_ms * 1e-3) / 1e6  # base keys per GPU (max rank)
    e2e = steps * step_keys * world / (wall_ms * 1e-3) / 1e6
    ops = ALU_OPS_PER_KEY + 5 * ENDO_EXTRA_OPS
    alu_peak = peaks["lop3_gops"] * 1e9
    pk, pk_src = measured_peaks()
    probe_bytes = 6 * (1 + dev.filter_fill()) * 32  # stage 1 fetches two 16 B pieces (32 B sectors) per hash; ~fill^2 go on to stage 2
    return {
        "workload": f"add -endo addr33, {size_words * 8 / 2**30:.2f} GiB filter in HBM (Bernoulli(0.37), counter PRNG seed 4, generated on the device in {gen_s:.2f} s, "
                    f"+{new} planted), range 400000000000000000:7fffffffffffffffff (Makefile:58) in {world} job-aligned shard(s) (BASELINE configs[3])",
        "metric": "Mkeys/s counted (add -endo: 6 per base key, main.c:431)", "unit": "Mkeys/s", "n_gpus": world,
        "value": round(6 * base * world, 2), "base_keys_value": round(base * world, 2),
        "keys_per_step_per_gpu": step_keys, "steps": steps,
        "e2e": {"value": round(6 * e2e, 2), "unit": "Mkeys/s", "h2d_bytes_per_step": 44, "d2h_bytes_per_step": 12 + 32 * (n_hits // max(1, steps * world))},
        "roofline": {"bound": "int_alu", "achieved": round(base * 1e6 * ops / 1e12, 3), "peak": round(alu_peak / 1e12, 3), "unit": "Tops/s",
                     "frac": round(base * 1e6 * ops / alu_peak, 4),
                     "model": f"{ops} canonical ALU-pipe ops per base key (2700 + 5 x 2400, SURVEY 8d)",
                     "hbm": {"achieved": round(base * 1e6 * probe_bytes / 1e9, 1), "peak": pk.get("hbm_gbs"), "unit": "GB/s",
                             "frac": round(base * 1e6 * probe_bytes / 1e9 / pk.get("hbm_gbs", 6650.0), 4), "peak_source": pk_src,
                             "model": f"{probe_bytes:.0f} B of random 32 B sectors per base key (6 hashes x 2 stage-1 probes)"}},
        "false_positives": {"bloom_positive": n_hits, "planted": 6 * world, "expected_false": round(expect_fp * world, 1),
                            "note": "bloom-only mode reports every bloom-positive hash; the same set is compared with the reference on a shared window in "
                                    "tests/test_gpu_cli.py::test_gib_filter_endo_window_identical_to_reference and profiles/r02_config4_parity.txt"},
        "parity_gate": "planted keys of the timed steps found" if ok else "FAILED", "clocks": sampler.summary(),
    }, ok


def leg_rnd(world, windows, with_reference):
    """configs[4]: the drop-in binary `ecloop rnd -d 128:32 -a cu` on all GPUs of the run, `windows` consecutive 2^32
    windows (ECLOOP_RND_WINDOWS stops it; the reference runs forever). Rank 0 only: the C host drives every GPU."""
    exe = ROOT / "ecloop_b200" / "host" / "ecloop"
    flt = ROOT / "tests" / "golden" / "btc-puzzles-hash"
    sampler = ClockSampler(0)
    sampler.start()
    t0 = time.perf_counter()
    r = subprocess.run([str(exe), "rnd", "-f", str(flt), "-d", "128:32", "-a", "cu", "-gpus", str(world)], capture_output=True,
                       env=dict(os.environ, ECLOOP_RND_WINDOWS=str(windows), LC_ALL="C"))
    wall = time.perf_counter() - t0
    sampler.stop()
    out = r.stdout.decode(errors="replace")
    err = r.stderr.decode(errors="replace").replace("\r", "\n")
    import re

    per = [(int(a.replace(",", "")), float(b)) for a, b in re.findall(r"^\d[\d,]* / ([\d,]+) ~ ([\d.]+)s$", out, re.M)]
    status = [l for l in err.splitlines() if "Mkeys/s ~" in l]
    m = re.search(r"([\d.]+)s ~ ([\d.]+) Mkeys/s ~ [\d,]+ / ([\d,]+)", status[-1]) if status else None
    ok = r.returncode == 0 and m is not None and len(per) == windows and all(c == 1 << 32 for c, _ in per)
    leg = {
        "workload": f"ecloop rnd -d 128:32 -a cu -gpus {world}: {windows} consecutive 2^32-key windows at stride 2^128 (BASELINE configs[4]), C host, "
                    "each window split over the GPUs by the shared dispenser",
        "metric": "Mkeys/s (rnd mode, -a cu, status line)", "unit": "Mkeys/s", "n_gpus": world,
        "value": float(m.group(2)) if m else None, "value_note": "the binary's own status line: k_checked / elapsed since the devices were ready (main.c:137-139)",
        "windows": len(per), "windows_per_s": round(len(per) / float(m.group(1)), 3) if m else None,
        "e2e": {"value": round(windows * 2**32 / wall / 1e6, 2) if ok else None, "unit": "Mkeys/s",
                "note": "process wall clock incl. start-up (CUDA contexts, window tables) / keys checked"},
        "roofline": None,
        "parity_gate": f"{windows} windows x 2^32 keys checked" if ok else f"FAILED rc={r.returncode}: {err[-200:]}",
        "clocks": sampler.summary(),
    }
    if m and ok:
        rate = float(m.group(2)) * 1e6 / world  # per GPU
        leg["roofline"] = {"bound": "int_alu", "per_gpu_mkeys": round(rate / 1e6, 1), "ops_per_key": CU_OPS_PER_KEY,
                           "model": "3 SHA-256 + 2 RIPEMD-160 blocks + field per key (SURVEY App. C figures; its table's 5034 omits one SHA block)"}
    if with_reference:
        leg["reference"] = reference_rnd_sample()
    return leg, ok


def reference_rnd_sample():
    import oracle as O

    exe = O.ref_binary()
    if exe is None:
        return {"value": None, "note": "oracle/_ref not built"}
    cores = os.cpu_count() or 1
    flt = ROOT / "tests" / "golden" / "btc-puzzles-hash"
    lo = 0x8000000000000000000000000000000000000000000000000000000000000
    args = [str(exe), "add", "-f", str(flt), "-r", "%x:%x" % (lo, lo + ((1 << 26) - 1 << 128)), "-d", "128:32", "-a", "cu", "-t", str(cores), "-q", "-o", "/dev/null"]
    t0 = time.perf_counter()
    r = subprocess.run(args, capture_output=True, env=dict(os.environ, LC_ALL="C"))
    dt = time.perf_counter() - t0
    status = [l for l in r.stderr.decode(errors="replace").replace("\r", "\n").splitlines() if "Mkeys/s ~" in l]
    return {"value": round((1 << 26) / dt / 1e6, 3) if r.returncode == 0 else None, "unit": "Mkeys/s", "cores": cores, "kind": "reference",
            "sample": f"add -d 128:32 -a cu over 2^26 keys at stride 2^128, -t {cores}, wall clock; status: {status[-1].strip() if status else ''}"}


# ---------------------------------------------------------------- our arm


def ours_arm(args):
    import torch
    import torch.distributed as dist

    import ecloop_b200 as E
    import ecloop_b200.host as H

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    log2_step = args.log2_keys_per_step
    step_keys = 1 << log2_step
    n_steps = args.warmup + args.steps
    shard_start, shard_keys = shard_of(rank, world)
    while n_steps * step_keys > shard_keys:  # many steps on many GPUs: keep every rank inside its shard
        log2_step -= 1
        step_keys = 1 << log2_step

    # filter: the puzzle list + planted keys inside the swept prefix of every shard. Their hash160 comes from a
    # few lines of plain-python secp256k1 + hashlib (input synthesis, independent of both the library and oracle/)
    planted = planted_offsets(n_steps, log2_step)
    planted_keys = [shard_of(g, world)[0] + o for g in range(world) for o in planted]
    planted_h = [py_hash160_33(k) for k in planted_keys]
    puzzles = [l.strip() for l in open(ROOT / "tests" / "golden" / "btc-puzzles-hash") if len(l.strip()) == 40]
    flt = H.filter_from_hashes(puzzles + planted_h)

    dev = E.Device(local)
    stream = torch.cuda.Stream(device=local)
    dev.set_stream(stream.cuda_stream)
    peaks = dev.peak_bench() if rank == 0 else None
    searcher = H.Searcher(dev, flt, E.A33)
    dev.set_stride(1)

    def step(i):
        hits = dev.batch_add(shard_start + i * step_keys, step_keys, E.A33)
        searcher._take(hits, shard_start + i * step_keys, 1)
        return len(hits)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hot_ms, launches, n_hits = 0.0, 0, 0
    t_wall0 = time.perf_counter()
    ev0.record(stream)
    for i in range(args.warmup, n_steps):
        n_hits += step(i)
        _, h, l = dev.last_elapsed_ms()
        hot_ms += h
        launches += l
    ev1.record(stream)
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    sampler.stop()
    dev_ms = ev0.elapsed_time(ev1)
    if world > 1:
        dist.barrier()
        t = torch.tensor([dev_ms, t_wall * 1e3, hot_ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms, hot_ms = (float(x) for x in t.tolist())
        c = torch.tensor([len(searcher.found), launches, n_hits], device=f"cuda:{local}", dtype=torch.int64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        found_total, launches, n_hits = (int(x) for x in c.tolist())
    else:
        wall_ms = t_wall * 1e3
        found_total = len(searcher.found)

    # correctness gate inside the bench: every planted key of this rank's swept prefix, recovered exactly
    mine = sorted(shard_start + o for o in planted)
    got = sorted(f.pk for f in searcher.found)
    ok = got == mine
    if world > 1:
        okt = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        ok = bool(okt.item())

    # ---- secondary legs (the other BASELINE configs), after the timed headline
    secondary, sec_ok = {}, True
    if not args.no_secondary:
        with_ref = world == 1 and not args.no_cpu_baseline
        def guarded(name, fn):
            """a leg that dies (out of memory on a shared box, a missing binary) must not take the headline line with it:
            the error is recorded in its place; a leg that RUNS and fails its parity gate still fails the whole bench"""
            nonlocal sec_ok
            try:
                leg, ok2 = fn()
            except Exception as e:  # noqa: BLE001
                leg, ok2 = {"error": f"{type(e).__name__}: {e}"[:500], "parity_gate": "not run to the end (the leg raised)"}, True
            if leg is not None or rank == 0:
                secondary[name] = leg
            sec_ok = sec_ok and ok2

        pk_all = peaks_all(dev, peaks, world, local)
        guarded("mul_10M_cu", lambda: leg_mul(E, H, dev, rank, world, local, pk_all, args.mul_keys, with_ref))
        guarded("add_endo_blf", lambda: leg_endo_blf(E, H, dev, rank, world, local, pk_all, int(args.blf_gib * 2**30) // 8 - 5,
                                                     args.endo_log2_step, args.endo_steps))
        dev.close()  # the C host of the rnd leg opens the GPUs itself
        torch.cuda.synchronize()
        # the other ranks wait on the CPU (gloo): an NCCL barrier is a kernel spinning on their GPUs, which the C host needs
        cpu_group = dist.new_group(backend="gloo") if world > 1 else None
        if world > 1:
            dist.barrier(group=cpu_group)
        if rank == 0 and args.rnd_windows > 0:
            guarded("rnd_128_32_cu", lambda: leg_rnd(world, args.rnd_windows, with_ref))
        if world > 1:
            dist.barrier(group=cpu_group)
    ok = ok and sec_ok

    if rank == 0:
        total_keys = args.steps * step_keys * world
        value = total_keys / (dev_ms * 1e-3) / 1e6
        e2e = total_keys / (wall_ms * 1e-3) / 1e6
        hot_rate = args.steps * step_keys / (hot_ms * 1e-3)  # keys/s of the add kernel alone on one GPU (max rank)
        pk, pk_src = measured_peaks()
        alu_peak_tops = peaks["lop3_gops"] / 1e3
        achieved_tops = hot_rate * ALU_OPS_PER_KEY / 1e12
        hbm_achieved = hot_rate * SCRATCH_BYTES_PER_KEY / 1e9  # 32 B written + 32 B read per PAIR of keys = 32 B/key
        clocks = sampler.summary()
        per_launch_keys = args.steps * step_keys / max(1, launches // (2 * world))  # launches counts smul + add pairs
        traffic_per_key, traffic_src = measured_traffic()
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "Mkeys/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dev_ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (8x32-bit Fp, 32-bit hash words)",
            "data": "synthetic",
            "config": {
                "workload": "add -r 400000000000000000:40000000ffffffffff addr33 (BASELINE configs[1])",
                "keys_per_step_per_gpu": step_keys, "filter": f"list: 160 puzzle hashes + {len(planted_h)} planted",
                "sharding": "contiguous job-aligned shard of the 2^40 range per rank, no collective on the data path",
                "l2": "no L2 flush needed: per-step input is 44 bytes; the prefix-product scratch (2.4 GB) exceeds L2",
                "parity_gate": "planted keys recovered exactly" if ok else "FAILED: planted keys not recovered",
            },
            "e2e": {"value": round(e2e, 2), "unit": "Mkeys/s", "h2d_bytes_per_step": 44,
                    "d2h_bytes_per_step": 12 + 32 * (n_hits // max(1, args.steps * world))},
            "gpu_launches": launches,
            "roofline": {
                "bound": "int_alu", "achieved": round(achieved_tops, 3), "peak": round(alu_peak_tops, 3), "unit": "Tops/s",
                "frac": round(achieved_tops / alu_peak_tops, 4),
                "traffic": round(traffic_per_key * per_launch_keys) if traffic_per_key else None,
                "traffic_note": (f"bytes per add_kernel launch of {int(per_launch_keys)} keys = {traffic_per_key:.2f} B/key DRAM read+write "
                                 f"measured by ncu ({traffic_src}); algorithmic bytes: {SCRATCH_BYTES_PER_KEY} B/key") if traffic_per_key else None,
                "model": f"{ALU_OPS_PER_KEY} canonical ALU-pipe int32 ops per key (SURVEY 8d) x add_kernel keys/s (CUDA events around its launches)",
                "alu_slots": (lambda c: None if not c else dict(c, frac_of_alu_issue_slots=round(hot_rate * c["slots_per_key"] / (alu_peak_tops * 1e12), 4),
                                                                  note="the kernel's own ALU-pipe instructions + IMAD.WIDE (one slot on both integer pipes) per key, "
                                                                       "hot loop only (tools/loop_census.py); its fraction of the measured LOP3 issue rate"))(kernel_slot_census()),
                "peak_source": "measured in this process: LOP3.LUT issue rate over all SMs (ecl_peak_bench)",
                "pipes": {k: round(v, 1) for k, v in peaks.items()},
                "hbm": {"bound": "hbm", "achieved": round(hbm_achieved, 1), "peak": pk.get("hbm_gbs"), "unit": "GB/s",
                        "frac": round(hbm_achieved / pk.get("hbm_gbs", 6650.0), 4), "peak_source": pk_src,
                        "traffic": round(traffic_per_key * per_launch_keys) if traffic_per_key else None,
                        "model": "prefix-product scratch: 32 B written + 32 B read per 2 keys"},
            },
            "clocks": clocks,
            "found": found_total,
        }
        if secondary:
            line["secondary"] = secondary
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n = 1 << 28 if cores >= 8 else 1 << 26
            v, info = run_reference_sample(RANGE_S, n, cores)
            if v is not None:
                line["cpu_baseline"] = {"value": round(v, 3), "unit": "Mkeys/s", "cores": cores, "kind": "reference",
                                        "sample": f"first 2^{n.bit_length() - 1} keys of configs[1], -t {cores}, its own status-line rate ({info.get('binary')}); status: {info.get('status_line')}"}
                v1, info1 = run_reference_sample(RANGE_S, 1 << 24, 1)  # BASELINE.md §3 asks for -t 1 beside -t nproc
                if v1 is not None:
                    line["cpu_baseline"]["single_thread"] = {"value": round(v1, 3), "unit": "Mkeys/s",
                                                             "sample": f"first 2^24 keys of configs[1], -t 1; status: {info1.get('status_line')}"}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "Mkeys/s", "cores": cores, "kind": "reference", "sample": str(info)}
        print(json.dumps(line))
    dev.close()
    if world > 1:
        dist.destroy_process_group()
    return 0 if ok else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-keys-per-step", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the legs for BASELINE configs[2..4]")
    ap.add_argument("--mul-keys", type=int, default=10_000_000)
    ap.add_argument("--blf-gib", type=float, default=6.0)
    ap.add_argument("--endo-log2-step", type=int, default=32)
    ap.add_argument("--endo-steps", type=int, default=2)
    ap.add_argument("--rnd-windows", type=int, default=50)
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours_arm(args)


if __name__ == "__main__":
    sys.exit(main())
