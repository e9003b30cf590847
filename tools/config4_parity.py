#!/usr/bin/env python3
"""BASELINE configs[3] as SURVEY §8d item 4 specifies it, against the unmodified reference: a 6 GiB `.blf`
(805 306 368 words, bits i.i.d. Bernoulli(0.37) from the counter PRNG with seed 4, generated on the device) + 64 planted
hashes (inserted with blf_add semantics by ecl_filter_add), `add -endo` over a 2^LOG2-key window inside
400000000000000000:7fffffffffffffffff (Makefile:58). The drop-in binary on every visible GPU and oracle/_ref on the
host cores read the SAME file and sweep the SAME window; the sorted `-o` files — planted keys and bloom false
positives alike — must be identical. Also times the streamed `.blf` load (ECLOOP_VERBOSE).

Usage: python tools/config4_parity.py [log2_window=30] [gib=6]     (run on the GPU box; writes to stdout)"""
import os
import random
import re
import struct
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ecloop_b200 as E  # noqa: E402
import oracle as O  # noqa: E402  (the reference binary lives in oracle/_ref)

log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 30
gib = float(sys.argv[2]) if len(sys.argv) > 2 else 6.0
size = int(gib * 2**30) // 8  # 805 306 368 words for 6 GiB
lo = 0x400000000000000000 + (0x1234 << 44)
r = random.Random(4)
planted = sorted(lo + r.randrange(1 << log2) for _ in range(64))
BIN = ROOT / "ecloop_b200" / "host" / "ecloop"
REF = O.ref_binary()

tmp = None
for base in (None, "/dev/shm"):
    try:
        d = tempfile.mkdtemp(dir=base)
        if os.statvfs(d).f_bavail * os.statvfs(d).f_frsize > size * 8 + (1 << 30):
            tmp = Path(d)
            break
    except OSError:
        pass
assert tmp, "no scratch space for the filter file"
blf = tmp / "cfg4.blf"

t0 = time.perf_counter()
with E.Device(0) as dev:
    dev.filter_generate(size, 0.37, 4)
    t_gen = time.perf_counter() - t0
    new = dev.filter_add([tuple(O.hex_to_h160(h33)) for _, _, h33, _ in O.pubkey_hashes(planted)])
    dev.filter_commit()
    fill = dev.filter_fill()
    t1 = time.perf_counter()
    with open(blf, "wb") as f:
        f.write(struct.pack("<IIQ", 0x45434246, 1, size))
        for off in range(0, size, 1 << 25):
            f.write(dev.filter_read(off, min(1 << 25, size - off)).tobytes())
    t_save = time.perf_counter() - t1
print(f"filter: {size} words = {size * 8 / 2**30:.2f} GiB, generated on the device in {t_gen:.2f} s, fill {fill:.5f}, {new} planted hashes new, "
      f"saved to {blf} in {t_save:.1f} s", flush=True)

rng = "%x:%x" % (lo, lo + (1 << log2) - 1)


def status(err):
    lines = [l for l in err.replace("\r", "\n").splitlines() if "Mkeys/s ~" in l]
    return lines[-1].strip() if lines else ""


o1, o2 = tmp / "ours.txt", tmp / "ref.txt"
t0 = time.perf_counter()
p = subprocess.run([str(BIN), "add", "-f", str(blf), "-r", rng, "-endo", "-q", "-o", str(o1)], capture_output=True, env=dict(os.environ, ECLOOP_VERBOSE="1", LC_ALL="C"))
t_ours = time.perf_counter() - t0
err = p.stderr.decode(errors="replace")
print(f"ours  : rc {p.returncode}, {t_ours:.1f} s wall, {E.device_count()} GPU(s); {' | '.join(l for l in err.replace(chr(13), chr(10)).splitlines() if l.startswith('filter:'))} | {status(err)}", flush=True)
cores = os.cpu_count() or 1
t0 = time.perf_counter()
q = subprocess.run([str(REF), "add", "-f", str(blf), "-r", rng, "-endo", "-t", str(cores), "-q", "-o", str(o2)], capture_output=True, env=dict(os.environ, LC_ALL="C"))
t_ref = time.perf_counter() - t0
print(f"ref   : rc {q.returncode}, {t_ref:.1f} s wall, -t {cores}; {status(q.stderr.decode(errors='replace'))}", flush=True)
ours, ref = sorted(o1.read_text().splitlines()), sorted(o2.read_text().splitlines())
keys_found = {int(l.split("\t")[2], 16) for l in ref}
n_planted = sum(1 for k in planted if k in keys_found)
print(f"window: 2^{log2} base keys x 6 images = {6 << log2} hashes; expected false positives {(6 << log2) * fill ** 20:.1f}")
print(f"lines : ours {len(ours)}, reference {len(ref)}, planted among them {n_planted} of 64, false positives {len(ref) - n_planted}")
print("RESULT:", "IDENTICAL (sorted -o files byte for byte)" if ours == ref and p.returncode == 0 and q.returncode == 0 else "DIFFERENT")
for l in ref[:6]:
    print("  ", l)
for p_ in (blf, o1, o2):
    p_.unlink(missing_ok=True)
sys.exit(0 if ours == ref else 1)
