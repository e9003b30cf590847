#!/usr/bin/env python3
"""Generate tests/golden/* from the UNMODIFIED reference (oracle/_ref, built from /root/reference).

Runs only in the build container (needs /root/reference). The fixtures it writes are committed so the
GPU box — which has no /root/reference — can check parity against the reference's real output:

  * data files the reference's own known-answer tests use (Makefile:26-30): copied as test inputs;
  * full dumps via the all-ones .blf trick: every visited key with its hash160 and
    recovered private key, in the reference's -t 1 emission order. Small dumps are stored whole (gzip);
    big ones as sha256 + head/tail lines;
  * known-answer sets: `make add` (9 keys), 8000:fffffff (13 keys), `make mul` (1080 keys).

Usage: python tools/gen_golden.py
"""
from __future__ import annotations

import gzip
import hashlib
import json
import os
import shutil
import struct
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import oracle as O  # noqa: E402

REF = Path("/root/reference")
GOLD = ROOT / "tests" / "golden"
N = O.N_ORDER


def run_ref(args, stdin=None):
    rc, out, err = O.run_ref(args, stdin_bytes=stdin, timeout=3600)
    assert rc == 0, (args, rc, err[-400:])
    return out, err


def allones(path):
    with open(path, "wb") as f:
        f.write(struct.pack("<IIQ", 0x45434246, 1, 1) + b"\xff" * 8)


def dump(tmp, name, args, stdin=None, store_full=False, note=""):
    outp = Path(tmp) / (name + ".txt")
    if outp.exists():
        outp.unlink()
    _, err = run_ref([*args, "-q", "-o", str(outp)], stdin=stdin)
    data = outp.read_bytes()
    lines = data.decode().splitlines()
    meta = {
        "name": name,
        "args": [a if not str(a).startswith(str(tmp)) else "<allones.blf>" for a in args],
        "note": note,
        "n_lines": len(lines),
        "sha256_emission_order": hashlib.sha256(data).hexdigest(),
        "sha256_sorted": hashlib.sha256(("\n".join(sorted(lines)) + "\n").encode()).hexdigest(),
        "md5_sorted": hashlib.md5(("\n".join(sorted(lines)) + "\n").encode()).hexdigest(),
        "head": lines[:16],
        "tail": lines[-16:],
        "status_line": [l for l in err.decode(errors="replace").replace("\r", "\n").splitlines() if "Mkeys/s ~" in l][-1:],
    }
    if store_full:
        with gzip.GzipFile(GOLD / (name + ".txt.gz"), "wb", mtime=0) as g:
            g.write(data)
        meta["full"] = name + ".txt.gz"
    print(f"{name}: {len(lines)} lines")
    return meta


def main():
    O.build()
    GOLD.mkdir(parents=True, exist_ok=True)
    for f in ("btc-puzzles-hash", "btc-bw-priv", "btc-bw-hash"):
        shutil.copyfile(REF / "data" / f, GOLD / f)
    metas = []
    with tempfile.TemporaryDirectory() as tmp:
        blf = str(Path(tmp) / "allones.blf")
        allones(blf)
        ph = str(REF / "data" / "btc-puzzles-hash")
        # --- known answers (the reference's own tests)
        metas.append(dump(tmp, "ka_add_8000_ffff", ["add", "-f", ph, "-r", "8000:ffff", "-t", "1"], store_full=True,
                          note="CI smoke (.github/workflows/ci.yml:27): 1 key"))
        metas.append(dump(tmp, "ka_add_8000_ffffff", ["add", "-f", ph, "-r", "8000:ffffff", "-t", "8"], store_full=True,
                          note="`make add` (Makefile:26-27): 9 keys"))
        metas.append(dump(tmp, "ka_add_8000_fffffff", ["add", "-f", ph, "-r", "8000:fffffff", "-t", "8"],
                          store_full=True, note="readme.md:200: 13 keys"))
        metas.append(dump(tmp, "ka_add_70bit", ["add", "-f", ph, "-r", "349b84b6431a000000:349b84b6431affffff", "-t", "8"],
                          store_full=True, note="puzzle 70 window"))
        metas.append(dump(tmp, "ka_mul_bw", ["mul", "-f", str(REF / "data" / "btc-bw-hash"), "-a", "cu", "-t", "1"],
                          stdin=(REF / "data" / "btc-bw-priv").read_bytes(), store_full=True,
                          note="`make mul` (Makefile:29-30): 1080 keys"))
        # --- full dumps (all-ones bloom): every visited key
        metas.append(dump(tmp, "dump_add_8000_cu", ["add", "-f", blf, "-r", "8000:8007", "-t", "1", "-a", "cu"],
                          store_full=True, note="SURVEY App. B dump heads; 2048 keys x {33,65}"))
        metas.append(dump(tmp, "dump_add_8000_endo_c", ["add", "-f", blf, "-r", "8000:8007", "-t", "1", "-a", "c", "-endo"],
                          note="2048 keys x 6 endo images"))
        metas.append(dump(tmp, "dump_add_8000_endo_cu", ["add", "-f", blf, "-r", "8000:8007", "-t", "1", "-a", "cu", "-endo"],
                          note="2048 keys x 6 x {33,65}"))
        metas.append(dump(tmp, "dump_add_2p70_c", ["add", "-f", blf, "-r", "400000000000000000:400000000000000007", "-t", "1"],
                          note="config-2 start, 2048 keys"))
        metas.append(dump(tmp, "dump_add_2p70_stride7_cu",
                          ["add", "-f", blf, "-r", "400000000000000000:400000000000000007", "-t", "1", "-a", "cu", "-d", "7:20"],
                          note="stride 2^7: keys 2^70 + 128 j"))
        metas.append(dump(tmp, "dump_add_multi_group", ["add", "-f", blf, "-r", "8000:9fff", "-t", "1"],
                          note="R=0x1fff -> 1 job of 8191 -> 4 groups = 8192 keys (A.1 overshoot)"))
        metas.append(dump(tmp, "dump_add_2jobs", ["add", "-f", blf, "-r", "10000:40ffff", "-t", "1"],
                          note="R=0x3fffff -> 2 jobs of 2^21 keys: 4194304 lines"))
        # mul: 24 keys (multiple of 8 so there are no phantom lanes, A.7), incl. edge scalars
        keys = [1, 2, 3, 0xC936, 2**70, N - 1, N - 2, 2**255, 2**256 - 1, N + 1, 7, 2**128]
        bw = (REF / "data" / "btc-bw-priv").read_text().split()[:12]
        text = "".join("%x\n" % k for k in keys) + "".join(l + "\n" for l in bw)
        (GOLD / "mul_keys_24.txt").write_text(text)
        metas.append(dump(tmp, "dump_mul_24_cu", ["mul", "-f", blf, "-a", "cu", "-t", "1"], stdin=text.encode(),
                          store_full=True, note="24 keys incl. edge scalars (k >= n is reduced once, ecc.c:264)"))
        raw = "hello\nworld\ncorrect horse battery staple\n\r\nabc\nsatoshi\nbitcoin\npassword\n1\n"
        (GOLD / "mul_raw_8.txt").write_text(raw)
        metas.append(dump(tmp, "dump_mul_raw_8_cu", ["mul", "-f", blf, "-a", "cu", "-t", "1", "-raw"], stdin=raw.encode(),
                          store_full=True, note="-raw: key = SHA-256(line) (main.c:506-527); empty lines skipped"))
    (GOLD / "golden.json").write_text(json.dumps({"reference": "vladkens/ecloop v0.5.0", "fixtures": metas}, indent=1))
    print("wrote", GOLD / "golden.json")


if __name__ == "__main__":
    main()
