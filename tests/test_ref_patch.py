"""The drop-in claim, mechanically: tools/ref_patch/main.patch applied to the reference's own main.c
builds against libecloop_b200.so, and — on the GPU box, where the prebuilt binary travels — the PATCHED REFERENCE passes
the reference's own `make add` / `make mul` checks (Makefile:26-30: 9 keys, 1080 keys) with every hit re-verified by
the reference's CPU code (pk_verify_hash)."""
import hashlib
import re
import subprocess
from pathlib import Path

import pytest

from conftest import GOLD, ROOT

PATCH_DIR = ROOT / "tools" / "ref_patch"
PATCHED = PATCH_DIR / "_build" / "ecloop_patched"
REF_SRC = Path("/root/reference/main.c")


@pytest.mark.skipif(not REF_SRC.exists(), reason="the reference sources are only in the build container")
def test_patch_applies_and_builds():
    import ecloop_b200 as E

    E.load_library()  # the .so the binary links against
    r = subprocess.run(["bash", str(PATCH_DIR / "build.sh")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert PATCHED.exists()
    syms = subprocess.run(["nm", "-D", "--undefined-only", str(PATCHED)], capture_output=True, text=True).stdout
    for s in ("ecl_open", "ecl_set_filter", "ecl_set_stride", "ecl_add_submit", "ecl_mul_submit", "ecl_collect"):
        assert re.search(rf"\bU {s}\b", syms), s
    # the patch is small and touches only the two seams + set-up (no reference source is carried in the repo)
    patch = (PATCH_DIR / "main.patch").read_text().splitlines()
    assert sum(1 for l in patch if l.startswith("-") and not l.startswith("---")) <= 6
    assert sum(1 for l in patch if l.startswith(" ")) <= 20


@pytest.mark.skipif(not REF_SRC.exists(), reason="the reference sources are only in the build container")
def test_patch_is_current():
    """main.patch is what tools/ref_patch/make_patch.py generates from the reference as it lies"""
    before = (PATCH_DIR / "main.patch").read_text()
    r = subprocess.run(["python", str(PATCH_DIR / "make_patch.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (PATCH_DIR / "main.patch").read_text() == before


def run(args, stdin=None):
    r = subprocess.run([str(PATCHED), *args], input=stdin, capture_output=True, timeout=900)
    return r.returncode, r.stdout.decode(), r.stderr.decode(errors="replace")


def status(err):
    lines = [l for l in err.replace("\r", "\n").splitlines() if "Mkeys/s ~" in l]
    m = re.search(r"~ ([\d,]+) / ([\d,]+)", lines[-1])
    return int(m.group(1).replace(",", "")), int(m.group(2).replace(",", ""))


@pytest.mark.gpu
@pytest.mark.skipif(not PATCHED.exists(), reason="tools/ref_patch/_build did not travel")
def test_patched_reference_make_add_and_make_mul(tmp_path):
    o1 = tmp_path / "add.txt"
    rc, out, err = run(["add", "-f", str(GOLD / "btc-puzzles-hash"), "-r", "8000:ffffff", "-t", "4", "-q", "-o", str(o1)])
    assert rc == 0, err
    lines = sorted(o1.read_text().splitlines())
    assert len(lines) == 9 and status(err) == (9, 16777216)
    assert hashlib.md5(("\n".join(lines) + "\n").encode()).hexdigest() == "6309efbef3fda727aac597db3a7f1a27"
    o2 = tmp_path / "mul.txt"
    rc, out, err = run(["mul", "-f", str(GOLD / "btc-bw-hash"), "-a", "cu", "-t", "4", "-q", "-o", str(o2)], (GOLD / "btc-bw-priv").read_bytes())
    assert rc == 0, err
    lines = sorted(o2.read_text().splitlines())
    assert len(lines) == 1080 and status(err) == (1080, 1080)
    assert hashlib.md5(("\n".join(lines) + "\n").encode()).hexdigest() == "d73787c22e3e626b1ab8b0e6ccf6d394"


@pytest.mark.gpu
@pytest.mark.skipif(not PATCHED.exists(), reason="tools/ref_patch/_build did not travel")
def test_patched_reference_endo_and_stride(tmp_path):
    """-endo -a cu and a stride: the patched reference's stdout equals the unmodified reference's (oracle/_ref)"""
    import oracle as O

    ref = O.REF_DIR / "ecloop_ref"
    if not ref.exists():
        pytest.skip("oracle/_ref did not travel")
    for args in (["add", "-f", str(GOLD / "btc-puzzles-hash"), "-r", "8000:1ffff", "-t", "1", "-a", "cu", "-endo"],
                 ["add", "-f", str(GOLD / "btc-puzzles-hash"), "-r", "100000:ffffff", "-t", "1", "-d", "3:24"]):
        rc1, out1, err1 = run(args)
        r2 = subprocess.run([str(ref), *args], capture_output=True, timeout=900)
        assert rc1 == 0 and r2.returncode == 0, err1
        assert out1 == r2.stdout.decode()
        assert status(err1) == status(r2.stderr.decode(errors="replace"))
