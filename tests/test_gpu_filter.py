"""GPU parity of the device-side bloom-filter tooling (lib/utils.c:362-475): chunked load, read-back,
the synthetic generator, and blf-gen's insert loop with its exact "new items" count — against the oracle's
restatements and, through the CLI, against the unmodified reference's blf-gen where oracle/_ref travelled."""
import random
import struct
import subprocess

import numpy as np
import pytest

import oracle as O
from conftest import GOLD, ROOT

pytestmark = pytest.mark.gpu
BIN = ROOT / "ecloop_b200" / "host" / "ecloop"
REF = O.REF_DIR / "ecloop_ref"


@pytest.fixture(scope="module")
def E():
    import ecloop_b200

    return ecloop_b200


@pytest.fixture(scope="module")
def dev(E):
    d = E.Device(0)
    yield d
    d.close()


def test_generate_matches_the_numpy_mirror(dev):
    for size, fill, seed in ((1021, 0.37, 4), (4099, 0.5, 0), (300, 1.0, 7), (300, 0.0, 7), (70001, 0.37109375, 2**63 + 5)):
        dev.filter_generate(size, fill, seed)
        got = dev.filter_read(0, size)
        want = O.synthetic_filter(size, fill, seed)
        assert np.array_equal(got, want)
    dev.filter_generate(1 << 20, 0.37, 4)  # 8 MB: stays in HBM, the fill is measured
    assert abs(dev.filter_fill() - 95 / 256) < 2e-4


def test_chunked_write_read_roundtrip(dev):
    rng = np.random.default_rng(3)
    size = (1 << 16) + 5
    bits = rng.integers(0, 2**63, size=size, dtype=np.uint64)
    dev.filter_alloc(size)
    for off in range(0, size, 9973):
        dev.filter_write(off, bits[off:off + 9973].copy())
    dev.filter_commit()
    assert np.array_equal(dev.filter_read(0, size), bits)
    assert np.array_equal(dev.filter_read(777, 4096), bits[777:777 + 4096])
    # and it is the filter the hot path probes
    hs = [tuple(int(x) for x in rng.integers(0, 2**32, size=5)) for _ in range(2000)]
    want = [all((int(bits[p >> 6]) >> (p & 63)) & 1 for p in O.blf_positions(h, size)) for h in hs]
    assert dev.bloom_has(hs) == want


@pytest.mark.parametrize("size,n,dups", [(509, 3000, 0.2), (6007, 20000, 0.05), (1 << 15, 60000, 0.3)])
def test_filter_add_has_the_reference_count_and_bits(dev, size, n, dups):
    """blf_gen's loop is sequential: a hash counts iff the filter as it was plus the hashes BEFORE it leave one of its
    bits clear. Small filters saturate (later hashes are skipped), duplicates are skipped: both must be counted
    exactly like the reference's single thread does."""
    r = random.Random(size)
    hashes = []
    for _ in range(n):
        if hashes and r.random() < dups:
            hashes.append(r.choice(hashes))
        else:
            hashes.append(tuple(r.getrandbits(32) for _ in range(5)))
    init = O.synthetic_filter(size, 0.1, 11)
    want_bits = [int(x) for x in init]
    # two calls: the second one sees the first one's bits ("updating bloom filter...")
    half = n // 3
    want1 = O.blf_gen_sequential(want_bits, hashes[:half])
    want2 = O.blf_gen_sequential(want_bits, hashes[half:])
    dev.filter_alloc(size)
    dev.filter_write(0, init)
    got1 = dev.filter_add(hashes[:half])
    got2 = dev.filter_add(np.array(hashes[half:], dtype=np.uint32))
    assert (got1, got2) == (want1, want2)
    assert [int(x) for x in dev.filter_read(0, size)] == want_bits
    dev.filter_commit()
    assert all(dev.bloom_has(hashes[:500]))


def test_peer_copy(E, dev):
    if E.device_count() < 2:
        pytest.skip("one GPU")
    dev.filter_generate((1 << 18) + 3, 0.37, 9)
    with E.Device(1) as d1:
        d1.filter_copy_peer(dev)
        assert np.array_equal(d1.filter_read(0, (1 << 18) + 3), dev.filter_read(0, (1 << 18) + 3))
        assert abs(d1.filter_fill() - dev.filter_fill()) < 1e-12


# ---------------------------------------------------------------- the blf-gen tool


def run(exe, args, stdin=None):
    r = subprocess.run([str(exe), *args], input=stdin, capture_output=True, timeout=900)
    return r.returncode, r.stdout.decode(), r.stderr.decode(errors="replace")


@pytest.fixture(scope="module", autouse=True)
def built():
    r = subprocess.run(["make", "-C", str(BIN.parent), "all"], capture_output=True, text=True)
    assert r.returncode == 0 and BIN.exists(), r.stderr


def test_blf_gen_on_the_gpu_is_byte_identical(tmp_path):
    """`make blf` (Makefile:35-44): create from the puzzle list, update with the brainwallet list — GPU tool vs the host
    tool (`-cpu`) vs the unmodified reference: same messages, same counts (160, then 1081: the comment-line quirk),
    same file bytes."""
    outs = {}
    for tag, exe, extra in (("gpu", BIN, []), ("cpu", BIN, ["-cpu"])) + ((("ref", REF, []),) if REF.exists() else ()):
        p = tmp_path / f"{tag}.blf"
        rc, o1, e1 = run(exe, ["blf-gen", "-n", "32768", "-o", str(p), *extra], (GOLD / "btc-puzzles-hash").read_bytes())
        assert rc == 0, e1
        rc, o2, e2 = run(exe, ["blf-gen", "-n", "32768", "-o", str(p), *extra], (GOLD / "btc-bw-hash").read_bytes())
        assert rc == 0, e2
        outs[tag] = (o1.replace(str(p), "F"), o2.replace(str(p), "F"), p.read_bytes())
    assert "added 160 new items" in outs["gpu"][0] and "added 1081 new items" in outs["gpu"][1].replace(",", "")
    assert outs["gpu"][2][:16] == struct.pack("<IIQ", 0x45434246, 1, 22084)
    for tag in outs:
        assert outs[tag] == outs["gpu"], tag


def test_blf_gen_gpu_large_input_with_duplicates(tmp_path):
    """300 000 lines, a third of them repeats, ragged pieces (short lines, 39-character lines, an over-long line): the
    GPU tool and the host tool agree on the count and on every byte"""
    r = random.Random(5)
    lines = []
    for i in range(300000):
        if lines and r.random() < 0.33:
            lines.append(r.choice(lines[-5000:]))
        else:
            lines.append("%040x" % r.getrandbits(160))
    lines[1000] = "abc"
    lines[2000] = "%039x" % r.getrandbits(150)
    lines[3000] = "f" * 100
    data = ("\n".join(lines) + "\n").encode()
    res = []
    for extra in ([], ["-cpu"]):
        p = tmp_path / ("g.blf" if not extra else "c.blf")
        rc, out, err = run(BIN, ["blf-gen", "-n", "400000", "-o", str(p), *extra], data)
        assert rc == 0, err
        res.append(([l for l in out.splitlines() if l.startswith("added")][0].split(";")[0], p.read_bytes()))
    assert res[0] == res[1]
