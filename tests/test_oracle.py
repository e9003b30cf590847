"""The oracle (oracle/ecl_oracle.c) against the reference's own fixtures and against dumps produced by the
unmodified reference binary (tests/golden, tools/gen_golden.py). CPU only."""
import ctypes as C
import hashlib
import random

import pytest

import oracle as O
from conftest import GOLD, golden_lines

P, N = O.P_FIELD, O.N_ORDER


def dump_lines(hits):
    return [O.format_found(kind, h, pk) for (_, _, kind, h, pk) in hits]


def sha_lines(lines):
    return hashlib.sha256(("\n".join(lines) + "\n").encode()).hexdigest()


# ---------------------------------------------------------------- SURVEY Appendix B golden scalars

APPENDIX_B = [
    (1, "751e76e8199196d454941c45d1b3a323f1433bd6", "91b24bf9f5288532960ac687abb035127b1d28a5"),
    (2, "06afd46bcdfd22ef94ac122aa11f241244a37ecc", "d6c8e828c1eca1bba065e1b83e1dc2a36e387a42"),
    (3, "7dd65592d0ab2fe0d0257d571abf032cd9db93dc", "ec7eced2c57ed1292bc4eb9bfd13c9f7603bc338"),
    (0xC936, "7025b4efb3ff42eb4d6d71fab6b53b4f4967e3dd", "16f39f4f19379a80533da9c81f25beb85d1ef06c"),
    (2**70, "0e137b1e6bb72c5c119a805e65c131b17044d88c", "8bbfa5fdce95eaafcb29a6f65d04482ed872d08c"),
    (N - 1, "adde4c73c7b9cee17da6c7b3e2b2eea1a0dcbe67", "bec08011c9e76dcc42e739a2d7752c2e3ac86e6e"),
    (0x23D4A09295BE678B21A5F1DCEAE1F634A69C1B41775F680EBF8165266471401B,
     "bf1c61ac19576d71d4623b185f3bae2a3d4df6bc", "24f98038e995ee03c4178bccaff1652223eba473"),
]


def test_appendix_b_vectors():
    for k, h33, h65 in APPENDIX_B:
        x, y = O.ec_mul_g(k)
        assert (y * y - x * x * x - 7) % P == 0
        assert O.hash160_33(x, y) == h33
        assert O.hash160_65(x, y) == h65


def test_field_vectors():
    assert O.fp_mul(P - 1, P - 2) == 2
    assert O.fp_mul(P - 1, P - 1) == 1
    gx = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
    assert O.fp_inv(gx) == 0x237AFDF1D2938D86870AAEB8AD77626A67B8E794ABFB076BE61D003687CA9EF6
    assert O.fp_inv(0) == 0
    r = random.Random(11)
    vals = [0, 1, 2, P - 1, P - 2, 2**256 - 1, 2**255, P + 5, 0x1000003D1, 2**128 - 1]
    vals += [r.getrandbits(256) for _ in range(500)]
    for a, b in zip(vals, vals[1:] + vals[:1]):
        assert O.fp_mul(a, b) == a * b % P
        assert O.fn_mul(a, b) == a * b % N
        if a < P and b < P:
            assert O.fp_sub(a, b) == (a - b) % P
            assert O.fp_add(a, b) == (a + b) % P


def test_hashlib_crosscheck():
    r = random.Random(1)
    for _ in range(50):
        k = r.getrandbits(256) % N or 1
        x, y = O.ec_mul_g(k)
        ser = bytes([2 + (y & 1)]) + x.to_bytes(32, "big")
        assert hashlib.new("ripemd160", hashlib.sha256(ser).digest()).hexdigest() == O.hash160_33(x, y)
        ser = b"\x04" + x.to_bytes(32, "big") + y.to_bytes(32, "big")
        assert hashlib.new("ripemd160", hashlib.sha256(ser).digest()).hexdigest() == O.hash160_65(x, y)


def test_bloom_positions_vector():
    pos = O.blf_positions(O.hex_to_h160("751e76e8199196d454941c45d1b3a323f1433bd6"), 320)
    assert pos == [1489, 5749, 1108, 9201, 18457, 5213, 3431, 3397, 16959, 3713,
                   12740, 5053, 2413, 10802, 14190, 1052, 9019, 16790, 931, 15990]


def test_endo_recovered_keys():
    # start 0x8000, first base key
    assert O.calc_priv(0x8000, 1, 0, 1) == N - 0x8000
    assert O.calc_priv(0x8000, 1, 0, 2) % 2**32 == 0x38C2790F
    assert O.calc_priv(0x8000, 1, 0, 3) % 2**32 == 0x9773C832
    for e in range(6):
        pk = O.calc_priv(0x8000, 1, 5, e)
        lam = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72
        base = 0x8005 * pow(lam, e // 2, N) % N
        assert pk == (base if e % 2 == 0 else N - base)


# ---------------------------------------------------------------- reference library, when built here


@pytest.fixture(scope="module")
def reflib():
    p = O.REF_DIR / "libecloop_ref.so"
    if not p.exists():
        pytest.skip("oracle/_ref/libecloop_ref.so not built")
    return C.CDLL(str(p))


def test_fp_against_reference_lib(reflib):
    r = random.Random(3)
    vals = [0, 1, 2, P - 1, P - 2, 0x1000003D1, 2**128 - 1, 2**255 % P] + [r.getrandbits(256) % P for _ in range(2000)]
    for a, b in zip(vals, vals[1:] + vals[:1]):
        out = O.FE()
        reflib.fe_modp_mul(out, O.to_fe(a), O.to_fe(b))
        assert O.from_fe(out) == O.fp_mul(a, b)
        reflib.fe_modp_sqr(out, O.to_fe(a))
        assert O.from_fe(out) == O.fp_mul(a, a)
        reflib.fe_modp_sub(out, O.to_fe(a), O.to_fe(b))
        assert O.from_fe(out) == O.fp_sub(a, b)
        reflib.fe_modn_mul(out, O.to_fe(a % N), O.to_fe(b % N))
        assert O.from_fe(out) == O.fn_mul(a % N, b % N)
    for a in vals[:200]:
        out = O.FE()
        reflib._fe_modp_inv_addchn(out, O.to_fe(a))
        assert O.from_fe(out) == O.fp_inv(a)


def test_bloom_against_reference_lib(reflib):
    class Blf(C.Structure):
        _fields_ = [("size", C.c_size_t), ("bits", C.POINTER(C.c_uint64))]

    r = random.Random(4)
    size = 331  # not a power of two
    bits_ref = (C.c_uint64 * size)()
    bits_orc = (C.c_uint64 * size)()
    blf = Blf(size, C.cast(bits_ref, C.POINTER(C.c_uint64)))
    hs = [[r.getrandbits(32) for _ in range(5)] for _ in range(300)]
    for h in hs[:150]:
        reflib.blf_add(C.byref(blf), (C.c_uint32 * 5)(*h))
        O.lib().orc_blf_add(bits_orc, C.c_uint64(size), (C.c_uint32 * 5)(*h))
    assert list(bits_ref) == list(bits_orc)
    reflib.blf_has.restype = C.c_bool
    for h in hs:
        a = bool(reflib.blf_has(C.byref(blf), (C.c_uint32 * 5)(*h)))
        b = bool(O.lib().orc_blf_has(bits_orc, C.c_uint64(size), (C.c_uint32 * 5)(*h)))
        assert a == b


# ---------------------------------------------------------------- known answers (reference's own tests)


def test_ci_smoke_range(puzzles_filter, golden):
    n, hits, kc = O.add_range(0x8000, 0xFFFF, 0, O.A33, puzzles_filter)
    assert dump_lines(hits) == golden_lines("ka_add_8000_ffff")
    assert kc == 32767  # counter adds job_size = R
    assert "1 / 32767" in golden["ka_add_8000_ffff"]["status_line"][0].replace(",", "")


def test_make_add_window(puzzles_filter):
    # first 2^19 keys of `make add` hold puzzles 16..19 (c936, 1764f, 3080d, 5749f); full 2^24 run is a GPU test
    n, hits, kc = O.add_range(0x8000, 0x8000 + 2**19, 0, O.A33, puzzles_filter)
    want = [l for l in golden_lines("ka_add_8000_ffffff") if int(l.split("\t")[2], 16) < 0x8000 + 2**19]
    assert sorted(dump_lines(hits)) == sorted(want) and len(want) == 4


def test_70bit_window(puzzles_filter):
    start = 0x349B84B6431A6C4EF1 - 3000
    n, hits = O.add_span(start, 1, 4096, O.A33, puzzles_filter)
    assert dump_lines(hits) == golden_lines("ka_add_70bit")


def test_make_mul():
    flt = O.filter_from_text_file(GOLD / "btc-bw-hash")
    assert len(flt.words) == 1081  # comment-line quirk, SURVEY A.7
    keys = [int(l, 16) for l in (GOLD / "btc-bw-priv").read_text().split()]
    n, hits = O.mul_batch(keys, O.A33 | O.A65, flt)
    assert n == 1080
    assert dump_lines(hits) == golden_lines("ka_mul_bw")


# ---------------------------------------------------------------- full dumps from the reference binary

ALL = O.filter_all_ones()


def check_dump(golden, name, lines):
    m = golden[name]
    assert len(lines) == m["n_lines"]
    assert lines[:16] == m["head"] and lines[-16:] == m["tail"]
    assert hashlib.sha256(("\n".join(lines) + "\n").encode()).hexdigest() == m["sha256_emission_order"]


def test_dump_add_cu(golden):
    n, hits = O.add_span(0x8000, 1, 2048, O.A33 | O.A65, ALL)
    lines = dump_lines(hits)
    assert lines == golden_lines("dump_add_8000_cu")
    check_dump(golden, "dump_add_8000_cu", lines)


def test_dump_add_endo(golden):
    n, hits = O.add_span(0x8000, 1, 2048, O.A33 | O.ENDO, ALL)
    check_dump(golden, "dump_add_8000_endo_c", dump_lines(hits))
    n, hits = O.add_span(0x8000, 1, 2048, O.A33 | O.A65 | O.ENDO, ALL)
    check_dump(golden, "dump_add_8000_endo_cu", dump_lines(hits))


def test_dump_add_2p70_and_stride(golden):
    n, hits = O.add_span(2**70, 1, 2048, O.A33, ALL)
    check_dump(golden, "dump_add_2p70_c", dump_lines(hits))
    n, hits, kc = O.add_range(2**70, 2**70 + 7, 7, O.A33 | O.A65, ALL)
    check_dump(golden, "dump_add_2p70_stride7_cu", dump_lines(hits))
    assert hits[1][0] == 0 and hits[2][0] == 1 and hits[2][4] == 2**70 + 128


def test_dump_multi_group_overshoot(golden):
    n, hits, kc = O.add_range(0x8000, 0x9FFF, 0, O.A33, ALL)
    assert kc == 0x1FFF and n == 8192
    check_dump(golden, "dump_add_multi_group", dump_lines(hits))


def test_dump_mul(golden):
    keys = [O.from_fe(_hex(l)) for l in (GOLD / "mul_keys_24.txt").read_text().split()]
    n, hits = O.mul_batch(keys, O.A33 | O.A65, ALL)
    assert dump_lines(hits) == golden_lines("dump_mul_24_cu")


def _hex(s):
    r = O.FE()
    O.lib().orc_fn_from_hex(r, s.encode())
    return r


def test_dump_mul_raw(golden):
    lines = [l.rstrip("\r") for l in (GOLD / "mul_raw_8.txt").read_text().split("\n")]
    keys = [int.from_bytes(hashlib.sha256(l.encode()).digest(), "big") for l in lines if l]
    n, hits = O.mul_batch(keys, O.A33 | O.A65, ALL)
    assert dump_lines(hits) == golden_lines("dump_mul_raw_8_cu")


def test_blf_gen_mirror_against_the_reference_tool(tmp_path):
    """oracle.blf_gen_sequential (what tests/test_gpu_filter.py holds ecl_filter_add to) against the unmodified reference's
    own `blf-gen` (lib/utils.c:409-475): a small filter that saturates and an input with repeats, so that the
    `if (blf_has) continue` branch decides the count; same "added N new items", same file bytes, create and update."""
    import random
    import re
    import struct
    import subprocess

    exe = O.ref_binary()
    if exe is None:
        pytest.skip("oracle/_ref not built")
    r = random.Random(11)
    for n_arg, n_hashes in ((40, 400), (3000, 6000)):
        hashes = []
        for _ in range(n_hashes):
            hashes.append(r.choice(hashes) if hashes and r.random() < 0.25 else tuple(r.getrandbits(32) for _ in range(5)))
        half = n_hashes // 2
        p = tmp_path / f"f{n_arg}.blf"
        counts = []
        for part in (hashes[:half], hashes[half:]):
            text = "".join("%08x%08x%08x%08x%08x\n" % h for h in part).encode()
            res = subprocess.run([str(exe), "blf-gen", "-n", str(n_arg), "-o", str(p)], input=text, capture_output=True, timeout=120)
            assert res.returncode == 0, res.stderr
            counts.append(int(re.search(rb"added ([\d,]+) new items", res.stdout).group(1).replace(b",", b"")))
        raw = p.read_bytes()
        magic, ver, size = struct.unpack("<IIQ", raw[:16])
        assert (magic, ver) == (0x45434246, 1)
        bits = [0] * size
        want = [O.blf_gen_sequential(bits, hashes[:half]), O.blf_gen_sequential(bits, hashes[half:])]
        assert counts == want
        assert want[0] < half  # the small filter saturates / repeats are skipped: the order-dependent branch was taken
        assert list(struct.unpack("<%dQ" % size, raw[16:])) == bits


def test_synthetic_filter_statistics():
    """the counter-based generator behind ecl_filter_generate (numpy mirror): fill = round(fill * 256) / 256 within sampling
    error, bits of a word and of neighbouring words uncorrelated, reproducible, seed-sensitive"""
    import numpy as np

    a = O.synthetic_filter(1 << 14, 0.37, 4)
    b = O.synthetic_filter(1 << 14, 0.37, 4)
    c = O.synthetic_filter(1 << 14, 0.37, 5)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    ones = np.unpackbits(a.view(np.uint8)).astype(np.float64)
    p = 95 / 256
    assert abs(ones.mean() - p) < 4 * (p * (1 - p) / ones.size) ** 0.5
    for lag in (1, 7, 8, 63, 64, 65):
        cov = np.mean(ones[:-lag] * ones[lag:]) - p * p
        assert abs(cov) < 5 * p * (1 - p) / ones.size ** 0.5, (lag, cov)
