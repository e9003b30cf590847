"""Host-side mirror (ecloop_b200/host.py) against the oracle and the reference's fixtures: filter loading, bloom
build, job plan, calc_priv (A.10), hex / -raw key parsing (A.6). CPU only, no device calls."""
import ctypes as C
import random

import numpy as np
import pytest

import oracle as O
from conftest import GOLD

import ecloop_b200.host as H


def test_load_filter_matches_reference_loader():
    f = H.load_filter(GOLD / "btc-puzzles-hash")
    o = O.filter_from_text_file(GOLD / "btc-puzzles-hash")
    assert f.hashes == o.words and len(f.hashes) == 160
    assert [int(x) for x in f.bits] == o.bits_list()
    f = H.load_filter(GOLD / "btc-bw-hash")
    o = O.filter_from_text_file(GOLD / "btc-bw-hash")
    assert len(f.hashes) == 1081 and f.hashes == o.words  # the 52-char comment line becomes one entry (A.7)
    assert [int(x) for x in f.bits] == o.bits_list()


def test_bloom_positions_vector():
    h = tuple(int("751e76e8199196d454941c45d1b3a323f1433bd6"[i:i + 8], 16) for i in range(0, 40, 8))
    assert H.blf_positions(h, 320) == [1489, 5749, 1108, 9201, 18457, 5213, 3431, 3397, 16959, 3713,
                                       12740, 5053, 2413, 10802, 14190, 1052, 9019, 16790, 931, 15990]


def test_blf_roundtrip_and_reference_compat(tmp_path):
    r = random.Random(5)
    bits = np.zeros(22084, dtype=np.uint64)  # `make blf` size (n=32768)
    hs = [tuple(r.getrandbits(32) for _ in range(5)) for _ in range(300)]
    for h in hs[:200]:
        H.blf_add(bits, h)
    p = tmp_path / "t.blf"
    H.blf_save(p, bits)
    raw = p.read_bytes()
    assert raw[:16] == bytes.fromhex("46424345" "01000000" "4456000000000000")  # SURVEY §8c header bytes
    back = H.blf_load(p)
    assert (back == bits).all()
    f = H.load_filter(p)
    assert f.hashes is None and f.label == "bloom"
    for h in hs:
        want = bool(O.lib().orc_blf_has((C.c_uint64 * bits.size)(*[int(x) for x in bits]), C.c_uint64(bits.size), (C.c_uint32 * 5)(*h)))
        assert H.blf_has(bits, h) == want
    if O.ref_binary() is not None:  # the reference accepts our file
        rc, out, err = O.run_ref(["blf-check", "-f", str(p)] if False else ["add", "-f", str(p), "-r", "8000:8007", "-t", "1", "-q", "-o", "/dev/null"])
        assert rc == 0


def test_calc_priv_all_endo():
    r = random.Random(6)
    for _ in range(200):
        start, offs, off, e = r.getrandbits(r.choice([30, 70, 255])) + 1, r.randrange(0, 40), r.getrandbits(40), r.randrange(6)
        assert H.calc_priv(start, 1 << offs, off, e) == O.calc_priv(start, 1 << offs, off, e)


@pytest.mark.parametrize("rs,re_,offs,jobs,job", [
    (0x8000, 0xFFFF, 0, 1, 0x7FFF),
    (0x8000, 0x9FFF, 0, 1, 0x1FFF),
    (0x8000, 0xFFFFFF, 0, 8, 2**21),
    (0x10000, 0x40FFFF, 0, 2, 2**21),
    (2**70, 2**70 + 7, 7, 1, 7),
    (2**70, 2**70 + 2**24, 3, 1, 2**21),  # stride 8: one job spans 2^24 in key value
    (2**70, 2**70 + 2**40 - 1, 0, 2**19, 2**21),
])
def test_job_plan(rs, re_, offs, jobs, job):
    j, starts = H.job_plan(rs, re_, offs)
    assert j == job and len(starts) == jobs
    assert starts[0] == rs
    if jobs > 1:
        assert starts[1] - starts[0] == job << offs


def test_job_plan_matches_oracle_counter():
    f = O.filter_from_hashes(["00" * 20])
    for rs, re_, offs in [(0x8000, 0xFFFF, 0), (0x8000, 0x9FFF, 0), (2**70, 2**70 + 7, 7), (0x8000, 0x8000 + 3 * 2**21 + 5, 0)]:
        job, starts = H.job_plan(rs, re_, offs)
        if len(starts) * job > 2**20:
            continue
        _, _, kc = O.add_range(rs, re_, offs, O.A33, f)
        assert kc == job * len(starts)


def test_hex_and_raw_parsing():
    assert H.fe_modn_from_hex("c936") == 0xC936
    assert H.fe_modn_from_hex("0x00c9 36\r") == 0xC936  # non-hex characters are skipped (x is not hex)
    assert H.fe_modn_from_hex("%x" % (H.N_ORDER + 5)) == 5
    assert H.fe_modn_from_hex("f" * 64) == 2**256 - 1 - H.N_ORDER
    r = O.FE()
    for s in ["c936", "zz12", "F" * 64, "%x" % (H.N_ORDER - 1), "1" * 70]:
        O.lib().orc_fn_from_hex(r, s.encode())
        assert H.fe_modn_from_hex(s) == O.from_fe(r)
    assert H.raw_to_key(b"hello") == 0x2CF24DBA5FB0A30E26E83B2AC5B9E29E1B161E5C1FA7425E73043362938B9824


def test_found_line_format():
    h = (0x7025B4EF, 0xB3FF42EB, 0x4D6D71FA, 0xB6B53B4F, 0x4967E3DD)
    assert H.format_found(0, h, 0xC936) == "addr33\t7025b4efb3ff42eb4d6d71fab6b53b4f4967e3dd\t" + "%064x" % 0xC936
    assert H.format_found(1, h, 1, tab=False) == "addr65: 7025b4efb3ff42eb4d6d71fab6b53b4f4967e3dd <- " + "%064x" % 1


def test_fp64_multiplication_model_is_exact():
    """tools/f64mul_model.py: the integer model of csrc/fp64mul.cuh (every intermediate below 2^53, accumulator inside
    its binade, result congruent to a*b mod p and weak enough to feed the next multiplication)"""
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "tools" / "f64mul_model.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    assert r.stdout.startswith("ok")
