"""GPU parity of the mul path (ecl_mul_submit/ecl_collect): `make mul`, the reference binary's dumps."""
import hashlib

import numpy as np
import pytest

from conftest import GOLD, golden_lines

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import ecloop_b200 as E
    import ecloop_b200.host as H

    d = E.Device(0)
    yield E, H, d
    d.close()


def test_make_mul_1080(env):
    E, H, dev = env
    flt = H.load_filter(GOLD / "btc-bw-hash")
    assert flt.label == "list (1,081)"  # comment-line quirk, SURVEY A.7
    keys = [H.fe_modn_from_hex(l) for l in (GOLD / "btc-bw-priv").read_text().split()]
    s = H.Searcher(dev, flt, E.A33 | E.A65)
    found = s.cmd_mul(keys)
    lines = [f.line() for f in found]
    assert len(lines) == 1080 and s.k_checked == len(keys)
    assert lines == golden_lines("ka_mul_bw")
    assert hashlib.md5(("\n".join(sorted(lines)) + "\n").encode()).hexdigest() == "d73787c22e3e626b1ab8b0e6ccf6d394"


def test_dump_mul_24(env):
    E, H, dev = env
    allones = H.Filter(np.full(1, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64), None)
    keys = [H.fe_modn_from_hex(l) for l in (GOLD / "mul_keys_24.txt").read_text().split()]
    found = H.Searcher(dev, allones, E.A33 | E.A65).cmd_mul(keys)
    assert [f.line() for f in found] == golden_lines("dump_mul_24_cu")


def test_dump_mul_raw(env):
    E, H, dev = env
    allones = H.Filter(np.full(1, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64), None)
    keys = []
    for line in (GOLD / "mul_raw_8.txt").read_bytes().split(b"\n"):
        line = line.rstrip(b"\r")
        if line:
            keys.append(H.raw_to_key(line))
    found = H.Searcher(dev, allones, E.A33 | E.A65).cmd_mul(keys)
    assert [f.line() for f in found] == golden_lines("dump_mul_raw_8_cu")


def test_zero_keys_are_skipped(env):
    E, H, dev = env
    allones = H.Filter(np.full(1, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64), None)
    found = H.Searcher(dev, allones, E.A33).cmd_mul([0, 1, H.N_ORDER, 2])
    assert [f.pk for f in found] == [1, 2]


def test_large_batch_linearity(env):
    """100k seeded keys: k*G hashes equal those of the add path walking the same keys (two independent kernels)"""
    E, H, dev = env
    start = 2**100 + 977
    n = 2048 * 48
    flt = H.Filter(np.full(3, 0x0F0F0F0F0F0F0F0F, dtype=np.uint64), None)  # half-full: ~1e-6 pass... use dump instead
    allones = H.Filter(np.full(1, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64), None)
    s1 = H.Searcher(dev, allones, E.A33)
    a = s1.cmd_mul([start + i for i in range(n)])
    dev.set_stride(1)
    dev.set_filter(allones.bits)
    b = dev.batch_add(start, n, E.A33, cap=n)
    assert [(f.pk - start, f.h160) for f in a] == [(k, h) for k, _, _, h in b]


def test_two_submits_in_flight_come_back_in_order(env):
    """ECL_MUL_DEPTH submits may be pending; ecl_collect returns them in submission order with their own key indices; a
    third submit is refused (ECL_E_STATE) and leaves the queue intact"""
    E, H, dev = env
    allones = H.Filter(np.full(1, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64), None)
    dev.set_filter(allones.bits)
    a = [2**200 + i for i in range(3000)]
    b = [2**90 + 7 * i for i in range(5000)]
    dev.mul_submit(a, E.A33)
    dev.mul_submit(b, E.A33 | E.A65)
    with pytest.raises(E.EclError) as ei:
        dev.mul_submit(a, E.A33)
    assert ei.value.code == -5
    hits_a, n_a = dev.collect(cap=1 << 14)
    hits_b, n_b = dev.collect(cap=1 << 14)
    assert (n_a, n_b) == (3000, 5000)
    assert [k for k, _, _, _ in hits_a] == list(range(3000))
    assert [(k, kd) for k, _, kd, _ in hits_b] == [(i, kd) for i in range(5000) for kd in (0, 1)]
    # the same hashes as one submit each
    assert hits_a == dev.mul_batch(a, E.A33, cap=1 << 14)
    assert hits_b == dev.mul_batch(b, E.A33 | E.A65, cap=1 << 14)
    with pytest.raises(E.EclError):
        dev.collect()


def test_keys_spanning_every_window_digit(env):
    """scalars that exercise the W-bit window boundaries of the table (all-ones, single bits, digits of 1 and 2^W - 1,
    n - 1, n + 1, 2^256 - 1) against the oracle's double-and-add"""
    import oracle as O

    E, H, dev = env
    ks = [1, 2, 3, 2**24 - 1, 2**24, 2**24 + 1, 2**48 - 1, 2**240, 2**240 - 1, 2**255, 2**256 - 1, O.N_ORDER - 1, O.N_ORDER + 1,
          O.N_ORDER - 2**24, int("01" * 128, 2), int("10" * 128, 2)] + [1 << b for b in range(0, 256, 7)] + [(1 << b) - 1 for b in range(20, 256, 11)]
    got = dev.scalar_mul(ks)
    for k, (x, y) in zip(ks, got):
        want = O.ec_mul_g(k % O.N_ORDER)
        assert (x, y) == (want if want else (0, 0)), hex(k)
