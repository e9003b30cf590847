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
