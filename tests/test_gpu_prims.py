"""GPU parity, primitive by primitive, through the C-ABI (ecl_prim_*): each device routine against the oracle
on the same seeded inputs, plus the reference's golden vectors. Bit-exact (integer work, no tolerance)."""
import random

import pytest

import oracle as O

pytestmark = pytest.mark.gpu
P, N = O.P_FIELD, O.N_ORDER


@pytest.fixture(scope="module")
def dev():
    import ecloop_b200 as E

    d = E.Device(0)
    yield d
    d.close()


def edge_values():
    return [0, 1, 2, 3, 977, 0x1000003D1, 2**32 - 1, 2**32, 2**64 - 1, 2**128 - 1, 2**255, 2**255 - 19,
            P - 1, P - 2, P - 977, P - 0x1000003D1, (P - 1) // 2, (P + 1) // 2]


def test_fp_mul_sqr(dev):
    import ecloop_b200 as E

    r = random.Random(101)
    a = edge_values() + [r.getrandbits(256) % P for _ in range(4000)]
    b = list(reversed(edge_values())) + [r.getrandbits(256) % P for _ in range(4000)]
    # all edge x edge pairs as well
    ev = edge_values()
    a += [x for x in ev for _ in ev]
    b += [y for _ in ev for y in ev]
    got = dev.fp(E.OP_MUL, a, b)
    assert got == [x * y % P for x, y in zip(a, b)]
    got = dev.fp(E.OP_SQR, a)
    assert got == [x * x % P for x in a]
    # the oracle agrees with python ints on a sample (ties the three together)
    for x, y in list(zip(a, b))[:200]:
        assert O.fp_mul(x, y) == x * y % P


def test_fp_mul_accepts_non_canonical_inputs(dev):
    import ecloop_b200 as E

    r = random.Random(102)
    a = [2**256 - 1, P, P + 1, 2**256 - 1, 2**256 - 2**32] + [r.getrandbits(256) for _ in range(1000)]
    b = [2**256 - 1, 2**256 - 1, P, 1, 2**256 - 977] + [r.getrandbits(256) for _ in range(1000)]
    assert dev.fp(E.OP_MUL, a, b) == [x * y % P for x, y in zip(a, b)]


def test_fp_add_sub_neg(dev):
    import ecloop_b200 as E

    r = random.Random(103)
    ev = edge_values()
    a = [x for x in ev for _ in ev] + [r.getrandbits(256) % P for _ in range(3000)]
    b = [y for _ in ev for y in ev] + [r.getrandbits(256) % P for _ in range(3000)]
    assert dev.fp(E.OP_SUB, a, b) == [(x - y) % P for x, y in zip(a, b)]
    assert dev.fp(E.OP_ADD, a, b) == [(x + y) % P for x, y in zip(a, b)]
    assert dev.fp(E.OP_NEG, a) == [(-x) % P for x in a]
    for x, y in list(zip(a, b))[:300]:
        assert O.fp_sub(x, y) == (x - y) % P


def test_fp_inv(dev):
    import ecloop_b200 as E

    r = random.Random(104)
    a = edge_values() + [r.getrandbits(256) % P for _ in range(500)]
    got = dev.fp(E.OP_INV, a)
    assert got == [pow(x, P - 2, P) for x in a]
    gx = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
    assert dev.fp(E.OP_INV, [gx]) == [0x237AFDF1D2938D86870AAEB8AD77626A67B8E794ABFB076BE61D003687CA9EF6]


def test_scalar_mul(dev):
    r = random.Random(105)
    ks = [1, 2, 3, 0xC936, 2**70, N - 1, N - 2, 2**255, 2**256 - 1, N + 1, 7, 2**128, 0xFFFF, 0x10000, 0xFFFF0000,
          2**16 - 1 << 240, 1 << 240, (1 << 256) - (1 << 240)]
    ks += [r.getrandbits(256) for _ in range(600)]
    ks += [r.getrandbits(16) << (16 * r.randrange(16)) for _ in range(200)]  # single-window scalars
    got = dev.scalar_mul(ks)
    want = [O.ec_mul_g(k % N) for k in ks]
    assert got == want
    assert got[0] == (0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
                      0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8)


def test_scalar_mul_infinity(dev):
    got = dev.scalar_mul([0, N, 5])
    assert got[0] == (0, 0) and got[1] == (0, 0)
    assert got[2] == O.ec_mul_g(5)


APPENDIX_B = [
    (1, "751e76e8199196d454941c45d1b3a323f1433bd6", "91b24bf9f5288532960ac687abb035127b1d28a5"),
    (2, "06afd46bcdfd22ef94ac122aa11f241244a37ecc", "d6c8e828c1eca1bba065e1b83e1dc2a36e387a42"),
    (3, "7dd65592d0ab2fe0d0257d571abf032cd9db93dc", "ec7eced2c57ed1292bc4eb9bfd13c9f7603bc338"),
    (0xC936, "7025b4efb3ff42eb4d6d71fab6b53b4f4967e3dd", "16f39f4f19379a80533da9c81f25beb85d1ef06c"),
    (2**70, "0e137b1e6bb72c5c119a805e65c131b17044d88c", "8bbfa5fdce95eaafcb29a6f65d04482ed872d08c"),
    (N - 1, "adde4c73c7b9cee17da6c7b3e2b2eea1a0dcbe67", "bec08011c9e76dcc42e739a2d7752c2e3ac86e6e"),
]


def test_hash160(dev):
    r = random.Random(106)
    ks = [k for k, _, _ in APPENDIX_B] + [r.getrandbits(256) % N or 1 for _ in range(1500)]
    info = O.pubkey_hashes(ks)
    pts = [(x, y) for x, y, _, _ in info]
    # synthetic coordinates exercise every byte lane, not only curve points
    pts += [(r.getrandbits(256), r.getrandbits(256)) for _ in range(500)] + [(0, 0), (2**256 - 1, 2**256 - 1), (1, 1), (1, 2)]
    h33, h65 = dev.hash160(pts)
    for i, (x, y) in enumerate(pts):
        assert h33[i] == O.hash160_33(x, y), i
        assert h65[i] == O.hash160_65(x, y), i
    for i, (_, a, b) in enumerate(APPENDIX_B):
        assert h33[i] == a and h65[i] == b


@pytest.mark.parametrize("size", [1, 2, 3, 320, 331, 4096, 22084, (1 << 20) + 7])
def test_bloom(dev, size):
    import ctypes as C

    r = random.Random(107 + size)
    bits = (C.c_uint64 * size)()
    members = [[r.getrandbits(32) for _ in range(5)] for _ in range(min(4000, max(1, size * 3)))]
    for h in members:
        O.lib().orc_blf_add(bits, C.c_uint64(size), (C.c_uint32 * 5)(*h))
    probes = members[:500] + [[r.getrandbits(32) for _ in range(5)] for _ in range(6000)]
    probes += [[0] * 5, [0xFFFFFFFF] * 5]
    dev.set_filter(bits)
    got = dev.bloom_has(probes)
    want = [bool(O.lib().orc_blf_has(bits, C.c_uint64(size), (C.c_uint32 * 5)(*h))) for h in probes]
    assert got == want
    assert all(got[: len(members[:500])])  # no false negatives
