"""GPU parity of the EXPERIMENTAL FP64-pipe field arithmetic (csrc/fp64mul.cuh). It is not on the hot path and not in
the product library: these tests load libecloop_b200_exp.so (the same sources built with -DECL_EXPERIMENTAL)."""
import random

import pytest

import oracle as O

pytestmark = pytest.mark.gpu
P = O.P_FIELD


@pytest.fixture(scope="module")
def dev():
    import ecloop_b200 as E

    d = E.Device(0, experimental=True)
    yield d
    d.close()


def edge_values():
    return [0, 1, 2, 3, 977, 0x1000003D1, 2**32 - 1, 2**32, 2**64 - 1, 2**128 - 1, 2**255, 2**255 - 19,
            P - 1, P - 2, P - 977, P - 0x1000003D1, (P - 1) // 2, (P + 1) // 2]


def test_fp_mul_on_the_fp64_pipe(dev):
    """experimental fe6_mul (csrc/fp64mul.cuh, 44-bit limbs as doubles, DFMA): same canonical residues as fe_mul for
    canonical, non-canonical and edge inputs, and 16 chained multiplications in its own weak-limb form"""
    import ecloop_b200 as E

    r = random.Random(131)
    ev = edge_values() + [2**256 - 1, P, P + 1, 2**256 - 2**32, 2**44 - 1, 2**44, 2**88 - 1, 2**220, (2**36 - 1) * (2**220)]
    a = [x for x in ev for _ in ev] + [r.getrandbits(256) for _ in range(6000)]
    b = [y for _ in ev for y in ev] + [r.getrandbits(256) for _ in range(6000)]
    got = dev.fp(E.OP_MUL_F64, a, b)
    assert got == [x * y % P for x, y in zip(a, b)]
    assert got == dev.fp(E.OP_MUL, a, b)
    got = dev.fp(E.OP_MUL_F64_CHAIN, a, b)
    assert got == [x * pow(y, 16, P) % P for x, y in zip(a, b)]


def test_affine_add_in_fp64_limb_form(dev):
    """batch_add's affine formula (main.c:378-386) with every product, square and difference in the FP64 limb form
    (fe6_mul / fe6_sqr / fe6_sub / fe6_norm): same canonical x, y as python ints, for random and edge operands"""
    import ecloop_b200 as E

    gx = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
    gy = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
    r = random.Random(137)
    xs = [v for v in edge_values() if v != gx] + [r.getrandbits(256) % P for _ in range(3000)]
    ys = list(reversed(edge_values()))[:len(edge_values()) - (1 if gx in edge_values() else 0)] + [r.getrandbits(256) % P for _ in range(3000)]
    ys = (ys + [0] * len(xs))[:len(xs)]
    want_x, want_y = [], []
    for x, y in zip(xs, ys):
        lam = (gy - y) * pow(gx - x, -1, P) % P
        rx = (lam * lam - x - gx) % P
        want_x.append(rx)
        want_y.append((lam * (x - rx) - y) % P)
    assert dev.fp(E.OP_AFFINE_F64_X, xs, ys) == want_x
    assert dev.fp(E.OP_AFFINE_F64_Y, xs, ys) == want_y
