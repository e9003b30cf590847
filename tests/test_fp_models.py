"""Limb-level python models of the two pieces of csrc/fp.cuh whose rare paths random GPU inputs never reach: the row
structure of the dedicated squaring (carry limbs may only ever hold carries) and the reduction with its branch-free final
correction (result >= p, carry out of the second fold). The models mirror the CUDA code operation for operation
(same rows, same limb indices, same masks) and are checked against python integers on adversarial values."""
import random

M = (1 << 32) - 1
P = 2**256 - 2**32 - 977
C0, P0, P1 = 977, 0xFFFFFC2F, 0xFFFFFFFE


def limbs(v, n=8):
    return [(v >> (32 * i)) & M for i in range(n)]


def value(l):
    return sum(x << (32 * i) for i, x in enumerate(l))


def mad_row(acc, s, b, alist):
    """fp_mad_row_c / row3 / row2 / row1: {mad.lo.cc, madc.hi.cc} pairs at limbs s, s+2, ... then addc into the limb above"""
    c = 0
    for n, a in enumerate(alist):
        p = a * b
        lo = acc[s + 2 * n] + (p & M) + c
        acc[s + 2 * n], c = lo & M, lo >> 32
        hi = acc[s + 2 * n + 1] + (p >> 32) + c
        acc[s + 2 * n + 1], c = hi & M, hi >> 32
    t = acc[s + 2 * len(alist)] + c
    assert t <= M, "a carry limb overflowed"
    acc[s + 2 * len(alist)] = t


def sqr_wide(a):
    """fp_sqr_wide: 13 rows of off-diagonal products, e + (o << 32), + the squares, + the off-diagonal part again"""
    e, o = [0] * 17, [0] * 17
    mad_row(e, 2, a[0], [a[2], a[4], a[6]])
    mad_row(e, 4, a[1], [a[3], a[5], a[7]])
    mad_row(e, 6, a[2], [a[4], a[6]])
    mad_row(e, 8, a[3], [a[5], a[7]])
    mad_row(e, 10, a[4], [a[6]])
    mad_row(e, 12, a[5], [a[7]])
    mad_row(o, 0, a[0], [a[1], a[3], a[5], a[7]])
    mad_row(o, 2, a[1], [a[2], a[4], a[6]])
    mad_row(o, 4, a[2], [a[3], a[5], a[7]])
    mad_row(o, 6, a[3], [a[4], a[6]])
    mad_row(o, 8, a[4], [a[5], a[7]])
    mad_row(o, 10, a[5], [a[6]])
    mad_row(o, 12, a[6], [a[7]])
    assert e[0] == e[1] == 0 and e[15] == e[16] == o[15] == o[16] == 0
    m = [0, o[0]] + [0] * 14
    c = 0
    for k in range(2, 16):
        t = e[k] + o[k - 1] + c
        m[k], c = t & M, t >> 32
    assert c == 0
    d, c = [0] * 16, 0
    for i in range(8):
        p = a[i] * a[i]
        lo = m[2 * i] + (p & M) + c
        d[2 * i], c = lo & M, lo >> 32
        hi = m[2 * i + 1] + (p >> 32) + c
        d[2 * i + 1], c = hi & M, hi >> 32
    assert c == 0
    t, c = [d[0]] + [0] * 15, 0
    for k in range(1, 16):
        x = d[k] + m[k] + c
        t[k], c = x & M, x >> 32
    assert c == 0
    return t


def reduce512(t, branchfree):
    """fp_reduce512_t: fold hi * (2^32 + 977) twice, then the final correction (branchy or fe_fix_branchfree)"""
    lo, hi = value(t[:8]), value(t[8:])
    a = lo + hi * C0 + (hi << 32)  # < 2^321
    r = (a & (2**256 - 1)) + (a >> 256) * C0 + ((a >> 256) << 32)
    cy, r = r >> 256, r & (2**256 - 1)
    assert cy in (0, 1)
    v = limbs(r)
    if branchfree:
        allhi = v[7] & v[6] & v[5] & v[4] & v[3] & v[2]
        ge = int(allhi == M and ((v[1] << 32) | v[0]) >= ((P1 << 32) | P0))
        assert not (ge and cy)
        m = (-(ge | cy)) & M
        add = (m & C0) | ((m & 1) << 32)
        r = (r + add) & (2**256 - 1)  # the 8-limb ripple add, carry out dropped
    else:
        if cy:
            r = (r + 2**32 + C0) & (2**256 - 1)
        if r >= P:
            r -= P
    return r


def adversarial():
    r = random.Random(5)
    vals = [0, 1, 2, P - 1, P - 2, P, P + 1, 2**256 - 1, 2**255, 2**256 - 2**32, 2**128 - 1, (P - 1) // 2, (P + 1) // 2, 0x1000003D1]
    vals += [2**256 - 1 - r.getrandbits(40) for _ in range(200)] + [P - r.getrandbits(34) for _ in range(200)]
    vals += [r.getrandbits(256) for _ in range(3000)]
    return vals


def test_squaring_rows():
    for v in adversarial():
        assert value(sqr_wide(limbs(v))) == v * v


def test_reduction_and_branchfree_correction():
    vals = adversarial()
    r = random.Random(6)
    # products and squares of adversarial values, plus 512-bit inputs built to land in [p, 2^256) and past 2^256 after the folds
    wide = [a * b for a, b in zip(vals, reversed(vals))] + [v * v for v in vals]
    CC = 2**32 + C0
    for k in range(2000):
        # the value the two folds should produce BEFORE the final correction: in [p, 2^256) or past 2^256
        target = r.choice([P + r.getrandbits(20), 2**256 - 1 - r.getrandbits(10), P, 2**256 + r.getrandbits(30)])
        a_hi = r.getrandbits(31)                 # what the first fold leaves above 2^256
        a = (a_hi << 256) + (target - a_hi * CC)  # second fold: a_lo + a_hi * C = target exactly
        hi = a // CC - r.getrandbits(8)          # first fold: lo + hi * C = a exactly
        lo = a - hi * CC
        assert 0 <= lo < 2**256 and 0 <= hi < 2**256
        wide.append((hi << 256) | lo)
    hit_ge = hit_cy = 0
    for w in wide:
        t = limbs(w, 16)
        lo, hi = value(t[:8]), value(t[8:])
        a = lo + hi * C0 + (hi << 32)
        pre = (a & (2**256 - 1)) + (a >> 256) * C0 + ((a >> 256) << 32)
        hit_cy += pre >> 256
        hit_ge += (pre >> 256) == 0 and pre >= P
        assert reduce512(t, True) == w % P == reduce512(t, False)
    assert hit_ge > 50 and hit_cy > 50  # both rare paths were really exercised
