"""Exact integer model of the FP64-pipe field multiplication in ecloop_b200/csrc/fp64mul.cuh (44-bit limbs held as
doubles): every intermediate is checked against the 53-bit mantissa and the binade of the biased accumulator."""
import random
P = 2**256 - 2**32 - 977
L = 44; B = 2**96; M = 2**40 + 250112
assert (2**264) % P == M
def fma_rz_bias(x, y, t):   # t multiple of 2^44 in [2^96, 2^97): result truncated to multiples of 2^44
    r = t + ((x*y) >> L << L)
    assert 2**96 <= r < 2**97, "binade"
    return r
def ck53(v): assert 0 <= v < 2**53, v; return v
def split(x, y):            # one product through the bias trick: (H, Lo)
    t = fma_rz_bias(x, y, B); return (t - B) >> L, x*y - (t - B)
def mul6(a, b):
    hi = [0]*11; lo = [0]*11
    for k in range(11):
        t = B; s = 0
        for i in range(6):
            j = k - i
            if 0 <= j < 6:
                tn = fma_rz_bias(a[i], b[j], t); d = t - tn
                r = a[i]*b[j] + d; assert 0 <= r < 2**L
                s = ck53(s + r); t = tn
        hi[k] = t - B; lo[k] = s
    c = [0]*12
    for k in range(12):
        c[k] = ck53((lo[k] if k < 11 else 0) + ((hi[k-1] >> L) if k >= 1 else 0))
    for k in range(6, 12):
        H, Lo = split(c[k], M)
        c[k-6] = ck53(c[k-6] + Lo)
        if k < 11: c[k-5] = ck53(c[k-5] + H)
        else: c6b = H
    H, Lo = split(c6b, M); c[0] = ck53(c[0] + Lo); c[1] = ck53(c[1] + H)
    for k in range(5):
        q = c[k] >> L; c[k] -= q << L; c[k+1] = ck53(c[k+1] + q)
    q = c[5] >> L; c[5] -= q << L
    c[0] = ck53(c[0] + q*M)
    q = c[0] >> L; c[0] -= q << L; c[1] = ck53(c[1] + q)
    out = c[:6]
    assert all(v < 2**45 for v in out), out
    return out
def to6(x): return [(x >> (L*i)) & (2**L-1) for i in range(6)]
def val(c): return sum(v << (L*i) for i, v in enumerate(c))
random.seed(1)
worst = 0
for it in range(20000):
    if it < 200:
        a = random.choice([0, 1, P-1, P-2, 2**256-1, 2**255, 2**264-1]); b = random.choice([0, 1, P-1, 2**256-1, 2**264-1, random.getrandbits(256)])
        A, Bv = to6(a), to6(b)
    else:
        A = [random.getrandbits(44) + random.choice([0, 0, 34]) for _ in range(6)]; Bv = [random.getrandbits(44) + random.choice([0, 0, 34]) for _ in range(6)]
        a, b = val(A), val(Bv)
    r = mul6(A, Bv)
    assert val(r) % P == (a*b) % P, it
    worst = max(worst, val(r).bit_length())
    # chain: result feeds another mul
    r2 = mul6(r, r); assert val(r2) % P == (a*b)**2 % P
print("ok, max result bits", worst)


# ---------------------------------------------------------------- round-2 groundwork: the other field operations in the
# same limb form, with the bounds a fused kernel has to respect (nothing below is on the GPU yet)
NORMAL = 2**44 + 64          # limb bound of a multiplication result ("normal")
def carry_norm(c):
    """the carry section of mul6 on its own: any limbs < 2^52 -> normal limbs, value preserved mod p"""
    c = list(c)
    for k in range(5):
        q = c[k] >> L; c[k] -= q << L; c[k+1] = ck53(c[k+1] + q)
    q = c[5] >> L; c[5] -= q << L
    c[0] = ck53(c[0] + q*M)
    q = c[0] >> L; c[0] -= q << L; c[1] = ck53(c[1] + q)
    assert all(v < NORMAL for v in c), c
    return c
def bias_limbs(min_limb_bits):
    """a multiple of p whose limbs are all >= 2^min_limb_bits (so that a - b + bias has no negative limb)"""
    assert min_limb_bits + 1 >= L
    need = 1 << (min_limb_bits + 1)             # borrowed into every low limb from the limb above it
    m = 1
    while True:
        v = m * P
        k = [(v >> (L*i)) & (2**L-1) for i in range(5)] + [v >> (L*5)]
        for i in range(5):
            k[i] += need; k[i+1] -= need >> L
        if all(2**min_limb_bits <= x < 2**52 for x in k):
            assert val(k) == v
            return k
        m *= 2
def sub6(a, b, bias):
    r = [a[i] - b[i] + bias[i] for i in range(6)]
    assert all(0 <= x < 2**52 for x in r), r
    return r
def sqr6(a):
    """21 products instead of 36: off-diagonal terms enter once with a doubled operand (limbs < 2^46 keep every column
    below the accumulator's binade)"""
    a2 = [2*x for x in a]
    hi = [0]*11; lo = [0]*11
    for k in range(11):
        t = B; s = 0
        for i in range(6):
            j = k - i
            if 0 <= j < 6 and i <= j:
                x, y = (a[i], a[j]) if i == j else (a2[i], a[j])
                tn = fma_rz_bias(x, y, t); r = x*y + (t - tn); assert 0 <= r < 2**L
                s = ck53(s + r); t = tn
        hi[k] = t - B; lo[k] = s
    c = [ck53((lo[k] if k < 11 else 0) + ((hi[k-1] >> L) if k >= 1 else 0)) for k in range(12)]
    for k in range(6, 12):
        H, Lo = split(c[k], M); c[k-6] = ck53(c[k-6] + Lo)
        if k < 11: c[k-5] = ck53(c[k-5] + H)
        else: c6b = H
    H, Lo = split(c6b, M); c[0] = ck53(c[0] + Lo); c[1] = ck53(c[1] + H)
    return carry_norm(c[:6])
def affine_add6(px, py, qx, qy, inv, bias):
    """batch_add's body (main.c:378-386) in limb form: which intermediates need a carry pass before the next product.
    One operand of a product may be 'wide' (a raw difference, limbs < 2^47.3) if the other is normal."""
    dy = sub6(qy, py, bias)                    # wide
    lam = mul6(dy, inv)                        # wide x normal -> normal
    l2 = sqr6(lam)
    rx = carry_norm(sub6(sub6(l2, px, bias), qx, bias))     # two differences stacked: normalise before reuse
    ry = carry_norm(sub6(mul6(sub6(px, rx, bias), lam), py, bias))
    return rx, ry
if True:  # run as a script and by tests/test_host.py
    bias = bias_limbs(45)
    random.seed(7)
    def ec_add(p1, p2):
        lam = (p2[1] - p1[1]) * pow(p2[0] - p1[0], -1, P) % P
        x = (lam*lam - p1[0] - p2[0]) % P
        return x, (lam*(p1[0] - x) - p1[1]) % P
    G = (0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798, 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8)
    pt = ec_add(G, (0xC6047F9441ED7D6D3045406E95C07CD85C778E4B8CEF3CA7ABAC09B95C709EE5, 0x1AE168FEA63DC339A3C58419466CEAEEF7F632653266D0E1236431A950CFE52A))
    cur = pt
    for it in range(300):
        inv = pow(G[0] - cur[0], -1, P)
        rx, ry = affine_add6(to6(cur[0]), to6(cur[1]), to6(G[0]), to6(G[1]), to6(inv), bias)
        want = ec_add(cur, G)
        assert (val(rx) % P, val(ry) % P) == want, it
        x = random.getrandbits(256); assert val(sqr6(to6(x))) % P == x*x % P
        cur = want
    print("ok, affine add chain in limb form (bias limbs >= 2^45:", [hex(v) for v in bias][:2], "...)")
