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
