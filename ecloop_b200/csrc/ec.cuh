// ec.cuh — secp256k1 group law on the device: Jacobian doubling / mixed addition, the fixed-base window table
// ("gtable") scalar multiply, scalar bookkeeping mod n, and the affine addition with a shared inverse that the
// add kernel is built on.
//
// Replaces lib/ecc.c:546-929 of the reference (pe, _ec_jacobi_add1/_dbl1/_rdc1, ec_gtable_init/ec_gtable_mul)
// and ctx_precompute_gpoints (main.c:219-246). The reference uses homogeneous projective coordinates and a
// w=14 table; any correct algorithm gives the same canonical affine (x, y), so we use Jacobian coordinates
// (a = 0 formulas) and a much wider window (W = 24 by default: 11 windows, a 10.7 GB table in HBM).
#pragma once
#include "fp.cuh"

// setup-path multiplies are not inlined (compile time / code size); the add kernel's hot loop uses fe_mul.
#define FE_MUL_S(a, b) fe_mul_noinline((a), (b))
#define FE_SQR_S(a) fe_mul_noinline((a), (a))

struct jac {
  fe x, y, z;
};

// Window width of the fixed-base table d * 2^(W w) * G. The reference uses W = 14 (lib/ecc.c:876: 19 windows, 30 MB);
// with 180 GB of HBM the width is a pure trade of table bytes against additions per key:
//   W = 16: 16 windows,  67 MB (L2-resident)      W = 22: 12 windows, 2.95 GB
//   W = 24: 11 windows, 10.7 GB                   W = 26: 10 windows, 38.9 GB
// One 64 B gather per window and key from a table far beyond L2 is cheap next to the ~2000 instructions of the
// addition it feeds; the table is built on the device in tens of milliseconds (kernels.cuh gtab_fill).
#ifndef GTAB_W
#define GTAB_W 24
#endif
#define GTAB_WINDOWS ((256 + GTAB_W - 1) / GTAB_W)
#define GTAB_TOP_BITS (256 - GTAB_W * (GTAB_WINDOWS - 1))  // bits of the top window
#define GTAB_STRIDE (1u << GTAB_W)                          // slots per window: slot d - 1 holds d * 2^(W w) * G
#define GTAB_ENTRIES ((size_t)(GTAB_WINDOWS - 1) * GTAB_STRIDE + ((size_t)1 << GTAB_TOP_BITS))
#define GTAB_CHUNK 256u  // entries one thread of the table builder fills: (c*256 + k) * B = c*256*B + k*B
static_assert(GTAB_W >= 9 && GTAB_W <= 28 && GTAB_TOP_BITS >= 8, "window width out of range");

// digit w of the scalar k: bits [W w, W w + W)
__device__ __forceinline__ u32 gtab_digit(const fe &k, int w) {
  const int lo = w * GTAB_W, limb = lo >> 5, sh = lo & 31;
  u32 v = k.v[limb] >> sh;
  if (sh + GTAB_W > 32 && limb < 7) v |= k.v[limb + 1] << (32 - sh);
  return v & ((1u << GTAB_W) - 1u);
}

__device__ __forceinline__ fe fe_dbl(const fe &a) { return fe_add(a, a); }

// r = 2p, Jacobian, a = 0 (2M + 5S). p must not be infinity; y = 0 cannot happen on secp256k1 (odd order).
static __device__ __noinline__ void jac_dbl(jac &r, const jac &p) {
  fe a = FE_SQR_S(p.x);
  fe b = FE_SQR_S(p.y);
  fe c = FE_SQR_S(b);
  fe t = fe_add(p.x, b);
  t = FE_SQR_S(t);
  t = fe_sub(fe_sub(t, a), c);
  fe d = fe_dbl(t);                // D = 2((X+B)^2 - A - C)
  fe e = fe_add(fe_dbl(a), a);     // E = 3A
  fe f = FE_SQR_S(e);              // F = E^2
  fe x3 = fe_sub(fe_sub(f, d), d);
  fe c8 = fe_dbl(fe_dbl(fe_dbl(c)));
  fe y3 = fe_sub(FE_MUL_S(e, fe_sub(d, x3)), c8);
  fe z3 = fe_dbl(FE_MUL_S(p.y, p.z));
  r.x = x3, r.y = y3, r.z = z3;
}

// r = p + (qx, qy), q affine (8M + 3S). Caller guarantees p != +-q and neither is infinity.
static __device__ __noinline__ void jac_madd(jac &r, const jac &p, const fe &qx, const fe &qy) {
  fe z2 = FE_SQR_S(p.z);
  fe u2 = FE_MUL_S(qx, z2);
  fe s2 = FE_MUL_S(FE_MUL_S(qy, z2), p.z);
  fe h = fe_sub(u2, p.x);
  fe rr = fe_sub(s2, p.y);
  fe h2 = FE_SQR_S(h);
  fe h3 = FE_MUL_S(h2, h);
  fe v = FE_MUL_S(p.x, h2);
  fe x3 = fe_sub(fe_sub(fe_sub(FE_SQR_S(rr), h3), v), v);
  fe y3 = fe_sub(FE_MUL_S(rr, fe_sub(v, x3)), FE_MUL_S(p.y, h3));
  fe z3 = FE_MUL_S(p.z, h);
  r.x = x3, r.y = y3, r.z = z3;
}

__device__ __forceinline__ void jac_to_affine(fe &x, fe &y, const jac &p) {
  fe zi = fe_inv(p.z);
  fe zi2 = FE_SQR_S(zi);
  x = FE_MUL_S(p.x, zi2);
  y = FE_MUL_S(p.y, FE_MUL_S(zi2, zi));
}

// ---------------------------------------------------------------- scalars mod n (lib/ecc.c:166-265)

__device__ __forceinline__ bool sc_ge_n(const u32 x[8]) {
  const u32 N[8] = {0xd0364141u, 0xbfd25e8cu, 0xaf48a03bu, 0xbaaedce6u, 0xfffffffeu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    if (x[i] != N[i]) return x[i] > N[i];
  }
  return true;
}

// k = k0 + j * step (mod n), canonical; k0, step < 2^256 (the start-point bookkeeping the reference does on the
// host with fe_modn_add_stride, ecc.c:255-260)
static __device__ __noinline__ fe sc_muladd_small(const fe &k0, const fe &step, u32 j) {
  const u32 CN[5] = {0x2fc9bebfu, 0x402da173u, 0x50b75fc4u, 0x45512319u, 0x1u};  // 2^256 - n
  const u32 N[8] = {0xd0364141u, 0xbfd25e8cu, 0xaf48a03bu, 0xbaaedce6u, 0xfffffffeu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
  u32 x[10];
  u64 c = 0;
  for (int i = 0; i < 8; ++i) {
    c += (u64)step.v[i] * j + k0.v[i];
    x[i] = (u32)c;
    c >>= 32;
  }
  x[8] = (u32)c;
  x[9] = (u32)(c >> 32);
  while (x[8] | x[9]) {
    const u32 h0 = x[8], h1 = x[9];
    x[8] = x[9] = 0;
    c = 0;
    for (int i = 0; i < 10; ++i) {  // x += h0 * CN
      const u64 t = (u64)x[i] + c + (i < 5 ? (u64)h0 * CN[i] : 0ull);
      x[i] = (u32)t;
      c = t >> 32;
    }
    c = 0;
    for (int i = 1; i < 10; ++i) {  // x += (h1 * CN) << 32
      const u64 t = (u64)x[i] + c + (i - 1 < 5 ? (u64)h1 * CN[i - 1] : 0ull);
      x[i] = (u32)t;
      c = t >> 32;
    }
  }
  if (sc_ge_n(x)) {
    u64 bw = 0;
    for (int i = 0; i < 8; ++i) {
      const u64 t = (u64)x[i] - N[i] - bw;
      x[i] = (u32)t;
      bw = (t >> 32) & 1;
    }
  }
  fe r;
  for (int i = 0; i < 8; ++i) r.v[i] = x[i];
  return r;
}

// ---------------------------------------------------------------- window table scalar multiply

__device__ __forceinline__ void gtab_load(fe &x, fe &y, const uint4 *__restrict__ gtab, u32 entry) {
  const uint4 *p = gtab + (size_t)entry * 4;
  const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
  x.v[0] = a.x, x.v[1] = a.y, x.v[2] = a.z, x.v[3] = a.w, x.v[4] = b.x, x.v[5] = b.y, x.v[6] = b.z, x.v[7] = b.w;
  y.v[0] = c.x, y.v[1] = c.y, y.v[2] = c.z, y.v[3] = c.w, y.v[4] = d.x, y.v[5] = d.y, y.v[6] = d.z, y.v[7] = d.w;
}

// acc = k*G in Jacobian coordinates (ec_gtable_mul, ecc.c:907-929: one table point per non-zero window).
// Returns false when k = 0 (mod n) (point at infinity; the reference poisons its batch there).
// Valid for any k < 2^256: partial sums stay below n until the last window, so the mixed addition never
// meets P = +-Q (DESIGN.md "degenerate cases").
static __device__ __noinline__ bool gtab_mul(jac &acc, const fe &k, const uint4 *__restrict__ gtab) {
  bool have = false;
  u32 dig[GTAB_WINDOWS];
#pragma unroll
  for (int w = 0; w < GTAB_WINDOWS; ++w) dig[w] = gtab_digit(k, w);
#pragma unroll 1
  for (int w = 0; w < GTAB_WINDOWS; ++w) {
    const u32 d = dig[w];
    if (d == 0) continue;
    fe qx, qy;
    gtab_load(qx, qy, gtab, (u32)w * GTAB_STRIDE + d - 1);
    if (!have) {
      acc.x = qx, acc.y = qy, acc.z = fe_one();
      have = true;
    } else {
      // k = n lands on -acc at the top window; the result is infinity
      jac t;
      jac_madd(t, acc, qx, qy);
      if (fe_is_zero(t.z)) return false;
      acc = t;
    }
  }
  return have;
}

// ---------------------------------------------------------------- affine addition with a shared inverse
// (the body of batch_add, main.c:378-386): inv = 1/(qx - px)
// inv may be any representative of the inverse below 2^256 (fe_mul_nc); px, py, qx, qy canonical; rx, ry canonical
__device__ __forceinline__ void affine_add_inv(fe &rx, fe &ry, const fe &px, const fe &py, const fe &qx, const fe &qy,
                                               const fe &inv) {
  const fe lam = fe_mul_nc(fe_sub(qy, py), inv);  // only multiplied and squared below
  rx = fe_sub(fe_sub(fe_sqr(lam), px), qx);
  ry = fe_sub(fe_mul(lam, fe_sub(px, rx)), py);
}
