/* ecl_oracle.c — TEST INFRASTRUCTURE ONLY (see ecl_oracle.h). Parity status: PINNED (tests/test_oracle.py).
 *
 * A CPU restatement of the reference's hot path written from its behaviour, not its code: every routine
 * produces the canonical value the reference produces (field elements in [0,p), affine points, digest
 * words in h160_t order), but via the simplest algorithm that is obviously correct (generic wide multiply
 * + fold, Jacobian double-and-add, table-driven RIPEMD-160), so it is an independent check.
 * Citations are file:line under /root/reference.
 */
#include "ecl_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;

/* p = 2^256 - 2^32 - 977, n = group order (lib/ecc.c:32-33) */
static const u64 P[4] = {0xfffffffefffffc2fULL, 0xffffffffffffffffULL, 0xffffffffffffffffULL, 0xffffffffffffffffULL};
static const u64 N[4] = {0xbfd25e8cd0364141ULL, 0xbaaedce6af48a03bULL, 0xfffffffffffffffeULL, 0xffffffffffffffffULL};
/* 2^256 - p and 2^256 - n */
static const u64 CP[3] = {0x1000003d1ULL, 0, 0};
static const u64 CN[3] = {0x402da1732fc9bebfULL, 0x4551231950b75fc4ULL, 0x1ULL};
/* endomorphism constants lambda, lambda^2 (mod n), beta, beta^2 (mod p) (lib/ecc.c:36-39) */
static const u64 LAM1[4] = {0xdf02967c1b23bd72ULL, 0x122e22ea20816678ULL, 0xa5261c028812645aULL, 0x5363ad4cc05c30e0ULL};
static const u64 LAM2[4] = {0xe0cfc810b51283ceULL, 0xa880b9fc8ec739c2ULL, 0x5ad9e3fd77ed9ba4ULL, 0xac9c52b33fa3cf1fULL};
static const u64 BETA1[4] = {0xc1396c28719501eeULL, 0x9cf0497512f58995ULL, 0x6e64479eac3434e9ULL, 0x7ae96a2b657c0710ULL};
static const u64 BETA2[4] = {0x3ec693d68e6afa40ULL, 0x630fb68aed0a766aULL, 0x919bb86153cbcb16ULL, 0x851695d49a83f8efULL};
/* generator (lib/ecc.c:550-554) */
static const u64 GX[4] = {0x59f2815b16f81798ULL, 0x029bfcdb2dce28d9ULL, 0x55a06295ce870b07ULL, 0x79be667ef9dcbbacULL};
static const u64 GY[4] = {0x9c47d08ffb10d4b8ULL, 0xfd17b448a6855419ULL, 0x5da4fbfc0e1108a8ULL, 0x483ada7726a3c465ULL};

/* ---------------------------------------------------------------- generic multi-limb helpers */

static int cmp4(const u64 *a, const u64 *b) {
  for (int i = 3; i >= 0; --i)
    if (a[i] != b[i]) return a[i] > b[i] ? 1 : -1;
  return 0;
}
static int is_zero4(const u64 *a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static u64 add_n(u64 *r, const u64 *a, const u64 *b, int n) {
  u64 c = 0;
  for (int i = 0; i < n; ++i) {
    u128 t = (u128)a[i] + b[i] + c;
    r[i] = (u64)t;
    c = (u64)(t >> 64);
  }
  return c;
}
static u64 sub_n(u64 *r, const u64 *a, const u64 *b, int n) {
  u64 bw = 0;
  for (int i = 0; i < n; ++i) {
    u128 t = (u128)a[i] - b[i] - bw;
    r[i] = (u64)t;
    bw = (u64)(t >> 64) & 1;
  }
  return bw;
}
/* r[0..na+nb) = a * b */
static void mul_wide(u64 *r, const u64 *a, int na, const u64 *b, int nb) {
  memset(r, 0, sizeof(u64) * (size_t)(na + nb));
  for (int i = 0; i < na; ++i) {
    u64 c = 0;
    for (int j = 0; j < nb; ++j) {
      u128 t = (u128)a[i] * b[j] + r[i + j] + c;
      r[i + j] = (u64)t;
      c = (u64)(t >> 64);
    }
    r[i + nb] = c;
  }
}
/* x (8 limbs) mod m, where m = 2^256 - c and c has nc limbs: fold hi*c + lo until hi == 0, then subtract */
static void reduce8(u64 r[4], const u64 x_in[8], const u64 m[4], const u64 *c, int nc) {
  u64 x[8];
  memcpy(x, x_in, sizeof x);
  while (x[4] | x[5] | x[6] | x[7]) {
    u64 t[8] = {0}, prod[8] = {0};
    mul_wide(prod, x + 4, 4, c, nc); /* 4+nc <= 7 limbs */
    memcpy(t, x, 4 * sizeof(u64));   /* lo */
    add_n(x, t, prod, 8);
  }
  while (cmp4(x, m) >= 0) sub_n(x, x, m, 4);
  memcpy(r, x, 4 * sizeof(u64));
}

/* ---------------------------------------------------------------- Fp  (lib/ecc.c:269-520) */

void orc_fp_add(orc_fe r, const orc_fe a, const orc_fe b) { /* ecc.c:292-305, canonical for canonical inputs */
  u64 t[8] = {0};
  t[4] = add_n(t, a, b, 4);
  reduce8(r, t, P, CP, 1);
}
void orc_fp_sub(orc_fe r, const orc_fe a, const orc_fe b) { /* ecc.c:277-290 */
  u64 t[4];
  if (sub_n(t, a, b, 4)) add_n(t, t, P, 4);
  memcpy(r, t, sizeof t);
}
void orc_fp_neg(orc_fe r, const orc_fe a) { /* ecc.c:269-275 (neg(0) = p there; we keep that quirk) */
  u64 t[4];
  sub_n(t, P, a, 4);
  memcpy(r, t, sizeof t);
}
void orc_fp_mul(orc_fe r, const orc_fe a, const orc_fe b) { /* ecc.c:307-347 */
  /* column-wise 4x4 product, then fold twice with 2^256 = 0x1000003D1 (mod p), then one exact
   * reduction; unlike ecc.c:341-344 the carry out of the second fold is not dropped (it cannot be
   * set for canonical operands except with probability ~2^-159). */
  u64 t[8];
  u128 acc = 0;
  u64 acchi = 0; /* 192-bit column accumulator: acchi:acc */
  for (int k = 0; k < 7; ++k) {
    int lo = k < 4 ? 0 : k - 3, hi = k < 4 ? k : 3;
    for (int i = lo; i <= hi; ++i) {
      u128 pr = (u128)a[i] * b[k - i];
      acc += pr;
      acchi += acc < pr;
    }
    t[k] = (u64)acc;
    acc = (acc >> 64) | ((u128)acchi << 64);
    acchi = 0;
  }
  t[7] = (u64)acc;
  const u64 C = 0x1000003d1ULL;
  u128 c = 0;
  u64 x[5];
  for (int i = 0; i < 4; ++i) {
    c += (u128)t[4 + i] * C + t[i];
    x[i] = (u64)c;
    c >>= 64;
  }
  x[4] = (u64)c; /* < 2^34 */
  c = (u128)x[4] * C;
  u64 carry = 0;
  for (int i = 0; i < 4; ++i) {
    c += x[i];
    x[i] = (u64)c;
    c >>= 64;
  }
  carry = (u64)c;
  u64 t8[8] = {x[0], x[1], x[2], x[3], carry, 0, 0, 0};
  reduce8(r, t8, P, CP, 1);
}
void orc_fp_sqr(orc_fe r, const orc_fe a) { orc_fp_mul(r, a, a); } /* ecc.c:349-444 */
void orc_fp_inv(orc_fe r, const orc_fe a) {                        /* ecc.c:463-520: a^(p-2); inv(0) = 0 */
  u64 e[4], acc[4] = {1, 0, 0, 0}, base[4];
  memcpy(e, P, sizeof e);
  e[0] -= 2;
  memcpy(base, a, sizeof base);
  for (int i = 0; i < 256; ++i) {
    if ((e[i / 64] >> (i % 64)) & 1) orc_fp_mul(acc, acc, base);
    orc_fp_sqr(base, base);
  }
  memcpy(r, acc, sizeof acc);
}
void orc_fp_grpinv(orc_fe *r, uint32_t n) { /* ecc.c:522-540 (Montgomery's trick) */
  if (n == 0) return;
  orc_fe *pre = (orc_fe *)malloc(sizeof(orc_fe) * n);
  memcpy(pre[0], r[0], sizeof(orc_fe));
  for (uint32_t i = 1; i < n; ++i) orc_fp_mul(pre[i], pre[i - 1], r[i]);
  orc_fe inv, t;
  orc_fp_inv(inv, pre[n - 1]);
  for (uint32_t i = n - 1; i > 0; --i) {
    orc_fp_mul(t, inv, pre[i - 1]);
    orc_fp_mul(inv, inv, r[i]);
    memcpy(r[i], t, sizeof t);
  }
  memcpy(r[0], inv, sizeof inv);
  free(pre);
}

/* ---------------------------------------------------------------- Fn  (lib/ecc.c:166-265) */

void orc_fn_add(orc_fe r, const orc_fe a, const orc_fe b) { /* ecc.c:174-187: subtract n only on 2^256 carry */
  u64 t[4];
  if (add_n(t, a, b, 4)) sub_n(t, t, N, 4);
  memcpy(r, t, sizeof t);
}
void orc_fn_sub(orc_fe r, const orc_fe a, const orc_fe b) { /* ecc.c:189-202 */
  u64 t[4];
  if (sub_n(t, a, b, 4)) add_n(t, t, N, 4);
  memcpy(r, t, sizeof t);
}
void orc_fn_neg(orc_fe r, const orc_fe a) { /* ecc.c:166-172 */
  u64 t[4];
  sub_n(t, N, a, 4);
  memcpy(r, t, sizeof t);
}
void orc_fn_mul(orc_fe r, const orc_fe a, const orc_fe b) { /* ecc.c:211-253 (Montgomery there; same value) */
  u64 t[8];
  mul_wide(t, a, 4, b, 4);
  reduce8(r, t, N, CN, 3);
}
void orc_fn_add_stride(orc_fe r, const orc_fe base, const orc_fe stride, uint64_t offset) { /* ecc.c:255-260 */
  u64 t[4] = {offset, 0, 0, 0};
  orc_fn_mul(t, t, stride);
  orc_fn_add(r, t, base);
}
void orc_fn_from_hex(orc_fe r, const char *hex) { /* ecc.c:81-95,262-265: right-to-left, skip non-hex */
  u64 t[4] = {0, 0, 0, 0};
  int cnt = 0;
  for (long i = (long)strlen(hex) - 1; i >= 0; --i) {
    int ch = (unsigned char)hex[i], v;
    if (ch >= '0' && ch <= '9') v = ch - '0';
    else if (ch >= 'a' && ch <= 'f') v = ch - 'a' + 10;
    else if (ch >= 'A' && ch <= 'F') v = ch - 'A' + 10;
    else continue;
    if (cnt < 64) t[cnt / 16] |= (u64)v << (4 * (cnt % 16));
    cnt++;
  }
  if (cmp4(t, N) >= 0) orc_fn_sub(t, t, N);
  memcpy(r, t, sizeof t);
}

/* ---------------------------------------------------------------- group law (lib/ecc.c:546-929)
 * Jacobian (X/Z^2, Y/Z^3), a = 0. The reference uses homogeneous projective formulas; affine results
 * are the same group elements, canonically reduced. */

typedef struct {
  orc_fe x, y, z;
  int inf;
} jac;

static void jac_dbl(jac *r, const jac *p) {
  if (p->inf || is_zero4(p->y)) {
    r->inf = 1;
    return;
  }
  orc_fe a, b, c, d, e, f, t;
  orc_fp_sqr(a, p->x);    /* A = X^2 */
  orc_fp_sqr(b, p->y);    /* B = Y^2 */
  orc_fp_sqr(c, b);       /* C = B^2 */
  orc_fp_add(t, p->x, b); /* D = 2((X+B)^2 - A - C) */
  orc_fp_sqr(t, t);
  orc_fp_sub(t, t, a);
  orc_fp_sub(t, t, c);
  orc_fp_add(d, t, t);
  orc_fp_add(e, a, a); /* E = 3A */
  orc_fp_add(e, e, a);
  orc_fp_sqr(f, e); /* F = E^2 */
  orc_fe x3, y3, z3;
  orc_fp_sub(x3, f, d);
  orc_fp_sub(x3, x3, d);
  orc_fp_sub(t, d, x3);
  orc_fp_mul(y3, e, t);
  orc_fp_add(c, c, c);
  orc_fp_add(c, c, c);
  orc_fp_add(c, c, c); /* 8C */
  orc_fp_sub(y3, y3, c);
  orc_fp_mul(z3, p->y, p->z);
  orc_fp_add(z3, z3, z3);
  memcpy(r->x, x3, sizeof x3);
  memcpy(r->y, y3, sizeof y3);
  memcpy(r->z, z3, sizeof z3);
  r->inf = 0;
}

/* r = p + (qx,qy) with q affine, not infinity */
static void jac_add_affine(jac *r, const jac *p, const orc_fe qx, const orc_fe qy) {
  if (p->inf) {
    memcpy(r->x, qx, sizeof(orc_fe));
    memcpy(r->y, qy, sizeof(orc_fe));
    memset(r->z, 0, sizeof(orc_fe));
    r->z[0] = 1;
    r->inf = 0;
    return;
  }
  orc_fe z2, u2, s2, h, rr, h2, h3, v, t;
  orc_fp_sqr(z2, p->z);
  orc_fp_mul(u2, qx, z2);
  orc_fp_mul(s2, qy, z2);
  orc_fp_mul(s2, s2, p->z);
  orc_fp_sub(h, u2, p->x);
  orc_fp_sub(rr, s2, p->y);
  if (is_zero4(h)) {
    if (is_zero4(rr)) {
      jac_dbl(r, p);
    } else {
      r->inf = 1;
    }
    return;
  }
  orc_fp_sqr(h2, h);
  orc_fp_mul(h3, h2, h);
  orc_fp_mul(v, p->x, h2);
  orc_fe x3, y3, z3;
  orc_fp_sqr(x3, rr);
  orc_fp_sub(x3, x3, h3);
  orc_fp_sub(x3, x3, v);
  orc_fp_sub(x3, x3, v);
  orc_fp_sub(t, v, x3);
  orc_fp_mul(y3, rr, t);
  orc_fp_mul(t, p->y, h3);
  orc_fp_sub(y3, y3, t);
  orc_fp_mul(z3, p->z, h);
  memcpy(r->x, x3, sizeof x3);
  memcpy(r->y, y3, sizeof y3);
  memcpy(r->z, z3, sizeof z3);
  r->inf = 0;
}

static int jac_to_affine(orc_fe x, orc_fe y, const jac *p) {
  if (p->inf || is_zero4(p->z)) {
    memset(x, 0, sizeof(orc_fe));
    memset(y, 0, sizeof(orc_fe));
    return 1;
  }
  orc_fe zi, zi2;
  orc_fp_inv(zi, p->z);
  orc_fp_sqr(zi2, zi);
  orc_fp_mul(x, p->x, zi2);
  orc_fp_mul(zi2, zi2, zi);
  orc_fp_mul(y, p->y, zi2);
  return 0;
}

/* 8-bit fixed-window table of G multiples: tab[w][d-1] = d * 256^w * G (affine), built once.
 * Stands in for ec_gtable_init/ec_gtable_mul (ecc.c:880-929) and ec_jacobi_mulrdc (ecc.c:821-853):
 * all of them compute k*G. */
static orc_fe (*g_tab)[255][2] = NULL;
static void gtab_init(void) {
  if (g_tab) return;
  g_tab = malloc(sizeof(orc_fe) * 2 * 255 * 32);
  orc_fe bx, by;
  memcpy(bx, GX, sizeof bx);
  memcpy(by, GY, sizeof by);
  for (int w = 0; w < 32; ++w) {
    jac acc;
    acc.inf = 1;
    for (int d = 1; d <= 255; ++d) {
      jac_add_affine(&acc, &acc, bx, by);
      jac_to_affine(g_tab[w][d - 1][0], g_tab[w][d - 1][1], &acc);
    }
    jac nb; /* next base = 256 * base */
    jac_add_affine(&nb, &acc, bx, by);
    jac_to_affine(bx, by, &nb);
  }
}

static void jac_mul_g(jac *r, const orc_fe k) {
  gtab_init();
  r->inf = 1;
  for (int w = 0; w < 32; ++w) {
    unsigned d = (unsigned)(k[w / 8] >> (8 * (w % 8))) & 0xff;
    if (d) jac_add_affine(r, r, g_tab[w][d - 1][0], g_tab[w][d - 1][1]);
  }
}

int orc_ec_mul_g(orc_fe x, orc_fe y, const orc_fe k) {
  jac r;
  jac_mul_g(&r, k);
  return jac_to_affine(x, y, &r);
}

int orc_ec_add(orc_fe rx, orc_fe ry, const orc_fe px, const orc_fe py, const orc_fe qx, const orc_fe qy) {
  jac p, r;
  memcpy(p.x, px, sizeof(orc_fe));
  memcpy(p.y, py, sizeof(orc_fe));
  memset(p.z, 0, sizeof(orc_fe));
  p.z[0] = 1;
  p.inf = 0;
  jac_add_affine(&r, &p, qx, qy);
  return jac_to_affine(rx, ry, &r);
}

/* ---------------------------------------------------------------- SHA-256 (lib/sha256.c:399-453) */

static const u32 SHA_K[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
static const u32 SHA_IV[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au,
                              0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};

static inline u32 rotr32(u32 x, int n) { return (x >> n) | (x << (32 - n)); }
static inline u32 rotl32(u32 x, int n) { return (x << n) | (x >> (32 - n)); }

/* starts from the IV each call and returns the 8 state words, like sha256_final (caller pads) */
void orc_sha256_blocks(uint32_t state[8], const uint8_t *data, size_t nblocks) {
  u32 h[8];
  memcpy(h, SHA_IV, sizeof h);
  for (size_t blk = 0; blk < nblocks; ++blk, data += 64) {
    u32 w[64];
    for (int i = 0; i < 16; ++i)
      w[i] = (u32)data[4 * i] << 24 | (u32)data[4 * i + 1] << 16 | (u32)data[4 * i + 2] << 8 | data[4 * i + 3];
    for (int i = 16; i < 64; ++i) {
      u32 s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
      u32 s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    u32 a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; ++i) {
      u32 t1 = hh + (rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25)) + ((e & f) ^ (~e & g)) + SHA_K[i] + w[i];
      u32 t2 = (rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
      hh = g, g = f, f = e, e = d + t1, d = c, c = b, b = a, a = t1 + t2;
    }
    h[0] += a, h[1] += b, h[2] += c, h[3] += d, h[4] += e, h[5] += f, h[6] += g, h[7] += hh;
  }
  memcpy(state, h, sizeof h);
}

/* ---------------------------------------------------------------- RIPEMD-160 (lib/rmd160s.c:122-336) */

static const u8 RMD_RL[80] = {0, 1, 2,  3,  4,  5,  6,  7, 8,  9, 10, 11, 12, 13, 14, 15, 7, 4,  13, 1,
                              10, 6, 15, 3,  12, 0,  9,  5, 2,  14, 11, 8,  3,  10, 14, 4,  9, 15, 8,  1,
                              2,  7, 0,  6,  13, 11, 5,  12, 1, 9,  11, 10, 0,  8,  12, 4,  13, 3,  7,  15,
                              14, 5, 6,  2,  4,  0,  5,  9, 7,  12, 2,  10, 14, 1,  3,  8,  11, 6,  15, 13};
static const u8 RMD_RR[80] = {5,  14, 7,  0, 9, 2,  11, 4,  13, 6,  15, 8,  1,  10, 3,  12, 6,  11, 3,  7,
                              0,  13, 5,  10, 14, 15, 8,  12, 4,  9,  1,  2,  15, 5,  1,  3,  7,  14, 6,  9,
                              11, 8,  12, 2, 10, 0,  4,  13, 8,  6,  4,  1,  3,  11, 15, 0,  5,  12, 2,  13,
                              9,  7,  10, 14, 12, 15, 10, 4, 1,  5,  8,  7,  6,  2,  13, 14, 0,  3,  9,  11};
static const u8 RMD_SL[80] = {11, 14, 15, 12, 5,  8,  7,  9,  11, 13, 14, 15, 6,  7,  9,  8,  7,  6,  8,  13,
                              11, 9,  7,  15, 7,  12, 15, 9,  11, 7,  13, 12, 11, 13, 6,  7,  14, 9,  13, 15,
                              14, 8,  13, 6,  5,  12, 7,  5,  11, 12, 14, 15, 14, 15, 9,  8,  9,  14, 5,  6,
                              8,  6,  5,  12, 9,  15, 5,  11, 6,  8,  13, 12, 5,  12, 13, 14, 11, 8,  5,  6};
static const u8 RMD_SR[80] = {8,  9,  9,  11, 13, 15, 15, 5,  7,  7,  8,  11, 14, 14, 12, 6,  9,  13, 15, 7,
                              12, 8,  9,  11, 7,  7,  12, 7,  6,  15, 13, 11, 9,  7,  15, 11, 8,  6,  6,  14,
                              12, 13, 5,  14, 13, 13, 7,  5,  15, 5,  8,  11, 14, 14, 6,  14, 6,  9,  12, 9,
                              12, 5,  15, 8,  8,  5,  12, 9,  12, 5,  14, 6,  8,  13, 6,  5,  15, 13, 11, 11};
static const u32 RMD_KL[5] = {0x00000000u, 0x5a827999u, 0x6ed9eba1u, 0x8f1bbcdcu, 0xa953fd4eu};
static const u32 RMD_KR[5] = {0x50a28be6u, 0x5c4dd124u, 0x6d703ef3u, 0x7a6d76e9u, 0x00000000u};

static u32 rmd_f(int j, u32 x, u32 y, u32 z) {
  switch (j) {
  case 0: return x ^ y ^ z;
  case 1: return (x & y) | (~x & z);
  case 2: return (x | ~y) ^ z;
  case 3: return (x & z) | (y & ~z);
  default: return x ^ (y | ~z);
  }
}

/* one compression from the standard IV; w = 16 little-endian message words; state = h0..h4 (host order) */
void orc_rmd160_block(uint32_t state[5], const uint32_t w[16]) {
  const u32 iv[5] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u, 0xc3d2e1f0u};
  u32 al = iv[0], bl = iv[1], cl = iv[2], dl = iv[3], el = iv[4];
  u32 ar = iv[0], br = iv[1], cr = iv[2], dr = iv[3], er = iv[4];
  for (int j = 0; j < 80; ++j) {
    u32 t = rotl32(al + rmd_f(j / 16, bl, cl, dl) + w[RMD_RL[j]] + RMD_KL[j / 16], RMD_SL[j]) + el;
    al = el, el = dl, dl = rotl32(cl, 10), cl = bl, bl = t;
    t = rotl32(ar + rmd_f(4 - j / 16, br, cr, dr) + w[RMD_RR[j]] + RMD_KR[j / 16], RMD_SR[j]) + er;
    ar = er, er = dr, dr = rotl32(cr, 10), cr = br, br = t;
  }
  state[0] = iv[1] + cl + dr;
  state[1] = iv[2] + dl + er;
  state[2] = iv[3] + el + ar;
  state[3] = iv[4] + al + br;
  state[4] = iv[0] + bl + cr;
}

/* ---------------------------------------------------------------- point -> hash160 (lib/addr.c:33-131) */

static void be_store(u8 *dst, const orc_fe v) { /* 32 bytes big-endian */
  for (int i = 0; i < 32; ++i) dst[i] = (u8)(v[3 - i / 8] >> (56 - 8 * (i % 8)));
}
static void sha_to_h160(uint32_t h[5], const u32 sha[8]) { /* addr.c:69-73,108-111; rmd160s.c:325-336 */
  u32 w[16] = {0}, st[5];
  for (int i = 0; i < 8; ++i) w[i] = __builtin_bswap32(sha[i]); /* digest bytes read as LE words */
  w[8] = 0x00000080u;
  w[14] = 256;
  orc_rmd160_block(st, w);
  for (int i = 0; i < 5; ++i) h[i] = __builtin_bswap32(st[i]); /* h160_t word = BE load of digest bytes */
}
void orc_hash160_33(uint32_t h[5], const orc_fe x, const orc_fe y) { /* addr.c:33-45,99-114 */
  u8 msg[64] = {0};
  u32 sha[8];
  msg[0] = (y[0] & 1) ? 0x03 : 0x02;
  be_store(msg + 1, x);
  msg[33] = 0x80;
  msg[62] = 0x01;
  msg[63] = 0x08;
  orc_sha256_blocks(sha, msg, 1);
  sha_to_h160(h, sha);
}
void orc_hash160_65(uint32_t h[5], const orc_fe x, const orc_fe y) { /* addr.c:47-67,116-131 */
  u8 msg[128] = {0};
  u32 sha[8];
  msg[0] = 0x04;
  be_store(msg + 1, x);
  be_store(msg + 33, y);
  msg[65] = 0x80;
  msg[126] = 0x02;
  msg[127] = 0x08;
  orc_sha256_blocks(sha, msg, 2);
  sha_to_h160(h, sha);
}

/* ---------------------------------------------------------------- bloom (lib/utils.c:282-326) */

void orc_blf_positions(uint64_t pos[20], const uint32_t h[5], uint64_t size_words) {
  const u64 a[5] = {(u64)h[0] << 32 | h[1], (u64)h[2] << 32 | h[3], (u64)h[4] << 32 | h[0],
                    (u64)h[1] << 32 | h[2], (u64)h[3] << 32 | h[4]};
  const int shifts[4] = {24, 28, 36, 40};
  for (int s = 0; s < 4; ++s)
    for (int i = 0; i < 5; ++i) {
      u64 v = a[i] << shifts[s] | a[(i + 1) % 5] >> shifts[s];
      pos[s * 5 + i] = v % (size_words * 64); /* word = pos/64, bit = pos%64 (= v%64) */
    }
}
void orc_blf_add(uint64_t *bits, uint64_t size_words, const uint32_t h[5]) {
  u64 pos[20];
  orc_blf_positions(pos, h, size_words);
  for (int i = 0; i < 20; ++i) bits[pos[i] / 64] |= (u64)1 << (pos[i] % 64);
}
int orc_blf_has(const uint64_t *bits, uint64_t size_words, const uint32_t h[5]) {
  u64 pos[20];
  orc_blf_positions(pos, h, size_words);
  for (int i = 0; i < 20; ++i)
    if (!(bits[pos[i] / 64] >> (pos[i] % 64) & 1)) return 0;
  return 1;
}
static int cmp160(const void *a, const void *b) { /* addr.c:18-26 */
  const u32 *x = a, *y = b;
  for (int i = 0; i < 5; ++i)
    if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
  return 0;
}
int orc_filter_check(const orc_filter *f, const uint32_t h[5]) { /* main.c:205-217 */
  if (!orc_blf_has(f->bits, f->size, h)) return 0;
  if (!f->list) return 1;
  return bsearch(h, f->list, f->count, 20, cmp160) != NULL;
}

/* ---------------------------------------------------------------- key recovery (main.c:267-276) */

void orc_calc_priv(orc_fe pk, const orc_fe start, const orc_fe stride, uint64_t off, uint8_t endo) {
  orc_fn_add_stride(pk, start, stride, off);
  if (endo == 0) return;
  if (endo == 1) orc_fn_neg(pk, pk);
  if (endo == 2 || endo == 3) orc_fn_mul(pk, pk, LAM1);
  if (endo == 3) orc_fn_neg(pk, pk);
  if (endo == 4 || endo == 5) orc_fn_mul(pk, pk, LAM2);
  if (endo == 5) orc_fn_neg(pk, pk);
}

/* ---------------------------------------------------------------- add path (main.c:287-403) */

#define GRP 2048u
#define HALF 1024u

typedef struct {
  orc_hit *hits;
  uint64_t cap, n;
} hit_sink;

static void emit(hit_sink *s, const orc_filter *f, const uint32_t h[5], uint64_t off, uint8_t endo, uint8_t kind,
                 const orc_fe start, const orc_fe stride, const orc_fe pk_direct) {
  if (!orc_filter_check(f, h)) return;
  if (s->n < s->cap) {
    orc_hit *o = &s->hits[s->n];
    memset(o, 0, sizeof *o);
    o->key_off = off;
    memcpy(o->h160, h, 20);
    o->endo = endo;
    o->kind = kind;
    if (pk_direct) memcpy(o->pk, pk_direct, sizeof(orc_fe));
    else orc_calc_priv(o->pk, start, stride, off, endo);
  }
  s->n++;
}

/* hash one affine point for the selected address kinds and probe (check_found_add inner, main.c:291-298) */
static void hash_probe(hit_sink *s, const orc_filter *f, uint32_t flags, const orc_fe x, const orc_fe y, uint64_t off,
                       uint8_t endo, const orc_fe start, const orc_fe stride) {
  uint32_t h[5];
  if (flags & ORC_A33) {
    orc_hash160_33(h, x, y);
    emit(s, f, h, off, endo, 0, start, stride, NULL);
  }
  if (flags & ORC_A65) {
    orc_hash160_65(h, x, y);
    emit(s, f, h, off, endo, 1, start, stride, NULL);
  }
}

uint64_t orc_add_span(const orc_fe start, const orc_fe stride, uint64_t n_keys, uint32_t flags, const orc_filter *f,
                      orc_hit *hits, uint64_t cap) {
  hit_sink sink = {hits, cap, 0};
  /* ctx_precompute_gpoints (main.c:219-246): gp[i] = (i+1)*stride*G, i < 1024; stride_p = 2048*stride*G */
  static orc_fe gpx[HALF], gpy[HALF];
  orc_fe sx, sy, t;
  orc_fe k1 = {0, 0, 0, 0};
  orc_fn_add_stride(k1, k1, stride, 1);
  orc_ec_mul_g(gpx[0], gpy[0], k1);
  for (uint32_t i = 1; i < HALF; ++i) {
    orc_fe ki = {0, 0, 0, 0};
    orc_fn_add_stride(ki, ki, stride, i + 1);
    orc_ec_mul_g(gpx[i], gpy[i], ki);
  }
  orc_fe ks = {0, 0, 0, 0};
  orc_fn_add_stride(ks, ks, stride, GRP);
  orc_ec_mul_g(sx, sy, ks);

  /* centre of the first group: (start + 1024*stride)*G (main.c:359-360) */
  orc_fe cx, cy, kc;
  orc_fn_add_stride(kc, start, stride, HALF);
  orc_ec_mul_g(cx, cy, kc);

  static orc_fe bx[GRP], by[GRP], dx[HALF];
  for (uint64_t base = 0; base < n_keys; base += GRP) {
    for (uint32_t i = 0; i < HALF; ++i) orc_fp_sub(dx[i], gpx[i], cx); /* main.c:369 */
    orc_fp_grpinv(dx, HALF);                                           /* main.c:370 */
    memcpy(bx[HALF], cx, sizeof cx);                                   /* main.c:372 */
    memcpy(by[HALF], cy, sizeof cy);
    for (int sign = 0; sign < 2; ++sign) {       /* main.c:374-396 */
      uint32_t gmax = sign == 0 ? HALF - 1 : HALF; /* plus side drops K+N/2 */
      for (uint32_t i = 0; i < gmax; ++i) {
        orc_fe gy, lam, rx, ry;
        if (sign == 0) memcpy(gy, gpy[i], sizeof gy);
        else orc_fp_neg(gy, gpy[i]);
        orc_fp_sub(lam, gy, cy);
        orc_fp_mul(lam, lam, dx[i]); /* lambda = (y2-y1)/(x2-x1) */
        orc_fp_sqr(rx, lam);
        orc_fp_sub(rx, rx, cx);
        orc_fp_sub(rx, rx, gpx[i]);
        orc_fp_sub(t, cx, rx);
        orc_fp_mul(t, lam, t);
        orc_fp_sub(ry, t, cy);
        uint32_t idx = sign == 0 ? HALF + i + 1 : HALF - 1 - i; /* main.c:391 */
        memcpy(bx[idx], rx, sizeof rx);
        memcpy(by[idx], ry, sizeof ry);
      }
    }
    /* check_found_add (main.c:287-347): plain points first, then endomorphism images */
    for (uint32_t j = 0; j < GRP; ++j) hash_probe(&sink, f, flags, bx[j], by[j], base + j, 0, start, stride);
    if (flags & ORC_ENDO) {
      for (uint32_t j = 0; j < GRP; ++j) {
        orc_fe ny, x1, x2;
        orc_fp_neg(ny, by[j]);
        orc_fp_mul(x1, bx[j], BETA1);
        orc_fp_mul(x2, bx[j], BETA2);
        hash_probe(&sink, f, flags, bx[j], ny, base + j, 1, start, stride);    /* (x,-y)      */
        hash_probe(&sink, f, flags, x1, by[j], base + j, 2, start, stride);    /* (bx, y)     */
        hash_probe(&sink, f, flags, x1, ny, base + j, 3, start, stride);       /* (bx,-y)     */
        hash_probe(&sink, f, flags, x2, by[j], base + j, 4, start, stride);    /* (b^2 x, y)  */
        hash_probe(&sink, f, flags, x2, ny, base + j, 5, start, stride);       /* (b^2 x,-y)  */
      }
    }
    /* next centre = centre + stride_p (main.c:400) */
    orc_fe nx, ny2;
    orc_ec_add(nx, ny2, cx, cy, sx, sy);
    memcpy(cx, nx, sizeof nx);
    memcpy(cy, ny2, sizeof ny2);
  }
  return sink.n;
}

/* cmd_add + cmd_add_worker, one thread (main.c:405-454) */
uint64_t orc_add_range(const orc_fe range_s, const orc_fe range_e, uint32_t ord_offs, uint32_t flags,
                       const orc_filter *f, orc_hit *hits, uint64_t cap, uint64_t *k_checked) {
  orc_fe stride = {0, 0, 0, 0};
  stride[ord_offs / 64] = (u64)1 << (ord_offs % 64); /* main.c:221-222 */
  orc_fe rsz;
  orc_fn_sub(rsz, range_e, range_s); /* main.c:441 */
  const u64 MAXJ = 2u * 1024 * 1024;
  u64 job = (rsz[3] | rsz[2] | rsz[1]) == 0 && rsz[0] < MAXJ ? rsz[0] : MAXJ; /* main.c:442 */
  orc_fe inc = {job, 0, 0, 0};
  orc_fn_mul(inc, inc, stride); /* main.c:413-415 */
  orc_fe cur, init;
  memcpy(cur, range_s, sizeof cur);
  memcpy(init, range_s, sizeof init);
  uint64_t total = 0, checked = 0;
  while (!(cmp4(cur, range_e) >= 0 || cmp4(cur, init) < 0)) { /* main.c:420-424 */
    u64 visit = (job + GRP - 1) / GRP * GRP;                    /* main.c:368,401 */
    orc_fe off256;
    orc_fn_sub(off256, cur, range_s);
    /* key_off relative to range_s, in units of stride (only meaningful while it fits 64 bits) */
    u64 off_units = 0;
    {
      /* off256 / stride: stride is a power of two */
      u64 tmp[4];
      memcpy(tmp, off256, sizeof tmp);
      for (uint32_t s = 0; s < ord_offs; ++s) {
        tmp[0] = tmp[0] >> 1 | tmp[1] << 63;
        tmp[1] = tmp[1] >> 1 | tmp[2] << 63;
        tmp[2] = tmp[2] >> 1 | tmp[3] << 63;
        tmp[3] >>= 1;
      }
      off_units = tmp[0];
    }
    uint64_t room = total < cap ? cap - total : 0;
    uint64_t got = orc_add_span(cur, stride, visit, flags, f, hits ? hits + total : NULL, room);
    uint64_t stored = got < room ? got : room;
    for (uint64_t i = 0; i < stored; ++i) hits[total + i].key_off += off_units;
    total += stored;
    if (got > room) total = cap; /* saturate */
    checked += (flags & ORC_ENDO) ? job * 6 : job; /* main.c:431 */
    orc_fn_add(cur, cur, inc);                     /* main.c:427 */
  }
  if (k_checked) *k_checked = checked;
  return total;
}

/* ---------------------------------------------------------------- mul path (main.c:458-540) */

uint64_t orc_mul_batch(const orc_fe *pks, uint64_t n, uint32_t flags, const orc_filter *f, orc_hit *hits,
                       uint64_t cap) {
  hit_sink sink = {hits, cap, 0};
  for (uint64_t i = 0; i < n; ++i) {
    orc_fe x, y;
    uint32_t h[5];
    if (orc_ec_mul_g(x, y, pks[i])) continue; /* k = 0 mod n: skipped (reference poisons the batch, A.7) */
    if (flags & ORC_A33) {
      orc_hash160_33(h, x, y);
      emit(&sink, f, h, i, 0, 0, NULL, NULL, pks[i]);
    }
    if (flags & ORC_A65) {
      orc_hash160_65(h, x, y);
      emit(&sink, f, h, i, 0, 1, NULL, NULL, pks[i]);
    }
  }
  return sink.n;
}

void orc_pubkey_hashes(const orc_fe *pks, uint64_t n, uint32_t *out33, uint32_t *out65, uint64_t *outxy) {
  for (uint64_t i = 0; i < n; ++i) {
    orc_fe x, y;
    orc_ec_mul_g(x, y, pks[i]);
    if (out33) orc_hash160_33(out33 + 5 * i, x, y);
    if (out65) orc_hash160_65(out65 + 5 * i, x, y);
    if (outxy) {
      memcpy(outxy + 8 * i, x, sizeof x);
      memcpy(outxy + 8 * i + 4, y, sizeof y);
    }
  }
}
