/* u256.c — see u256.h. Host bookkeeping only: nothing here touches a curve point or a hash. */
#include "u256.h"

#include <string.h>

typedef unsigned __int128 u128;

const u256 SECP_N = {0xbfd25e8cd0364141ULL, 0xbaaedce6af48a03bULL, 0xfffffffffffffffeULL, 0xffffffffffffffffULL};
const u256 SECP_P = {0xfffffffefffffc2fULL, 0xffffffffffffffffULL, 0xffffffffffffffffULL, 0xffffffffffffffffULL};
const u256 SECP_LAMBDA = {0xdf02967c1b23bd72ULL, 0x122e22ea20816678ULL, 0xa5261c028812645aULL, 0x5363ad4cc05c30e0ULL};
const u256 SECP_LAMBDA2 = {0xe0cfc810b51283ceULL, 0xa880b9fc8ec739c2ULL, 0x5ad9e3fd77ed9ba4ULL, 0xac9c52b33fa3cf1fULL};
/* 2^256 - n, 129 bits */
static const uint64_t N_COMPL[3] = {0x402da1732fc9bebfULL, 0x4551231950b75fc4ULL, 0x1ULL};

void u256_set64(u256 r, uint64_t v) { r[0] = v, r[1] = r[2] = r[3] = 0; }
void u256_copy(u256 r, const u256 a) { memmove(r, a, sizeof(u256)); }
bool u256_is_zero(const u256 a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }

int u256_cmp(const u256 a, const u256 b) {
  for (int i = 3; i >= 0; --i)
    if (a[i] != b[i]) return a[i] > b[i] ? 1 : -1;
  return 0;
}

unsigned u256_bitlen(const u256 a) {
  for (int i = 3; i >= 0; --i)
    if (a[i]) return (unsigned)(64 * i + 64 - __builtin_clzll(a[i]));
  return 0;
}

uint64_t u256_add_raw(u256 r, const u256 a, const u256 b) {
  u128 c = 0;
  for (int i = 0; i < 4; ++i) {
    c += (u128)a[i] + b[i];
    r[i] = (uint64_t)c;
    c >>= 64;
  }
  return (uint64_t)c;
}

uint64_t u256_sub_raw(u256 r, const u256 a, const u256 b) {
  uint64_t bw = 0;
  for (int i = 0; i < 4; ++i) {
    const u128 t = (u128)a[i] - b[i] - bw;
    r[i] = (uint64_t)t;
    bw = (uint64_t)(t >> 64) & 1;
  }
  return bw;
}

void modn_add(u256 r, const u256 a, const u256 b) {
  if (u256_add_raw(r, a, b)) u256_sub_raw(r, r, SECP_N);
}

void modn_sub(u256 r, const u256 a, const u256 b) {
  if (u256_sub_raw(r, a, b)) u256_add_raw(r, r, SECP_N);
}

void modn_neg(u256 r, const u256 a) { u256_sub_raw(r, SECP_N, a); }

/* x[0..7] (512 bit) -> canonical residue mod n: fold the high half with 2^256 = N_COMPL (mod n) until it is gone */
static void reduce512(u256 r, uint64_t x[8]) {
  for (;;) {
    if (!(x[4] | x[5] | x[6] | x[7])) break;
    uint64_t hi[4] = {x[4], x[5], x[6], x[7]}, t[8] = {x[0], x[1], x[2], x[3], 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) { /* t += hi[i] * N_COMPL << 64 i */
      u128 c = 0;
      for (int j = 0; j < 3; ++j) {
        c += (u128)hi[i] * N_COMPL[j] + t[i + j];
        t[i + j] = (uint64_t)c;
        c >>= 64;
      }
      for (int k = i + 3; k < 8 && c; ++k) {
        c += t[k];
        t[k] = (uint64_t)c;
        c >>= 64;
      }
    }
    memcpy(x, t, sizeof t);
  }
  while (u256_cmp(x, SECP_N) >= 0) u256_sub_raw(x, x, SECP_N);
  memcpy(r, x, sizeof(u256));
}

void modn_mul(u256 r, const u256 a, const u256 b) {
  uint64_t x[8] = {0};
  for (int i = 0; i < 4; ++i) {
    u128 c = 0;
    for (int j = 0; j < 4; ++j) {
      c += (u128)a[i] * b[j] + x[i + j];
      x[i + j] = (uint64_t)c;
      c >>= 64;
    }
    x[i + 4] = (uint64_t)c;
  }
  reduce512(r, x);
}

void modn_add_stride(u256 r, const u256 base, const u256 stride, uint64_t off) {
  u256 t;
  u256_set64(t, off);
  modn_mul(t, t, stride);
  modn_add(r, t, base);
}

/* nibble value of a character, 0xFF for anything that is not a hex digit */
static uint8_t HEXVAL[256];
static void hexval_init(void) __attribute__((constructor));
static void hexval_init(void) {
  memset(HEXVAL, 0xFF, sizeof HEXVAL);
  for (int c = '0'; c <= '9'; ++c) HEXVAL[c] = (uint8_t)(c - '0');
  for (int c = 'a'; c <= 'f'; ++c) HEXVAL[c] = (uint8_t)(c - 'a' + 10);
  for (int c = 'A'; c <= 'F'; ++c) HEXVAL[c] = (uint8_t)(c - 'A' + 10);
}

void u256_from_hex_n(u256 r, const char *hex, size_t len) {
  u256_set64(r, 0);
  unsigned cnt = 0;
  while (len-- > 0) {
    const uint8_t v = HEXVAL[(unsigned char)hex[len]];
    if (v == 0xFF) continue;
    if (cnt < 64) r[cnt / 16] |= (uint64_t)v << (4 * (cnt % 16)); /* the reference writes past the array here */
    cnt++;
  }
}

void u256_from_hex(u256 r, const char *hex) { u256_from_hex_n(r, hex, strlen(hex)); }

void modn_from_hex_n(u256 r, const char *hex, size_t len) {
  u256_from_hex_n(r, hex, len);
  if (u256_cmp(r, SECP_N) >= 0) modn_sub(r, r, SECP_N);
}

void modn_from_hex(u256 r, const char *hex) { modn_from_hex_n(r, hex, strlen(hex)); }
