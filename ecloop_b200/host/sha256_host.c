/* sha256_host.c — FIPS 180-4 SHA-256, one-shot, for the `-raw` line -> key parse (see sha256_host.h) */
#include "sha256_host.h"

#include <string.h>

static const uint32_t RC[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static inline uint32_t ror(uint32_t x, unsigned n) { return x >> n | x << (32 - n); }

static void block(uint32_t s[8], const uint8_t *p) {
  uint32_t w[64];
  for (int i = 0; i < 16; ++i) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
  for (int i = 16; i < 64; ++i) {
    const uint32_t a = w[i - 15], b = w[i - 2];
    w[i] = w[i - 16] + (ror(a, 7) ^ ror(a, 18) ^ a >> 3) + w[i - 7] + (ror(b, 17) ^ ror(b, 19) ^ b >> 10);
  }
  uint32_t v[8];
  memcpy(v, s, sizeof v);
  for (int i = 0; i < 64; ++i) {
    const uint32_t e = v[4], a = v[0];
    const uint32_t t1 = v[7] + (ror(e, 6) ^ ror(e, 11) ^ ror(e, 25)) + ((e & v[5]) ^ (~e & v[6])) + RC[i] + w[i];
    const uint32_t t2 = (ror(a, 2) ^ ror(a, 13) ^ ror(a, 22)) + ((a & v[1]) ^ (a & v[2]) ^ (v[1] & v[2]));
    memmove(v + 1, v, 7 * sizeof(uint32_t));
    v[4] += t1;
    v[0] = t1 + t2;
  }
  for (int i = 0; i < 8; ++i) s[i] += v[i];
}

void sha256_bytes(uint32_t digest_words[8], const uint8_t *msg, size_t len) {
  uint32_t s[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  size_t off = 0;
  for (; off + 64 <= len; off += 64) block(s, msg + off);
  uint8_t tail[128] = {0};
  const size_t rem = len - off;
  memcpy(tail, msg + off, rem);
  tail[rem] = 0x80;
  const size_t tlen = rem + 9 <= 64 ? 64 : 128;
  const uint64_t bits = (uint64_t)len * 8;
  for (int j = 0; j < 8; ++j) tail[tlen - 1 - j] = (uint8_t)(bits >> (8 * j));
  block(s, tail);
  if (tlen == 128) block(s, tail + 64);
  memcpy(digest_words, s, sizeof s);
}
