// bloom.cuh — the reference's 20-probe bloom test (blf_has/blf_getbit, lib/utils.c:286-326), bit-exact
// including false positives: five overlapping 64-bit lanes of the hash160, four shifts {24,28,36,40},
// pos = v mod (size*64), early exit on the first clear bit.
//
// v mod (size*64) is computed as ((v >> 6) mod size)*64 + (v & 63): the word index only needs a 58-bit by
// `size` remainder, done with a precomputed reciprocal (floor((2^64-1)/size)) and at most a few corrections
// instead of a 64-bit hardware-less division.
#pragma once
#include <stdint.h>

typedef uint32_t u32;
typedef uint64_t u64;

struct BloomView {
  const u64 *bits;  // global or shared (generic address)
  u64 size;         // words
  u64 magic;        // floor((2^64 - 1) / size)
};

__device__ __forceinline__ u64 bloom_word_index(u64 q, u64 size, u64 magic) {
  u64 r = q - __umul64hi(q, magic) * size;  // quotient estimate is never too large, at most 2 too small
  while (r >= size) r -= size;
  return r;
}

// probes FIRST..19 of the 20, in the reference's order (shifts outer, lanes inner); FIRST = 2 for candidates whose
// first two probes were already seen set by stage 1 of the asynchronous probe (probe_pipe.cuh)
template <int FIRST>
static __device__ __noinline__ bool bloom_has_from(const u64 *bits, u64 size, u64 magic, u64 a0, u64 a1, u64 a2, u64 a3, u64 a4) {
  const u64 a[5] = {a0, a1, a2, a3, a4};
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int S = s == 0 ? 24 : s == 1 ? 28 : s == 2 ? 36 : 40;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      if (s * 5 + i < FIRST) continue;
      const u64 v = (a[i] << S) | (a[(i + 1) % 5] >> S);
      const u64 word = bits[bloom_word_index(v >> 6, size, magic)];
      if (!((word >> (v & 63)) & 1)) return false;
    }
  }
  return true;
}

// The first probe rejects most keys (1 - fill of them), so only it is inlined into the hot loop.
__device__ __forceinline__ bool bloom_has(const BloomView &b, const u32 h[5]) {
  const u64 a0 = (u64)h[0] << 32 | h[1], a1 = (u64)h[2] << 32 | h[3];
  const u64 v = (a0 << 24) | (a1 >> 24);
  const u64 word = b.bits[bloom_word_index(v >> 6, b.size, b.magic)];
  if (!((word >> (v & 63)) & 1)) return false;
  return bloom_has_from<0>(b.bits, b.size, b.magic, a0, a1, (u64)h[4] << 32 | h[0], (u64)h[1] << 32 | h[2],
                           (u64)h[3] << 32 | h[4]);
}

// the verdict on a hash whose probes 0 and 1 are known to be set
__device__ __forceinline__ bool bloom_has_after_two(const BloomView &b, const u32 h[5]) {
  return bloom_has_from<2>(b.bits, b.size, b.magic, (u64)h[0] << 32 | h[1], (u64)h[2] << 32 | h[3], (u64)h[4] << 32 | h[0],
                           (u64)h[1] << 32 | h[2], (u64)h[3] << 32 | h[4]);
}
