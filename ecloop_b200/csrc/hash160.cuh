// hash160.cuh — SHA-256 -> RIPEMD-160 of a serialised secp256k1 point, NW independent keys per thread
// interleaved at source level (every step is applied to all NW lanes back to back, so the scheduler always has
// NW independent dependency chains to issue from).
//
// Replaces prepare33/prepare65 + sha256_final + rmd160_batch (lib/addr.c:33-131, lib/sha256.c, lib/rmd160s.c).
// Message layouts are hard-wired: for the 33-byte key the block words W9..W14 are zero and
// W15 = 264, so the compiler folds them through the first schedule rounds; the RIPEMD block is 8 digest words
// + fixed padding. Output is h160_t word order (big-endian load of the digest bytes, lib/addr.c:16).
#pragma once
#include <stdint.h>

typedef uint32_t u32;
typedef uint64_t u64;

template <int N>
struct vw {  // N independent 32-bit lanes owned by one thread
  u32 l[N];
};
#define VW_BINOP(op)                                                            \
  template <int N>                                                              \
  __device__ __forceinline__ vw<N> operator op(const vw<N> &a, const vw<N> &b) { \
    vw<N> r;                                                                    \
    _Pragma("unroll") for (int i = 0; i < N; ++i) r.l[i] = a.l[i] op b.l[i];   \
    return r;                                                                   \
  }                                                                             \
  template <int N>                                                              \
  __device__ __forceinline__ vw<N> operator op(const vw<N> &a, u32 b) {         \
    vw<N> r;                                                                    \
    _Pragma("unroll") for (int i = 0; i < N; ++i) r.l[i] = a.l[i] op b;        \
    return r;                                                                   \
  }
VW_BINOP(+)
VW_BINOP(^)
VW_BINOP(&)
VW_BINOP(|)
template <int N>
__device__ __forceinline__ vw<N> operator~(const vw<N> &a) {
  vw<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i) r.l[i] = ~a.l[i];
  return r;
}
template <int N>
__device__ __forceinline__ vw<N> vrotr(const vw<N> &a, int n) {
  vw<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i) r.l[i] = __funnelshift_r(a.l[i], a.l[i], n);
  return r;
}
template <int N>
__device__ __forceinline__ vw<N> vrotl(const vw<N> &a, int n) {
  vw<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i) r.l[i] = __funnelshift_l(a.l[i], a.l[i], n);
  return r;
}
template <int N>
__device__ __forceinline__ vw<N> vshr(const vw<N> &a, int n) {
  vw<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i) r.l[i] = a.l[i] >> n;
  return r;
}
template <int N>
__device__ __forceinline__ vw<N> vbswap(const vw<N> &a) {
  vw<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i) r.l[i] = __byte_perm(a.l[i], 0, 0x0123);
  return r;
}
template <int N>
__device__ __forceinline__ vw<N> vset(u32 c) {
  vw<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i) r.l[i] = c;
  return r;
}


// ---------------------------------------------------------------- pipe steering
// The fused add kernel is bound by the ALU pipe (LOP3/SHF/IADD3: 64 lanes/SM/clk) while the FMA pipe (IMAD,
// another 64 lanes/SM/clk, co-issued) idles during the hashes. ptxas chooses IADD3 for every addition it sees,
// so the additions we want on the FMA pipe are written as a*ONE+b with ONE read from constant memory (opaque
// to the compiler, folded into the IMAD's constant operand, no register).
// Levels (compile-time, see DESIGN.md K1 "pipe balance"):
//   ECL_SHA_FMA  0 none | 1 w+K, h+wk | 3 + e', a', S0+maj (ALU 1, FMA 5 adds per round) | 4 all 7 adds
//   ECL_SHS_FMA  0 none | 1 w[i]+w[i+9] | 2 all three schedule adds
//   ECL_RMD_FMA  0 none | 1 a+(w+K) | 2 + F
#ifndef ECL_SHA_FMA
#define ECL_SHA_FMA 4
#endif
#ifndef ECL_SHS_FMA
#define ECL_SHS_FMA 2
#endif
#ifndef ECL_RMD_FMA
#define ECL_RMD_FMA 2
#endif
static __constant__ u32 ecl_k_one = 1u;
// Measured and dropped (round 1): shifts as IMAD.HI (half rate, peak.cuh kind 7), sigma rotations and RIPEMD's
// rotl(c, 10) as products (-5 % .. -13 %): rotations and shifts stay on SHF.
__device__ __forceinline__ u32 fma_add(u32 a, u32 b) {
  u32 d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(ecl_k_one), "r"(b));
  return d;
}
template <int N>
__device__ __forceinline__ vw<N> vfadd(const vw<N> &a, const vw<N> &b) {
  vw<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i) r.l[i] = fma_add(a.l[i], b.l[i]);
  return r;
}
template <int N>
__device__ __forceinline__ vw<N> vfadd(const vw<N> &a, u32 k) {
  vw<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i) r.l[i] = fma_add(a.l[i], k);
  return r;
}

// ---------------------------------------------------------------- scheduling hooks
// A hook is called at fixed places inside a hash (before SHA-256 round i uses its message word; at the four round
// boundaries of RIPEMD-160) with a reference to a live word of the hash state. The pipelined add kernel uses it to run
// one field multiplication there and XOR (result & 0) into the word: a data dependency that costs one LOP3 and makes
// ptxas place that multiplication inside the hash instead of after it (DESIGN.md K1 "pins"). NoHook compiles to nothing.
struct NoHook {
  __device__ __forceinline__ void sha(int, u32 &) {}
  __device__ __forceinline__ void rmd(int, u32 &) {}
};

// ---------------------------------------------------------------- SHA-256 (FIPS 180-4; lib/sha256.c:399-453)

// one compression; st = chaining value in/out, w = 16 message words (big-endian loads), clobbered
// SYNC != 0: the caller guarantees that every thread of the CTA runs this function the same number of times, and
// the warps re-align at a CTA barrier every 16 rounds / 32 steps so that they share instruction fetches (the
// unrolled hashes are ~100 KB of code, far beyond the instruction caches; see DESIGN.md K1 "lockstep")
template <int SYNC>
__device__ __forceinline__ void hash_sync_point() {
  if (SYNC) __syncthreads();
}

// CMASK: bit i set = message word i is a compile-time constant (padding); additions with such words stay plain C
// so that the compiler folds them, everything else is steered by the ECL_*_FMA levels above.
template <int N, int SYNC = 0, u32 CMASK = 0, class HOOK = NoHook>
__device__ __forceinline__ void sha256_compress(vw<N> st[8], vw<N> w[16], HOOK &hook, int hook_base = 0) {
  constexpr u32 K[64] = {
      0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
      0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
      0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
      0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
      0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
      0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
      0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
      0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
  vw<N> a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
  bool wc[64];  // wc[i]: schedule word i is a compile-time constant (evaluated by the compiler after unrolling)
#pragma unroll
  for (int i = 0; i < 16; ++i) wc[i] = (CMASK >> i) & 1u;
#pragma unroll
  for (int i = 16; i < 64; ++i) wc[i] = wc[i - 16] && wc[i - 15] && wc[i - 7] && wc[i - 2];
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    if (SYNC > 1 && i > 0 && (i % (64 / SYNC)) == 0) hash_sync_point<SYNC>();
    if (i >= 16) {
      const vw<N> x = w[(i + 1) & 15], y = w[(i + 14) & 15];
      const bool cA = wc[i - 16], cB = wc[i - 15], cC = wc[i - 7], cD = wc[i - 2];
      const vw<N> s0 = vrotr(x, 7) ^ vrotr(x, 18) ^ vshr(x, 3);
      const vw<N> s1 = vrotr(y, 17) ^ vrotr(y, 19) ^ vshr(y, 10);
#if ECL_SHS_FMA == 0
      w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
#else
      // constant terms are summed in plain C (folded), the others on the FMA pipe
      vw<N> kc = vset<N>(0), r = vset<N>(0);
      bool have = false;
      const vw<N> term[4] = {w[i & 15], w[(i + 9) & 15], s0, s1};
      const bool tc[4] = {cA, cC, cB, cD};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (tc[q]) kc = kc + term[q];
        else if (!have) r = term[q], have = true;
        else if (ECL_SHS_FMA == 1 && q >= 2) r = r + term[q];
        else r = vfadd(r, term[q]);
      }
      w[i & 15] = have ? r + kc : kc;
#endif
    }
    hook.sha(hook_base + i, w[i & 15].l[0]);
    const vw<N> S1 = vrotr(e, 6) ^ vrotr(e, 11) ^ vrotr(e, 25), ch = (e & f) ^ (~e & g);
    const vw<N> S0 = vrotr(a, 2) ^ vrotr(a, 13) ^ vrotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
#if ECL_SHA_FMA == 0
    const vw<N> t1 = h + S1 + ch + (w[i & 15] + K[i]);
    const vw<N> ne = d + t1, na = t1 + (S0 + mj);
#elif ECL_SHA_FMA == 1
    const vw<N> p = vfadd(h, wc[i] ? w[i & 15] + K[i] : vfadd(w[i & 15], K[i]));
    const vw<N> t1 = p + S1 + ch;
    const vw<N> ne = d + t1, na = t1 + S0 + mj;
#elif ECL_SHA_FMA == 3
    const vw<N> p = vfadd(h, wc[i] ? w[i & 15] + K[i] : vfadd(w[i & 15], K[i]));
    const vw<N> t1 = p + S1 + ch;
    const vw<N> ne = vfadd(d, t1), na = vfadd(t1, vfadd(S0, mj));
#else
    const vw<N> p = vfadd(h, wc[i] ? w[i & 15] + K[i] : vfadd(w[i & 15], K[i]));
    const vw<N> t1 = vfadd(p, vfadd(S1, ch));
    const vw<N> ne = vfadd(d, t1), na = vfadd(t1, vfadd(S0, mj));
#endif
    h = g, g = f, f = e, e = ne, d = c, c = b, b = a, a = na;
  }
  st[0] = st[0] + a, st[1] = st[1] + b, st[2] = st[2] + c, st[3] = st[3] + d;
  st[4] = st[4] + e, st[5] = st[5] + f, st[6] = st[6] + g, st[7] = st[7] + h;
}

template <int N, int SYNC = 0, u32 CMASK = 0>
__device__ __forceinline__ void sha256_compress(vw<N> st[8], vw<N> w[16]) {
  NoHook none;
  sha256_compress<N, SYNC, CMASK, NoHook>(st, w, none);
}

template <int N>
__device__ __forceinline__ void sha256_iv(vw<N> st[8]) {
  st[0] = vset<N>(0x6a09e667u), st[1] = vset<N>(0xbb67ae85u), st[2] = vset<N>(0x3c6ef372u), st[3] = vset<N>(0xa54ff53au);
  st[4] = vset<N>(0x510e527fu), st[5] = vset<N>(0x9b05688cu), st[6] = vset<N>(0x1f83d9abu), st[7] = vset<N>(0x5be0cd19u);
}

// ---------------------------------------------------------------- RIPEMD-160 (lib/rmd160s.c:122-336)

#define RMD_F1(x, y, z) ((x) ^ (y) ^ (z))
#define RMD_F2(x, y, z) (((x) & (y)) | (~(x) & (z)))
#define RMD_F3(x, y, z) (((x) | ~(y)) ^ (z))
#define RMD_F4(x, y, z) (((x) & (z)) | ((y) & ~(z)))
#define RMD_F5(x, y, z) ((x) ^ ((y) | ~(z)))
#define RMD_WK(wi, k) ((k) ? vfadd(w[wi], (u32)(k)) : w[wi])
// message words 8..15 are padding constants: a + F + (w + K) is then one IADD3 with an immediate
#define RMD_CONSTW(wi) ((wi) >= 8)
#if ECL_RMD_FMA == 0
#define RMD_STEP(F, a, b, c, d, e, wi, k, s)      \
  a = vrotl(a + F(b, c, d) + (w[wi] + (k)), s) + e; \
  c = vrotl(c, 10);
#elif ECL_RMD_FMA == 1
#define RMD_STEP(F, a, b, c, d, e, wi, k, s)                                           \
  a = RMD_CONSTW(wi) ? vrotl(a + F(b, c, d) + (w[wi] + (k)), s) + e                    \
                     : vrotl(vfadd(a, RMD_WK(wi, k)) + F(b, c, d), s) + e;             \
  c = vrotl(c, 10);
#else
#define RMD_STEP(F, a, b, c, d, e, wi, k, s)                                           \
  a = RMD_CONSTW(wi) ? vrotl(a + F(b, c, d) + (w[wi] + (k)), s) + e                    \
                     : vrotl(vfadd(vfadd(a, RMD_WK(wi, k)), F(b, c, d)), s) + e;       \
  c = vrotl(c, 10);
#endif

// digest words of SHA-256 (sha[0..7], big-endian word values) -> h160_t words
template <int N, int SYNC = 0, class HOOK = NoHook>
__device__ __forceinline__ void rmd160_of_sha(vw<N> out[5], const vw<N> sha[8], HOOK &hook) {
  vw<N> w[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = vbswap(sha[i]);  // digest bytes read back as little-endian words
  w[8] = vset<N>(0x00000080u);
#pragma unroll
  for (int i = 9; i < 16; ++i) w[i] = vset<N>(0);
  w[14] = vset<N>(256u);  // message length in bits
  const u32 h0 = 0x67452301u, h1 = 0xefcdab89u, h2 = 0x98badcfeu, h3 = 0x10325476u, h4 = 0xc3d2e1f0u;
  vw<N> al = vset<N>(h0), bl = vset<N>(h1), cl = vset<N>(h2), dl = vset<N>(h3), el = vset<N>(h4);
  vw<N> ar = al, br = bl, cr = cl, dr = dl, er = el;
  int rmd_boundary = 0;
#define RMD_ROUND_BOUNDARY               \
  if (SYNC > 1) hash_sync_point<SYNC>(); \
  hook.rmd(rmd_boundary++, al.l[0]);
#include "rmd160_steps.inc"
#undef RMD_ROUND_BOUNDARY
  out[0] = vbswap(cl + dr + h1);
  out[1] = vbswap(dl + er + h2);
  out[2] = vbswap(el + ar + h3);
  out[3] = vbswap(al + br + h4);
  out[4] = vbswap(bl + cr + h0);
}

template <int N, int SYNC = 0>
__device__ __forceinline__ void rmd160_of_sha(vw<N> out[5], const vw<N> sha[8]) {
  NoHook none;
  rmd160_of_sha<N, SYNC, NoHook>(out, sha, none);
}

// ---------------------------------------------------------------- point -> hash160

// X[i] = big-endian word i of the coordinate = limb 7-i (little-endian 32-bit limbs)
// compressed key 02|03 || X (lib/addr.c:33-45): one block
template <int N, int SYNC, class HOOK>
__device__ __forceinline__ void hash160_33(vw<N> out[5], const u32 (&x)[N][8], const u32 (&y_odd)[N], HOOK &hook) {
  vw<N> w[16], st[8];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    w[0].l[n] = ((0x02u | (y_odd[n] & 1u)) << 24) | (x[n][7] >> 8);
#pragma unroll
    for (int i = 1; i < 8; ++i) w[i].l[n] = __funnelshift_r(x[n][7 - i], x[n][8 - i], 8);
    w[8].l[n] = (x[n][0] << 24) | 0x00800000u;
  }
#pragma unroll
  for (int i = 9; i < 15; ++i) w[i] = vset<N>(0);
  w[15] = vset<N>(33 * 8);
  sha256_iv(st);
  sha256_compress<N, SYNC, 0xFE00u, HOOK>(st, w, hook);
  hash_sync_point<SYNC>();
  rmd160_of_sha<N, SYNC, HOOK>(out, st, hook);
}
template <int N, int SYNC = 0>
__device__ __forceinline__ void hash160_33(vw<N> out[5], const u32 (&x)[N][8], const u32 (&y_odd)[N]) {
  NoHook none;
  hash160_33<N, SYNC, NoHook>(out, x, y_odd, none);
}

// uncompressed key 04 || X || Y (lib/addr.c:47-67): two blocks
template <int N, int SYNC, class HOOK>
__device__ __forceinline__ void hash160_65(vw<N> out[5], const u32 (&x)[N][8], const u32 (&y)[N][8], HOOK &hook) {
  vw<N> w[16], st[8];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    w[0].l[n] = (0x04u << 24) | (x[n][7] >> 8);
#pragma unroll
    for (int i = 1; i < 8; ++i) w[i].l[n] = __funnelshift_r(x[n][7 - i], x[n][8 - i], 8);
    w[8].l[n] = __funnelshift_r(y[n][7], x[n][0], 8);
#pragma unroll
    for (int i = 9; i < 16; ++i) w[i].l[n] = __funnelshift_r(y[n][15 - i], y[n][16 - i], 8);
  }
  sha256_iv(st);
  sha256_compress<N, SYNC, 0u, HOOK>(st, w, hook);  // the hook sees rounds 0..63 of the first block only
  hash_sync_point<SYNC>();
#pragma unroll
  for (int n = 0; n < N; ++n) w[0].l[n] = (y[n][0] << 24) | 0x00800000u;
#pragma unroll
  for (int i = 1; i < 15; ++i) w[i] = vset<N>(0);
  w[15] = vset<N>(65 * 8);
  sha256_compress<N, SYNC, 0xFFFEu>(st, w);
  hash_sync_point<SYNC>();
  rmd160_of_sha<N, SYNC, HOOK>(out, st, hook);
}
template <int N, int SYNC = 0>
__device__ __forceinline__ void hash160_65(vw<N> out[5], const u32 (&x)[N][8], const u32 (&y)[N][8]) {
  NoHook none;
  hash160_65<N, SYNC, NoHook>(out, x, y, none);
}
