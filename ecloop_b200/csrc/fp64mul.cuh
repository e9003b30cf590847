// fp64mul.cuh — EXPERIMENTAL (not on the hot path yet): secp256k1 field multiplication on the FP64 pipe.
//
// Why: the fused add kernel sits on the ALU pipe. Every IMAD.WIDE of fe_mul takes an ALU-pipe slot
// besides its FMA-pipe slot, and the carry chains are ALU work too, while B200's FP64 pipe is full rate and
// co-issues with both integer pipes (peak.cuh kinds 16-18: 16.9 T DFMA/s, 30.5 T/s mixed with LOP3 or IMAD).
// This file moves the multiplication there; parity with fe_mul is tested through ecl_prim_fp (ECL_OP_MUL_F64,
// ECL_OP_MUL_F64_CHAIN) and the throughput of both forms under an ALU-pipe load is measured by mulbench_kernel.
//
// Representation: a field element is 6 limbs of 44 bits held as doubles ("weak": limbs < 2^45, value < 2^265,
// not canonical). A limb product x*y < 2^90 goes through a biased accumulator t in the binade [2^96, 2^97), whose
// ulp is 2^44:   t' = fma_rz(x, y, t)  adds floor(x*y / 2^44) * 2^44 exactly,   r = fma(x, y, t - t')  is the exact
// low part x*y mod 2^44. Column sums of at most six products stay below 2^53 in each half, so nothing rounds.
// Reduction: 2^264 = 2^40 + 250112 (mod p); the high columns are folded through the same product trick, carries are
// taken with floor(c / 2^44) = fma_rz(c, 2^-44, 2^52) - 2^52. tools/f64mul_model.py is the exact integer model of
// every step (bounds asserted); each function below follows it line by line.
#pragma once
#include "fp.cuh"

struct fe6 {
  double v[6];
};

#define F6_BIAS 0x1p96
#define F6_M 1099511877888.0  // 2^40 + 250112 = 2^264 mod p
#define F6_2P44 0x1p44
#define F6_2M44 0x1p-44
#define F6_2P52 0x1p52
// limbs of 1024 p after borrowing 2^46 into each low limb: all >= 2^45 (tools/f64mul_model.py bias_limbs(45))
#define F6_SUB_BIAS0 83562882710528.0
#define F6_SUB_BIAS1 87960930222075.0
#define F6_SUB_BIAS2 87960930222075.0
#define F6_SUB_BIAS3 87960930222075.0
#define F6_SUB_BIAS4 87960930222075.0
#define F6_SUB_BIAS5 70368744177659.0

// 8 x 32-bit words (canonical or not) -> 6 x 44-bit limbs
__device__ __forceinline__ fe6 fe6_from_fe(const fe &a) {
  const u64 x0 = (u64)a.v[1] << 32 | a.v[0], x1 = (u64)a.v[3] << 32 | a.v[2];
  const u64 x2 = (u64)a.v[5] << 32 | a.v[4], x3 = (u64)a.v[7] << 32 | a.v[6];
  const u64 m44 = (1ull << 44) - 1;
  fe6 r;
  r.v[0] = (double)(long long)(x0 & m44);
  r.v[1] = (double)(long long)((x0 >> 44) | (x1 & ((1ull << 24) - 1)) << 20);
  r.v[2] = (double)(long long)((x1 >> 24) | (x2 & 0xfull) << 40);
  r.v[3] = (double)(long long)((x2 >> 4) & m44);
  r.v[4] = (double)(long long)((x2 >> 48) | (x3 & ((1ull << 28) - 1)) << 16);
  r.v[5] = (double)(long long)(x3 >> 28);
  return r;
}

// weak limbs -> the canonical residue in 8 x 32-bit words
__device__ __forceinline__ fe fe6_to_fe(const fe6 &c) {
  const u64 m44 = (1ull << 44) - 1;
  u64 l[6], carry = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const u64 t = (u64)__double2ll_rz(c.v[k]) + carry;
    l[k] = t & m44, carry = t >> 44;
  }
  u64 x[4];
  x[0] = l[0] | l[1] << 44;
  x[1] = l[1] >> 20 | l[2] << 24;
  x[2] = l[2] >> 40 | l[3] << 4 | l[4] << 48;
  x[3] = l[4] >> 16 | l[5] << 28;
  u64 top = l[5] >> 36 | carry << 8;  // bits 256 and up
  // fold top * 2^256 = top * (2^32 + 977); repeat while the addition wraps past 2^256 (at most once more)
  while (top) {
    unsigned __int128 acc = (unsigned __int128)x[0] + (unsigned __int128)top * 977u + ((unsigned __int128)top << 32);
    x[0] = (u64)acc;
    acc >>= 64;
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      acc += x[k];
      x[k] = (u64)acc;
      acc >>= 64;
    }
    top = (u64)acc;
  }
  fe r;
#pragma unroll
  for (int k = 0; k < 4; ++k) r.v[2 * k] = (u32)x[k], r.v[2 * k + 1] = (u32)(x[k] >> 32);
  fe_canon(r);
  return r;
}

// one product through the biased accumulator: hi = floor(x*y / 2^44), lo = x*y mod 2^44 (both exact)
__device__ __forceinline__ void f6_split(double &hi, double &lo, double x, double y) {
  const double t = __fma_rz(x, y, F6_BIAS);
  lo = __fma_rn(x, y, F6_BIAS - t);
  hi = (t - F6_BIAS) * F6_2M44;
}

// carries: any limbs below 2^52 -> "normal" limbs (< 2^44 + 64), value preserved mod p (model: carry_norm)
__device__ __forceinline__ void f6_carry(double c[6]) {
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const double q = __fma_rz(c[k], F6_2M44, F6_2P52) - F6_2P52;
    c[k] = __fma_rn(q, -F6_2P44, c[k]);
    c[k + 1] += q;
  }
  const double q = __fma_rz(c[5], F6_2M44, F6_2P52) - F6_2P52;
  c[5] = __fma_rn(q, -F6_2P44, c[5]);
  c[0] = __fma_rn(q, F6_M, c[0]);
  const double q0 = __fma_rz(c[0], F6_2M44, F6_2P52) - F6_2P52;
  c[0] = __fma_rn(q0, -F6_2P44, c[0]);
  c[1] += q0;
}

// columns c[0..11] of a product -> normal limbs: fold columns 6..11 with 2^264 = M (mod p), then carries
__device__ __forceinline__ fe6 f6_fold(double c[12]) {
  double c6b = 0.0;
#pragma unroll
  for (int k = 6; k < 12; ++k) {
    double h, l;
    f6_split(h, l, c[k], F6_M);
    c[k - 6] += l;
    if (k < 11) c[k - 5] += h;
    else c6b = h;
  }
  {
    double h, l;
    f6_split(h, l, c6b, F6_M);
    c[0] += l, c[1] += h;
  }
  f6_carry(c);
  fe6 r;
#pragma unroll
  for (int k = 0; k < 6; ++k) r.v[k] = c[k];
  return r;
}

// One operand may be "wide" (a raw difference from fe6_sub, limbs < 2^47.3) if the other is normal.
__device__ __forceinline__ fe6 fe6_mul(const fe6 &a, const fe6 &b) {
  double c[12];
  double hi_prev = 0.0;  // high half of column k-1, already divided by 2^44
#pragma unroll
  for (int k = 0; k < 11; ++k) {
    double t = F6_BIAS, s = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int j = k - i;
      if (j >= 0 && j < 6) {
        const double tn = __fma_rz(a.v[i], b.v[j], t);
        s += __fma_rn(a.v[i], b.v[j], t - tn);
        t = tn;
      }
    }
    c[k] = s + hi_prev;
    hi_prev = (t - F6_BIAS) * F6_2M44;
  }
  c[11] = hi_prev;
  return f6_fold(c);
}

// a^2 for a normal a: 21 products, off-diagonal terms once with a doubled operand (model: sqr6)
__device__ __forceinline__ fe6 fe6_sqr(const fe6 &a) {
  double a2[6], c[12];
#pragma unroll
  for (int i = 0; i < 6; ++i) a2[i] = a.v[i] + a.v[i];
  double hi_prev = 0.0;
#pragma unroll
  for (int k = 0; k < 11; ++k) {
    double t = F6_BIAS, s = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int j = k - i;
      if (j >= 0 && j < 6 && i <= j) {
        const double x = i == j ? a.v[i] : a2[i];
        const double tn = __fma_rz(x, a.v[j], t);
        s += __fma_rn(x, a.v[j], t - tn);
        t = tn;
      }
    }
    c[k] = s + hi_prev;
    hi_prev = (t - F6_BIAS) * F6_2M44;
  }
  c[11] = hi_prev;
  return f6_fold(c);
}

// a - b + (a multiple of p whose limbs are all >= 2^45): no negative limb for normal a, b; the result is "wide"
// (limbs < 2^47.3). The bias is 1024 p (model: bias_limbs(45)).
__device__ __forceinline__ fe6 fe6_sub(const fe6 &a, const fe6 &b) {
  const double bias[6] = {F6_SUB_BIAS0, F6_SUB_BIAS1, F6_SUB_BIAS2, F6_SUB_BIAS3, F6_SUB_BIAS4, F6_SUB_BIAS5};
  fe6 r;
#pragma unroll
  for (int i = 0; i < 6; ++i) r.v[i] = (a.v[i] - b.v[i]) + bias[i];
  return r;
}
__device__ __forceinline__ fe6 fe6_norm(const fe6 &a) {
  double c[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) c[i] = a.v[i];
  f6_carry(c);
  fe6 r;
#pragma unroll
  for (int i = 0; i < 6; ++i) r.v[i] = c[i];
  return r;
}

// batch_add's body (main.c:378-386) in limb form: P + Q with inv = 1/(qx - px) (model: affine_add6)
__device__ __forceinline__ void fe6_affine_add(fe6 &rx, fe6 &ry, const fe6 &px, const fe6 &py, const fe6 &qx, const fe6 &qy,
                                               const fe6 &inv) {
  const fe6 lam = fe6_mul(fe6_sub(qy, py), inv);
  rx = fe6_norm(fe6_sub(fe6_sub(fe6_sqr(lam), px), qx));
  ry = fe6_norm(fe6_sub(fe6_mul(fe6_sub(px, rx), lam), py));
}

// ---------------------------------------------------------------- throughput of the two multiplications
// Each thread runs two dependent multiplication chains plus FILL independent LOP3/SHF steps per multiplication (the
// ALU-pipe load the hashes put beside the field arithmetic in the fused kernel: ~380 ALU instructions per fe_mul).
// KIND 0: fe_mul (IMAD.WIDE), KIND 1: fe6_mul (DFMA). Launched like the add kernel: 512 threads per SM.
#define MULBENCH_ITERS 512
template <int KIND, int FILL>
__global__ void __launch_bounds__(512, 1) mulbench_kernel(u32 *out, u32 seed) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  fe a0, a1, b;
#pragma unroll
  for (int i = 0; i < 8; ++i) a0.v[i] = seed * (i + 1) + t, a1.v[i] = seed * (i + 9) ^ t, b.v[i] = seed * (i + 17) + 3 * t;
  u32 f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = seed + i * t;
  fe6 d0 = fe6_from_fe(a0), d1 = fe6_from_fe(a1), db = fe6_from_fe(b);
  // Odd warps start half an iteration late (one filler block first): all warps of an SM run the same deterministic
  // code, so without this every warp does its multiplications at the same moment and the pipes take turns instead
  // of overlapping (the two-phase effect DESIGN.md describes for K1).
  const bool late = FILL > 0 && ((threadIdx.x >> 5) & 1);
#pragma unroll 1
  for (int it = 0; it < MULBENCH_ITERS + 1; ++it) {
    if (it < MULBENCH_ITERS && !(late && it == 0)) {
      if (KIND == 0) a0 = fe_mul(a0, b), a1 = fe_mul(a1, b);
      else d0 = fe6_mul(d0, db), d1 = fe6_mul(d1, db);
    } else if (late && it == MULBENCH_ITERS) {
      if (KIND == 0) a0 = fe_mul(a0, b), a1 = fe_mul(a1, b);
      else d0 = fe6_mul(d0, db), d1 = fe6_mul(d1, db);
    }
    if (it < MULBENCH_ITERS) {
#pragma unroll
      for (int k = 0; k < 2 * FILL / 8; ++k) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (k & 1) asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(f[i]));
          else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(f[i]) : "r"(f[(i + 1) & 7]), "r"(seed));
        }
      }
    }
  }
  if (KIND == 1) a0 = fe6_to_fe(d0), a1 = fe6_to_fe(d1);
  u32 acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= a0.v[i] ^ a1.v[i] ^ f[i];
  if (acc == 0x12345678u) out[t & 1023] = acc;
}
