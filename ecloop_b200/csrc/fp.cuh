// fp.cuh — secp256k1 base-field arithmetic for sm_100a, 8 x 32-bit limbs held in registers.
//
// Replaces the reference's fe_modp_* (lib/ecc.c:269-540). Memory image of an `fe` is identical to the
// reference's `fe` (4 x u64 little-endian == 8 x u32 little-endian), so host buffers pass through unchanged.
//
// Design notes (B200): the SM has no 64-bit integer multiplier; PTX mad.{lo,hi}.u64 is lowered to 32-bit
// IMADs plus ALU-pipe carry fix-ups. The hot kernel is bound by the ALU pipe (LOP3/SHF/IADD3 of the two
// hashes), so the multiplier is written to live on the FMA pipe: 32x32->64 IMAD.WIDE.U32 with the carry
// riding the predicate (.X form), which ptxas emits for {mad.lo.cc, madc.hi.cc} pairs on an aligned
// register pair. Even and odd columns are accumulated separately so that every pair is aligned.
//
// All results are canonical (in [0,p)) for canonical inputs, like the reference; fe_mul/fe_sqr accept ANY
// 256-bit inputs and still return the canonical residue.
#pragma once
#include <stdint.h>

typedef uint32_t u32;
typedef uint64_t u64;

struct fe {
  u32 v[8];
};

// p = 2^256 - 2^32 - 977 (lib/ecc.c:32)
#define FP_P0 0xfffffc2fu
#define FP_P1 0xfffffffeu
#define FP_C0 977u  // 2^256 - p = 2^32 + 977

__device__ __forceinline__ fe fe_zero() {
  fe r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = 0;
  return r;
}
__device__ __forceinline__ fe fe_one() {
  fe r = fe_zero();
  r.v[0] = 1;
  return r;
}
__device__ __forceinline__ bool fe_is_zero(const fe &a) {
  return (a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7]) == 0;
}
__device__ __forceinline__ bool fe_eq(const fe &a, const fe &b) {
  u32 d = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) d |= a.v[i] ^ b.v[i];
  return d == 0;
}

// r = a - b (mod p), canonical inputs -> canonical output (lib/ecc.c:277-290).
// On borrow the wrapped value is a-b+2^256; subtracting 2^256-p = 2^32+977 from it gives a-b+p.
__device__ __forceinline__ fe fe_sub(const fe &a, const fe &b) {
  fe r;
  u32 bw;
  asm("sub.cc.u32  %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32    %8, 0, 0;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
        "=r"(r.v[7]), "=r"(bw)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  // bw = 0 or 0xffffffff
  const u32 c0 = bw & FP_C0, c1 = bw & 1u;
  asm("sub.cc.u32  %0, %0, %8;\n\t"
      "subc.cc.u32 %1, %1, %9;\n\t"
      "subc.cc.u32 %2, %2, 0;\n\t"
      "subc.cc.u32 %3, %3, 0;\n\t"
      "subc.cc.u32 %4, %4, 0;\n\t"
      "subc.cc.u32 %5, %5, 0;\n\t"
      "subc.cc.u32 %6, %6, 0;\n\t"
      "subc.u32    %7, %7, 0;"
      : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
        "+r"(r.v[7])
      : "r"(c0), "r"(c1));
  return r;
}

// true iff a >= p (a is any 256-bit value)
__device__ __forceinline__ bool fe_ge_p_full(const fe &a) {
  if ((a.v[7] & a.v[6] & a.v[5] & a.v[4] & a.v[3] & a.v[2]) != 0xffffffffu) return false;
  if (a.v[1] == 0xffffffffu) return true;
  if (a.v[1] == FP_P1) return a.v[0] >= FP_P0;
  return false;
}
// a -= p (only called when a >= p): a - p = a + (2^32 + 977) - 2^256
__device__ __forceinline__ void fe_sub_p(fe &a) {
  asm("add.cc.u32  %0, %0, 977;\n\t"
      "addc.cc.u32 %1, %1, 1;\n\t"
      "addc.cc.u32 %2, %2, 0;\n\t"
      "addc.cc.u32 %3, %3, 0;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t"
      "addc.cc.u32 %5, %5, 0;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\t"
      "addc.u32    %7, %7, 0;"
      : "+r"(a.v[0]), "+r"(a.v[1]), "+r"(a.v[2]), "+r"(a.v[3]), "+r"(a.v[4]), "+r"(a.v[5]), "+r"(a.v[6]),
        "+r"(a.v[7]));
}
__device__ __forceinline__ void fe_canon(fe &a) {
  if (a.v[7] == 0xffffffffu) {  // rare path: only values within 2^224 of 2^256 can be >= p
    if (fe_ge_p_full(a)) fe_sub_p(a);
  }
}

// Branch-free final correction of a product (ECL_FE_BRANCHFREE, used by the pipelined add kernel): r is the low 256 bits,
// cy the carry out of the second fold. Both rare cases — cy = 1 (the value is 2^256 + r, r < 2^66) and r >= p — are
// fixed by the same step, r += 2^32 + 977 (mod 2^256), applied under a mask instead of behind a branch: a branch ends the
// basic block, and ptxas only interleaves the hash (ALU pipe) with the field arithmetic (FMA pipe) INSIDE a basic
// block. 17 ALU-pipe instructions per product instead of ~4, paid for by the overlap.
#ifndef ECL_FE_BRANCHFREE
#define ECL_FE_BRANCHFREE 0
#endif
// ECL_FE_NC = 1: fe_mul_nc really skips the canonical correction (12 instructions less per such product). Measured
// (r02_f): no gain on the addr33 instance with the filter in shared memory (6 566 vs 6 583 Mkeys/s) and -9 % on its HBM
// twin (ptxas schedules the 7 000-instruction loop differently), so it is off; fe_mul_nc is then fe_mul.
#ifndef ECL_FE_NC
#define ECL_FE_NC 0
#endif
__device__ __forceinline__ void fe_fix_branchfree(fe &r, u32 cy) {
  const u32 hi = r.v[7] & r.v[6] & r.v[5] & r.v[4] & r.v[3] & r.v[2];
  const u64 lo = (u64)r.v[1] << 32 | r.v[0];
  const u32 ge = (hi == 0xffffffffu) & (lo >= (((u64)FP_P1 << 32) | FP_P0));
  const u32 m = 0u - (ge | cy);  // all ones: add 2^32 + 977
  const u32 c0 = m & FP_C0, c1 = m & 1u;
  asm("add.cc.u32  %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, 0;\n\t"
      "addc.cc.u32 %3, %3, 0;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t"
      "addc.cc.u32 %5, %5, 0;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\t"
      "addc.u32    %7, %7, 0;"
      : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7])
      : "r"(c0), "r"(c1));
}

// r = a + b (mod p), canonical inputs -> canonical output (the reference's fe_modp_add, lib/ecc.c:292-305,
// only reduces on a 2^256 carry; on the hot path it is never used, we return the canonical residue).
__device__ __forceinline__ fe fe_add(const fe &a, const fe &b) {
  fe r;
  u32 cy;
  asm("add.cc.u32  %0, %9, %17;\n\t"
      "addc.cc.u32 %1, %10, %18;\n\t"
      "addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\t"
      "addc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\t"
      "addc.cc.u32 %7, %16, %24;\n\t"
      "addc.u32    %8, 0, 0;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
        "=r"(r.v[7]), "=r"(cy)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  if (cy) fe_sub_p(r);  // a+b-2^256+C = a+b-p, and it is < p because a+b < 2p
  else fe_canon(r);
  return r;
}

// r = -a (mod p) for canonical a; neg(0) = 0 (the reference returns p for 0, lib/ecc.c:269-275 — never hit:
// it is only applied to y coordinates of curve points, and y = 0 is not on the curve)
__device__ __forceinline__ fe fe_neg(const fe &a) {
  fe p;
  p.v[0] = FP_P0, p.v[1] = FP_P1;
#pragma unroll
  for (int i = 2; i < 8; ++i) p.v[i] = 0xffffffffu;
  fe r;
  asm("sub.cc.u32  %0, %8, %16;\n\t"
      "subc.cc.u32 %1, %9, %17;\n\t"
      "subc.cc.u32 %2, %10, %18;\n\t"
      "subc.cc.u32 %3, %11, %19;\n\t"
      "subc.cc.u32 %4, %12, %20;\n\t"
      "subc.cc.u32 %5, %13, %21;\n\t"
      "subc.cc.u32 %6, %14, %22;\n\t"
      "subc.u32    %7, %15, %23;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
        "=r"(r.v[7])
      : "r"(p.v[0]), "r"(p.v[1]), "r"(p.v[2]), "r"(p.v[3]), "r"(p.v[4]), "r"(p.v[5]), "r"(p.v[6]), "r"(p.v[7]),
        "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]));
  if (fe_is_zero(a)) r = a;
  return r;
}

// r = p - a for 0 < a < p (coordinates of curve points: y = 0 is not on the curve): no zero test
__device__ __forceinline__ fe fe_neg_nz(const fe &a) {
  fe r;
  asm("sub.cc.u32  %0, %8, %9;\n\t"
      "subc.cc.u32 %1, %10, %11;\n\t"
      "subc.cc.u32 %2, -1, %12;\n\t"
      "subc.cc.u32 %3, -1, %13;\n\t"
      "subc.cc.u32 %4, -1, %14;\n\t"
      "subc.cc.u32 %5, -1, %15;\n\t"
      "subc.cc.u32 %6, -1, %16;\n\t"
      "subc.u32    %7, -1, %17;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
      : "r"(FP_P0), "r"(a.v[0]), "r"(FP_P1), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]));
  return r;
}

// acc[0..7] += {a0,a1,a2,a3} * b, the four 64-bit products sitting side by side (an aligned "row"); the carry
// out of acc[7] is added into acc[8]. Each {mad.lo.cc, madc.hi.cc} pair becomes one IMAD.WIDE.U32(.X) and the
// trailing addc one IADD3.X, so a row is 4 FMA-pipe + 1 ALU-pipe instructions.
__device__ __forceinline__ void fp_mad_row_c(u32 *acc, u32 a0, u32 a1, u32 a2, u32 a3, u32 b) {
  asm("mad.lo.cc.u32  %0, %9, %13, %0;\n\t"
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32       %8, %8, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
        "+r"(acc[7]), "+r"(acc[8])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
// same, for a row whose top limb is the top of the whole product (no carry can leave it)
__device__ __forceinline__ void fp_mad_row_top(u32 *acc, u32 a0, u32 a1, u32 a2, u32 a3, u32 b) {
  asm("mad.lo.cc.u32  %0, %8, %12, %0;\n\t"
      "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
      "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
      "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
      "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
      "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
      "madc.hi.u32    %7, %11, %12, %7;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
        "+r"(acc[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
// returns the carry instead of adding it somewhere (used by the reduction)
__device__ __forceinline__ u32 fp_mad_row(u32 *acc, u32 a0, u32 a1, u32 a2, u32 a3, u32 b) {
  u32 c = 0;
  u32 tmp[9];
#pragma unroll
  for (int k = 0; k < 8; ++k) tmp[k] = acc[k];
  tmp[8] = 0;
  fp_mad_row_c(tmp, a0, a1, a2, a3, b);
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = tmp[k];
  c = tmp[8];
  return c;
}

// t[0..15] = a * b  (schoolbook, 64 IMAD.WIDE). Position k of the even accumulator e is limb k; position k of
// the odd accumulator o is limb k+1. Rows are issued in order of b[i]; a row's carry lands in the limb just
// above it, which only ever holds such carries until the next row of the same parity covers it.
__device__ __forceinline__ void fp_mul_wide(u32 t[16], const fe &a, const fe &b) {
  u32 e[16], o[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) e[k] = 0, o[k] = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    {  // products a[j]*b[i] with i+j even -> limbs i+j, i+j+1 of e
      const int j0 = i & 1, s = i + j0;
      if (s + 8 < 16) fp_mad_row_c(e + s, a.v[j0], a.v[j0 + 2], a.v[j0 + 4], a.v[j0 + 6], b.v[i]);
      else fp_mad_row_top(e + s, a.v[j0], a.v[j0 + 2], a.v[j0 + 4], a.v[j0 + 6], b.v[i]);
    }
    {  // products with i+j odd -> limbs i+j, i+j+1 = o[i+j-1], o[i+j]
      const int j0 = (i + 1) & 1, s = i + j0 - 1;
      if (s + 8 < 16) fp_mad_row_c(o + s, a.v[j0], a.v[j0 + 2], a.v[j0 + 4], a.v[j0 + 6], b.v[i]);
      else fp_mad_row_top(o + s, a.v[j0], a.v[j0 + 2], a.v[j0 + 4], a.v[j0 + 6], b.v[i]);
    }
  }
  // t = e + (o << 32)
  t[0] = e[0];
  asm("add.cc.u32  %0, %15, %30;\n\t"
      "addc.cc.u32 %1, %16, %31;\n\t"
      "addc.cc.u32 %2, %17, %32;\n\t"
      "addc.cc.u32 %3, %18, %33;\n\t"
      "addc.cc.u32 %4, %19, %34;\n\t"
      "addc.cc.u32 %5, %20, %35;\n\t"
      "addc.cc.u32 %6, %21, %36;\n\t"
      "addc.cc.u32 %7, %22, %37;\n\t"
      "addc.cc.u32 %8, %23, %38;\n\t"
      "addc.cc.u32 %9, %24, %39;\n\t"
      "addc.cc.u32 %10, %25, %40;\n\t"
      "addc.cc.u32 %11, %26, %41;\n\t"
      "addc.cc.u32 %12, %27, %42;\n\t"
      "addc.cc.u32 %13, %28, %43;\n\t"
      "addc.u32    %14, %29, %44;"
      : "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]), "=r"(t[9]),
        "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
      : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]), "r"(e[9]),
        "r"(e[10]), "r"(e[11]), "r"(e[12]), "r"(e[13]), "r"(e[14]), "r"(e[15]), "r"(o[0]), "r"(o[1]), "r"(o[2]),
        "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]),
        "r"(o[12]), "r"(o[13]), "r"(o[14]));
}

// t (512 bit) mod p, canonical. 2^256 = 2^32 + 977 (mod p): fold the high half twice (lib/ecc.c:331-346),
// then the rare final corrections. Unlike ecc.c:341-344 the carry out of the second fold is honoured.
template <int FIX>  // 0: canonical result (branchy or branch-free per ECL_FE_BRANCHFREE); 1: any representative < 2^256;
                    // 2: canonical, always branch-free (kernels that want independent products in one basic block)
__device__ __forceinline__ fe fp_reduce512_t(const u32 t[16]) {
  // a[0..9] = lo + hi*977 + (hi << 32)
  u32 a[10], o[9];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = t[k], o[k] = t[8 + k];  // odd-row addends carry the (hi << 32) term
  a[8] = fp_mad_row(a, t[8], t[10], t[12], t[14], FP_C0);      // limbs 0..7 (+ carry -> limb 8)
  o[8] = fp_mad_row(o, t[9], t[11], t[13], t[15], FP_C0);      // limbs 1..8 (+ carry -> limb 9)
  asm("add.cc.u32  %0, %0, %9;\n\t"
      "addc.cc.u32 %1, %1, %10;\n\t"
      "addc.cc.u32 %2, %2, %11;\n\t"
      "addc.cc.u32 %3, %3, %12;\n\t"
      "addc.cc.u32 %4, %4, %13;\n\t"
      "addc.cc.u32 %5, %5, %14;\n\t"
      "addc.cc.u32 %6, %6, %15;\n\t"
      "addc.cc.u32 %7, %7, %16;\n\t"
      "addc.u32    %8, %17, 0;"
      : "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "=r"(a[9])
      : "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]));
  // second fold: H = a[8] + a[9]*2^32 < 2^34;  q = H * (2^32 + 977) < 2^67
  const u64 w = (u64)a[8] * FP_C0 + ((u64)(a[9] * FP_C0) << 32);  // H*977 (a[9] <= 3: no overflow)
  const u32 w0 = (u32)w, w1 = (u32)(w >> 32);
  fe r;
  u32 cy;
  asm("add.cc.u32  %0, %9, %17;\n\t"   // + q0 = w0
      "addc.cc.u32 %1, %10, %18;\n\t"  // + w1
      "addc.cc.u32 %2, %11, 0;\n\t"
      "addc.cc.u32 %3, %12, 0;\n\t"
      "addc.cc.u32 %4, %13, 0;\n\t"
      "addc.cc.u32 %5, %14, 0;\n\t"
      "addc.cc.u32 %6, %15, 0;\n\t"
      "addc.cc.u32 %7, %16, 0;\n\t"
      "addc.u32    %8, 0, 0;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
        "=r"(r.v[7]), "=r"(cy)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(w0), "r"(w1));
  u32 cy2;
  asm("add.cc.u32  %0, %0, %7;\n\t"  // + (H << 32): limb 1 += a[8], limb 2 += a[9]
      "addc.cc.u32 %1, %1, %8;\n\t"
      "addc.cc.u32 %2, %2, 0;\n\t"
      "addc.cc.u32 %3, %3, 0;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t"
      "addc.cc.u32 %5, %5, 0;\n\t"
      "addc.u32    %6, 0, 0;"
      : "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "=r"(cy2)
      : "r"(a[8]), "r"(a[9]));
  // propagate cy2 into limb 7 (kept out of the asm above to stay within operand limits)
  const u32 r7 = r.v[7] + cy2;
  cy += (r7 < cy2);
  r.v[7] = r7;
  constexpr bool branchfree = (FIX == 2) || ECL_FE_BRANCHFREE;
  if (branchfree) {
    if (FIX == 1 && ECL_FE_NC) {  // only the wrap past 2^256 (then r < 2^66): the three low limbs take 2^32 + 977
      const u32 c0 = (0u - cy) & FP_C0;
      asm("add.cc.u32  %0, %0, %3;\n\t"
          "addc.cc.u32 %1, %1, %4;\n\t"
          "addc.u32    %2, %2, 0;"
          : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2])
          : "r"(c0), "r"(cy));
    } else {
      fe_fix_branchfree(r, cy);
    }
  } else {
    if (cy) fe_sub_p(r);  // value wrapped past 2^256 (prob ~2^-190): add 2^32+977; cannot carry again
    fe_canon(r);
  }
  return r;
}
__device__ __forceinline__ fe fp_reduce512(const u32 t[16]) { return fp_reduce512_t<0>(t); }

__device__ __forceinline__ fe fe_mul(const fe &a, const fe &b) {
  u32 t[16];
  fp_mul_wide(t, a, b);
  return fp_reduce512(t);
}
// canonical a * b / a^2 with the branch-free correction whatever ECL_FE_BRANCHFREE says: K2a (kernels.cuh) is bound by the
// latency of the carry chains, and only products inside one basic block can overlap
__device__ __forceinline__ fe fe_mul_bf(const fe &a, const fe &b) {
  u32 t[16];
  fp_mul_wide(t, a, b);
  return fp_reduce512_t<2>(t);
}
// a * b as ANY representative of the residue below 2^256 (possibly >= p, with probability 2^-192): for products that only
// feed further multiplications (which accept any 256-bit operand). Saves the canonical correction where the
// branch-free form makes it cost 17 instructions; identical to fe_mul otherwise.
__device__ __forceinline__ fe fe_mul_nc(const fe &a, const fe &b) {
  u32 t[16];
  fp_mul_wide(t, a, b);
  return fp_reduce512_t<1>(t);
}
// a == 0 (mod p) for a representative below 2^256: 0 or p
__device__ __forceinline__ bool fe_is_zero_modp(const fe &a) {
  const u32 all = a.v[7] & a.v[6] & a.v[5] & a.v[4] & a.v[3] & a.v[2];
  return fe_is_zero(a) || (all == 0xffffffffu && a.v[1] == FP_P1 && a.v[0] == FP_P0);
}

// rows of 3, 2 and 1 products for the squaring (same conventions as fp_mad_row_c: aligned pairs, carry into the limb above)
__device__ __forceinline__ void fp_mad_row3(u32 *acc, u32 a0, u32 a1, u32 a2, u32 b) {
  asm("mad.lo.cc.u32  %0, %7, %10, %0;\n\t"
      "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
      "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
      "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
      "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
      "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
      "addc.u32       %6, %6, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6])
      : "r"(a0), "r"(a1), "r"(a2), "r"(b));
}
__device__ __forceinline__ void fp_mad_row2(u32 *acc, u32 a0, u32 a1, u32 b) {
  asm("mad.lo.cc.u32  %0, %5, %7, %0;\n\t"
      "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
      "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
      "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
      "addc.u32       %4, %4, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4])
      : "r"(a0), "r"(a1), "r"(b));
}
__device__ __forceinline__ void fp_mad_row1(u32 *acc, u32 a0, u32 b) {
  asm("mad.lo.cc.u32  %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32       %2, %2, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2])
      : "r"(a0), "r"(b));
}

// t[0..15] = a^2 (lib/ecc.c:349-444 computes each off-diagonal product once as well): the 28 products a[i]*a[j], i < j,
// in 13 aligned rows (even columns in e, odd columns in o, like fp_mul_wide; a row's carry lands in a limb that
// holds nothing but such carries until a later row covers it), then 2*(e + (o << 32)) + the 8 squares.
// 36 IMAD.WIDE + 44 carry-chain adds instead of 64 + 31: IMAD.WIDE occupies an issue slot on BOTH integer pipes of
// sm_100, so this is 15 ALU-pipe slots and 28 FMA-pipe slots less per squaring.
__device__ __forceinline__ void fp_sqr_wide(u32 t[16], const fe &a) {
  u32 e[16], o[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) e[k] = 0, o[k] = 0;
  fp_mad_row3(e + 2, a.v[2], a.v[4], a.v[6], a.v[0]);
  fp_mad_row3(e + 4, a.v[3], a.v[5], a.v[7], a.v[1]);
  fp_mad_row2(e + 6, a.v[4], a.v[6], a.v[2]);
  fp_mad_row2(e + 8, a.v[5], a.v[7], a.v[3]);
  fp_mad_row1(e + 10, a.v[6], a.v[4]);
  fp_mad_row1(e + 12, a.v[7], a.v[5]);
  fp_mad_row_c(o + 0, a.v[1], a.v[3], a.v[5], a.v[7], a.v[0]);
  fp_mad_row3(o + 2, a.v[2], a.v[4], a.v[6], a.v[1]);
  fp_mad_row3(o + 4, a.v[3], a.v[5], a.v[7], a.v[2]);
  fp_mad_row2(o + 6, a.v[4], a.v[6], a.v[3]);
  fp_mad_row2(o + 8, a.v[5], a.v[7], a.v[4]);
  fp_mad_row1(o + 10, a.v[6], a.v[5]);
  fp_mad_row1(o + 12, a.v[7], a.v[6]);
  // m = e + (o << 32): limbs 0 and 1 of e are zero, so m[0] = 0, m[1] = o[0]
  u32 m[16];
  m[0] = 0, m[1] = o[0];
  asm("add.cc.u32  %0, %14, %28;\n\t"
      "addc.cc.u32 %1, %15, %29;\n\t"
      "addc.cc.u32 %2, %16, %30;\n\t"
      "addc.cc.u32 %3, %17, %31;\n\t"
      "addc.cc.u32 %4, %18, %32;\n\t"
      "addc.cc.u32 %5, %19, %33;\n\t"
      "addc.cc.u32 %6, %20, %34;\n\t"
      "addc.cc.u32 %7, %21, %35;\n\t"
      "addc.cc.u32 %8, %22, %36;\n\t"
      "addc.cc.u32 %9, %23, %37;\n\t"
      "addc.cc.u32 %10, %24, %38;\n\t"
      "addc.cc.u32 %11, %25, %39;\n\t"
      "addc.cc.u32 %12, %26, %40;\n\t"
      "addc.u32    %13, %27, %41;"
      : "=r"(m[2]), "=r"(m[3]), "=r"(m[4]), "=r"(m[5]), "=r"(m[6]), "=r"(m[7]), "=r"(m[8]), "=r"(m[9]), "=r"(m[10]), "=r"(m[11]),
        "=r"(m[12]), "=r"(m[13]), "=r"(m[14]), "=r"(m[15])
      : "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]), "r"(e[9]), "r"(e[10]), "r"(e[11]), "r"(e[12]),
        "r"(e[13]), "r"(e[14]), "r"(e[15]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]),
        "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]));
  // d = m + the squares a[i]^2 at limbs 2i, 2i+1 (one carry chain through all 8 products)
  u32 d[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) d[k] = m[k];
  asm("mad.lo.cc.u32  %0, %16, %16, %0;\n\t"
      "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
      "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
      "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
      "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
      "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
      "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
      "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
      "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
      "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
      "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
      "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
      "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
      "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
      "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
      "madc.hi.u32    %15, %23, %23, %15;"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(d[4]), "+r"(d[5]), "+r"(d[6]), "+r"(d[7]), "+r"(d[8]), "+r"(d[9]),
        "+r"(d[10]), "+r"(d[11]), "+r"(d[12]), "+r"(d[13]), "+r"(d[14]), "+r"(d[15])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]));
  // t = d + m  (the second copy of the off-diagonal part); limb 0 of m is zero
  t[0] = d[0];
  asm("add.cc.u32  %0, %15, %30;\n\t"
      "addc.cc.u32 %1, %16, %31;\n\t"
      "addc.cc.u32 %2, %17, %32;\n\t"
      "addc.cc.u32 %3, %18, %33;\n\t"
      "addc.cc.u32 %4, %19, %34;\n\t"
      "addc.cc.u32 %5, %20, %35;\n\t"
      "addc.cc.u32 %6, %21, %36;\n\t"
      "addc.cc.u32 %7, %22, %37;\n\t"
      "addc.cc.u32 %8, %23, %38;\n\t"
      "addc.cc.u32 %9, %24, %39;\n\t"
      "addc.cc.u32 %10, %25, %40;\n\t"
      "addc.cc.u32 %11, %26, %41;\n\t"
      "addc.cc.u32 %12, %27, %42;\n\t"
      "addc.cc.u32 %13, %28, %43;\n\t"
      "addc.u32    %14, %29, %44;"
      : "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]), "=r"(t[9]), "=r"(t[10]),
        "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
      : "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(d[7]), "r"(d[8]), "r"(d[9]), "r"(d[10]), "r"(d[11]),
        "r"(d[12]), "r"(d[13]), "r"(d[14]), "r"(d[15]), "r"(m[1]), "r"(m[2]), "r"(m[3]), "r"(m[4]), "r"(m[5]), "r"(m[6]), "r"(m[7]),
        "r"(m[8]), "r"(m[9]), "r"(m[10]), "r"(m[11]), "r"(m[12]), "r"(m[13]), "r"(m[14]), "r"(m[15]));
}

#ifndef ECL_FE_SQR_DEDICATED
#define ECL_FE_SQR_DEDICATED 1  // 0: fe_sqr = fe_mul(a, a) (the round-1 form, kept for A/B timing of the big kernels)
#endif
__device__ __forceinline__ fe fe_sqr(const fe &a) {
#if ECL_FE_SQR_DEDICATED
  u32 t[16];
  fp_sqr_wide(t, a);
  return fp_reduce512(t);
#else
  return fe_mul(a, a);
#endif
}

__device__ __forceinline__ fe fe_sqr_bf(const fe &a) {
  u32 t[16];
  fp_sqr_wide(t, a);
  return fp_reduce512_t<2>(t);
}

static __device__ __noinline__ fe fe_sqr_n(fe x, int n) {
#pragma unroll 1
  for (int i = 0; i < n; ++i) x = fe_sqr(x);
  return x;
}
static __device__ __noinline__ fe fe_mul_noinline(const fe &a, const fe &b) { return fe_mul(a, b); }

// a^(p-2) (Fermat), 255 squarings + 15 multiplications along the standard x2,x3,x6,x9,x11,x22,x44,x88,
// x176,x220,x223 chain for p (same exponent the reference uses, lib/ecc.c:463-518). inv(0) = 0.
static __device__ __noinline__ fe fe_inv(const fe &a) {
  fe x2 = fe_mul_noinline(fe_sqr_n(a, 1), a);
  fe x3 = fe_mul_noinline(fe_sqr_n(x2, 1), a);
  fe x6 = fe_mul_noinline(fe_sqr_n(x3, 3), x3);
  fe x9 = fe_mul_noinline(fe_sqr_n(x6, 3), x3);
  fe x11 = fe_mul_noinline(fe_sqr_n(x9, 2), x2);
  fe x22 = fe_mul_noinline(fe_sqr_n(x11, 11), x11);
  fe x44 = fe_mul_noinline(fe_sqr_n(x22, 22), x22);
  fe x88 = fe_mul_noinline(fe_sqr_n(x44, 44), x44);
  fe x176 = fe_mul_noinline(fe_sqr_n(x88, 88), x88);
  fe x220 = fe_mul_noinline(fe_sqr_n(x176, 44), x44);
  fe x223 = fe_mul_noinline(fe_sqr_n(x220, 3), x3);
  // p-2 = 2^256 - 2^32 - 979: 223 ones, 0, 22 ones, 0000, 1, 0, 11, 0, 1  (binary, msb first)
  fe t = fe_mul_noinline(fe_sqr_n(x223, 23), x22);
  t = fe_mul_noinline(fe_sqr_n(t, 5), a);
  t = fe_mul_noinline(fe_sqr_n(t, 3), x2);
  t = fe_mul_noinline(fe_sqr_n(t, 2), a);
  return t;
}
