// peak.cuh — integer-pipe throughput microbenchmarks (build-plan step 0b).
// The fused add kernel is bound by the SM's integer ALU pipe (LOP3 / SHF / IADD3), not by HBM or the tensor
// cores, so the roofline denominator has to be measured: each kernel below issues ITER x UNROLL x 8 independent
// chains of ONE instruction kind from every resident warp of every SM; ops/s = lanes * instructions / time.
// The instruction mix is pinned with inline PTX that ptxas maps 1:1 (lop3.b32 -> LOP3.LUT, shf.l.wrap -> SHF,
// mad.lo.u32 -> IMAD, mad.wide.u32 -> IMAD.WIDE.U32); profiles/ holds the SASS check.
#pragma once
#include <stdint.h>

#define PEAK_ITERS 4096
#define PEAK_CHAINS 8
#define PEAK_UNROLL 4

static __constant__ uint32_t peak_k_one = 1u;  // same trick as hash160.cuh: an opaque multiplicand in constant memory

template <int KIND>
__global__ void __launch_bounds__(256) peak_kernel(uint32_t *out, uint32_t seed, unsigned long long *cycles) {
  uint32_t r[PEAK_CHAINS];
  uint64_t q[PEAK_CHAINS];
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int c = 0; c < PEAK_CHAINS; ++c) r[c] = seed * (c + 1) + t, q[c] = (uint64_t)r[c] << 7;
  const uint32_t m = seed | 1u, z = seed ^ 0x9e3779b9u;
  double f[PEAK_CHAINS];  // FP64 chains (kinds 16-18): is the DFMA pipe a third lane beside ALU and FMA?
#pragma unroll
  for (int c = 0; c < PEAK_CHAINS; ++c) f[c] = 1.0 + (double)(r[c] & 1023u) * 1e-9;
  const double fm = 1.0 + (double)(seed & 255u) * 1e-12, fz = (double)(seed & 15u) * 1e-15;
  unsigned long long g0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  const long long c0 = clock64();
#pragma unroll 1
  for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < PEAK_UNROLL; ++u) {
#pragma unroll
      for (int c = 0; c < PEAK_CHAINS; ++c) {
        if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(m), "r"(z));
        if (KIND == 1) asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(r[c]) : "r"(m), "r"(z));  // one IADD3
        if (KIND == 2) asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(r[c]));
        if (KIND == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(m), "r"(z));
        if (KIND == 4) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(q[c]) : "r"(r[c]), "r"(m));
        if (KIND == 5) {  // half the chains on the ALU pipe, half on the FMA pipe: do they co-issue?
          if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(m), "r"(z));
          else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(m), "r"(z));
        }
        // ---- the instruction forms the hashes steer onto the FMA pipe (hash160.cuh)
        if (KIND == 6) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(peak_k_one), "r"(z));  // IMAD, constant-bank operand
        if (KIND == 7) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(r[c]) : "r"(m));                      // IMAD.HI
        if (KIND == 8) {  // LOP3 + IMAD(const) co-issue
          if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(m), "r"(z));
          else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(peak_k_one), "r"(z));
        }
        if (KIND == 9) {  // SHF + IMAD.WIDE co-issue
          if (c & 1) asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(r[c]));
          else asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(q[c]) : "r"(r[c]), "r"(m));
        }
        if (KIND == 10) {  // LOP3 + IMAD.HI co-issue
          if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(m), "r"(z));
          else asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(r[c]) : "r"(m));
        }
        if (KIND == 11) {  // the hash kernels' ratio: 5 ALU-pipe : 3 FMA-pipe
          if (c < 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(m), "r"(z));
          else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(peak_k_one), "r"(z));
        }
        if (KIND == 13) {  // LOP3 + IMAD.WIDE: does the 64-bit multiply-add take an ALU-pipe slot as well?
          if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(m), "r"(z));
          else asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(q[c]) : "r"(r[c]), "r"(m));
        }
        if (KIND == 14) {  // SHF + IMAD
          if (c & 1) asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(r[c]));
          else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(m), "r"(z));
        }
        if (KIND == 15) {  // LOP3 + SHF (both ALU pipe: no gain expected)
          if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(m), "r"(z));
          else asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(r[c]));
        }
        if (KIND == 16) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[c]) : "d"(fm), "d"(fz));  // DFMA
        if (KIND == 17) {  // DFMA + LOP3
          if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(m), "r"(z));
          else asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[c]) : "d"(fm), "d"(fz));
        }
        if (KIND == 18) {  // DFMA + IMAD
          if (c & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(m), "r"(z));
          else asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[c]) : "d"(fm), "d"(fz));
        }
        if (KIND == 12) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[c]) : "r"(z));  // 2-input add: IADD3 or IMAD.IADD, ptxas decides
      }
    }
  }
  const long long c1 = clock64();
  unsigned long long g1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < PEAK_CHAINS; ++c) acc ^= r[c] ^ (uint32_t)q[c] ^ (uint32_t)(q[c] >> 32) ^ (uint32_t)__double2ll_rn(f[c] * 1e6);
  if (acc == 0x12345678u) out[t & 1023] = acc;  // keep the chains alive
  if (t == 0) cycles[0] = (unsigned long long)(c1 - c0), cycles[1] = g1 - g0;  // SM cycles and nanoseconds of this thread's loop
}
