// common.cuh — pieces shared by the add and mul kernels: hit ring, mbarrier/TMA bulk-copy helpers, and the
// hash + bloom-probe step (check_found_add's inner loop, main.c:291-344).
#pragma once
#include "../../include/ecloop_b200.h"
#include "bloom.cuh"
#include "ec.cuh"
#include "fp.cuh"
#include "hash160.cuh"

// One CTA of 512 threads per SM: all 16 warps of the SM sit behind the same barriers (ECL_HASH_SYNC), which is
// what lets them share instruction fetches; 128 registers per thread either way.
#ifndef ADD_THREADS
#define ADD_THREADS 512
#endif
#ifndef ADD_MIN_BLOCKS
#define ADD_MIN_BLOCKS 1
#endif
#ifndef ECL_HASH_SYNC
#define ECL_HASH_SYNC 1  // 0: no lockstep; 1: barrier per step and between SHA-256 and RIPEMD-160; 4: + every 16 rounds
#endif
#ifndef ADD_H
#define ADD_H 1024  // half group: 2*ADD_H keys share one inversion; 2048 = the reference's GROUP_INV_SIZE
#endif

// ---------------------------------------------------------------- small helpers

struct HitSink {
  ecl_hit *hits;
  u32 *count;
  u32 cap;
};

__device__ __forceinline__ void emit_hit(const HitSink &s, u64 key_off, const u32 h[5], u32 endo, u32 kind) {
  const u32 idx = atomicAdd(s.count, 1u);
  if (idx < s.cap) {
    uint4 *o = reinterpret_cast<uint4 *>(s.hits + idx);
    o[0] = make_uint4((u32)key_off, (u32)(key_off >> 32), h[0], h[1]);
    o[1] = make_uint4(h[2], h[3], h[4], endo | (kind << 8));
  }
}

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(smem_u32(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 phase) {
  u32 done;
  do {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, u32 bytes, u64 *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ fe fe_from_u4(const uint4 a, const uint4 b) {
  fe r;
  r.v[0] = a.x, r.v[1] = a.y, r.v[2] = a.z, r.v[3] = a.w, r.v[4] = b.x, r.v[5] = b.y, r.v[6] = b.z, r.v[7] = b.w;
  return r;
}

// beta (lib/ecc.c:38, B1): x -> beta*x is the curve endomorphism; beta^2 = B2
__device__ __forceinline__ fe fe_beta() {
  fe b;
  b.v[0] = 0x719501eeu, b.v[1] = 0xc1396c28u, b.v[2] = 0x12f58995u, b.v[3] = 0x9cf04975u;
  b.v[4] = 0xac3434e9u, b.v[5] = 0x6e64479eu, b.v[6] = 0x657c0710u, b.v[7] = 0x7ae96a2bu;
  return b;
}

// ---------------------------------------------------------------- hash + probe of NW points
// check_found_add's inner loop (main.c:291-298) and its endomorphism block (main.c:300-344) for NW points at
// once. Emission order is restored on the host (ecl_collect sorts), so lanes may report in any order.
// SYNC: CTA-barrier density inside the hashes (hash160.cuh); only legal when the whole CTA calls this in lockstep.
// `active[n]` masks the reporting of lanes that only run along to keep the CTA in lockstep (threads without work,
// keys past the end of the launch).
// PIPE: nullptr_t-like tag type NoPipe (probe inline: filter in shared memory, or a parity kernel) or a ProbePipe
// (probe_pipe.cuh: filter in HBM, probes in flight while the next hash is computed).
struct NoPipe {};
// SLOT: which of the pipe's two in-flight slots this call site uses (call sites must alternate 0, 1, 0, 1, ...)
template <u32 SLOT>
__device__ __forceinline__ void probe_hash(NoPipe &, const BloomView &bv, const HitSink &sink, const u32 (&hh)[5], u64 off,
                                           u32 endo, u32 kind, bool active) {
  if (bloom_has(bv, hh) && active) emit_hit(sink, off, hh, endo, kind);
}

// the single call site of an instance that hashes one point at a time (NW = 1)
__device__ __forceinline__ void probe_hash_one(NoPipe &np, const BloomView &bv, const HitSink &sink, const u32 (&hh)[5], u64 off,
                                              u32 endo, u32 kind, bool active) {
  probe_hash<0>(np, bv, sink, hh, off, endo, kind, active);
}

template <int NW, bool A33, bool A65, bool ENDO, int SYNC = 0, class PIPE = NoPipe>
__device__ __forceinline__ void check_points(  // with a ProbePipe: NW == 2 -> lane n uses slot n; NW == 1 -> one slot, judged a hash later
    const BloomView &bv, const HitSink &sink, u32 (&x)[NW][8], u32 (&y)[NW][8],
                                             const u64 (&off)[NW], const bool (&active)[NW], PIPE &pipe) {
  constexpr int NE = ENDO ? 6 : 1;
#pragma unroll 1
  for (int e = 0; e < NE; ++e) {
    if (ENDO && e > 0) {  // images 1..5: (x,-y) (bx,y) (bx,-y) (b^2x,y) (b^2x,-y)  
#pragma unroll
      for (int n = 0; n < NW; ++n) {
        fe t;
#pragma unroll
        for (int i = 0; i < 8; ++i) t.v[i] = y[n][i];
        t = fe_neg(t);
#pragma unroll
        for (int i = 0; i < 8; ++i) y[n][i] = t.v[i];
        if ((e & 1) == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) t.v[i] = x[n][i];
          t = fe_mul(t, fe_beta());
#pragma unroll
          for (int i = 0; i < 8; ++i) x[n][i] = t.v[i];
        }
      }
    }
    if (A33) {
      vw<NW> h[5];
      u32 odd[NW];
#pragma unroll
      for (int n = 0; n < NW; ++n) odd[n] = y[n][0];
      hash160_33<NW, SYNC>(h, x, odd);
#pragma unroll
      for (int n = 0; n < NW; ++n) {
        const u32 hh[5] = {h[0].l[n], h[1].l[n], h[2].l[n], h[3].l[n], h[4].l[n]};
        if (NW == 1) probe_hash_one(pipe, bv, sink, hh, off[n], (u32)e, 0u, active[n]);
        else if (n & 1) probe_hash<1>(pipe, bv, sink, hh, off[n], (u32)e, 0u, active[n]);
        else probe_hash<0>(pipe, bv, sink, hh, off[n], (u32)e, 0u, active[n]);
      }
    }
    if (A65) {
      vw<NW> h[5];
      hash160_65<NW, SYNC>(h, x, y);
#pragma unroll
      for (int n = 0; n < NW; ++n) {
        const u32 hh[5] = {h[0].l[n], h[1].l[n], h[2].l[n], h[3].l[n], h[4].l[n]};
        if (NW == 1) probe_hash_one(pipe, bv, sink, hh, off[n], (u32)e, 1u, active[n]);
        else if (n & 1) probe_hash<1>(pipe, bv, sink, hh, off[n], (u32)e, 1u, active[n]);
        else probe_hash<0>(pipe, bv, sink, hh, off[n], (u32)e, 1u, active[n]);
      }
    }
  }
}

template <int NW, bool A33, bool A65, bool ENDO, int SYNC = 0>
__device__ __forceinline__ void check_points(const BloomView &bv, const HitSink &sink, u32 (&x)[NW][8], u32 (&y)[NW][8],
                                             const u64 (&off)[NW]) {
  NoPipe none;
  bool active[NW];
#pragma unroll
  for (int n = 0; n < NW; ++n) active[n] = true;
  check_points<NW, A33, A65, ENDO, SYNC, NoPipe>(bv, sink, x, y, off, active, none);
}
