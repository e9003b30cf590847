// probe_pipe.cuh — asynchronous bloom probing for filters that live in HBM (`.blf` files of hundreds of MB to GBs).
//
// blf_has (lib/utils.c:308-326) is a chain of up to 20 dependent random 8-byte reads with early exit. With the
// filter in shared memory that is a few cycles; in HBM every read is ~1 us, and because the fused add kernel runs
// its 16 warps in lockstep nobody hides it: the inline probe costs 30 % of the kernel (profiles/r01_d_large_bloom.txt).
// Here the kernel never waits for the filter:
//   stage 1 (in the add kernel)  the first TWO probe words of a hash are fetched with cp.async straight into shared
//                                memory (no register, no stall) and looked at PP_DEPTH hashes later; a hash whose
//                                two bits are set (fill^2 of them) goes to a per-CTA candidate queue in HBM;
//   stage 2 (cand_verify_kernel) one thread per candidate runs the full 20-probe test at full occupancy and
//                                reports the hits.
// The decision is blf_has's, bit for bit: stage 1 only drops hashes that blf_has would drop at probe 1 or 2.
#pragma once
#include "common.cuh"

#define PP_DEPTH 2  // hashes in flight per thread

struct CandQueue {
  uint4 *entries;   // n_cta regions of cap_per_cta candidates, 32 B each (same layout as ecl_hit)
  u32 *counts;      // candidates pushed by each CTA (may exceed cap_per_cta: then *overflow is set)
  u32 *overflow;
  u32 cap_per_cta;
};

// 16 bytes global -> shared without touching a register; .cg keeps the line out of L1, so a probe costs one 32 B
// DRAM sector (with .ca every probe pulled a whole 128 B line: 267 B/key of DRAM reads in the first version)
// The probes stream through a filter far larger than L2: they are marked evict-first so that they do not push the
// kernel's own instructions out of L2 (an ncu capture of the `-a cu` instance showed 5.6 no_instruction stalls per
// issued instruction with plain probes: the 230 KB loop body was being refetched from DRAM).
__device__ __forceinline__ u64 l2_evict_first_policy() {
  u64 pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// dst is a 32-bit shared-memory address held in ONE opaque register (ProbePipe::init): when ptxas could see it as
// uniform base + thread offset it emitted LDGSTS [R+UR+imm], desc[UR] with undefined uniform registers for some
// instances (CUDA 12.9, "illegal instruction" at run time).
__device__ __forceinline__ void cp_async_16(u32 smem_dst, const void *gmem_src, u64 policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gmem_src), "l"(policy) : "memory");
}
// plain form (no cache hint), destination given as a pointer
__device__ __forceinline__ void cp_async_16_plain(void *smem_dst, const void *gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Per-thread view of the CTA's pending-probe storage in shared memory. Layout (128-bit accesses, thread-minor):
//   uint4 probe[PP_DEPTH][2][THREADS]   the aligned 16 bytes around each of the two probe words
//   uint4 meta [PP_DEPTH][2][THREADS]   {h0,h1,h2,h3}, {h4, key_off lo, key_off hi, flags}
// flags: endo | kind << 8 | active << 16 | (bit position, which half) of probe 1 << 17 | same of probe 2 << 24
template <int THREADS>
struct ProbePipe {
  static constexpr u32 BYTES = PP_DEPTH * 4 * 16 * THREADS;
  uint4 *probe;
  uint4 *meta;
  u32 *cta_count;  // shared
  BloomView bv;    // the filter in global memory
  CandQueue q;
  u32 n;           // hashes submitted so far by this thread
  u64 policy;      // L2 evict-first for the probe fetches
  u32 probe_s32;   // shared-memory address of probe[0] for this thread, opaque to the compiler

  __device__ __forceinline__ void init(unsigned char *smem, u32 *count, const BloomView &b, const CandQueue &cq) {
    probe = reinterpret_cast<uint4 *>(smem) + threadIdx.x;
    meta = probe + PP_DEPTH * 2 * THREADS;
    cta_count = count, bv = b, q = cq, n = 0;
    policy = l2_evict_first_policy();
    asm("{\n\t.reg .u64 t;\n\tcvta.to.shared.u64 t, %1;\n\tcvt.u32.u64 %0, t;\n\t}" : "=r"(probe_s32) : "l"(probe));
  }

  __device__ __forceinline__ void retire(u32 slot) {
    const uint4 m1 = meta[(slot * 2 + 1) * THREADS];
    const u32 flags = m1.w;
    // the probed 64-bit word is the low or high half of the fetched 16 bytes
    const u32 p0 = (flags >> 17) & 127u, p1 = (flags >> 24) & 127u;
    const u64 *w = reinterpret_cast<const u64 *>(probe + (slot * 2) * THREADS);
    const u64 w0 = w[p0 >> 6], w1 = reinterpret_cast<const u64 *>(probe + (slot * 2 + 1) * THREADS)[p1 >> 6];
    if ((((w0 >> (p0 & 63u)) & (w1 >> (p1 & 63u))) & 1ull) && ((flags >> 16) & 1u)) {
      const uint4 m0 = meta[(slot * 2) * THREADS];
      const u32 idx = atomicAdd(cta_count, 1u);
      if (idx < q.cap_per_cta) {
        uint4 *o = q.entries + ((size_t)blockIdx.x * q.cap_per_cta + idx) * 2;
        o[0] = make_uint4(m1.y, m1.z, m0.x, m0.y);
        o[1] = make_uint4(m0.z, m0.w, m1.x, flags & 0xffffu);
      } else {
        *q.overflow = 1u;
      }
    }
  }

  // hand one hash to the pipe; the verdict on the hash that used this slot before (PP_DEPTH submissions ago) is
  // taken first. Call sites alternate the slots 0, 1, 0, 1, ... with a compile-time SLOT, which keeps every
  // shared-memory address of the pipe in the form [register + immediate].
  // KEEP = younger groups that may still be in flight when this slot's previous occupant is judged: PP_DEPTH - 1 when
  // the call sites alternate slots 0, 1 (two points hashed side by side, NW = 2); 0 when ONE call site serves every
  // hash (NW = 1: the previous submission is a whole hash old, far longer than a DRAM access, so waiting for it is free
  // and the single static slot keeps the address in the safe [R + imm] form).
  template <u32 slot, int KEEP = PP_DEPTH - 1>
  __device__ __forceinline__ void submit(const u32 h[5], u64 off, u32 endo, u32 kind, bool active) {
    static_assert(slot < PP_DEPTH, "slot out of range");
    if (n >= (KEEP ? PP_DEPTH : 1u)) {
      cp_async_wait<KEEP>();
      retire(slot);
    }
    const u64 a0 = (u64)h[0] << 32 | h[1], a1 = (u64)h[2] << 32 | h[3], a2 = (u64)h[4] << 32 | h[0];
    const u64 v0 = (a0 << 24) | (a1 >> 24), v1 = (a1 << 24) | (a2 >> 24);
    const u64 i0 = bloom_word_index(v0 >> 6, bv.size, bv.magic), i1 = bloom_word_index(v1 >> 6, bv.size, bv.magic);
    cp_async_16(probe_s32 + (slot * 2 + 0) * THREADS * 16, bv.bits + (i0 & ~1ull), policy);
    cp_async_16(probe_s32 + (slot * 2 + 1) * THREADS * 16, bv.bits + (i1 & ~1ull), policy);
    cp_async_commit();
    const u32 p0 = ((u32)v0 & 63u) | (((u32)i0 & 1u) << 6), p1 = ((u32)v1 & 63u) | (((u32)i1 & 1u) << 6);
    meta[(slot * 2) * THREADS] = make_uint4(h[0], h[1], h[2], h[3]);
    meta[(slot * 2 + 1) * THREADS] =
        make_uint4(h[4], (u32)off, (u32)(off >> 32), endo | (kind << 8) | ((active ? 1u : 0u) << 16) | (p0 << 17) | (p1 << 24));
    ++n;
  }

  // Variant for the software-pipelined kernel: run-time slot, no cache hint. That kernel's 7 000-instruction basic
  // blocks are sensitive to what ptxas makes of any change here (5 840 vs 4 560 Mkeys/s for equivalent forms), its
  // loop is small enough to stay in L2 beside the probe traffic, and this is the form measured at 5 840.
  __device__ __forceinline__ void submit_dyn(const u32 h[5], u64 off, u32 endo, u32 kind, bool active) {
    const u32 slot = n % PP_DEPTH;
    if (n >= PP_DEPTH) {
      cp_async_wait<PP_DEPTH - 1>();
      retire(slot);
    }
    const u64 a0 = (u64)h[0] << 32 | h[1], a1 = (u64)h[2] << 32 | h[3], a2 = (u64)h[4] << 32 | h[0];
    const u64 v0 = (a0 << 24) | (a1 >> 24), v1 = (a1 << 24) | (a2 >> 24);
    const u64 i0 = bloom_word_index(v0 >> 6, bv.size, bv.magic), i1 = bloom_word_index(v1 >> 6, bv.size, bv.magic);
    cp_async_16_plain(&probe[(slot * 2 + 0) * THREADS], bv.bits + (i0 & ~1ull));
    cp_async_16_plain(&probe[(slot * 2 + 1) * THREADS], bv.bits + (i1 & ~1ull));
    cp_async_commit();
    const u32 p0 = ((u32)v0 & 63u) | (((u32)i0 & 1u) << 6), p1 = ((u32)v1 & 63u) | (((u32)i1 & 1u) << 6);
    meta[(slot * 2) * THREADS] = make_uint4(h[0], h[1], h[2], h[3]);
    meta[(slot * 2 + 1) * THREADS] =
        make_uint4(h[4], (u32)off, (u32)(off >> 32), endo | (kind << 8) | ((active ? 1u : 0u) << 16) | (p0 << 17) | (p1 << 24));
    ++n;
  }

  // end of the kernel: take the verdict on everything still in flight and publish the CTA's candidate count
  template <bool SINGLE_SLOT = false>
  __device__ __forceinline__ void finish() {
    cp_async_wait<0>();
    if (SINGLE_SLOT) {
      if (n >= 1) retire(0);
    } else {
      if (n >= 1) retire((n - 1) % PP_DEPTH);
      if (n >= 2) retire((n - 2) % PP_DEPTH);
    }
    static_assert(PP_DEPTH == 2, "finish() retires exactly two slots");
    __syncthreads();
    if (threadIdx.x == 0) q.counts[blockIdx.x] = *cta_count;
  }
};

template <u32 SLOT, int THREADS>
__device__ __forceinline__ void probe_hash(ProbePipe<THREADS> &pipe, const BloomView &, const HitSink &, const u32 (&hh)[5],
                                           u64 off, u32 endo, u32 kind, bool active) {
  pipe.template submit<SLOT>(hh, off, endo, kind, active);
}

template <int THREADS>
__device__ __forceinline__ void probe_hash_one(ProbePipe<THREADS> &pipe, const BloomView &, const HitSink &, const u32 (&hh)[5],
                                               u64 off, u32 endo, u32 kind, bool active) {
  pipe.template submit<0, 0>(hh, off, endo, kind, active);
}

template <int THREADS>
__device__ __forceinline__ void probe_hash_dyn(ProbePipe<THREADS> &pipe, const BloomView &, const HitSink &, const u32 (&hh)[5],
                                               u64 off, u32 endo, u32 kind, bool active) {
  pipe.submit_dyn(hh, off, endo, kind, active);
}
__device__ __forceinline__ void probe_hash_dyn(NoPipe &np, const BloomView &bv, const HitSink &sink, const u32 (&hh)[5], u64 off,
                                               u32 endo, u32 kind, bool active) {
  probe_hash<0>(np, bv, sink, hh, off, endo, kind, active);
}

// stage 2: the rest of blf_has (probes 2..19, early exit) on every queued candidate: stage 1 saw probes 0 and 1 set.
// grid = (x, number of source CTAs)
static __global__ void __launch_bounds__(256) cand_verify_kernel(const CandQueue q, const BloomView bv, const HitSink sink) {
  const u32 src = blockIdx.y;
  u32 cnt = q.counts[src];
  if (cnt > q.cap_per_cta) cnt = q.cap_per_cta;
  for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) {
    const uint4 *e = q.entries + ((size_t)src * q.cap_per_cta + j) * 2;
    const uint4 a = e[0], b = e[1];
    const u32 hh[5] = {a.z, a.w, b.x, b.y, b.z};
    if (bloom_has_after_two(bv, hh)) emit_hit(sink, (u64)a.x | (u64)a.y << 32, hh, b.w & 0xffu, (b.w >> 8) & 0xffu);
  }
}

// set bits of the filter (grid-stride popcount): the launch planner sizes the candidate queue by fill^2
static __global__ void __launch_bounds__(256) bloom_popcount_kernel(const u64 *bits, u64 n, unsigned long long *total) {
  unsigned long long acc = 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) acc += __popcll(bits[i]);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(total, acc);
}
