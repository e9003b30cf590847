// add_kernel.cuh — K1, the fused add kernel: batch_add + check_found_add + addr33/65_batch + blf_has
// (main.c:287-403, lib/addr.c:99-131, lib/utils.c:308-326) in one launch.
#pragma once
#include "common.cuh"
#include "probe_pipe.cuh"

struct AddParams {
  const u32 *cx, *cy;  // thread centres, SoA: limb l of thread t at [l*T + t]
  const uint4 *table;  // H affine points of 64 B: entry i = (i+1)*s*G
  const uint4 *step_pt;  // the group step 2*Hr*s*G (64 B): an entry of `table` when 2*Hr <= H, else computed per launch
  uint4 *scratch;      // prefix products: element i, half h of thread t at [(2i+h)*T + t]
  BloomView bloom;     // device-global filter
  u32 bloom_smem_words;  // != 0: the filter is staged into shared memory (then == bloom.size)
  HitSink sink;
  u32 T;                  // threads that own work (also the SoA stride)
  u32 groups_per_thread;  // consecutive groups of 2*Hr keys owned by one thread
  u32 Hr;                 // half group of THIS launch, 2 <= Hr <= H: chosen by the launch planner (ecl_api.cu plan_launch)
                          // so that T * groups_per_thread * 2*Hr tiles the launch's keys with (almost) no idle lanes
  u64 n_keys;             // keys of this launch; a key index (relative to the launch) >= n_keys is not reported
  u64 key_off0;           // index (in keys) of the first key of this launch inside the submitted span
  u32 zero;               // 0, opaque to the compiler: the pinned form XORs (field result & zero) into the hash state
  u32 *err;               // set to 1 when a group's batch product is zero (a centre equal to +-m*s*G: the keys of the
                          // span reach 0 or n; the reference asserts there, lib/ecc.c:666)
  CandQueue cand;         // HBM kernels only: where stage 1 of the asynchronous probe queues its candidates
};

// keys of group `g` (0-based inside the launch) of thread t that lie inside the launch: 0 .. 2*Hr
__device__ __forceinline__ u32 group_keys_inside(const AddParams &p, u64 first_key) {
  if (first_key >= p.n_keys) return 0u;
  const u64 left = p.n_keys - first_key;
  return left < (u64)(2u * p.Hr) ? (u32)left : 2u * p.Hr;
}

// table (H entries) + the step point behind it, by TMA bulk copies counted on one mbarrier
#define ADD_STAGE_TABLES(H_)                                                       \
  extern __shared__ __align__(128) unsigned char smem_raw[];                       \
  __shared__ __align__(8) u64 mbar;                                                \
  uint4 *tab = reinterpret_cast<uint4 *>(smem_raw);                                \
  u64 *sbloom = reinterpret_cast<u64 *>(smem_raw + ((H_) + 1) * 64);               \
  const u32 tab_bytes = (H_) * 64;                                                 \
  const u32 bloom_bytes = ((p.bloom_smem_words * 8u + 15u) / 16u) * 16u;           \
  if (threadIdx.x == 0) mbar_init(&mbar, 1);                                       \
  __syncthreads();                                                                 \
  if (threadIdx.x == 0) {                                                          \
    mbar_expect_tx(&mbar, tab_bytes + 64u + bloom_bytes);                          \
    bulk_g2s(tab, p.table, tab_bytes, &mbar);                                      \
    bulk_g2s(tab + (H_) * 4, p.step_pt, 64u, &mbar);                               \
    if (bloom_bytes) bulk_g2s(sbloom, p.bloom.bits, bloom_bytes, &mbar);           \
  }                                                                                \
  mbar_wait(&mbar, 0);

// The probe pipe of a kernel instance: a ProbePipe in the dynamic shared memory behind the table for HBM filters,
// an empty tag otherwise. `pipe` is what check_points / probe_hash take.
#define PROBE_PIPE_SETUP(HBM_, SMEM_OFFSET_)                                                  \
  __shared__ u32 cand_count;                                                                  \
  typename PipeOf<HBM_>::type pipe;                                                           \
  if (HBM_) {                                                                                 \
    if (threadIdx.x == 0) cand_count = 0;                                                     \
    __syncthreads();                                                                          \
  }                                                                                           \
  pipe_init(pipe, smem_raw + (SMEM_OFFSET_), &cand_count, p.bloom, p.cand);
#define PROBE_PIPE_FINISH(HBM_, SINGLE_) pipe_finish<SINGLE_>(pipe);

template <bool HBM>
struct PipeOf {
  typedef NoPipe type;
};
template <>
struct PipeOf<true> {
  typedef ProbePipe<ADD_THREADS> type;
};
__device__ __forceinline__ void pipe_init(NoPipe &, unsigned char *, u32 *, const BloomView &, const CandQueue &) {}
__device__ __forceinline__ void pipe_init(ProbePipe<ADD_THREADS> &pp, unsigned char *smem, u32 *cnt, const BloomView &bv,
                                          const CandQueue &q) {
  pp.init(smem, cnt, bv, q);
}
template <bool SINGLE>
__device__ __forceinline__ void pipe_finish(NoPipe &) {}
template <bool SINGLE>
__device__ __forceinline__ void pipe_finish(ProbePipe<ADD_THREADS> &pp) {
  pp.template finish<SINGLE>();
}

// Thread t owns the consecutive groups [t*c, (t+1)*c) of 2*Hr keys. For one group with centre point
// P = (start + (g*2Hr + Hr)*s)*G it forms every P +- (i+1)*s*G, i < Hr, sharing ONE field inversion through
// Montgomery's trick (fe_modp_grpinv, lib/ecc.c:522-540): prefix products go to a coalesced global scratch
// (32 B per element, written once, read once), the running inverse stays in registers. The group step 2Hr*s*G
// rides in the same batch as element 0, so moving to the next group costs one affine addition and no
// extra inversion (the reference pays a second inversion per group for that, main.c:400).
// Key order inside a group matches the reference: K-Hr .. K-1, K, K+1 .. K+Hr-1 (main.c:363,391).
// NW = points hashed side by side at source level (2: both points of a step; 1: one after the other, half the code).
template <int H, bool A33, bool A65, bool ENDO, bool HBM, int NW = 2>
__global__ void __launch_bounds__(ADD_THREADS, ADD_MIN_BLOCKS) add_kernel(const AddParams p) {
  static_assert(NW == 1 || NW == 2, "NW is 1 or 2");
  ADD_STAGE_TABLES(H)

  BloomView bv = p.bloom;
  if (p.bloom_smem_words) bv.bits = sbloom;
  PROBE_PIPE_SETUP(HBM, (H + 1) * 64)

  // Every thread of the CTA runs the same number of steps so that the CTA can sit behind barriers (lockstep,
  // see common.cuh): threads without work of their own (past T, or past the last key) run along on the last
  // owner's centre and only their reporting is masked.
  const u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 T = p.T;
  const bool owner = tid < T;
  const u32 t = owner ? tid : T - 1;
  const int Hr = (int)p.Hr;

  fe px, py;
#pragma unroll
  for (int l = 0; l < 8; ++l) px.v[l] = p.cx[(size_t)l * T + t], py.v[l] = p.cy[(size_t)l * T + t];
  if (owner && fe_is_zero(px) && fe_is_zero(py)) *p.err = 1u;  // the centre key is 0 (mod n): no point to start from

  uint4 *scr = p.scratch + tid;  // scratch is sized for whole CTAs
  const size_t TS = (size_t)gridDim.x * blockDim.x;  // scratch stride

#pragma unroll 1
  for (u32 g = 0; g < p.groups_per_thread; ++g) {
    const u64 first = ((u64)t * p.groups_per_thread + g) * (u64)(2 * Hr);  // first key of the group, launch-relative
    const u32 inside = owner ? group_keys_inside(p, first) : 0u;           // keys of this group to report
    const u64 kc = p.key_off0 + first + (u64)Hr;                           // span-relative index of the centre key

    // ---- pass 1: prefix products e_0, e_0 e_1, ...   e_0 = step.x - px, e_{i+1} = table[i].x - px
    fe acc = fe_sub(fe_from_u4(tab[H * 4 + 0], tab[H * 4 + 1]), px);
#pragma unroll 1
    for (int i = 0; i < Hr; ++i) {
      scr[(size_t)(2 * i) * TS] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
      scr[(size_t)(2 * i + 1) * TS] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
      const fe d = fe_sub(fe_from_u4(tab[i * 4 + 0], tab[i * 4 + 1]), px);
      acc = fe_mul(acc, d);
    }
    if (inside && fe_is_zero(acc)) *p.err = 1u;
    fe inv = fe_inv(acc);  // 1 / (e_0 ... e_Hr)

    // ---- pass 2: peel the inverses off from the far end; two points per step
    // With the filter in HBM thousands of TLB-missing probe fetches are in flight per SM and a scratch load issued
    // behind them takes ~10 us: there the prefix of the NEXT step is fetched before this step's hashes.
    fe pre_next = fe_from_u4(scr[(size_t)(2 * (Hr - 1)) * TS], scr[(size_t)(2 * (Hr - 1) + 1) * TS]);
#pragma unroll 1
    for (int i = Hr - 1; i >= 0; --i) {
      if (ECL_HASH_SYNC) __syncthreads();
      fe pre;  // e_0 ... e_i
      if (HBM) {
        pre = pre_next;
        const int j = i > 0 ? i - 1 : 0;
        pre_next = fe_from_u4(scr[(size_t)(2 * j) * TS], scr[(size_t)(2 * j + 1) * TS]);
      } else {
        pre = fe_from_u4(scr[(size_t)(2 * i) * TS], scr[(size_t)(2 * i + 1) * TS]);
      }
      const fe gx = fe_from_u4(tab[i * 4 + 0], tab[i * 4 + 1]);
      const fe gy = fe_from_u4(tab[i * 4 + 2], tab[i * 4 + 3]);
      const fe inv_i = fe_mul(inv, pre);  // 1 / (gx - px)
      inv = fe_mul(inv, fe_sub(gx, px));  // 1 / (e_0 ... e_i)

      u32 x[2][8], y[2][8];
      u64 off[2];
      bool act[2];
      fe rx, ry;
      // lane 0: P - (i+1)sG  -> key K - (i+1), index Hr - (i+1) inside the group
      affine_add_inv(rx, ry, px, py, gx, fe_neg(gy), inv_i);
#pragma unroll
      for (int l = 0; l < 8; ++l) x[0][l] = rx.v[l], y[0][l] = ry.v[l];
      off[0] = kc - (u64)(i + 1);
      act[0] = (u32)(Hr - (i + 1)) < inside;
      // lane 1: P + (i+1)sG -> key K + (i+1); the far end K+Hr is outside the group, its slot takes K itself
      if (i == Hr - 1) {
        rx = px, ry = py;
        off[1] = kc;
        act[1] = (u32)Hr < inside;
      } else {
        affine_add_inv(rx, ry, px, py, gx, gy, inv_i);
        off[1] = kc + (u64)(i + 1);
        act[1] = (u32)(Hr + (i + 1)) < inside;
      }
#pragma unroll
      for (int l = 0; l < 8; ++l) x[1][l] = rx.v[l], y[1][l] = ry.v[l];

      if (NW == 2) {
        check_points<2, A33, A65, ENDO, ECL_HASH_SYNC>(bv, p.sink, x, y, off, act, pipe);
      } else {
#pragma unroll 1
        for (int n = 0; n < 2; ++n) {
          u32 x1[1][8], y1[1][8];
#pragma unroll
          for (int l = 0; l < 8; ++l) x1[0][l] = n ? x[1][l] : x[0][l], y1[0][l] = n ? y[1][l] : y[0][l];
          const u64 off1[1] = {n ? off[1] : off[0]};
          const bool act1[1] = {n ? act[1] : act[0]};
          check_points<1, A33, A65, ENDO, ECL_HASH_SYNC>(bv, p.sink, x1, y1, off1, act1, pipe);
        }
      }
    }

    // ---- next group's centre: P + 2Hr*s*G with inv = 1/(step.x - px)
    {
      const fe sx = fe_from_u4(tab[H * 4 + 0], tab[H * 4 + 1]);
      const fe sy = fe_from_u4(tab[H * 4 + 2], tab[H * 4 + 3]);
      fe nx, ny;
      affine_add_inv(nx, ny, px, py, sx, sy, inv);
      px = nx, py = ny;
    }
  }
  PROBE_PIPE_FINISH(HBM, NW == 1)
}

// ---------------------------------------------------------------- K1-sp: the software-pipelined variant (no endomorphism)
// Same work as add_kernel<H, A33, A65, false>, restructured so that every basic block of the hot loop holds
// the hash160(s) of one point (ALU-pipe: LOP3/SHF/IADD3) AND the field arithmetic that produces the next point
// (FMA-pipe: IMAD.WIDE): ptxas interleaves the two independent streams, so the two integer pipes work at the same
// time instead of taking turns (in add_kernel the whole CTA alternates between an FMA-bound field phase and an
// ALU-bound hash phase). One pass-2 step = block X: hash(P - (i+1)G) || form P + (i+1)G, then block Y:
// hash(P + (i+1)G) || peel the inverse of step i-1 and form P - iG.

// hash160(s) of one point, probe, report; used where nothing is pipelined (three times per group)
template <bool A33, bool A65>
static __device__ __noinline__ void check_one_slow(const BloomView &bv, const HitSink &sink, const fe &x, const fe &y, u64 off,
                                                   bool active) {
  u32 xx[1][8], yy[1][8];
#pragma unroll
  for (int l = 0; l < 8; ++l) xx[0][l] = x.v[l], yy[0][l] = y.v[l];
  vw<1> h[5];
  if (A33) {
    const u32 odd[1] = {y.v[0]};
    hash160_33<1, 0>(h, xx, odd);
    const u32 hh[5] = {h[0].l[0], h[1].l[0], h[2].l[0], h[3].l[0], h[4].l[0]};
    if (bloom_has(bv, hh) && active) emit_hit(sink, off, hh, 0, 0);
  }
  if (A65) {
    hash160_65<1, 0>(h, xx, yy);
    const u32 hh[5] = {h[0].l[0], h[1].l[0], h[2].l[0], h[3].l[0], h[4].l[0]};
    if (bloom_has(bv, hh) && active) emit_hit(sink, off, hh, 0, 1);
  }
}

// the hash(es) of the point (ax, ay) inside a pipelined block: digest(s) out, probing is done by the caller after the
// field work of the block so that the probe's branches do not cut the block in two
// the hook (hash160.cuh) rides in the first hash of the point
template <bool A33, bool A65, class HOOK>
__device__ __forceinline__ void hash_point(vw<1> (&h33)[5], vw<1> (&h65)[5], const fe &ax, const fe &ay, HOOK &hook) {
  u32 xx[1][8], yy[1][8];
#pragma unroll
  for (int l = 0; l < 8; ++l) xx[0][l] = ax.v[l], yy[0][l] = ay.v[l];
  if (A33) {
    const u32 odd[1] = {ay.v[0]};
    hash160_33<1, 0, HOOK>(h33, xx, odd, hook);
  }
  if (A65) {
    if (A33) hash160_65<1, 0>(h65, xx, yy);
    else hash160_65<1, 0, HOOK>(h65, xx, yy, hook);
  }
}
template <bool A33, bool A65>
__device__ __forceinline__ void hash_point(vw<1> (&h33)[5], vw<1> (&h65)[5], const fe &ax, const fe &ay) {
  NoHook none;
  hash_point<A33, A65, NoHook>(h33, h65, ax, ay, none);
}

#ifndef ECL_SP_SYNC_EVERY
#define ECL_SP_SYNC_EVERY 1  // lockstep barrier every n-th pass-2 step of the pipelined kernel
#endif
#ifndef ECL_SP_PINS
#define ECL_SP_PINS 0  // 1: every field multiplication of a pass-2 step is pinned into the hash of its block (see HookX / HookY)
#endif

// Block X's field work as a hook: form P + (i+1)G from lam = (gy - py) * inv_i, and one prefix product of the next group.
// Each piece runs where the hash calls the hook and XORs (result & zero) into a word of the hash state (zero is a kernel
// parameter that is 0: ptxas cannot know), so the hash cannot pass that point before the piece is done and the scheduler
// places the multiplication INSIDE the hash instead of behind it.
struct HookX {
  const fe &px, &py, &gx, &gy, &inv_i, &nx;
  fe &bx, &by, &accn;
  const uint4 *tab;
  uint4 *scr_nxt;
  size_t TS;
  int k;
  u32 zero;
  fe lam;
  __device__ __forceinline__ HookX(const fe &px_, const fe &py_, const fe &gx_, const fe &gy_, const fe &inv_i_, const fe &nx_, fe &bx_, fe &by_,
                                   fe &accn_, const uint4 *tab_, uint4 *scr_, size_t TS_, int k_, u32 zero_)
      : px(px_), py(py_), gx(gx_), gy(gy_), inv_i(inv_i_), nx(nx_), bx(bx_), by(by_), accn(accn_), tab(tab_), scr_nxt(scr_), TS(TS_), k(k_), zero(zero_) {}
  __device__ __forceinline__ void sha(int i, u32 &w) {
    if (i == 4) {
      lam = fe_mul_nc(fe_sub(gy, py), inv_i);
      w ^= lam.v[0] & zero;
    }
    if (i == 26) {
      bx = fe_sub(fe_sub(fe_sqr(lam), px), gx);
      w ^= bx.v[0] & zero;
    }
    if (i == 48) {
      by = fe_sub(fe_mul(lam, fe_sub(px, bx)), py);
      w ^= by.v[0] & zero;
    }
  }
  __device__ __forceinline__ void rmd(int r, u32 &a) {
    if (r == 1) {
      scr_nxt[(size_t)(2 * k) * TS] = make_uint4(accn.v[0], accn.v[1], accn.v[2], accn.v[3]);
      scr_nxt[(size_t)(2 * k + 1) * TS] = make_uint4(accn.v[4], accn.v[5], accn.v[6], accn.v[7]);
      accn = fe_mul_nc(accn, fe_sub(fe_from_u4(tab[k * 4 + 0], tab[k * 4 + 1]), nx));
      a ^= accn.v[0] & zero;
    }
  }
};

// Block Y's field work: peel the inverse of step i-1 (two independent products) and form P - iG.
struct HookY {
  const fe &px, &py, &gx, &gy, &q;
  fe &inv, &inv_i, &ax, &ay;
  u32 zero;
  fe lam;
  __device__ __forceinline__ HookY(const fe &px_, const fe &py_, const fe &gx_, const fe &gy_, const fe &q_, fe &inv_, fe &inv_i_, fe &ax_, fe &ay_, u32 zero_)
      : px(px_), py(py_), gx(gx_), gy(gy_), q(q_), inv(inv_), inv_i(inv_i_), ax(ax_), ay(ay_), zero(zero_) {}
  __device__ __forceinline__ void sha(int i, u32 &w) {
    if (i == 2) {
      inv_i = fe_mul_nc(inv, q);
      w ^= inv_i.v[0] & zero;
    }
    if (i == 18) {
      inv = fe_mul_nc(inv, fe_sub(gx, px));
      w ^= inv.v[0] & zero;
    }
    if (i == 34) {
      lam = fe_mul_nc(fe_sub(fe_neg_nz(gy), py), inv_i);
      w ^= lam.v[0] & zero;
    }
    if (i == 52) {
      ax = fe_sub(fe_sub(fe_sqr(lam), px), gx);
      w ^= ax.v[0] & zero;
    }
  }
  __device__ __forceinline__ void rmd(int r, u32 &a) {
    if (r == 1) {
      ay = fe_sub(fe_mul(lam, fe_sub(px, ax)), py);
      a ^= ay.v[0] & zero;
    }
  }
};
template <bool A33, bool A65, class PIPE>
__device__ __forceinline__ void probe_point(PIPE &pipe, const BloomView &bv, const HitSink &sink, const vw<1> (&h33)[5],
                                            const vw<1> (&h65)[5], u64 off, bool active) {
  if (A33) {
    const u32 hh[5] = {h33[0].l[0], h33[1].l[0], h33[2].l[0], h33[3].l[0], h33[4].l[0]};
    probe_hash_dyn(pipe, bv, sink, hh, off, 0u, 0u, active);
  }
  if (A65) {
    const u32 hh[5] = {h65[0].l[0], h65[1].l[0], h65[2].l[0], h65[3].l[0], h65[4].l[0]};
    probe_hash_dyn(pipe, bv, sink, hh, off, 0u, 1u, active);
  }
}

// Pass 1 of the NEXT group (its prefix products) rides in block X as well, so after the first group of a launch
// there is no field-only phase left: the group step is the element peeled FIRST (it is multiplied in last), which
// makes the next centre known at the start of pass 2, and the prefixes go to the other half of a ping-pong scratch.
// Elements of a group: f_i = table[i].x - px (i < Hr), f_Hr = step.x - px; scratch entry k holds q_k = f_0 ... f_{k-1}.
template <int H, bool A33, bool A65, bool HBM>
__global__ void __launch_bounds__(ADD_THREADS, ADD_MIN_BLOCKS) add_kernel_sp(const AddParams p) {
  static_assert(H >= 2, "the pipelined loop needs a prologue step and an epilogue step");
  ADD_STAGE_TABLES(H)

  BloomView bv = p.bloom;
  if (p.bloom_smem_words) bv.bits = sbloom;
  PROBE_PIPE_SETUP(HBM, (H + 1) * 64)

  const u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 T = p.T;
  const bool owner = tid < T;
  const u32 t = owner ? tid : T - 1;
  const int Hr = (int)p.Hr;  // >= 2 (plan_launch)

  fe px, py;
#pragma unroll
  for (int l = 0; l < 8; ++l) px.v[l] = p.cx[(size_t)l * T + t], py.v[l] = p.cy[(size_t)l * T + t];
  if (owner && fe_is_zero(px) && fe_is_zero(py)) *p.err = 1u;  // the centre key is 0 (mod n): no point to start from

  const size_t TS = (size_t)gridDim.x * blockDim.x;  // scratch stride (scratch is sized for whole CTAs)
  uint4 *scr_cur = p.scratch + tid;                  // entry k, half h at [(2k + h) * TS]
  uint4 *scr_nxt = scr_cur + (size_t)2 * (H + 1) * TS;
  const fe sx = fe_from_u4(tab[H * 4 + 0], tab[H * 4 + 1]);

  // ---- pass 1 of this thread's first group (the only one that is not hidden behind hashing)
  fe tot;
  {
    fe acc = fe_one();
#pragma unroll 1
    for (int k = 0; k < Hr; ++k) {
      scr_cur[(size_t)(2 * k) * TS] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
      scr_cur[(size_t)(2 * k + 1) * TS] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
      acc = fe_mul_nc(acc, fe_sub(fe_from_u4(tab[k * 4 + 0], tab[k * 4 + 1]), px));
    }
    scr_cur[(size_t)(2 * Hr) * TS] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
    scr_cur[(size_t)(2 * Hr + 1) * TS] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
    tot = fe_mul_nc(acc, fe_sub(sx, px));
  }

#pragma unroll 1
  for (u32 g = 0; g < p.groups_per_thread; ++g) {
    const u64 first = ((u64)t * p.groups_per_thread + g) * (u64)(2 * Hr);
    const u32 inside = owner ? group_keys_inside(p, first) : 0u;
    const u64 kc = p.key_off0 + first + (u64)Hr;

    if (inside && fe_is_zero_modp(tot)) *p.err = 1u;
    fe inv = fe_inv(tot);  // 1 / (f_0 ... f_Hr)

    // ---- the group step first: next centre N = P + 2Hr*s*G
    fe nx, ny;
    {
      const fe qH = fe_from_u4(scr_cur[(size_t)(2 * Hr) * TS], scr_cur[(size_t)(2 * Hr + 1) * TS]);
      const fe inv_s = fe_mul_nc(inv, qH);  // 1 / f_Hr
      inv = fe_mul_nc(inv, fe_sub(sx, px));  // 1 / q_Hr
      const fe sy = fe_from_u4(tab[H * 4 + 2], tab[H * 4 + 3]);
      affine_add_inv(nx, ny, px, py, sx, sy, inv_s);
    }

    // the centre itself (key K) takes the slot of the far end K+Hr, which lies outside the group
    check_one_slow<A33, A65>(bv, p.sink, px, py, kc, (u32)Hr < inside);

    // ---- prologue: operands and inverse of step Hr-1, and its first point P - Hr*G
    fe gx = fe_from_u4(tab[(Hr - 1) * 4 + 0], tab[(Hr - 1) * 4 + 1]);
    fe gy = fe_from_u4(tab[(Hr - 1) * 4 + 2], tab[(Hr - 1) * 4 + 3]);
    fe inv_i;
    {
      const fe q = fe_from_u4(scr_cur[(size_t)(2 * (Hr - 1)) * TS], scr_cur[(size_t)(2 * (Hr - 1) + 1) * TS]);
      inv_i = fe_mul_nc(inv, q);
      inv = fe_mul_nc(inv, fe_sub(gx, px));
    }
    fe ax, ay;  // the point waiting to be hashed
    affine_add_inv(ax, ay, px, py, gx, fe_neg_nz(gy), inv_i);
    fe accn = fe_one();  // q'_k of the next group

    // ---- pass 2, pipelined
#pragma unroll 1
    for (int i = Hr - 1; i >= 1; --i) {
      if (ECL_HASH_SYNC && (i % ECL_SP_SYNC_EVERY) == 0) __syncthreads();
#if ECL_SP_PINS
      // block X / block Y with every field multiplication pinned into the block's hash (HookX / HookY above)
      vw<1> h33[5], h65[5];
      fe bx, by;
      {
        HookX hx(px, py, gx, gy, inv_i, nx, bx, by, accn, tab, scr_nxt, TS, Hr - 1 - i, p.zero);
        hash_point<A33, A65, HookX>(h33, h65, ax, ay, hx);
      }
      probe_point<A33, A65>(pipe, bv, p.sink, h33, h65, kc - (u64)(i + 1), (u32)(Hr - (i + 1)) < inside);
      {
        const fe q = fe_from_u4(scr_cur[(size_t)(2 * (i - 1)) * TS], scr_cur[(size_t)(2 * (i - 1) + 1) * TS]);
        gx = fe_from_u4(tab[(i - 1) * 4 + 0], tab[(i - 1) * 4 + 1]);
        gy = fe_from_u4(tab[(i - 1) * 4 + 2], tab[(i - 1) * 4 + 3]);
        HookY hy(px, py, gx, gy, q, inv, inv_i, ax, ay, p.zero);
        hash_point<A33, A65, HookY>(h33, h65, bx, by, hy);
      }
      probe_point<A33, A65>(pipe, bv, p.sink, h33, h65, kc + (u64)(i + 1), i != Hr - 1 && (u32)(Hr + (i + 1)) < inside);
    }
#else
      // block X: hash P - (i+1)G  ||  form P + (i+1)G, and one pass-1 step of the next group
      vw<1> h33[5], h65[5];
      hash_point<A33, A65>(h33, h65, ax, ay);
      fe bx, by;
      affine_add_inv(bx, by, px, py, gx, gy, inv_i);
      {
        const int k = Hr - 1 - i;
        scr_nxt[(size_t)(2 * k) * TS] = make_uint4(accn.v[0], accn.v[1], accn.v[2], accn.v[3]);
        scr_nxt[(size_t)(2 * k + 1) * TS] = make_uint4(accn.v[4], accn.v[5], accn.v[6], accn.v[7]);
        accn = fe_mul_nc(accn, fe_sub(fe_from_u4(tab[k * 4 + 0], tab[k * 4 + 1]), nx));
      }
      probe_point<A33, A65>(pipe, bv, p.sink, h33, h65, kc - (u64)(i + 1), (u32)(Hr - (i + 1)) < inside);
      // block Y: hash P + (i+1)G  ||  peel step i-1 and form P - iG
      hash_point<A33, A65>(h33, h65, bx, by);
      {
        const fe q = fe_from_u4(scr_cur[(size_t)(2 * (i - 1)) * TS], scr_cur[(size_t)(2 * (i - 1) + 1) * TS]);
        gx = fe_from_u4(tab[(i - 1) * 4 + 0], tab[(i - 1) * 4 + 1]);
        gy = fe_from_u4(tab[(i - 1) * 4 + 2], tab[(i - 1) * 4 + 3]);
        inv_i = fe_mul_nc(inv, q);
        inv = fe_mul_nc(inv, fe_sub(gx, px));
        affine_add_inv(ax, ay, px, py, gx, fe_neg_nz(gy), inv_i);
      }
      // the far end K+Hr (i == Hr-1) lies outside the group: computed along, never reported
      probe_point<A33, A65>(pipe, bv, p.sink, h33, h65, kc + (u64)(i + 1), i != Hr - 1 && (u32)(Hr + (i + 1)) < inside);
    }
#endif
    // ---- epilogue: step 0 (keys K-1 and K+1), the last two prefixes of the next group
    {
      fe bx, by;
      affine_add_inv(bx, by, px, py, gx, gy, inv_i);
      check_one_slow<A33, A65>(bv, p.sink, ax, ay, kc - 1, (u32)(Hr - 1) < inside);
      check_one_slow<A33, A65>(bv, p.sink, bx, by, kc + 1, (u32)(Hr + 1) < inside);
      scr_nxt[(size_t)(2 * (Hr - 1)) * TS] = make_uint4(accn.v[0], accn.v[1], accn.v[2], accn.v[3]);
      scr_nxt[(size_t)(2 * (Hr - 1) + 1) * TS] = make_uint4(accn.v[4], accn.v[5], accn.v[6], accn.v[7]);
      accn = fe_mul_nc(accn, fe_sub(fe_from_u4(tab[(Hr - 1) * 4 + 0], tab[(Hr - 1) * 4 + 1]), nx));
      scr_nxt[(size_t)(2 * Hr) * TS] = make_uint4(accn.v[0], accn.v[1], accn.v[2], accn.v[3]);
      scr_nxt[(size_t)(2 * Hr + 1) * TS] = make_uint4(accn.v[4], accn.v[5], accn.v[6], accn.v[7]);
      tot = fe_mul_nc(accn, fe_sub(sx, nx));
    }
    px = nx, py = ny;
    uint4 *sw = scr_cur;
    scr_cur = scr_nxt, scr_nxt = sw;
  }
  PROBE_PIPE_FINISH(HBM, false)
}
