// add_kernel.cuh — K1, the fused add kernel: batch_add + check_found_add + addr33/65_batch + blf_has
// (main.c:287-403, lib/addr.c:99-131, lib/utils.c:308-326) in one launch.
#pragma once
#include "common.cuh"
#include "probe_pipe.cuh"

struct AddParams {
  const u32 *cx, *cy;  // thread centres, SoA: limb l of thread t at [l*T + t]
  const uint4 *table;  // H+1 affine points of 64 B: entry i<H = (i+1)*s*G, entry H = 2H*s*G (group step)
  uint4 *scratch;      // prefix products: element i, half h of thread t at [(2i+h)*T + t]
  BloomView bloom;     // device-global filter
  u32 bloom_smem_words;  // != 0: the filter is staged into shared memory (then == bloom.size)
  HitSink sink;
  u32 T;                  // threads that own work (also the SoA stride)
  u32 groups_per_thread;  // consecutive groups of 2H keys owned by one thread
  u64 n_groups;           // groups in this launch
  u64 key_off0;           // index (in keys) of the first key of this launch inside the submitted span
  CandQueue cand;         // HBM kernels only: where stage 1 of the asynchronous probe queues its candidates
};

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
#define PROBE_PIPE_FINISH(HBM_) pipe_finish(pipe);

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
__device__ __forceinline__ void pipe_finish(NoPipe &) {}
__device__ __forceinline__ void pipe_finish(ProbePipe<ADD_THREADS> &pp) { pp.finish(); }

// Thread t owns the consecutive groups [t*c, (t+1)*c) of 2H keys. For one group with centre point
// P = (start + (g*2H + H)*s)*G it forms every P +- (i+1)*s*G, i < H, sharing ONE field inversion through
// Montgomery's trick (fe_modp_grpinv, lib/ecc.c:522-540): prefix products go to a coalesced global scratch
// (32 B per element, written once, read once), the running inverse stays in registers. The group step 2H*s*G
// rides in the same batch as element 0, so moving to the next group costs one affine addition and no
// extra inversion (the reference pays a second inversion per group for that, main.c:400).
// Key order inside a group matches the reference: K-H .. K-1, K, K+1 .. K+H-1 (main.c:363,391).
template <int H, bool A33, bool A65, bool ENDO, bool HBM>
__global__ void __launch_bounds__(ADD_THREADS, ADD_MIN_BLOCKS) add_kernel(const AddParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) u64 mbar;
  uint4 *tab = reinterpret_cast<uint4 *>(smem_raw);
  u64 *sbloom = reinterpret_cast<u64 *>(smem_raw + (H + 1) * 64);
  const u32 tab_bytes = (H + 1) * 64;
  const u32 bloom_bytes = ((p.bloom_smem_words * 8u + 15u) / 16u) * 16u;

  if (threadIdx.x == 0) mbar_init(&mbar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&mbar, tab_bytes + bloom_bytes);
    bulk_g2s(tab, p.table, tab_bytes, &mbar);
    if (bloom_bytes) bulk_g2s(sbloom, p.bloom.bits, bloom_bytes, &mbar);
  }
  mbar_wait(&mbar, 0);

  BloomView bv = p.bloom;
  if (p.bloom_smem_words) bv.bits = sbloom;
  PROBE_PIPE_SETUP(HBM, (H + 1) * 64)

  // Every thread of the CTA runs the same number of steps so that the CTA can sit behind barriers (lockstep,
  // see common.cuh): threads without work of their own (past T, or past the last group) run along on the last
  // owner's centre and only their reporting is masked.
  const u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 T = p.T;
  const bool owner = tid < T;
  const u32 t = owner ? tid : T - 1;

  fe px, py;
#pragma unroll
  for (int l = 0; l < 8; ++l) px.v[l] = p.cx[(size_t)l * T + t], py.v[l] = p.cy[(size_t)l * T + t];

  const u64 g0 = (u64)t * p.groups_per_thread;
  const u64 g1 = g0 + p.groups_per_thread;
  uint4 *scr = p.scratch + tid;  // scratch is sized for whole CTAs
  const size_t TS = (size_t)gridDim.x * blockDim.x;  // scratch stride

#pragma unroll 1
  for (u64 g = g0; g < g1; ++g) {
    const bool active = owner && g < p.n_groups;
    const u64 kc = p.key_off0 + g * (2 * H) + H;  // key index of the centre

    // ---- pass 1: prefix products e_0, e_0 e_1, ...   e_0 = step.x - px, e_{i+1} = table[i].x - px
    fe acc = fe_sub(fe_from_u4(tab[H * 4 + 0], tab[H * 4 + 1]), px);
#pragma unroll 1
    for (int i = 0; i < H; ++i) {
      scr[(size_t)(2 * i) * TS] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
      scr[(size_t)(2 * i + 1) * TS] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
      const fe d = fe_sub(fe_from_u4(tab[i * 4 + 0], tab[i * 4 + 1]), px);
      acc = fe_mul(acc, d);
    }
    fe inv = fe_inv(acc);  // 1 / (e_0 ... e_H)

    // ---- pass 2: peel the inverses off from the far end; two points per step
    // With the filter in HBM thousands of TLB-missing probe fetches are in flight per SM and a scratch load issued
    // behind them takes ~10 us: there the prefix of the NEXT step is fetched before this step's hashes.
    fe pre_next = fe_from_u4(scr[(size_t)(2 * (H - 1)) * TS], scr[(size_t)(2 * (H - 1) + 1) * TS]);
#pragma unroll 1
    for (int i = H - 1; i >= 0; --i) {
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
      fe rx, ry;
      // lane 0: P - (i+1)sG  -> key K - (i+1)
      affine_add_inv(rx, ry, px, py, gx, fe_neg(gy), inv_i);
#pragma unroll
      for (int l = 0; l < 8; ++l) x[0][l] = rx.v[l], y[0][l] = ry.v[l];
      off[0] = kc - (u64)(i + 1);
      // lane 1: P + (i+1)sG -> key K + (i+1); the far end K+H is outside the group, its slot takes K itself
      if (i == H - 1) {
        rx = px, ry = py;
        off[1] = kc;
      } else {
        affine_add_inv(rx, ry, px, py, gx, gy, inv_i);
        off[1] = kc + (u64)(i + 1);
      }
#pragma unroll
      for (int l = 0; l < 8; ++l) x[1][l] = rx.v[l], y[1][l] = ry.v[l];

      check_points<2, A33, A65, ENDO, ECL_HASH_SYNC>(bv, p.sink, x, y, off, active, pipe);
    }

    // ---- next group's centre: P + 2H*s*G with inv = 1/(step.x - px)
    {
      const fe sx = fe_from_u4(tab[H * 4 + 0], tab[H * 4 + 1]);
      const fe sy = fe_from_u4(tab[H * 4 + 2], tab[H * 4 + 3]);
      fe nx, ny;
      affine_add_inv(nx, ny, px, py, sx, sy, inv);
      px = nx, py = ny;
    }
  }
  PROBE_PIPE_FINISH(HBM)
}

// ---------------------------------------------------------------- K1-sp: software-pipelined addr33 variant
// Same work as add_kernel<H, true, false, false>, restructured so that every basic block of the hot loop holds
// one hash160 (ALU-pipe: LOP3/SHF/IADD3) AND the field arithmetic that produces the next point (FMA-pipe:
// IMAD.WIDE): ptxas interleaves the two independent streams, so the two integer pipes work at the same time
// instead of taking turns (in add_kernel the whole CTA alternates between an FMA-bound field phase and an
// ALU-bound hash phase). One pass-2 step = block X: hash(P - (i+1)G) || form P + (i+1)G, then block Y:
// hash(P + (i+1)G) || peel the inverse of step i-1 and form P - iG.

// hash160 of one compressed point, probe, report; used where nothing is pipelined (once per group)
static __device__ __noinline__ void check_one_slow(const BloomView &bv, const HitSink &sink, const fe &x, u32 y0, u64 off,
                                                   bool active) {
  u32 xx[1][8], odd[1] = {y0};
#pragma unroll
  for (int l = 0; l < 8; ++l) xx[0][l] = x.v[l];
  vw<1> h[5];
  hash160_33<1, 0>(h, xx, odd);
  const u32 hh[5] = {h[0].l[0], h[1].l[0], h[2].l[0], h[3].l[0], h[4].l[0]};
  if (bloom_has(bv, hh) && active) emit_hit(sink, off, hh, 0, 0);
}

// Pass 1 of the NEXT group (its prefix products) rides in block X as well, so after the first group of a launch
// there is no field-only phase left: the group step is the element peeled FIRST (it is multiplied in last), which
// makes the next centre known at the start of pass 2, and the prefixes go to the other half of a ping-pong scratch.
// Elements of a group: f_i = table[i].x - px (i < H), f_H = step.x - px; scratch entry k holds q_k = f_0 ... f_{k-1}.
template <int H, bool HBM>
__global__ void __launch_bounds__(ADD_THREADS, ADD_MIN_BLOCKS) add_kernel_sp(const AddParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) u64 mbar;
  uint4 *tab = reinterpret_cast<uint4 *>(smem_raw);
  u64 *sbloom = reinterpret_cast<u64 *>(smem_raw + (H + 1) * 64);
  const u32 tab_bytes = (H + 1) * 64;
  const u32 bloom_bytes = ((p.bloom_smem_words * 8u + 15u) / 16u) * 16u;

  if (threadIdx.x == 0) mbar_init(&mbar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&mbar, tab_bytes + bloom_bytes);
    bulk_g2s(tab, p.table, tab_bytes, &mbar);
    if (bloom_bytes) bulk_g2s(sbloom, p.bloom.bits, bloom_bytes, &mbar);
  }
  mbar_wait(&mbar, 0);

  BloomView bv = p.bloom;
  if (p.bloom_smem_words) bv.bits = sbloom;
  PROBE_PIPE_SETUP(HBM, (H + 1) * 64)

  const u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 T = p.T;
  const bool owner = tid < T;
  const u32 t = owner ? tid : T - 1;

  fe px, py;
#pragma unroll
  for (int l = 0; l < 8; ++l) px.v[l] = p.cx[(size_t)l * T + t], py.v[l] = p.cy[(size_t)l * T + t];

  const u64 g0 = (u64)t * p.groups_per_thread;
  const u64 g1 = g0 + p.groups_per_thread;
  const size_t TS = (size_t)gridDim.x * blockDim.x;  // scratch stride (scratch is sized for whole CTAs)
  uint4 *scr_cur = p.scratch + tid;                  // entry k, half h at [(2k + h) * TS]
  uint4 *scr_nxt = scr_cur + (size_t)2 * (H + 1) * TS;
  const fe sx = fe_from_u4(tab[H * 4 + 0], tab[H * 4 + 1]);

  // ---- pass 1 of this thread's first group (the only one that is not hidden behind hashing)
  fe tot;
  {
    fe acc = fe_one();
#pragma unroll 1
    for (int k = 0; k < H; ++k) {
      scr_cur[(size_t)(2 * k) * TS] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
      scr_cur[(size_t)(2 * k + 1) * TS] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
      acc = fe_mul(acc, fe_sub(fe_from_u4(tab[k * 4 + 0], tab[k * 4 + 1]), px));
    }
    scr_cur[(size_t)(2 * H) * TS] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
    scr_cur[(size_t)(2 * H + 1) * TS] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
    tot = fe_mul(acc, fe_sub(sx, px));
  }

#pragma unroll 1
  for (u64 g = g0; g < g1; ++g) {
    const bool active = owner && g < p.n_groups;
    const u64 kc = p.key_off0 + g * (2 * H) + H;

    fe inv = fe_inv(tot);  // 1 / (f_0 ... f_H)

    // ---- the group step first: next centre N = P + 2H*s*G
    fe nx, ny;
    {
      const fe qH = fe_from_u4(scr_cur[(size_t)(2 * H) * TS], scr_cur[(size_t)(2 * H + 1) * TS]);
      const fe inv_s = fe_mul(inv, qH);  // 1 / f_H
      inv = fe_mul(inv, fe_sub(sx, px));  // 1 / q_H
      const fe sy = fe_from_u4(tab[H * 4 + 2], tab[H * 4 + 3]);
      affine_add_inv(nx, ny, px, py, sx, sy, inv_s);
    }

    // the centre itself (key K) takes the slot of the far end K+H, which lies outside the group
    check_one_slow(bv, p.sink, px, py.v[0], kc, active);

    // ---- prologue: operands and inverse of step H-1, and its first point P - H*G
    fe gx = fe_from_u4(tab[(H - 1) * 4 + 0], tab[(H - 1) * 4 + 1]);
    fe gy = fe_from_u4(tab[(H - 1) * 4 + 2], tab[(H - 1) * 4 + 3]);
    fe inv_i;
    {
      const fe q = fe_from_u4(scr_cur[(size_t)(2 * (H - 1)) * TS], scr_cur[(size_t)(2 * (H - 1) + 1) * TS]);
      inv_i = fe_mul(inv, q);
      inv = fe_mul(inv, fe_sub(gx, px));
    }
    fe ax, ay;  // the point waiting to be hashed
    affine_add_inv(ax, ay, px, py, gx, fe_neg(gy), inv_i);
    fe accn = fe_one();  // q'_k of the next group

    // ---- pass 2, pipelined
#pragma unroll 1
    for (int i = H - 1; i >= 1; --i) {
      if (ECL_HASH_SYNC) __syncthreads();
      // block X: hash P - (i+1)G  ||  form P + (i+1)G, and one pass-1 step of the next group
      u32 xx[1][8], odd[1];
      vw<1> h[5];
#pragma unroll
      for (int l = 0; l < 8; ++l) xx[0][l] = ax.v[l];
      odd[0] = ay.v[0];
      hash160_33<1, 0>(h, xx, odd);
      fe bx, by;
      affine_add_inv(bx, by, px, py, gx, gy, inv_i);
      {
        const int k = H - 1 - i;
        scr_nxt[(size_t)(2 * k) * TS] = make_uint4(accn.v[0], accn.v[1], accn.v[2], accn.v[3]);
        scr_nxt[(size_t)(2 * k + 1) * TS] = make_uint4(accn.v[4], accn.v[5], accn.v[6], accn.v[7]);
        accn = fe_mul(accn, fe_sub(fe_from_u4(tab[k * 4 + 0], tab[k * 4 + 1]), nx));
      }
      {
        const u32 hh[5] = {h[0].l[0], h[1].l[0], h[2].l[0], h[3].l[0], h[4].l[0]};
        probe_hash_dyn(pipe, bv, p.sink, hh, kc - (u64)(i + 1), 0u, 0u, active);
      }
      // block Y: hash P + (i+1)G  ||  peel step i-1 and form P - iG
#pragma unroll
      for (int l = 0; l < 8; ++l) xx[0][l] = bx.v[l];
      odd[0] = by.v[0];
      hash160_33<1, 0>(h, xx, odd);
      {
        const fe q = fe_from_u4(scr_cur[(size_t)(2 * (i - 1)) * TS], scr_cur[(size_t)(2 * (i - 1) + 1) * TS]);
        gx = fe_from_u4(tab[(i - 1) * 4 + 0], tab[(i - 1) * 4 + 1]);
        gy = fe_from_u4(tab[(i - 1) * 4 + 2], tab[(i - 1) * 4 + 3]);
        inv_i = fe_mul(inv, q);
        inv = fe_mul(inv, fe_sub(gx, px));
        affine_add_inv(ax, ay, px, py, gx, fe_neg(gy), inv_i);
      }
      {
        const u32 hh[5] = {h[0].l[0], h[1].l[0], h[2].l[0], h[3].l[0], h[4].l[0]};
        probe_hash_dyn(pipe, bv, p.sink, hh, kc + (u64)(i + 1), 0u, 0u, active && i != H - 1);
      }
    }
    // ---- epilogue: step 0 (keys K-1 and K+1), the last two prefixes of the next group
    {
      fe bx, by;
      affine_add_inv(bx, by, px, py, gx, gy, inv_i);
      check_one_slow(bv, p.sink, ax, ay.v[0], kc - 1, active);
      check_one_slow(bv, p.sink, bx, by.v[0], kc + 1, active);
      scr_nxt[(size_t)(2 * (H - 1)) * TS] = make_uint4(accn.v[0], accn.v[1], accn.v[2], accn.v[3]);
      scr_nxt[(size_t)(2 * (H - 1) + 1) * TS] = make_uint4(accn.v[4], accn.v[5], accn.v[6], accn.v[7]);
      accn = fe_mul(accn, fe_sub(fe_from_u4(tab[(H - 1) * 4 + 0], tab[(H - 1) * 4 + 1]), nx));
      scr_nxt[(size_t)(2 * H) * TS] = make_uint4(accn.v[0], accn.v[1], accn.v[2], accn.v[3]);
      scr_nxt[(size_t)(2 * H + 1) * TS] = make_uint4(accn.v[4], accn.v[5], accn.v[6], accn.v[7]);
      tot = fe_mul(accn, fe_sub(sx, nx));
    }
    px = nx, py = ny;
    uint4 *sw = scr_cur;
    scr_cur = scr_nxt, scr_nxt = sw;
  }
  PROBE_PIPE_FINISH(HBM)
}

