// kernels.cuh — the small sm_100a kernels behind the C-ABI (the fused add kernel lives in add_kernel.cuh and is
// compiled one variant per translation unit, add_inst.cu):
//
//   mul_points_kernel            K2a: ec_gtable_mul + ec_jacobi_grprdc (main.c:531-532)
//   mul_hash_kernel<A33,A65,NW>  K2b: check_found_mul (main.c:458-484): hash160 + blf_has
//   smul_kernel                  K3: k*G for generated or given scalars -> +-i*s*G table / thread centres
//                                (ctx_precompute_gpoints main.c:219-246, GStart main.c:359-360)
//   gtab_bases/gtab_fill         one-off window table d * 2^(16 w) * G (ec_gtable_init, lib/ecc.c:880-905)
//   prim_* kernels               per-routine parity entry points
#pragma once
#include "add_kernel.cuh"
#include "common.cuh"
#ifdef ECL_EXPERIMENTAL
#include "fp64mul.cuh"  // FP64-pipe field arithmetic: parity-tested experiment, not on the hot path, not in the product library
#endif

// ---------------------------------------------------------------- K3: scalar multiples of G

struct SmulParams {
  fe k0, step;         // generated scalars: k = k0 + m(j)*step (mod n)
  const fe *scalars;   // mode 2: explicit scalars instead
  const uint4 *gtab;
  u32 count;
  u32 mode;  // 0: add-kernel table (m = j+1 for j < count-1, m = 2*(count-1) for the last = group step), AoS out
             // 1: thread centres (m = j), SoA out with stride `count`
             // 2: explicit scalars, AoS out (zeros for the point at infinity)
  u32 *out;
  fe extra_k;      // thread `count` (one past the last) computes extra_k * G, AoS, into extra_out when that is set:
  u32 *extra_out;  // the group step 2*Hr*s*G of an add launch whose Hr has no table entry (ecl_api.cu plan_launch)
};

__global__ void __launch_bounds__(128) smul_kernel(const SmulParams p) {
  const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == p.count && p.extra_out) {
    jac a;
    fe x = fe_zero(), y = fe_zero();
    if (gtab_mul(a, p.extra_k, p.gtab)) jac_to_affine(x, y, a);
    for (int l = 0; l < 8; ++l) p.extra_out[l] = x.v[l], p.extra_out[8 + l] = y.v[l];
    return;
  }
  if (j >= p.count) return;
  fe k;
  if (p.mode == 2) k = p.scalars[j];
  else {
    u32 m = j;
    if (p.mode == 0) m = (j + 1 < p.count) ? j + 1 : 2 * (p.count - 1);
    k = sc_muladd_small(p.k0, p.step, m);
  }
  jac a;
  fe x = fe_zero(), y = fe_zero();
  if (gtab_mul(a, k, p.gtab)) jac_to_affine(x, y, a);
  if (p.mode == 1) {
    for (int l = 0; l < 8; ++l) p.out[(size_t)l * p.count + j] = x.v[l], p.out[(size_t)(8 + l) * p.count + j] = y.v[l];
  } else {
    for (int l = 0; l < 8; ++l) p.out[(size_t)j * 16 + l] = x.v[l], p.out[(size_t)j * 16 + 8 + l] = y.v[l];
  }
}

// ---------------------------------------------------------------- K2: mul path
// ec_gtable_mul x n + ec_jacobi_grprdc + check_found_mul (main.c:531-534) as two kernels with different shapes:
//   K2a mul_points_kernel   field work only (FMA + ALU pipes, 100 registers): k*G per key from the window table in
//                           Jacobian coordinates, then the batch normalisation of lib/ecc.c:695-707 per thread (B keys
//                           share one Fermat inversion); affine (x, y) go back into the key's scratch slot;
//   K2b mul_hash_kernel     hashing only (ALU pipe): SHA-256 -> RIPEMD-160 of the 33 / 65 byte encodings + bloom probe,
//                           one CTA of 512 threads per SM in lockstep like the add kernel, filter in shared memory when
//                           it fits.
// Between them 64 B per key are written and read once (coalesced): 0.13 KB/key against ~25 k integer operations.

struct MulParams {
  const fe *scalars;
  const uint4 *gtab;
  uint4 *scratch;  // per key 8 x 16 B (X, Y, Z, prefix product; X, Y become affine x, y), slot-major: [(m*8 + q)*T + t]
  u32 count;  // keys
  u32 T;      // threads that own keys
  u32 B;      // keys per thread: thread t owns keys m*T + t, m < B (coalesced scalar loads)
};

__device__ __forceinline__ void st_fe(uint4 *p, size_t stride, const fe &a) {
  p[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
  p[stride] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
__device__ __forceinline__ fe ld_fe(const uint4 *p, size_t stride) { return fe_from_u4(p[0], p[stride]); }

// acc = k*G like gtab_mul (ec.cuh) with the field multiplications inlined into ONE loop body (the setup-path gtab_mul
// calls fe_mul out of line through memory: fine for 75 000 centres per launch, 40 % of the time here).
// The first non-zero window is a load; the second is an affine + affine addition (4M + 2S); the rest are mixed
// additions (8M + 3S). Returns false for k = 0 (mod n).
#ifndef ECL_MULPTS_BF
#define ECL_MULPTS_BF 1  // K2a's products use the branch-free correction: independent products of an addition overlap
#endif
#if ECL_MULPTS_BF
#define FE_MUL_K2 fe_mul_bf
#define FE_SQR_K2 fe_sqr_bf
#else
#define FE_MUL_K2 fe_mul
#define FE_SQR_K2 fe_sqr
#endif
#ifndef ECL_MULPTS_PREFETCH
#define ECL_MULPTS_PREFETCH 1
#endif
__device__ __forceinline__ void gtab_prefetch(const uint4 *gtab, u32 entry) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(gtab + (size_t)entry * 4));
}
static __device__ __noinline__ bool gtab_mul_fast(jac &acc, const fe &k, const uint4 *__restrict__ gtab) {
  int have = 0, dead = 0;  // dead: the sum hit infinity (k = 0 mod n); the loop still runs to its end (lockstep barriers)
  u32 dig[GTAB_WINDOWS];
#pragma unroll
  for (int w = 0; w < GTAB_WINDOWS; ++w) dig[w] = gtab_digit(k, w);
#if ECL_MULPTS_PREFETCH
  if (dig[0]) gtab_prefetch(gtab, dig[0] - 1);
#endif
#pragma unroll 1
  for (int w = 0; w < GTAB_WINDOWS; ++w) {
#if ECL_MULPTS_SYNC
    __syncthreads();
#endif
#if ECL_MULPTS_PREFETCH
    // the table is far larger than L2: the next window's entry is on its way from HBM while this window's addition
    // (~2000 instructions) runs; a prefetch costs no register
    if (w + 1 < GTAB_WINDOWS && dig[w + 1]) gtab_prefetch(gtab, (u32)(w + 1) * GTAB_STRIDE + dig[w + 1] - 1);
#endif
    const u32 d = dig[w];
    if (d == 0 || dead) continue;
    fe qx, qy;
    gtab_load(qx, qy, gtab, (u32)w * GTAB_STRIDE + d - 1);
    if (!have) {
      acc.x = qx, acc.y = qy, acc.z = fe_one();
      have = 1;
      continue;
    }
    // mixed addition; with Z1 = 1 the first three products are copies, kept in one code path by multiplying by one:
    // the second window is 1 of 15 additions, a separate body would double the loop for a 3 % gain
    const fe z2 = FE_SQR_K2(acc.z);
    const fe u2 = FE_MUL_K2(qx, z2);
    const fe s2 = FE_MUL_K2(FE_MUL_K2(qy, z2), acc.z);
    const fe h = fe_sub(u2, acc.x);
    const fe rr = fe_sub(s2, acc.y);
    const fe h2 = FE_SQR_K2(h);
    const fe h3 = FE_MUL_K2(h2, h);
    const fe v = FE_MUL_K2(acc.x, h2);
    const fe x3 = fe_sub(fe_sub(fe_sub(FE_SQR_K2(rr), h3), v), v);
    const fe y3 = fe_sub(FE_MUL_K2(rr, fe_sub(v, x3)), FE_MUL_K2(acc.y, h3));
    const fe z3 = FE_MUL_K2(acc.z, h);
    if (fe_is_zero(z3)) dead = 1;  // k = n lands on -acc at the top window: infinity
    acc.x = x3, acc.y = y3, acc.z = z3;
  }
  return have != 0 && !dead;
}

#ifndef ECL_MULPTS_MINBLOCKS
#define ECL_MULPTS_MINBLOCKS 2
#endif
#ifndef ECL_MULPTS_SYNC
#define ECL_MULPTS_SYNC 0  // 1: the CTA walks the windows in lockstep (barrier per window) to share instruction fetches
#endif
__global__ void __launch_bounds__(256, ECL_MULPTS_MINBLOCKS) mul_points_kernel(const MulParams p) {
  const u32 t0 = blockIdx.x * blockDim.x + threadIdx.x;
#if ECL_MULPTS_SYNC
  const u32 t = t0 < p.T ? t0 : p.T - 1;  // threads past the end run along on the last thread's keys and store nothing
  const bool store = t0 < p.T;
#else
  if (t0 >= p.T) return;
  const u32 t = t0;
  const bool store = true;
#endif
  const size_t T = p.T;
  uint4 *scr = p.scratch + t;
  fe acc = fe_one();
#pragma unroll 1
  for (u32 m = 0; m < p.B; ++m) {
    const u32 j = m * p.T + t;
    jac a;
    bool ok = false;
#if ECL_MULPTS_SYNC
    ok = gtab_mul_fast(a, p.scalars[j < p.count ? j : p.count - 1], p.gtab) && j < p.count;
#else
    if (j < p.count) ok = gtab_mul_fast(a, p.scalars[j], p.gtab);
#endif
    if (!ok) a.x = fe_zero(), a.y = fe_zero(), a.z = fe_one();
    uint4 *slot = scr + (size_t)m * 8 * T;
    if (store) st_fe(slot, T, a.x), st_fe(slot + 2 * T, T, a.y), st_fe(slot + 4 * T, T, a.z), st_fe(slot + 6 * T, T, acc);
    acc = FE_MUL_K2(acc, a.z);
  }
  fe inv = fe_inv(acc);
#pragma unroll 1
  for (int m = (int)p.B - 1; m >= 0; --m) {
    uint4 *slot = scr + (size_t)m * 8 * T;
    const fe z = ld_fe(slot + 4 * T, T), pre = ld_fe(slot + 6 * T, T);
    const fe zi = FE_MUL_K2(inv, pre);
    inv = FE_MUL_K2(inv, z);
    const fe ax = ld_fe(slot, T), ay = ld_fe(slot + 2 * T, T);
    const fe zi2 = FE_SQR_K2(zi);
    const fe fx = FE_MUL_K2(ax, zi2), fy = FE_MUL_K2(ay, FE_MUL_K2(zi2, zi));  // (0, 0) stays (0, 0): "no point for this key"
    if (store) st_fe(slot, T, fx), st_fe(slot + 2 * T, T, fy);
  }
}

struct MulHashParams {
  const uint4 *scratch;  // affine x, y of key j = m*T + t at [(m*8 + {0,1 | 2,3})*T + t]
  BloomView bloom;
  u32 bloom_smem_words;  // != 0: staged into shared memory
  HitSink sink;
  u32 T;
  u32 begin, end;  // keys [begin, end) of the batch
};

template <bool A33, bool A65, int NW>
__global__ void __launch_bounds__(512, 1) mul_hash_kernel(const MulHashParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) u64 mbar;
  u64 *sbloom = reinterpret_cast<u64 *>(smem_raw);
  BloomView bv = p.bloom;
  if (p.bloom_smem_words) {
    const u32 bytes = ((p.bloom_smem_words * 8u + 15u) / 16u) * 16u;
    if (threadIdx.x == 0) mbar_init(&mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&mbar, bytes);
      bulk_g2s(sbloom, p.bloom.bits, bytes, &mbar);
    }
    mbar_wait(&mbar, 0);
    bv.bits = sbloom;
  }
  const u32 lanes = gridDim.x * blockDim.x * NW;
  const u32 n = p.end - p.begin;
  const u32 rounds = (n + lanes - 1) / lanes;  // the same for every thread: the CTA stays in lockstep
#pragma unroll 1
  for (u32 r = 0; r < rounds; ++r) {
    __syncthreads();
    u32 x[NW][8], y[NW][8];
    u64 off[NW];
    bool act[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const u32 i = r * lanes + (u32)w * (lanes / NW) + blockIdx.x * blockDim.x + threadIdx.x;
      const u32 j = p.begin + (i < n ? i : n - 1);
      const u32 m = j / p.T, t = j - m * p.T;
      const uint4 *slot = p.scratch + (size_t)m * 8 * p.T + t;
      const fe fx = ld_fe(slot, p.T), fy = ld_fe(slot + 2 * (size_t)p.T, p.T);
#pragma unroll
      for (int l = 0; l < 8; ++l) x[w][l] = fx.v[l], y[w][l] = fy.v[l];
      off[w] = j;
      act[w] = i < n && !(fe_is_zero(fx) && fe_is_zero(fy));
    }
    NoPipe none;
    check_points<NW, A33, A65, false, 1>(bv, p.sink, x, y, off, act, none);
  }
}

// ---------------------------------------------------------------- window table build (one-off per device)
// ec_gtable_init (lib/ecc.c:880-905) on the device, in three steps: the window bases B_w = 2^(W w) G, a small table
// k * B_w (k <= 256) per window, and the fill: thread (w, c) owns the 256 entries (256 c + k) * B_w = S + k * B_w with
// S = 256 c * B_w — 255 affine additions that share ONE inversion (Montgomery's trick, prefix products parked in the
// x field of the entries they belong to). ~9 field multiplications per entry: 46 M entries (W = 22) take ~10 ms.

// bases[w] = 2^(W w) * G, affine (one thread per window)
__global__ void gtab_bases_kernel(u32 *bases) {
  const u32 w = threadIdx.x;
  if (w >= GTAB_WINDOWS) return;
  jac a;
  a.x.v[0] = 0x16f81798u, a.x.v[1] = 0x59f2815bu, a.x.v[2] = 0x2dce28d9u, a.x.v[3] = 0x029bfcdbu;
  a.x.v[4] = 0xce870b07u, a.x.v[5] = 0x55a06295u, a.x.v[6] = 0xf9dcbbacu, a.x.v[7] = 0x79be667eu;
  a.y.v[0] = 0xfb10d4b8u, a.y.v[1] = 0x9c47d08fu, a.y.v[2] = 0xa6855419u, a.y.v[3] = 0xfd17b448u;
  a.y.v[4] = 0x0e1108a8u, a.y.v[5] = 0x5da4fbfcu, a.y.v[6] = 0x26a3c465u, a.y.v[7] = 0x483ada77u;
  a.z = fe_one();
  for (u32 i = 0; i < w * GTAB_W; ++i) {
    jac t;
    jac_dbl(t, a);
    a = t;
  }
  fe x, y;
  jac_to_affine(x, y, a);
  for (int l = 0; l < 8; ++l) bases[w * 16 + l] = x.v[l], bases[w * 16 + 8 + l] = y.v[l];
}

// m * (bx, by), m >= 1, by left-to-right double-and-add, affine result
static __device__ __noinline__ void small_multiple(fe &x, fe &y, const fe &bx, const fe &by, u32 m) {
  jac a;
  a.x = bx, a.y = by, a.z = fe_one();
  int bit = 31 - __clz(m);
  for (--bit; bit >= 0; --bit) {
    jac t;
    jac_dbl(t, a);
    a = t;
    if ((m >> bit) & 1) {
      jac_madd(t, a, bx, by);
      a = t;
    }
  }
  jac_to_affine(x, y, a);
}

// small[w][k-1] = k * bases[w], k = 1..GTAB_CHUNK
__global__ void __launch_bounds__(128) gtab_small_kernel(u32 *small, const u32 *bases) {
  const u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= GTAB_WINDOWS * GTAB_CHUNK) return;
  const u32 w = idx / GTAB_CHUNK, k = idx % GTAB_CHUNK + 1;
  fe bx, by, x, y;
  for (int l = 0; l < 8; ++l) bx.v[l] = bases[w * 16 + l], by.v[l] = bases[w * 16 + 8 + l];
  small_multiple(x, y, bx, by, k);
  for (int l = 0; l < 8; ++l) small[(size_t)idx * 16 + l] = x.v[l], small[(size_t)idx * 16 + 8 + l] = y.v[l];
}

__device__ __forceinline__ void st_point(uint4 *e, const fe &x, const fe &y) {
  e[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]), e[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
  e[2] = make_uint4(y.v[0], y.v[1], y.v[2], y.v[3]), e[3] = make_uint4(y.v[4], y.v[5], y.v[6], y.v[7]);
}

// thread (w, c): entries d = 256 c + k, k < 256, of window w (slot w * GTAB_STRIDE + d - 1)
__global__ void __launch_bounds__(128) gtab_fill_kernel(uint4 *gtab, const uint4 *small) {
  const u32 full = GTAB_STRIDE / GTAB_CHUNK, top = (1u << GTAB_TOP_BITS) / GTAB_CHUNK;  // chunks per window
  const u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (GTAB_WINDOWS - 1) * full + top) return;
  const u32 w = idx / full < GTAB_WINDOWS - 1 ? idx / full : GTAB_WINDOWS - 1;
  const u32 c = idx - w * full;
  const uint4 *tk = small + (size_t)w * GTAB_CHUNK * 4;           // tk[(k-1)*4 ..] = k * B_w
  uint4 *out = gtab + ((size_t)w * GTAB_STRIDE + (size_t)c * GTAB_CHUNK) * 4 - 4;  // out[k*4 ..] = slot of d = 256 c + k
  if (c == 0) {
    for (u32 k = 1; k < GTAB_CHUNK; ++k)
      for (int q = 0; q < 4; ++q) out[k * 4 + q] = tk[(k - 1) * 4 + q];
    return;
  }
  fe bx, by, sx, sy;  // S = c * (256 B_w)
  bx = fe_from_u4(tk[(GTAB_CHUNK - 1) * 4 + 0], tk[(GTAB_CHUNK - 1) * 4 + 1]);
  by = fe_from_u4(tk[(GTAB_CHUNK - 1) * 4 + 2], tk[(GTAB_CHUNK - 1) * 4 + 3]);
  small_multiple(sx, sy, bx, by, c);
  st_point(out, sx, sy);  // k = 0
  fe acc = fe_one();
  for (u32 k = 1; k < GTAB_CHUNK; ++k) {  // prefix products, parked in the x field of entry k
    out[k * 4 + 0] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
    out[k * 4 + 1] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
    acc = fe_mul_noinline(acc, fe_sub(fe_from_u4(tk[(k - 1) * 4 + 0], tk[(k - 1) * 4 + 1]), sx));
  }
  fe inv = fe_inv(acc);
  for (u32 k = GTAB_CHUNK - 1; k >= 1; --k) {
    const fe pre = fe_from_u4(out[k * 4 + 0], out[k * 4 + 1]);
    const fe qx = fe_from_u4(tk[(k - 1) * 4 + 0], tk[(k - 1) * 4 + 1]), qy = fe_from_u4(tk[(k - 1) * 4 + 2], tk[(k - 1) * 4 + 3]);
    const fe inv_k = fe_mul_noinline(inv, pre);
    inv = fe_mul_noinline(inv, fe_sub(qx, sx));
    const fe lam = fe_mul_noinline(fe_sub(qy, sy), inv_k);
    const fe rx = fe_sub(fe_sub(fe_mul_noinline(lam, lam), sx), qx);
    const fe ry = fe_sub(fe_mul_noinline(lam, fe_sub(sx, rx)), sy);
    st_point(out + k * 4, rx, ry);
  }
}

// ---------------------------------------------------------------- primitive parity kernels

__global__ void prim_fp_kernel(int op, const fe *a, const fe *b, fe *out, u32 n) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const fe x = a[i];
  fe y = fe_zero();
  if (b) y = b[i];
  fe r;
  switch (op) {
  case ECL_OP_MUL: r = fe_mul(x, y); break;
  case ECL_OP_SQR: r = fe_sqr(x); break;
  case ECL_OP_ADD: r = fe_add(x, y); break;
  case ECL_OP_SUB: r = fe_sub(x, y); break;
  case ECL_OP_NEG: r = fe_neg(x); break;
#ifdef ECL_EXPERIMENTAL
  case ECL_OP_MUL_F64: r = fe6_to_fe(fe6_mul(fe6_from_fe(x), fe6_from_fe(y))); break;
  case ECL_OP_MUL_F64_CHAIN: {  // x * y^16 without leaving the weak 44-bit-limb form in between
    fe6 acc = fe6_from_fe(x);
    const fe6 m = fe6_from_fe(y);
    for (int i = 0; i < 16; ++i) acc = fe6_mul(acc, m);
    r = fe6_to_fe(acc);
    break;
  }
  case ECL_OP_AFFINE_F64_X:
  case ECL_OP_AFFINE_F64_Y: {  // (x, y) + G by the affine formula, arithmetic in the FP64 limb form (inverse from fe_inv)
    fe gx, gy;
    gx.v[0] = 0x16f81798u, gx.v[1] = 0x59f2815bu, gx.v[2] = 0x2dce28d9u, gx.v[3] = 0x029bfcdbu;
    gx.v[4] = 0xce870b07u, gx.v[5] = 0x55a06295u, gx.v[6] = 0xf9dcbbacu, gx.v[7] = 0x79be667eu;
    gy.v[0] = 0xfb10d4b8u, gy.v[1] = 0x9c47d08fu, gy.v[2] = 0xa6855419u, gy.v[3] = 0xfd17b448u;
    gy.v[4] = 0x0e1108a8u, gy.v[5] = 0x5da4fbfcu, gy.v[6] = 0x26a3c465u, gy.v[7] = 0x483ada77u;
    fe xc = x, yc = y;
    fe_canon(xc), fe_canon(yc);
    const fe inv = fe_inv(fe_sub(gx, xc));
    fe6 rx, ry;
    fe6_affine_add(rx, ry, fe6_from_fe(xc), fe6_from_fe(yc), fe6_from_fe(gx), fe6_from_fe(gy), fe6_from_fe(inv));
    r = fe6_to_fe(op == ECL_OP_AFFINE_F64_X ? rx : ry);
    break;
  }
#endif
  default: r = fe_inv(x); break;
  }
  out[i] = r;
}

__global__ void prim_hash160_kernel(const u32 *xy, u32 *out33, u32 *out65, u32 n) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 x[1][8], y[1][8];
  for (int l = 0; l < 8; ++l) x[0][l] = xy[(size_t)i * 16 + l], y[0][l] = xy[(size_t)i * 16 + 8 + l];
  vw<1> h[5];
  if (out33) {
    const u32 odd[1] = {y[0][0]};
    hash160_33<1>(h, x, odd);
    for (int k = 0; k < 5; ++k) out33[(size_t)i * 5 + k] = h[k].l[0];
  }
  if (out65) {
    hash160_65<1>(h, x, y);
    for (int k = 0; k < 5; ++k) out65[(size_t)i * 5 + k] = h[k].l[0];
  }
}

__global__ void prim_bloom_kernel(BloomView bv, const u32 *h160, uint8_t *out, u32 n) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u32 hh[5] = {h160[i * 5], h160[i * 5 + 1], h160[i * 5 + 2], h160[i * 5 + 3], h160[i * 5 + 4]};
  out[i] = bloom_has(bv, hh) ? 1 : 0;
}
