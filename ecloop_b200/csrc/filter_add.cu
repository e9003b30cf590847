// filter_add.cu — blf-gen's insert loop on the GPU (lib/utils.c:453-465), byte-identical filter and exact count.
//
// The reference walks its input once: `if (blf_has(h)) continue; blf_add(h); count++`. The resulting bits are the OR
// of all hashes' 20 positions whatever the order (a skipped hash had nothing left to add); the count is not: hash i
// counts iff one of its positions is clear in the filter as it was AND is touched by no hash before i. So per chunk:
//   1. probe:  blf_has against the filter as it is; for a hash that fails, its still-clear positions become sort keys
//              (value = index of the hash; pairs are laid out in input order), everything else a sentinel key;
//   2. sort:   stable radix sort by position (CUB) — within a run of equal positions the first pair is the earliest hash;
//   3. heads:  the head of every run marks its hash as "new";
//   4. insert: atomicOr of all 20 positions of every hash that failed step 1; count the marked ones.
// Chunks are processed in order, so "the filter as it was" for chunk c includes chunks < c.
#include <cub/device/device_radix_sort.cuh>

#include "filter_kernels.cuh"

#define FA_CHUNK (1u << 22)  // hashes per pass: 84 M pairs, ~2 GB of sort buffers

__device__ __forceinline__ void blf_positions(u64 pos[20], const u32 h[5], u64 size, u64 magic) {
  const u64 a[5] = {(u64)h[0] << 32 | h[1], (u64)h[2] << 32 | h[3], (u64)h[4] << 32 | h[0], (u64)h[1] << 32 | h[2],
                    (u64)h[3] << 32 | h[4]};
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int S = s == 0 ? 24 : s == 1 ? 28 : s == 2 ? 36 : 40;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const u64 v = (a[i] << S) | (a[(i + 1) % 5] >> S);
      pos[s * 5 + i] = bloom_word_index(v >> 6, size, magic) * 64 + (v & 63);
    }
  }
}

static __global__ void __launch_bounds__(256) fa_probe_kernel(BloomView bv, const u32 *h160, u32 n, u64 *keys, u32 *vals, uint8_t *fresh,
                                                              uint8_t *is_new, u64 sentinel) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u32 h[5] = {h160[i * 5], h160[i * 5 + 1], h160[i * 5 + 2], h160[i * 5 + 3], h160[i * 5 + 4]};
  u64 pos[20];
  blf_positions(pos, h, bv.size, bv.magic);
  u32 clear = 0;
#pragma unroll
  for (int k = 0; k < 20; ++k)
    if (!((bv.bits[pos[k] >> 6] >> (pos[k] & 63)) & 1)) clear |= 1u << k;
  fresh[i] = clear != 0;  // blf_has is false
  is_new[i] = 0;
#pragma unroll
  for (int k = 0; k < 20; ++k) {
    keys[(size_t)i * 20 + k] = ((clear >> k) & 1) ? pos[k] : sentinel;
    vals[(size_t)i * 20 + k] = i;
  }
}

static __global__ void __launch_bounds__(256) fa_heads_kernel(const u64 *keys, const u32 *vals, size_t m, u64 sentinel, uint8_t *is_new) {
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (size_t)gridDim.x * blockDim.x) {
    const u64 k = keys[j];
    if (k != sentinel && (j == 0 || keys[j - 1] != k)) is_new[vals[j]] = 1;
  }
}

static __global__ void __launch_bounds__(256) fa_insert_kernel(u64 *bits, u64 size, u64 magic, const u32 *h160, u32 n, const uint8_t *fresh,
                                                               const uint8_t *is_new, unsigned long long *count) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  u32 mine = 0;
  if (i < n && fresh[i]) {
    const u32 h[5] = {h160[i * 5], h160[i * 5 + 1], h160[i * 5 + 2], h160[i * 5 + 3], h160[i * 5 + 4]};
    u64 pos[20];
    blf_positions(pos, h, size, magic);
#pragma unroll
    for (int k = 0; k < 20; ++k) atomicOr((unsigned long long *)&bits[pos[k] >> 6], 1ull << (pos[k] & 63));
    mine = is_new[i];
  }
  const u32 warp_total = __reduce_add_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && warp_total) atomicAdd(count, (unsigned long long)warp_total);
}

#define FA_CK(call)                      \
  do {                                   \
    if ((call) != cudaSuccess) goto bad; \
  } while (0)

int filter_add_device(cudaStream_t stream, BloomView view, u64 *bits, const uint32_t (*h160)[5], uint32_t n, unsigned long long *n_new) {
  const u32 chunk = n < FA_CHUNK ? n : FA_CHUNK;
  const size_t pairs = (size_t)chunk * 20;
  const u64 sentinel = view.size * 64;  // one past the largest position
  int end_bit = 1;
  while (end_bit < 64 && (sentinel >> end_bit) != 0) ++end_bit;
  u32 *d_h = nullptr, *d_v0 = nullptr, *d_v1 = nullptr;
  u64 *d_k0 = nullptr, *d_k1 = nullptr;
  uint8_t *d_fresh = nullptr, *d_new = nullptr;
  unsigned long long *d_count = nullptr, total = 0;
  void *d_tmp = nullptr;
  size_t tmp_bytes = 0;
  int rc = -1;
  FA_CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_k0, d_k1, d_v0, d_v1, pairs, 0, end_bit, stream));
  FA_CK(cudaMalloc(&d_h, (size_t)chunk * 20));
  FA_CK(cudaMalloc(&d_k0, pairs * 8));
  FA_CK(cudaMalloc(&d_k1, pairs * 8));
  FA_CK(cudaMalloc(&d_v0, pairs * 4));
  FA_CK(cudaMalloc(&d_v1, pairs * 4));
  FA_CK(cudaMalloc(&d_fresh, chunk));
  FA_CK(cudaMalloc(&d_new, chunk));
  FA_CK(cudaMalloc(&d_count, sizeof total));
  FA_CK(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 16));
  FA_CK(cudaMemsetAsync(d_count, 0, sizeof total, stream));
  for (u32 b = 0; b < n; b += chunk) {
    const u32 m = n - b < chunk ? n - b : chunk;
    FA_CK(cudaMemcpyAsync(d_h, h160 + b, (size_t)m * 20, cudaMemcpyHostToDevice, stream));
    fa_probe_kernel<<<(m + 255) / 256, 256, 0, stream>>>(view, d_h, m, d_k0, d_v0, d_fresh, d_new, sentinel);
    FA_CK(cudaGetLastError());
    size_t tb = tmp_bytes;
    FA_CK(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_k0, d_k1, d_v0, d_v1, (size_t)m * 20, 0, end_bit, stream));
    fa_heads_kernel<<<1184, 256, 0, stream>>>(d_k1, d_v1, (size_t)m * 20, sentinel, d_new);
    FA_CK(cudaGetLastError());
    fa_insert_kernel<<<(m + 255) / 256, 256, 0, stream>>>(bits, view.size, view.magic, d_h, m, d_fresh, d_new, d_count);
    FA_CK(cudaGetLastError());
    FA_CK(cudaStreamSynchronize(stream));  // the host buffer of this chunk is free again; keeps memory use bounded
  }
  FA_CK(cudaMemcpyAsync(&total, d_count, sizeof total, cudaMemcpyDeviceToHost, stream));
  FA_CK(cudaStreamSynchronize(stream));
  *n_new = total;
  rc = 0;
bad:
  cudaFree(d_h), cudaFree(d_k0), cudaFree(d_k1), cudaFree(d_v0), cudaFree(d_v1), cudaFree(d_fresh), cudaFree(d_new), cudaFree(d_count),
      cudaFree(d_tmp);
  return rc;
}
