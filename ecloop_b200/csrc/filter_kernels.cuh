// filter_kernels.cuh — bloom-filter tooling on the device: the synthetic filter generator used by the
// config-4 benchmark, and the entry point of the GPU blf-gen insert loop (filter_add.cu; lib/utils.c:409-475).
#pragma once
#include <stdint.h>

#include <cuda_runtime.h>

#include "bloom.cuh"

// splitmix64 output function over a counter: r(c) = mix(seed + (c + 1) * gamma). Stateless, so any word of the filter
// can be produced anywhere (the tests mirror it in numpy).
__host__ __device__ __forceinline__ u64 filter_mix64(u64 seed, u64 counter) {
  u64 z = seed + (counter + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

// word i: bit 8k + b is set iff byte b of r(8i + k) is below thr (0..256): every bit i.i.d. with p = thr / 256
static __global__ void __launch_bounds__(256) filter_generate_kernel(u64 *bits, u64 n, u32 thr, u64 seed) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    u64 w = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const u64 r = filter_mix64(seed, i * 8 + k);
#pragma unroll
      for (int b = 0; b < 8; ++b) w |= (u64)(((u32)(r >> (8 * b)) & 255u) < thr) << (8 * k + b);
    }
    bits[i] = w;
  }
}

// blf_gen's `if (blf_has) continue; blf_add; count++` over n hashes (host pointer) in input order, exact count.
// Returns 0, or -1 after a CUDA error (cudaGetLastError has it).
int filter_add_device(cudaStream_t stream, BloomView view, u64 *bits, const uint32_t (*h160)[5], uint32_t n, unsigned long long *n_new);
