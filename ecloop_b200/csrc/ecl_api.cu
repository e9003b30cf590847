// ecl_api.cu — host side of libecloop_b200.so: the C-ABI of include/ecloop_b200.h over the kernels in
// kernels.cuh. Host work here is bookkeeping only (launch geometry, 256-bit scalar offsets mod n, hit sorting);
// every field/curve/hash/bloom operation runs on the GPU. There is no CPU fallback.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "filter_kernels.cuh"
#include "launch_plan.h"
#include "kernels.cuh"
#include "peak.cuh"

#define GROUP_KEYS (2u * ADD_H)
#define HR_MIN 64u                        // smallest half group the planner picks (tiny spans still use many threads)
#define DEFAULT_GROUPS_PER_THREAD 64u     // bounds one launch to Tmax x 64 x 2048 keys (~1.6 s of addr33 work)
#define DEFAULT_HIT_CAP (1u << 20)
#define MAX_HITS_PER_KEY 12u  // {33,65} x 6 endomorphism images
#define SMEM_FILTER_MAX (44u * 1024u)  // filters up to this size ride in shared memory beside the table

typedef unsigned __int128 u128;

// ---------------------------------------------------------------- device object

struct mul_slot {  // one mul submit in flight (ECL_MUL_DEPTH of them)
  fe *h_keys = nullptr;  // pinned staging copy of the caller's keys
  fe *d_keys = nullptr;
  uint4 *d_scratch = nullptr;
  u32 cap = 0;  // keys the three buffers hold
  ecl_hit *d_hits = nullptr;
  u32 *d_hit_count = nullptr;
  u32 *h_count = nullptr;  // pinned: the hit count comes back with the submit's own stream work
  u32 n = 0, flags = 0;
  bool busy = false;
  cudaEvent_t ev_begin = nullptr, ev_k0 = nullptr, ev_k1 = nullptr, ev_end = nullptr;
};

struct ecl_dev {
  int ordinal = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaStream_t io_stream = nullptr;  // result downloads of an older mul submit while a younger one computes
  char err[512] = {0};

  uint4 *gtab = nullptr;  // window table, GTAB_ENTRIES x 64 B
  u32 *bases = nullptr;

  uint4 *add_table = nullptr;  // (ADD_H+1) x 64 B for the current stride: (i+1)*s*G, i < H; entry H = 2H*s*G
  uint4 *step_buf = nullptr;   // 64 B: the group step of a launch whose Hr has no table entry
  bool table_valid = false;
  u64 stride[4] = {1, 0, 0, 0};

  u64 *bloom_bits = nullptr;
  u64 bloom_size = 0, bloom_magic = 0;
  double bloom_fill = 0.5;  // fraction of set bits (measured for filters that stay in HBM)

  u32 Tmax = 0;
  u32 *centres = nullptr;  // 16 x Tmax u32 (SoA x then y)
  uint4 *scratch = nullptr;

  ecl_hit *d_hits = nullptr;
  u32 *d_hit_count = nullptr;
  u32 *d_err = nullptr;  // degenerate-group flag of the add kernels
  u32 hit_cap = DEFAULT_HIT_CAP;
  u32 groups_per_thread = DEFAULT_GROUPS_PER_THREAD;

  // asynchronous probing of filters that do not fit shared memory (probe_pipe.cuh)
  uint4 *cand_entries = nullptr;
  u32 *cand_counts = nullptr;  // [grid_max] counts + 1 overflow flag behind them
  u32 grid_max = 0;
  u64 cand_cap = 0;            // entries in total
  bool force_inline = false;   // the last span overflowed the queue and is being redone with inline probes

  // pending work: one add submit, or up to ECL_MUL_DEPTH mul submits (collected in submission order)
  int pending = 0;  // 0 none, 1 add, 2 mul
  u64 p_start[4] = {0, 0, 0, 0};
  u64 p_keys = 0;
  u32 p_flags = 0;
  mul_slot mslot[ECL_MUL_DEPTH];
  u32 m_head = 0, m_count = 0;  // FIFO of busy mul slots: oldest = m_head
  std::vector<ecl_hit> result;  // filled when a collect had to re-run in exact mode
  bool result_ready = false;

  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  std::vector<cudaEvent_t> ev_pool;  // pairs around hot-kernel launches
  size_t ev_used = 0;
  float last_total_ms = 0, last_hot_ms = 0;
  u32 last_launches = 0, launches = 0;
};

static char g_open_err[512] = "";

static int fail(ecl_dev *d, int code, const char *fmt, ...) {
  char *buf = d ? d->err : g_open_err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, 512, fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(dev, ECL_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// After a failed submit / collect the handle goes back to "nothing pending" with the stream drained, so that the
// caller can go on (ECL_E_ARG, ECL_E_OVERFLOW, ECL_E_DEGENERATE) or close the device (ECL_E_CUDA: sticky CUDA errors
// need a new process). Used on every error exit of the submit / collect functions.
static int abandon(ecl_dev *dev, int rc) {
  cudaStreamSynchronize(dev->stream);
  cudaGetLastError();
  dev->pending = 0;
  dev->result_ready = false;
  dev->force_inline = false;
  for (auto &s : dev->mslot) s.busy = false;
  dev->m_head = dev->m_count = 0;
  return rc;
}
#define CKA(call)                                                                                                        \
  do {                                                                                                                   \
    cudaError_t e_ = (call);                                                                                             \
    if (e_ != cudaSuccess)                                                                                               \
      return abandon(dev, fail(dev, ECL_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); \
  } while (0)

// ---------------------------------------------------------------- host scalars mod n (bookkeeping only)

static const u64 N_ORDER[4] = {0xbfd25e8cd0364141ULL, 0xbaaedce6af48a03bULL, 0xfffffffffffffffeULL, 0xffffffffffffffffULL};
static const u64 N_COMP[3] = {0x402da1732fc9bebfULL, 0x4551231950b75fc4ULL, 0x1ULL};  // 2^256 - n

static bool ge_n(const u64 a[4]) {
  for (int i = 3; i >= 0; --i)
    if (a[i] != N_ORDER[i]) return a[i] > N_ORDER[i];
  return true;
}
// r = (a * m + b) mod n, m a 64-bit multiplier; a, b any 256-bit values
static void sc_muladd64(u64 r[4], const u64 a[4], u64 m, const u64 b[4]) {
  u64 x[6] = {0, 0, 0, 0, 0, 0};
  u128 c = 0;
  for (int i = 0; i < 4; ++i) {
    c += (u128)a[i] * m + b[i];
    x[i] = (u64)c;
    c >>= 64;
  }
  x[4] = (u64)c;
  x[5] = (u64)(c >> 64);
  while (x[4] | x[5]) {  // fold: hi * (2^256 - n) + lo
    const u64 h0 = x[4], h1 = x[5];
    x[4] = x[5] = 0;
    u128 cc = 0;
    for (int i = 0; i < 6; ++i) {
      cc += (u128)x[i] + (i < 3 ? (u128)h0 * N_COMP[i] : 0);
      x[i] = (u64)cc;
      cc >>= 64;
    }
    cc = 0;
    for (int i = 1; i < 6; ++i) {
      cc += (u128)x[i] + (i - 1 < 3 ? (u128)h1 * N_COMP[i - 1] : 0);
      x[i] = (u64)cc;
      cc >>= 64;
    }
  }
  if (ge_n(x)) {
    u128 bw = 0;
    for (int i = 0; i < 4; ++i) {
      u128 t = (u128)x[i] - N_ORDER[i] - (u64)bw;
      x[i] = (u64)t;
      bw = (t >> 64) & 1;
    }
  }
  memcpy(r, x, 32);
}
static fe to_fe(const u64 a[4]) {
  fe r;
  for (int i = 0; i < 4; ++i) r.v[2 * i] = (u32)a[i], r.v[2 * i + 1] = (u32)(a[i] >> 32);
  return r;
}

// ---------------------------------------------------------------- lifetime

extern "C" int ecl_abi_version(void) { return ECL_ABI_VERSION; }

extern "C" int ecl_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" const char *ecl_last_error(const ecl_dev *dev) { return dev ? dev->err : g_open_err; }

static int build_gtab(ecl_dev *dev) {
  CK(cudaMalloc(&dev->gtab, (size_t)GTAB_ENTRIES * 64));
  CK(cudaMalloc(&dev->bases, (size_t)GTAB_WINDOWS * 64 + (size_t)GTAB_WINDOWS * GTAB_CHUNK * 64));
  u32 *small = dev->bases + GTAB_WINDOWS * 16;  // k * B_w, k <= 256, behind the bases
  gtab_bases_kernel<<<1, 32, 0, dev->stream>>>(dev->bases);
  gtab_small_kernel<<<(GTAB_WINDOWS * GTAB_CHUNK + 127) / 128, 128, 0, dev->stream>>>(small, dev->bases);
  const u32 chunks = (GTAB_WINDOWS - 1) * (GTAB_STRIDE / GTAB_CHUNK) + (1u << GTAB_TOP_BITS) / GTAB_CHUNK;
  gtab_fill_kernel<<<(chunks + 127) / 128, 128, 0, dev->stream>>>(dev->gtab, (const uint4 *)small);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(dev->stream));
  return ECL_OK;
}

static void free_mul_slot(mul_slot &s) {
  cudaFreeHost(s.h_keys), cudaFreeHost(s.h_count), cudaFree(s.d_keys), cudaFree(s.d_scratch), cudaFree(s.d_hits), cudaFree(s.d_hit_count);
  for (cudaEvent_t ev : {s.ev_begin, s.ev_k0, s.ev_k1, s.ev_end})
    if (ev) cudaEventDestroy(ev);
  s = mul_slot();
}

extern "C" int ecl_open(ecl_dev **out, int ordinal) {
  if (!out) return fail(nullptr, ECL_E_ARG, "ecl_open: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, ECL_E_NODEV, "no CUDA device (%s); this library has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
  if (ordinal < 0 || ordinal >= n) return fail(nullptr, ECL_E_ARG, "device ordinal %d out of range (0..%d)", ordinal, n - 1);
  ecl_dev *dev = new ecl_dev();
  dev->ordinal = ordinal;
  int rc = [&]() -> int {
    CK(cudaSetDevice(ordinal));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ordinal));
    if (prop.major < 10) return fail(dev, ECL_E_NODEV, "device %d is sm_%d%d; this build targets sm_100a only", ordinal, prop.major, prop.minor);
    dev->sm_count = prop.multiProcessorCount;
    dev->grid_max = (u32)dev->sm_count * ADD_MIN_BLOCKS;
    dev->Tmax = dev->grid_max * ADD_THREADS;
    CK(cudaStreamCreateWithFlags(&dev->own_stream, cudaStreamNonBlocking));
    dev->stream = dev->own_stream;
    CK(cudaStreamCreateWithFlags(&dev->io_stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&dev->ev_begin));
    CK(cudaEventCreate(&dev->ev_end));
    CK(cudaMalloc(&dev->d_hit_count, 2 * sizeof(u32)));
    dev->d_err = dev->d_hit_count + 1;
    CK(cudaMemsetAsync(dev->d_hit_count, 0, 2 * sizeof(u32), dev->stream));
    CK(cudaMalloc(&dev->d_hits, (size_t)dev->hit_cap * sizeof(ecl_hit)));
    CK(cudaMalloc(&dev->add_table, (size_t)(ADD_H + 1) * 64));
    CK(cudaMalloc(&dev->step_buf, 64));
    return build_gtab(dev);
  }();
  if (rc != ECL_OK) {
    snprintf(g_open_err, sizeof g_open_err, "%s", dev->err);
    ecl_close(dev);
    return rc;
  }
  *out = dev;
  return ECL_OK;
}

extern "C" void ecl_close(ecl_dev *dev) {
  if (!dev) return;
  cudaSetDevice(dev->ordinal);
  cudaDeviceSynchronize();
  cudaFree(dev->gtab), cudaFree(dev->bases), cudaFree(dev->add_table), cudaFree(dev->step_buf), cudaFree(dev->bloom_bits);
  cudaFree(dev->centres), cudaFree(dev->scratch), cudaFree(dev->d_hits), cudaFree(dev->d_hit_count);
  cudaFree(dev->cand_entries), cudaFree(dev->cand_counts);
  for (auto &s : dev->mslot) free_mul_slot(s);
  for (auto ev : dev->ev_pool) cudaEventDestroy(ev);
  if (dev->ev_begin) cudaEventDestroy(dev->ev_begin);
  if (dev->ev_end) cudaEventDestroy(dev->ev_end);
  if (dev->own_stream) cudaStreamDestroy(dev->own_stream);
  if (dev->io_stream) cudaStreamDestroy(dev->io_stream);
  delete dev;
}

extern "C" int ecl_set_stream(ecl_dev *dev, void *cuda_stream) {
  if (!dev) return ECL_E_ARG;
  if (dev->pending) return fail(dev, ECL_E_STATE, "ecl_set_stream with work pending");
  dev->stream = cuda_stream ? (cudaStream_t)cuda_stream : dev->own_stream;
  return ECL_OK;
}

extern "C" int ecl_set_tuning(ecl_dev *dev, uint32_t groups_per_thread, uint32_t hit_capacity) {
  if (!dev) return ECL_E_ARG;
  if (dev->pending) return fail(dev, ECL_E_STATE, "ecl_set_tuning with work pending");
  CK(cudaSetDevice(dev->ordinal));
  if (groups_per_thread > 4096) return fail(dev, ECL_E_ARG, "groups_per_thread %u > 4096", groups_per_thread);
  dev->groups_per_thread = groups_per_thread ? groups_per_thread : DEFAULT_GROUPS_PER_THREAD;
  const u32 cap = hit_capacity ? hit_capacity : DEFAULT_HIT_CAP;
  if (cap < MAX_HITS_PER_KEY * GROUP_KEYS) return fail(dev, ECL_E_ARG, "hit_capacity %u < %u", cap, MAX_HITS_PER_KEY * GROUP_KEYS);
  if (cap != dev->hit_cap) {
    CK(cudaFree(dev->d_hits));
    dev->d_hits = nullptr;
    CK(cudaMalloc(&dev->d_hits, (size_t)cap * sizeof(ecl_hit)));
    dev->hit_cap = cap;
    for (auto &s : dev->mslot) {  // the mul slots follow on their next use
      cudaFree(s.d_hits);
      s.d_hits = nullptr;
    }
  }
  return ECL_OK;
}

// ---------------------------------------------------------------- filters

static int filter_set_size(ecl_dev *dev, u64 size_words) {
  CK(cudaFree(dev->bloom_bits));
  dev->bloom_bits = nullptr, dev->bloom_size = 0;
  const size_t padded = (size_t)((size_words + 1) / 2 * 2);  // 16-byte multiple for the bulk copy
  CK(cudaMalloc(&dev->bloom_bits, padded * 8));
  CK(cudaMemsetAsync(dev->bloom_bits, 0, padded * 8, dev->stream));
  dev->bloom_size = size_words;
  dev->bloom_magic = ~0ULL / size_words;
  dev->bloom_fill = 0.5;
  return ECL_OK;
}

static int filter_measure(ecl_dev *dev) {
  dev->bloom_fill = 0.5;
  const bool in_hbm = dev->bloom_size * 8 > SMEM_FILTER_MAX;
  if (in_hbm) {  // stays in HBM: measure its fill for the candidate-queue planner
    unsigned long long *d_total = nullptr, total = 0;
    CK(cudaMalloc(&d_total, sizeof total));
    CK(cudaMemsetAsync(d_total, 0, sizeof total, dev->stream));
    bloom_popcount_kernel<<<dev->sm_count * 8, 256, 0, dev->stream>>>(dev->bloom_bits, dev->bloom_size, d_total);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&total, d_total, sizeof total, cudaMemcpyDeviceToHost, dev->stream));
    CK(cudaStreamSynchronize(dev->stream));
    cudaFree(d_total);
    dev->bloom_fill = (double)total / ((double)dev->bloom_size * 64.0);
  } else {
    CK(cudaStreamSynchronize(dev->stream));
  }
  return ECL_OK;
}

extern "C" int ecl_set_filter(ecl_dev *dev, const uint64_t *bits, uint64_t size_words) {
  if (!dev || !bits || size_words == 0) return fail(dev, ECL_E_ARG, "ecl_set_filter: empty filter");
  if (dev->pending) return fail(dev, ECL_E_STATE, "ecl_set_filter with work pending");
  CK(cudaSetDevice(dev->ordinal));
  int rc = filter_set_size(dev, size_words);
  if (rc) return rc;
  CK(cudaMemcpyAsync(dev->bloom_bits, bits, (size_t)size_words * 8, cudaMemcpyHostToDevice, dev->stream));
  CK(cudaStreamSynchronize(dev->stream));  // `bits` may be pageable and is the caller's again after this call
  return filter_measure(dev);
}

extern "C" int ecl_filter_alloc(ecl_dev *dev, uint64_t size_words) {
  if (!dev || size_words == 0) return fail(dev, ECL_E_ARG, "ecl_filter_alloc: empty filter");
  if (dev->pending) return fail(dev, ECL_E_STATE, "ecl_filter_alloc with work pending");
  CK(cudaSetDevice(dev->ordinal));
  return filter_set_size(dev, size_words);
}

extern "C" int ecl_filter_write(ecl_dev *dev, uint64_t offset_words, const uint64_t *bits, uint64_t n_words) {
  if (!dev || !bits) return ECL_E_ARG;
  if (!dev->bloom_bits || offset_words + n_words > dev->bloom_size || offset_words + n_words < offset_words)
    return fail(dev, ECL_E_ARG, "ecl_filter_write: words [%llu, +%llu) outside the filter of %llu words", (unsigned long long)offset_words,
                (unsigned long long)n_words, (unsigned long long)dev->bloom_size);
  CK(cudaSetDevice(dev->ordinal));
  CK(cudaMemcpyAsync(dev->bloom_bits + offset_words, bits, (size_t)n_words * 8, cudaMemcpyHostToDevice, dev->stream));
  return ECL_OK;
}

extern "C" int ecl_filter_flush(ecl_dev *dev) {
  if (!dev) return ECL_E_ARG;
  CK(cudaSetDevice(dev->ordinal));
  CK(cudaStreamSynchronize(dev->stream));
  return ECL_OK;
}

extern "C" int ecl_filter_commit(ecl_dev *dev) {
  if (!dev || !dev->bloom_bits) return fail(dev, ECL_E_ARG, "ecl_filter_commit: no filter");
  CK(cudaSetDevice(dev->ordinal));
  return filter_measure(dev);
}

extern "C" int ecl_filter_read(ecl_dev *dev, uint64_t offset_words, uint64_t *bits, uint64_t n_words) {
  if (!dev || !bits) return ECL_E_ARG;
  if (!dev->bloom_bits || offset_words + n_words > dev->bloom_size || offset_words + n_words < offset_words)
    return fail(dev, ECL_E_ARG, "ecl_filter_read: words outside the filter");
  CK(cudaSetDevice(dev->ordinal));
  CK(cudaMemcpyAsync(bits, dev->bloom_bits + offset_words, (size_t)n_words * 8, cudaMemcpyDeviceToHost, dev->stream));
  CK(cudaStreamSynchronize(dev->stream));
  return ECL_OK;
}

extern "C" int ecl_filter_copy_peer(ecl_dev *dev, ecl_dev *src) {
  if (!dev || !src || !src->bloom_bits) return fail(dev, ECL_E_ARG, "ecl_filter_copy_peer: no source filter");
  if (dev->pending) return fail(dev, ECL_E_STATE, "ecl_filter_copy_peer with work pending");
  CK(cudaSetDevice(src->ordinal));
  CK(cudaStreamSynchronize(src->stream));
  CK(cudaSetDevice(dev->ordinal));
  int rc = filter_set_size(dev, src->bloom_size);
  if (rc) return rc;
  int can = 0;
  if (cudaDeviceCanAccessPeer(&can, dev->ordinal, src->ordinal) == cudaSuccess && can) {
    cudaError_t e = cudaDeviceEnablePeerAccess(src->ordinal, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
    cudaGetLastError();
  }
  // with peer access this is a direct NVLink copy; without, the runtime stages it through the host
  CK(cudaMemcpyPeerAsync(dev->bloom_bits, dev->ordinal, src->bloom_bits, src->ordinal, (size_t)src->bloom_size * 8, dev->stream));
  CK(cudaStreamSynchronize(dev->stream));
  dev->bloom_fill = src->bloom_fill;
  return ECL_OK;
}

extern "C" int ecl_filter_generate(ecl_dev *dev, uint64_t size_words, double fill, uint64_t seed) {
  if (!dev || size_words == 0 || !(fill >= 0.0 && fill <= 1.0)) return fail(dev, ECL_E_ARG, "ecl_filter_generate: bad arguments");
  if (dev->pending) return fail(dev, ECL_E_STATE, "ecl_filter_generate with work pending");
  CK(cudaSetDevice(dev->ordinal));
  int rc = filter_set_size(dev, size_words);
  if (rc) return rc;
  const u32 thr = (u32)(fill * 256.0 + 0.5);
  filter_generate_kernel<<<dev->sm_count * 8, 256, 0, dev->stream>>>(dev->bloom_bits, size_words, thr, seed);
  CK(cudaGetLastError());
  return filter_measure(dev);
}

extern "C" double ecl_filter_fill(const ecl_dev *dev) { return dev ? dev->bloom_fill : 0.0; }

extern "C" void *ecl_host_alloc(uint64_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
extern "C" void ecl_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

extern "C" int ecl_set_stride(ecl_dev *dev, const uint64_t stride_k[4]) {
  if (!dev || !stride_k) return ECL_E_ARG;
  if (dev->pending) return fail(dev, ECL_E_STATE, "ecl_set_stride with work pending");
  if ((stride_k[0] | stride_k[1] | stride_k[2] | stride_k[3]) == 0) return fail(dev, ECL_E_ARG, "stride is zero");
  if (memcmp(dev->stride, stride_k, 32) != 0) dev->table_valid = false;
  memcpy(dev->stride, stride_k, 32);
  return ECL_OK;
}

static BloomView bloom_view(const ecl_dev *dev) {
  BloomView b;
  b.bits = dev->bloom_bits, b.size = dev->bloom_size, b.magic = dev->bloom_magic;
  return b;
}

// blf_gen's insert loop (lib/utils.c:453-465) on the device, exact count included. The final bits do not depend on
// the order (a skipped hash would have added nothing), the count does: hash i counts iff at least one of its 20 bits
// is set neither in the filter as it was nor by a hash before i. Per call: (1) blf_has against the filter as it is,
// and for every hash that fails it the positions of its still-clear bits as sort keys (i rides as the value, pairs
// are generated in input order); (2) a stable radix sort by position; (3) the head of every run of equal positions
// is the FIRST hash to set that bit: it counts; (4) all bits are OR-ed in.
extern "C" int ecl_filter_add(ecl_dev *dev, const uint32_t (*h160)[5], uint32_t n, uint64_t *n_new) {
  if (!dev || !h160) return ECL_E_ARG;
  if (!dev->bloom_bits) return fail(dev, ECL_E_ARG, "ecl_filter_add: no filter (ecl_filter_alloc / ecl_set_filter)");
  if (dev->pending) return fail(dev, ECL_E_STATE, "ecl_filter_add with work pending");
  if (n_new) *n_new = 0;
  if (n == 0) return ECL_OK;
  if (n > (1u << 26)) return fail(dev, ECL_E_ARG, "ecl_filter_add: at most 2^26 hashes per call");
  CK(cudaSetDevice(dev->ordinal));
  unsigned long long count = 0;
  int rc = filter_add_device(dev->stream, bloom_view(dev), dev->bloom_bits, h160, n, &count);
  if (rc == -1) return fail(dev, ECL_E_CUDA, "ecl_filter_add: %s", cudaGetErrorString(cudaGetLastError()));
  if (n_new) *n_new = count;
  return ECL_OK;
}

// ---------------------------------------------------------------- add path

static int ensure_add_resources(ecl_dev *dev) {
  if (!dev->centres) CK(cudaMalloc(&dev->centres, (size_t)dev->Tmax * 16 * sizeof(u32)));
  // prefix products: (ADD_H + 1) entries of 32 B per thread, twice (add_kernel_sp ping-pongs between the current
  // group's prefixes and the next group's)
  if (!dev->scratch) CK(cudaMalloc(&dev->scratch, (size_t)dev->Tmax * (ADD_H + 1) * 32 * 2));
  if (!dev->table_valid) {  // ctx_precompute_gpoints (main.c:219-246) on the device
    SmulParams sp;
    memset(&sp, 0, sizeof sp);
    const u64 zero[4] = {0, 0, 0, 0};
    sp.k0 = to_fe(zero), sp.step = to_fe(dev->stride);
    sp.gtab = dev->gtab, sp.count = ADD_H + 1, sp.mode = 0, sp.out = (u32 *)dev->add_table;
    smul_kernel<<<(sp.count + 127) / 128, 128, 0, dev->stream>>>(sp);
    CK(cudaGetLastError());
    dev->table_valid = true;
    dev->launches++;
  }
  return ECL_OK;
}

// the add_kernel variants live in add_inst.cu, one translation unit each (parallel build)
#define DECL_ADD(v) cudaError_t ecl_add_launch_##v(const AddParams &p, unsigned grid, unsigned smem, cudaStream_t stream);
DECL_ADD(1) DECL_ADD(2) DECL_ADD(3) DECL_ADD(5) DECL_ADD(6) DECL_ADD(7)
DECL_ADD(hbm_1) DECL_ADD(hbm_2) DECL_ADD(hbm_3) DECL_ADD(hbm_5) DECL_ADD(hbm_6) DECL_ADD(hbm_7)
typedef cudaError_t (*add_launch_fn)(const AddParams &, unsigned, unsigned, cudaStream_t);
static add_launch_fn pick_add_kernel(u32 flags, bool hbm) {
  switch (flags & (ECL_A33 | ECL_A65 | ECL_ENDO)) {
  case 1: return hbm ? ecl_add_launch_hbm_1 : ecl_add_launch_1;
  case 2: return hbm ? ecl_add_launch_hbm_2 : ecl_add_launch_2;
  case 3: return hbm ? ecl_add_launch_hbm_3 : ecl_add_launch_3;
  case 5: return hbm ? ecl_add_launch_hbm_5 : ecl_add_launch_5;
  case 6: return hbm ? ecl_add_launch_hbm_6 : ecl_add_launch_6;
  case 7: return hbm ? ecl_add_launch_hbm_7 : ecl_add_launch_7;
  default: return nullptr;
  }
}

static double cand_per_key(const ecl_dev *dev) {  // expected stage-1 survivors per key, with head-room
  const u32 hashes_per_key = ((dev->p_flags & ECL_A33) ? 1u : 0u) + ((dev->p_flags & ECL_A65) ? 1u : 0u);
  const double hashes = (double)hashes_per_key * ((dev->p_flags & ECL_ENDO) ? 6.0 : 1.0);
  return std::max(1e-6, hashes * dev->bloom_fill * dev->bloom_fill * 1.5);
}

// The candidate queue of the asynchronous probe, sized for what is asked of it: `want_keys` keys per launch at the
// measured fill of the current filter and the pending flags (stage 1 passes fill^2 of all hashes), at most 2^29
// entries (16 GB) and a third of the free memory. It only ever grows.
static int ensure_cand_queue(ecl_dev *dev, u64 want_keys) {
  double need = (double)want_keys * cand_per_key(dev);
  u64 want = 1ull << 16;
  while (want < (1ull << 29) && (double)want < need) want <<= 1;
  if (const char *env = getenv("ECLOOP_B200_CAND_LOG2")) want = 1ull << std::min(31, std::max(8, atoi(env)));  // test hook
  if (!dev->cand_counts) {
    CK(cudaMalloc(&dev->cand_counts, ((size_t)dev->grid_max + 1) * sizeof(u32)));
    CK(cudaMemsetAsync(dev->cand_counts, 0, ((size_t)dev->grid_max + 1) * sizeof(u32), dev->stream));
  }
  if (dev->cand_entries && dev->cand_cap >= want) return ECL_OK;
  CK(cudaStreamSynchronize(dev->stream));
  CK(cudaFree(dev->cand_entries));
  dev->cand_entries = nullptr, dev->cand_cap = 0;
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  while (want > 256 && want * 32 > free_b / 3) want >>= 1;
  for (;; want >>= 1) {
    if (cudaMalloc(&dev->cand_entries, want * 32) == cudaSuccess) break;
    cudaGetLastError();
    dev->cand_entries = nullptr;
    if (want <= 256) return fail(dev, ECL_E_CUDA, "cannot allocate the candidate queue");
  }
  dev->cand_cap = want;
  return ECL_OK;
}

static cudaEvent_t next_event(ecl_dev *dev) {
  if (dev->ev_used == dev->ev_pool.size()) {
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return nullptr;
    dev->ev_pool.push_back(ev);
  }
  return dev->ev_pool[dev->ev_used++];
}

// Queue the launches covering keys [k_begin, k_end) of the pending span. max_keys_per_launch bounds one launch.
static int launch_add(ecl_dev *dev, u64 k_begin, u64 k_end, u64 max_keys_per_launch, bool drain_each, std::vector<ecl_hit> *drain_to) {
  const u32 smem_table = (ADD_H + 1) * 64;
  // the filter rides in shared memory when it fits beside the table; otherwise it stays in HBM and is probed
  // asynchronously (probe_pipe.cuh) unless a span is being redone after a candidate-queue overflow, or drained
  // launch by launch for a dense filter (then the inline probe is exact and simple)
  const u64 bloom_bytes = (dev->bloom_size + 1) / 2 * 16;
  const bool bloom_smem = dev->bloom_size * 8 <= SMEM_FILTER_MAX;
  static const bool no_pipe = getenv("ECLOOP_B200_INLINE_PROBE") != nullptr;  // measurement hook: probe HBM filters inline
  const bool hbm = !bloom_smem && !dev->force_inline && !drain_each && !no_pipe;
  add_launch_fn fn = pick_add_kernel(dev->p_flags, hbm);
  if (!fn) return fail(dev, ECL_E_ARG, "flags select no address type");
  const u32 smem = smem_table + (bloom_smem ? (u32)bloom_bytes : 0u) + (hbm ? ProbePipe<ADD_THREADS>::BYTES : 0u);
  max_keys_per_launch = std::max<u64>(max_keys_per_launch, GROUP_KEYS);
  if (hbm) {
    int rc = ensure_cand_queue(dev, std::min(k_end - k_begin, max_keys_per_launch));
    if (rc) return rc;
    // a launch is sized so that the expected candidates use 2/3 of the queue
    const u64 fit = std::max<u64>(GROUP_KEYS, (u64)((double)dev->cand_cap / cand_per_key(dev)));
    max_keys_per_launch = std::min(max_keys_per_launch, fit);
  }
  // equal launches (whole reference groups each) rather than full ones and a remainder
  if (const char *env = getenv("ECLOOP_B200_MAX_LAUNCH_KEYS"))  // test hook: force several launches for a small span
    max_keys_per_launch = std::max<u64>(GROUP_KEYS, strtoull(env, nullptr, 10) / GROUP_KEYS * GROUP_KEYS);
  const u64 span = k_end - k_begin;
  const u64 n_launches = (span + max_keys_per_launch - 1) / max_keys_per_launch;
  const u64 per_launch = ((span + n_launches - 1) / n_launches + GROUP_KEYS - 1) / GROUP_KEYS * GROUP_KEYS;

  u64 k = k_begin;
  while (k < k_end) {
    const u64 L = std::min<u64>(k_end - k, per_launch);
    u32 threads = dev->Tmax;
    if (const char *env = getenv("ECLOOP_B200_MAX_THREADS"))  // test hook: few threads make small spans use large half groups
      threads = std::min<u32>(dev->Tmax, std::max<u32>(1u, (u32)strtoul(env, nullptr, 10)));
    const launch_plan lp = plan_launch(L, threads, ADD_H, HR_MIN);
    // centres: (start + (k + Hr + t*c*2Hr) * stride) * G   (GStart, main.c:359-360)
    u64 k0[4], step[4], kstep[4];
    const u64 zero[4] = {0, 0, 0, 0};
    sc_muladd64(k0, dev->stride, k + lp.Hr, dev->p_start);
    sc_muladd64(step, dev->stride, (u64)lp.c * 2 * lp.Hr, zero);
    SmulParams sp;
    memset(&sp, 0, sizeof sp);
    sp.k0 = to_fe(k0), sp.step = to_fe(step), sp.gtab = dev->gtab, sp.count = lp.T, sp.mode = 1, sp.out = dev->centres;
    const uint4 *step_pt;
    if (lp.Hr == ADD_H) step_pt = dev->add_table + (size_t)ADD_H * 4;
    else if (2 * lp.Hr <= ADD_H) step_pt = dev->add_table + (size_t)(2 * lp.Hr - 1) * 4;
    else {  // 2*Hr*s*G is not in the table: one more thread of the centre kernel computes it
      sc_muladd64(kstep, dev->stride, 2 * (u64)lp.Hr, zero);
      sp.extra_k = to_fe(kstep), sp.extra_out = (u32 *)dev->step_buf;
      step_pt = dev->step_buf;
    }
    smul_kernel<<<(lp.T + 1 + 127) / 128, 128, 0, dev->stream>>>(sp);
    CK(cudaGetLastError());

    AddParams ap;
    memset(&ap, 0, sizeof ap);
    ap.cx = dev->centres, ap.cy = dev->centres + (size_t)8 * lp.T;
    ap.table = dev->add_table, ap.step_pt = step_pt, ap.scratch = dev->scratch;
    ap.bloom = bloom_view(dev);
    ap.bloom_smem_words = bloom_smem ? (u32)dev->bloom_size : 0u;
    ap.sink.hits = dev->d_hits, ap.sink.count = dev->d_hit_count, ap.sink.cap = dev->hit_cap;
    ap.T = lp.T, ap.groups_per_thread = lp.c, ap.Hr = lp.Hr, ap.n_keys = L, ap.key_off0 = k, ap.err = dev->d_err;
    const u32 grid = (lp.T + ADD_THREADS - 1) / ADD_THREADS;
    if (hbm) {
      ap.cand.entries = dev->cand_entries, ap.cand.counts = dev->cand_counts;
      ap.cand.overflow = dev->cand_counts + dev->grid_max;
      ap.cand.cap_per_cta = (u32)std::min<u64>(dev->cand_cap / grid, 0xffffffffu);
      CK(cudaMemsetAsync(dev->cand_counts, 0, (size_t)dev->grid_max * sizeof(u32), dev->stream));  // not the flag
    }
    cudaEvent_t e0 = next_event(dev), e1 = next_event(dev);
    if (!e0 || !e1) return fail(dev, ECL_E_CUDA, "cudaEventCreate failed");
    CK(cudaEventRecord(e0, dev->stream));
    CK(fn(ap, grid, smem, dev->stream));
    if (hbm) {  // stage 2: the full test on what stage 1 queued
      cand_verify_kernel<<<dim3(32, grid), 256, 0, dev->stream>>>(ap.cand, ap.bloom, ap.sink);
      CK(cudaGetLastError());
      dev->launches++;
    }
    CK(cudaEventRecord(e1, dev->stream));
    dev->launches += 2;
    k += L;

    if (drain_each) {
      CK(cudaStreamSynchronize(dev->stream));
      u32 cnt = 0;
      CK(cudaMemcpyAsync(&cnt, dev->d_hit_count, sizeof cnt, cudaMemcpyDeviceToHost, dev->stream));
      CK(cudaStreamSynchronize(dev->stream));
      if (cnt > dev->hit_cap) return fail(dev, ECL_E_OVERFLOW, "hit buffer overflow in exact mode (%u > %u)", cnt, dev->hit_cap);
      const size_t old = drain_to->size();
      drain_to->resize(old + cnt);
      if (cnt) {
        CK(cudaMemcpyAsync(drain_to->data() + old, dev->d_hits, (size_t)cnt * sizeof(ecl_hit), cudaMemcpyDeviceToHost, dev->stream));
        CK(cudaStreamSynchronize(dev->stream));
      }
      CK(cudaMemsetAsync(dev->d_hit_count, 0, sizeof(u32), dev->stream));
    }
  }
  return ECL_OK;
}

extern "C" int ecl_add_submit(ecl_dev *dev, const uint64_t start_pk[4], uint64_t n_keys, uint32_t flags) {
  if (!dev || !start_pk) return ECL_E_ARG;
  if (dev->pending) return fail(dev, ECL_E_STATE, "a submit is already pending; call ecl_collect first");
  if (!dev->bloom_bits) return fail(dev, ECL_E_ARG, "no filter set (ecl_set_filter)");
  if (n_keys == 0 || n_keys % ECL_GROUP) return fail(dev, ECL_E_ARG, "n_keys %llu is not a positive multiple of %u", (unsigned long long)n_keys, ECL_GROUP);
  if (!(flags & (ECL_A33 | ECL_A65))) return fail(dev, ECL_E_ARG, "flags select no address type");
  CKA(cudaSetDevice(dev->ordinal));
  dev->ev_used = 0, dev->launches = 0;
  dev->result.clear(), dev->result_ready = false;
  CKA(cudaEventRecord(dev->ev_begin, dev->stream));
  int rc = ensure_add_resources(dev);
  if (rc) return abandon(dev, rc);
  memcpy(dev->p_start, start_pk, 32);
  dev->p_keys = n_keys, dev->p_flags = flags;
  CKA(cudaMemsetAsync(dev->d_hit_count, 0, 2 * sizeof(u32), dev->stream));  // hit cursor + degenerate-group flag
  dev->force_inline = false;
  if (dev->cand_counts) CKA(cudaMemsetAsync(dev->cand_counts + dev->grid_max, 0, sizeof(u32), dev->stream));
  rc = launch_add(dev, 0, n_keys, (u64)dev->Tmax * dev->groups_per_thread * GROUP_KEYS, false, nullptr);
  if (rc) return abandon(dev, rc);
  CKA(cudaEventRecord(dev->ev_end, dev->stream));
  dev->pending = 1;
  return ECL_OK;
}

// ---------------------------------------------------------------- mul path

typedef void (*mul_hash_fn)(const MulHashParams);
#ifndef ECL_MUL_NW
#define ECL_MUL_NW 1  // keys hashed side by side per thread in K2b: 1 is 1.5 % faster than 2 (profiles/r02_b_mul_variants.txt)
#endif
static mul_hash_fn pick_mul_hash(u32 flags) {
  const bool c = flags & ECL_A33, u = flags & ECL_A65;
  if (c && u) return mul_hash_kernel<true, true, ECL_MUL_NW>;
  if (c) return mul_hash_kernel<true, false, ECL_MUL_NW>;
  if (u) return mul_hash_kernel<false, true, ECL_MUL_NW>;
  return nullptr;
}

static int ensure_mul_slot(ecl_dev *dev, mul_slot &s, u32 n) {
  if (!s.ev_begin) {
    CK(cudaEventCreate(&s.ev_begin));
    CK(cudaEventCreate(&s.ev_k0));
    CK(cudaEventCreate(&s.ev_k1));
    CK(cudaEventCreate(&s.ev_end));
    CK(cudaMalloc(&s.d_hit_count, sizeof(u32)));
    CK(cudaMallocHost(&s.h_count, sizeof(u32)));
  }
  if (!s.d_hits) CK(cudaMalloc(&s.d_hits, (size_t)dev->hit_cap * sizeof(ecl_hit)));
  if (n > s.cap) {
    const u32 cap = std::max<u32>(n, 1u << 16);
    cudaFreeHost(s.h_keys), cudaFree(s.d_keys), cudaFree(s.d_scratch);
    s.h_keys = nullptr, s.d_keys = nullptr, s.d_scratch = nullptr, s.cap = 0;
    CK(cudaMallocHost(&s.h_keys, (size_t)cap * sizeof(fe)));
    CK(cudaMalloc(&s.d_keys, (size_t)cap * sizeof(fe)));
    CK(cudaMalloc(&s.d_scratch, ((size_t)cap + 64) * 128));
    s.cap = cap;
  }
  return ECL_OK;
}

// geometry of K2a for n keys: thread t owns keys m*T + t, m < B
static void mul_geometry(const ecl_dev *dev, u32 n, u32 *T, u32 *B) {
  // keys per thread: as many as it takes to fill the threads that are resident at once (K2a: 2 CTAs of 256 per SM),
  // so that small batches still spread over the whole GPU and large ones amortise the per-thread inversion
  // (270 multiplications) and run as one wave
  const u32 want_threads = (u32)dev->sm_count * 512u;
  u32 b = std::min<u32>(64u, (n + want_threads - 1) / want_threads);
  if (b == 0) b = 1;
  *B = b, *T = (n + b - 1) / b;
}

extern "C" int ecl_mul_reserve(ecl_dev *dev, uint32_t n) {
  if (!dev || n == 0) return ECL_E_ARG;
  if (dev->pending) return fail(dev, ECL_E_STATE, "ecl_mul_reserve with work pending");
  CK(cudaSetDevice(dev->ordinal));
  for (auto &s : dev->mslot) {
    int rc = ensure_mul_slot(dev, s, n);
    if (rc) return rc;
  }
  return ECL_OK;
}

// K2b over keys [begin, end) of a slot whose points are already affine
static int launch_mul_hash(ecl_dev *dev, mul_slot &s, u32 begin, u32 end) {
  u32 T, B;
  mul_geometry(dev, s.n, &T, &B);
  MulHashParams hp;
  memset(&hp, 0, sizeof hp);
  hp.scratch = s.d_scratch, hp.bloom = bloom_view(dev);
  const bool bloom_smem = dev->bloom_size * 8 <= SMEM_FILTER_MAX;
  hp.bloom_smem_words = bloom_smem ? (u32)dev->bloom_size : 0u;
  hp.sink.hits = s.d_hits, hp.sink.count = s.d_hit_count, hp.sink.cap = dev->hit_cap;
  hp.T = T, hp.begin = begin, hp.end = end;
  const u32 per_cta = 512u * ECL_MUL_NW;
  const u32 grid = std::max<u32>(1u, std::min<u32>((u32)dev->sm_count, (end - begin + per_cta - 1) / per_cta));
  const u32 smem = bloom_smem ? (u32)((dev->bloom_size + 1) / 2 * 16) : 0u;
  mul_hash_fn fn = pick_mul_hash(s.flags);
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fn<<<grid, 512, smem, dev->stream>>>(hp);
  CK(cudaGetLastError());
  dev->launches++;
  return ECL_OK;
}

extern "C" int ecl_mul_submit(ecl_dev *dev, const uint64_t (*pks)[4], uint32_t n, uint32_t flags) {
  if (!dev || !pks || n == 0) return fail(dev, ECL_E_ARG, "ecl_mul_submit: no keys");
  if (dev->pending == 1) return fail(dev, ECL_E_STATE, "an add submit is pending; call ecl_collect first");
  if (dev->m_count == ECL_MUL_DEPTH) return fail(dev, ECL_E_STATE, "%d mul submits are already pending; call ecl_collect first", ECL_MUL_DEPTH);
  if (!dev->bloom_bits) return fail(dev, ECL_E_ARG, "no filter set (ecl_set_filter)");
  if (!pick_mul_hash(flags)) return fail(dev, ECL_E_ARG, "flags select no address type");
  CKA(cudaSetDevice(dev->ordinal));
  mul_slot &s = dev->mslot[(dev->m_head + dev->m_count) % ECL_MUL_DEPTH];
  int rc = ensure_mul_slot(dev, s, n);
  if (rc) return abandon(dev, rc);
  // the (hi, lo) limb image of uint64_t[4] equals fe's 8 x u32 on a little-endian host
  memcpy(s.h_keys, pks, (size_t)n * 32);
  s.n = n, s.flags = flags;
  CKA(cudaEventRecord(s.ev_begin, dev->stream));
  CKA(cudaMemcpyAsync(s.d_keys, s.h_keys, (size_t)n * 32, cudaMemcpyHostToDevice, dev->stream));
  CKA(cudaMemsetAsync(s.d_hit_count, 0, sizeof(u32), dev->stream));
  MulParams mp;
  memset(&mp, 0, sizeof mp);
  mp.scalars = s.d_keys, mp.gtab = dev->gtab, mp.scratch = s.d_scratch, mp.count = n;
  mul_geometry(dev, n, &mp.T, &mp.B);
  dev->launches = 0;
  CKA(cudaEventRecord(s.ev_k0, dev->stream));
  mul_points_kernel<<<(mp.T + 255) / 256, 256, 0, dev->stream>>>(mp);
  CKA(cudaGetLastError());
  dev->launches++;
  rc = launch_mul_hash(dev, s, 0, n);
  if (rc) return abandon(dev, rc);
  CKA(cudaEventRecord(s.ev_k1, dev->stream));
  CKA(cudaMemcpyAsync(s.h_count, s.d_hit_count, sizeof(u32), cudaMemcpyDeviceToHost, dev->stream));
  CKA(cudaEventRecord(s.ev_end, dev->stream));
  s.busy = true;
  dev->m_count++;
  dev->pending = 2;
  dev->result.clear(), dev->result_ready = false;
  return ECL_OK;
}

// ---------------------------------------------------------------- collect

static bool hit_less_add(const ecl_hit &a, const ecl_hit &b) {  // reference -t 1 emission order
  const u64 ga = a.key_off / ECL_GROUP, gb = b.key_off / ECL_GROUP;
  if (ga != gb) return ga < gb;
  const int ea = a.endo != 0, eb = b.endo != 0;
  if (ea != eb) return ea < eb;
  if (a.key_off != b.key_off) return a.key_off < b.key_off;
  if (a.endo != b.endo) return a.endo < b.endo;
  return a.kind < b.kind;
}
static bool hit_less_mul(const ecl_hit &a, const ecl_hit &b) {
  if (a.key_off != b.key_off) return a.key_off < b.key_off;
  return a.kind < b.kind;
}

// blocking copy of a device word / buffer on the device's stream (the stream is non-blocking: the legacy default
// stream does not order with it)
static cudaError_t fetch(ecl_dev *dev, void *dst, const void *src, size_t bytes) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, dev->stream);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(dev->stream);
}

static int collect_add(ecl_dev *dev) {
  CKA(cudaStreamSynchronize(dev->stream));
  float ms = 0;
  CKA(cudaEventElapsedTime(&ms, dev->ev_begin, dev->ev_end));
  dev->last_total_ms = ms;
  dev->last_hot_ms = 0;
  for (size_t i = 0; i + 1 < dev->ev_used; i += 2) {
    CKA(cudaEventElapsedTime(&ms, dev->ev_pool[i], dev->ev_pool[i + 1]));
    dev->last_hot_ms += ms;
  }
  dev->last_launches = dev->launches;
  u32 head[2] = {0, 0};  // hit cursor, degenerate flag
  CKA(fetch(dev, head, dev->d_hit_count, sizeof head));
  if (head[1])
    return abandon(dev, fail(dev, ECL_E_DEGENERATE, "the span reaches key 0 or n (a group centre equals +-m*stride*G): no inverse there; "
                                                     "the reference asserts at this point (lib/ecc.c:666)"));
  if (dev->cand_counts && !dev->force_inline) {
    // the asynchronous probe's candidate queue overflowed (a filter far denser than a bloom filter should be):
    // nothing may be lost, so the span is redone with the probes inline
    u32 ovf = 0;
    CKA(fetch(dev, &ovf, dev->cand_counts + dev->grid_max, sizeof ovf));
    if (ovf) {
      dev->force_inline = true;
      CKA(cudaMemsetAsync(dev->d_hit_count, 0, sizeof(u32), dev->stream));
      CKA(cudaMemsetAsync(dev->cand_counts + dev->grid_max, 0, sizeof(u32), dev->stream));
      int rc2 = launch_add(dev, 0, dev->p_keys, (u64)dev->Tmax * dev->groups_per_thread * GROUP_KEYS, false, nullptr);
      if (rc2) return abandon(dev, rc2);
      CKA(fetch(dev, head, dev->d_hit_count, sizeof head));
    }
  }
  const u32 cnt = head[0];
  dev->result.clear();
  if (cnt <= dev->hit_cap) {
    dev->result.resize(cnt);
    if (cnt) CKA(fetch(dev, dev->result.data(), dev->d_hits, (size_t)cnt * sizeof(ecl_hit)));
  } else {
    // Dense filter (e.g. the all-ones dump filter): redo the span in slices whose worst case fits the ring.
    CKA(cudaMemsetAsync(dev->d_hit_count, 0, sizeof(u32), dev->stream));
    const u64 slice = std::max<u64>(1, dev->hit_cap / (MAX_HITS_PER_KEY * GROUP_KEYS)) * GROUP_KEYS;
    int rc = launch_add(dev, 0, dev->p_keys, slice, true, &dev->result);
    if (rc) return abandon(dev, rc);
  }
  std::sort(dev->result.begin(), dev->result.end(), hit_less_add);
  return ECL_OK;
}

static int collect_mul(ecl_dev *dev, mul_slot &s) {
  CKA(cudaEventSynchronize(s.ev_end));
  float ms = 0;
  CKA(cudaEventElapsedTime(&ms, s.ev_begin, s.ev_end));
  dev->last_total_ms = ms;
  CKA(cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1));
  dev->last_hot_ms = ms;
  dev->last_launches = 2;
  const u32 cnt = *s.h_count;
  dev->result.clear();
  if (cnt <= dev->hit_cap) {
    dev->result.resize(cnt);
    if (cnt) {  // on the io stream: a younger submit may be computing on dev->stream, this batch is complete (ev_end)
      CKA(cudaMemcpyAsync(dev->result.data(), s.d_hits, (size_t)cnt * sizeof(ecl_hit), cudaMemcpyDeviceToHost, dev->io_stream));
      CKA(cudaStreamSynchronize(dev->io_stream));
    }
  } else {
    // dense filter: the points are still in the slot's scratch, only the hashing is redone, in slices that fit the ring
    const u32 slice = dev->hit_cap / 2;
    for (u32 b = 0; b < s.n; b += slice) {
      CKA(cudaMemsetAsync(s.d_hit_count, 0, sizeof(u32), dev->stream));
      int rc = launch_mul_hash(dev, s, b, std::min<u32>(s.n, b + slice));
      if (rc) return abandon(dev, rc);
      u32 c2 = 0;
      CKA(fetch(dev, &c2, s.d_hit_count, sizeof c2));
      const size_t old = dev->result.size();
      dev->result.resize(old + c2);
      if (c2) CKA(fetch(dev, dev->result.data() + old, s.d_hits, (size_t)c2 * sizeof(ecl_hit)));
    }
  }
  std::sort(dev->result.begin(), dev->result.end(), hit_less_mul);
  return ECL_OK;
}

extern "C" int ecl_collect(ecl_dev *dev, ecl_hit *hits, uint32_t cap, uint32_t *n_hits, uint64_t *keys_done) {
  if (!dev) return ECL_E_ARG;
  if (!dev->pending) return fail(dev, ECL_E_STATE, "ecl_collect without a pending submit");
  CKA(cudaSetDevice(dev->ordinal));
  if (!dev->result_ready) {
    int rc = dev->pending == 1 ? collect_add(dev) : collect_mul(dev, dev->mslot[dev->m_head]);
    if (rc) return rc;
    dev->result_ready = true;
  }
  const u64 done = dev->pending == 1 ? dev->p_keys : dev->mslot[dev->m_head].n;
  if (n_hits) *n_hits = (u32)std::min<size_t>(dev->result.size(), cap);
  if (keys_done) *keys_done = done;
  if (dev->result.size() > cap)  // the work is kept: the caller comes back with a larger buffer
    return fail(dev, ECL_E_OVERFLOW, "%zu hits do not fit the caller's buffer of %u; call ecl_collect again with a larger one", dev->result.size(), cap);
  if (hits && !dev->result.empty()) memcpy(hits, dev->result.data(), dev->result.size() * sizeof(ecl_hit));
  dev->result_ready = false;
  if (dev->pending == 2) {
    dev->mslot[dev->m_head].busy = false;
    dev->m_head = (dev->m_head + 1) % ECL_MUL_DEPTH;
    if (--dev->m_count == 0) dev->pending = 0;
  } else {
    dev->pending = 0;
  }
  return ECL_OK;
}

extern "C" int ecl_last_elapsed_ms(ecl_dev *dev, float *total_ms, float *hot_kernel_ms, uint32_t *kernel_launches) {
  if (!dev) return ECL_E_ARG;
  if (total_ms) *total_ms = dev->last_total_ms;
  if (hot_kernel_ms) *hot_kernel_ms = dev->last_hot_ms;
  if (kernel_launches) *kernel_launches = dev->last_launches;
  return ECL_OK;
}

// ---------------------------------------------------------------- primitives (parity entry points)
// Uploads, kernel and downloads all go through dev->stream (created non-blocking: it does not order with the legacy
// default stream, and a pageable cudaMemcpy may return before its DMA has landed).

template <typename T>
struct DevBuf {
  T *p = nullptr;
  ~DevBuf() { cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, n * sizeof(T)); }
};
#define H2D(dst, src, bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, dev->stream))
#define D2H(dst, src, bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, dev->stream))

extern "C" int ecl_prim_fp(ecl_dev *dev, int op, const uint64_t (*a)[4], const uint64_t (*b)[4], uint64_t (*out)[4], uint32_t n) {
  if (!dev || !a || !out || n == 0) return fail(dev, ECL_E_ARG, "ecl_prim_fp: bad arguments");
  CK(cudaSetDevice(dev->ordinal));
  DevBuf<fe> da, db, dout;
  CK(da.alloc(n));
  CK(dout.alloc(n));
  H2D(da.p, a, (size_t)n * 32);
  if (b) {
    CK(db.alloc(n));
    H2D(db.p, b, (size_t)n * 32);
  }
  prim_fp_kernel<<<(n + 127) / 128, 128, 0, dev->stream>>>(op, da.p, db.p, dout.p, n);
  CK(cudaGetLastError());
  D2H(out, dout.p, (size_t)n * 32);
  CK(cudaStreamSynchronize(dev->stream));
  return ECL_OK;
}

extern "C" int ecl_prim_scalar_mul(ecl_dev *dev, const uint64_t (*k)[4], uint64_t (*out_xy)[8], uint32_t n) {
  if (!dev || !k || !out_xy || n == 0) return fail(dev, ECL_E_ARG, "ecl_prim_scalar_mul: bad arguments");
  CK(cudaSetDevice(dev->ordinal));
  DevBuf<fe> dk;
  DevBuf<u32> dout;
  CK(dk.alloc(n));
  CK(dout.alloc((size_t)n * 16));
  H2D(dk.p, k, (size_t)n * 32);
  SmulParams sp;
  memset(&sp, 0, sizeof sp);
  sp.scalars = dk.p, sp.gtab = dev->gtab, sp.count = n, sp.mode = 2, sp.out = dout.p;
  smul_kernel<<<(n + 127) / 128, 128, 0, dev->stream>>>(sp);
  CK(cudaGetLastError());
  D2H(out_xy, dout.p, (size_t)n * 64);
  CK(cudaStreamSynchronize(dev->stream));
  return ECL_OK;
}

extern "C" int ecl_prim_hash160(ecl_dev *dev, const uint64_t (*xy)[8], uint32_t (*out33)[5], uint32_t (*out65)[5], uint32_t n) {
  if (!dev || !xy || n == 0) return fail(dev, ECL_E_ARG, "ecl_prim_hash160: bad arguments");
  CK(cudaSetDevice(dev->ordinal));
  DevBuf<u32> dxy, d33, d65;
  CK(dxy.alloc((size_t)n * 16));
  H2D(dxy.p, xy, (size_t)n * 64);
  if (out33) CK(d33.alloc((size_t)n * 5));
  if (out65) CK(d65.alloc((size_t)n * 5));
  prim_hash160_kernel<<<(n + 127) / 128, 128, 0, dev->stream>>>(dxy.p, d33.p, d65.p, n);
  CK(cudaGetLastError());
  if (out33) D2H(out33, d33.p, (size_t)n * 20);
  if (out65) D2H(out65, d65.p, (size_t)n * 20);
  CK(cudaStreamSynchronize(dev->stream));
  return ECL_OK;
}

extern "C" int ecl_prim_bloom(ecl_dev *dev, const uint32_t (*h160)[5], uint8_t *out, uint32_t n) {
  if (!dev || !h160 || !out || n == 0) return fail(dev, ECL_E_ARG, "ecl_prim_bloom: bad arguments");
  if (!dev->bloom_bits) return fail(dev, ECL_E_ARG, "no filter set (ecl_set_filter)");
  CK(cudaSetDevice(dev->ordinal));
  DevBuf<u32> dh;
  DevBuf<uint8_t> dout;
  CK(dh.alloc((size_t)n * 5));
  CK(dout.alloc(n));
  H2D(dh.p, h160, (size_t)n * 20);
  prim_bloom_kernel<<<(n + 127) / 128, 128, 0, dev->stream>>>(bloom_view(dev), dh.p, dout.p, n);
  CK(cudaGetLastError());
  D2H(out, dout.p, n);
  CK(cudaStreamSynchronize(dev->stream));
  return ECL_OK;
}

// ---------------------------------------------------------------- integer-pipe peaks

template <int KIND>
static int run_peak(ecl_dev *dev, double *gops, double *mhz) {
  DevBuf<u32> out;
  DevBuf<unsigned long long> cyc;
  CK(out.alloc(1024));
  CK(cyc.alloc(2));
  const int blocks = dev->sm_count * 8;  // 8 x 256 threads = 2048 threads per SM: full occupancy
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  unsigned long long cycles[2] = {0, 1};
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0, dev->stream));
    peak_kernel<KIND><<<blocks, 256, 0, dev->stream>>>(out.p, 0x1234567u + rep, cyc.p);
    CK(cudaEventRecord(e1, dev->stream));
    CK(cudaStreamSynchronize(dev->stream));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) {
      best = ms;
      CK(cudaMemcpyAsync(cycles, cyc.p, sizeof cycles, cudaMemcpyDeviceToHost, dev->stream));
      CK(cudaStreamSynchronize(dev->stream));
    }
  }
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  const double insts = (double)blocks * 256.0 * PEAK_ITERS * PEAK_UNROLL * PEAK_CHAINS;
  *gops = insts / (best * 1e-3) / 1e9;
  // SM clock: cycle counter over the nanosecond timer, both read by the same thread around its loop (the ratio of one
  // block's cycles to the kernel's wall time is only the clock when all blocks run as one wave)
  *mhz = cycles[1] ? (double)cycles[0] / (double)cycles[1] * 1e3 : 0.0;
  return ECL_OK;
}

#ifdef ECL_EXPERIMENTAL
template <int KIND, int FILL>
static int run_mulbench(ecl_dev *dev, double *gmuls) {
  DevBuf<u32> out;
  CK(out.alloc(1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0, dev->stream));
    mulbench_kernel<KIND, FILL><<<dev->sm_count, 512, 0, dev->stream>>>(out.p, 0x9e3779b9u + rep);
    CK(cudaEventRecord(e1, dev->stream));
    CK(cudaStreamSynchronize(dev->stream));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  *gmuls = (double)dev->sm_count * 512.0 * MULBENCH_ITERS * 2.0 / (best * 1e-3) / 1e9;
  return ECL_OK;
}
#endif

// one kind of peak.cuh by number (0..18); see PEAK_KINDS in ecloop_b200/__init__.py for the names
extern "C" int ecl_peak_bench_kind(ecl_dev *dev, int kind, double *gops, double *sm_mhz) {
  if (!dev || !gops) return ECL_E_ARG;
  CK(cudaSetDevice(dev->ordinal));
  double mhz = 0;
  int rc;
  switch (kind) {
  case 0: rc = run_peak<0>(dev, gops, &mhz); break;
  case 1: rc = run_peak<1>(dev, gops, &mhz); break;
  case 2: rc = run_peak<2>(dev, gops, &mhz); break;
  case 3: rc = run_peak<3>(dev, gops, &mhz); break;
  case 4: rc = run_peak<4>(dev, gops, &mhz); break;
  case 5: rc = run_peak<5>(dev, gops, &mhz); break;
  case 6: rc = run_peak<6>(dev, gops, &mhz); break;
  case 7: rc = run_peak<7>(dev, gops, &mhz); break;
  case 8: rc = run_peak<8>(dev, gops, &mhz); break;
  case 9: rc = run_peak<9>(dev, gops, &mhz); break;
  case 10: rc = run_peak<10>(dev, gops, &mhz); break;
  case 11: rc = run_peak<11>(dev, gops, &mhz); break;
  case 12: rc = run_peak<12>(dev, gops, &mhz); break;
  case 13: rc = run_peak<13>(dev, gops, &mhz); break;
  case 14: rc = run_peak<14>(dev, gops, &mhz); break;
  case 15: rc = run_peak<15>(dev, gops, &mhz); break;
  case 16: rc = run_peak<16>(dev, gops, &mhz); break;
  case 17: rc = run_peak<17>(dev, gops, &mhz); break;
  case 18: rc = run_peak<18>(dev, gops, &mhz); break;
#ifdef ECL_EXPERIMENTAL
  // field multiplications per second (G mul/s) of fe_mul (IMAD.WIDE) / fe6_mul (DFMA), alone (19, 20) and with
  // 384 LOP3/SHF per multiplication beside them (21, 22), 512 threads per SM like the add kernel
  case 19: rc = run_mulbench<0, 0>(dev, gops); break;
  case 20: rc = run_mulbench<1, 0>(dev, gops); break;
  case 21: rc = run_mulbench<0, 384>(dev, gops); break;
  case 22: rc = run_mulbench<1, 384>(dev, gops); break;
#endif
  default: return fail(dev, ECL_E_ARG, "unknown peak kind %d", kind);
  }
  if (sm_mhz) *sm_mhz = mhz;
  return rc;
}

extern "C" int ecl_peak_bench(ecl_dev *dev, double out[8]) {
  if (!dev || !out) return ECL_E_ARG;
  CK(cudaSetDevice(dev->ordinal));
  double mhz = 0, m2 = 0;
  int rc;
  for (int i = 0; i < 8; ++i) out[i] = 0;
  if ((rc = run_peak<0>(dev, &out[0], &mhz))) return rc;
  if ((rc = run_peak<1>(dev, &out[1], &m2))) return rc;
  if ((rc = run_peak<2>(dev, &out[2], &m2))) return rc;
  if ((rc = run_peak<3>(dev, &out[3], &m2))) return rc;
  if ((rc = run_peak<4>(dev, &out[4], &m2))) return rc;
  if ((rc = run_peak<5>(dev, &out[5], &m2))) return rc;
  out[6] = mhz;
  return ECL_OK;
}
