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

#include "kernels.cuh"
#include "peak.cuh"

static_assert(ECL_GROUP % (2 * ADD_H) == 0, "a reference group must be a whole number of device groups");

#define GROUP_KEYS (2u * ADD_H)
#define DEFAULT_GROUPS_PER_THREAD 8u
#define DEFAULT_HIT_CAP (1u << 20)
#define MAX_HITS_PER_KEY 12u  // {33,65} x 6 endomorphism images

typedef unsigned __int128 u128;

// ---------------------------------------------------------------- device object

struct ecl_dev {
  int ordinal = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  char err[512] = {0};

  uint4 *gtab = nullptr;  // window table, GTAB_ENTRIES x 64 B
  u32 *bases = nullptr;

  uint4 *add_table = nullptr;  // (ADD_H+1) x 64 B for the current stride
  bool table_valid = false;
  u64 stride[4] = {1, 0, 0, 0};

  u64 *bloom_bits = nullptr;
  u64 bloom_size = 0, bloom_magic = 0;
  double bloom_fill = 0.5;  // fraction of set bits (measured by ecl_set_filter for filters that stay in HBM)

  u32 Tmax = 0;
  u32 *centres = nullptr;  // 16 x Tmax u32 (SoA x then y)
  uint4 *scratch = nullptr;

  ecl_hit *d_hits = nullptr;
  u32 *d_hit_count = nullptr;
  u32 hit_cap = DEFAULT_HIT_CAP;
  u32 groups_per_thread = DEFAULT_GROUPS_PER_THREAD;

  fe *d_scalars = nullptr;
  u32 scalars_cap = 0;
  // asynchronous probing of filters that do not fit shared memory (probe_pipe.cuh)
  uint4 *cand_entries = nullptr;
  u32 *cand_counts = nullptr;  // [sm_count] + 1 overflow flag
  u64 cand_cap = 0;            // entries in total
  bool force_inline = false;   // the last span overflowed the queue and is being redone with inline probes

  uint4 *mul_scratch = nullptr;  // 128 B per key of a mul batch (X, Y, Z, prefix product)
  u32 mul_scratch_cap = 0;

  // pending work (one submit at a time)
  int pending = 0;  // 0 none, 1 add, 2 mul
  u64 p_start[4] = {0, 0, 0, 0};
  u64 p_keys = 0;
  u32 p_flags = 0;
  std::vector<fe> p_scalars;
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
  CK(cudaMalloc(&dev->bases, GTAB_WINDOWS * 64));
  gtab_bases_kernel<<<1, 32, 0, dev->stream>>>(dev->bases);
  gtab_fill_kernel<<<(GTAB_ENTRIES + 127) / 128, 128, 0, dev->stream>>>((u32 *)dev->gtab, dev->bases);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(dev->stream));
  return ECL_OK;
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
    dev->Tmax = (u32)dev->sm_count * ADD_THREADS * ADD_MIN_BLOCKS;
    // random 8-byte filter probes should cost one 32 B sector of DRAM traffic, not a 128 B line
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    CK(cudaStreamCreateWithFlags(&dev->own_stream, cudaStreamNonBlocking));
    dev->stream = dev->own_stream;
    CK(cudaEventCreate(&dev->ev_begin));
    CK(cudaEventCreate(&dev->ev_end));
    CK(cudaMalloc(&dev->d_hit_count, sizeof(u32)));
    CK(cudaMalloc(&dev->d_hits, (size_t)dev->hit_cap * sizeof(ecl_hit)));
    CK(cudaMalloc(&dev->add_table, (size_t)(ADD_H + 1) * 64));
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
  cudaFree(dev->gtab), cudaFree(dev->bases), cudaFree(dev->add_table), cudaFree(dev->bloom_bits);
  cudaFree(dev->centres), cudaFree(dev->scratch), cudaFree(dev->d_hits), cudaFree(dev->d_hit_count);
  cudaFree(dev->d_scalars), cudaFree(dev->mul_scratch);
  cudaFree(dev->cand_entries), cudaFree(dev->cand_counts);
  for (auto ev : dev->ev_pool) cudaEventDestroy(ev);
  if (dev->ev_begin) cudaEventDestroy(dev->ev_begin);
  if (dev->ev_end) cudaEventDestroy(dev->ev_end);
  if (dev->own_stream) cudaStreamDestroy(dev->own_stream);
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
  }
  return ECL_OK;
}

extern "C" int ecl_set_filter(ecl_dev *dev, const uint64_t *bits, uint64_t size_words) {
  if (!dev || !bits || size_words == 0) return fail(dev, ECL_E_ARG, "ecl_set_filter: empty filter");
  if (dev->pending) return fail(dev, ECL_E_STATE, "ecl_set_filter with work pending");
  CK(cudaSetDevice(dev->ordinal));
  CK(cudaFree(dev->bloom_bits));
  dev->bloom_bits = nullptr;
  const size_t padded = (size_t)((size_words + 1) / 2 * 2);  // 16-byte multiple for the bulk copy
  CK(cudaMalloc(&dev->bloom_bits, padded * 8));
  CK(cudaMemsetAsync(dev->bloom_bits, 0, padded * 8, dev->stream));
  CK(cudaMemcpyAsync(dev->bloom_bits, bits, (size_t)size_words * 8, cudaMemcpyHostToDevice, dev->stream));
  CK(cudaStreamSynchronize(dev->stream));
  dev->bloom_size = size_words;
  dev->bloom_magic = ~0ULL / size_words;
  dev->bloom_fill = 0.5;
  if (size_words * 8 > 64 * 1024) {  // stays in HBM: measure its fill for the candidate-queue planner
    unsigned long long *d_total = nullptr, total = 0;
    CK(cudaMalloc(&d_total, sizeof total));
    CK(cudaMemsetAsync(d_total, 0, sizeof total, dev->stream));
    bloom_popcount_kernel<<<dev->sm_count * 8, 256, 0, dev->stream>>>(dev->bloom_bits, size_words, d_total);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&total, d_total, sizeof total, cudaMemcpyDeviceToHost, dev->stream));
    CK(cudaStreamSynchronize(dev->stream));
    cudaFree(d_total);
    dev->bloom_fill = (double)total / ((double)size_words * 64.0);
  }
  return ECL_OK;
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

// the six add_kernel variants live in add_inst.cu, one translation unit each (parallel build)
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

// The candidate queue of the asynchronous probe: as large as is reasonable (the add kernel wants >= 75 776 threads
// x 2048 keys per launch and fill^2 of all hashes become candidates), sized once per device.
static int ensure_cand_queue(ecl_dev *dev) {
  if (dev->cand_entries) return ECL_OK;
  u64 want = 1ull << 29;  // 16 GB
  if (const char *env = getenv("ECLOOP_B200_CAND_LOG2")) want = 1ull << std::min(31, std::max(8, atoi(env)));  // test hook
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
  CK(cudaMalloc(&dev->cand_counts, ((size_t)dev->sm_count + 1) * sizeof(u32)));
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

// Queue the launches covering groups [g_begin, g_end) of the pending span. max_groups_per_launch bounds one launch.
static int launch_add(ecl_dev *dev, u64 g_begin, u64 g_end, u64 max_groups_per_launch, bool drain_each,
                      std::vector<ecl_hit> *drain_to) {
  const u32 smem_table = (ADD_H + 1) * 64;
  // the filter rides in shared memory when it fits beside the table; otherwise it stays in HBM and is probed
  // asynchronously (probe_pipe.cuh) unless a span is being redone after a candidate-queue overflow, or drained
  // launch by launch for a dense filter (then the inline probe is exact and simple)
  const u64 bloom_bytes = (dev->bloom_size + 1) / 2 * 16;
  const bool bloom_smem = smem_table + bloom_bytes <= 110u * 1024u;
  static const bool no_pipe = getenv("ECLOOP_B200_INLINE_PROBE") != nullptr;  // measurement hook: probe HBM filters inline
  const bool hbm = !bloom_smem && !dev->force_inline && !drain_each && !no_pipe;
  add_launch_fn fn = pick_add_kernel(dev->p_flags, hbm);
  if (!fn) return fail(dev, ECL_E_ARG, "flags select no address type");
  const u32 smem = smem_table + (bloom_smem ? (u32)bloom_bytes : 0u) + (hbm ? ProbePipe<ADD_THREADS>::BYTES : 0u);
  if (hbm) {
    int rc = ensure_cand_queue(dev);
    if (rc) return rc;
    // stage 1 passes fill^2 of the hashes: a launch is sized so that the expected candidates use 2/3 of the queue
    const u32 hashes_per_key = ((dev->p_flags & ECL_A33) ? 1u : 0u) + ((dev->p_flags & ECL_A65) ? 1u : 0u);
    const double hashes_per_group = (double)GROUP_KEYS * hashes_per_key * ((dev->p_flags & ECL_ENDO) ? 6.0 : 1.0);
    const double per_group = std::max(1.0, hashes_per_group * dev->bloom_fill * dev->bloom_fill * 1.5);
    u64 fit = std::max<u64>(1, (u64)((double)dev->cand_cap / per_group));
    if (fit >= dev->Tmax) fit -= fit % dev->Tmax;  // whole rounds of the grid: every SM keeps its CTA busy
    max_groups_per_launch = std::min(max_groups_per_launch, fit);
  }

  u64 g = g_begin;
  while (g < g_end) {
    const u64 L = std::min<u64>(g_end - g, max_groups_per_launch);
    const u64 c = (L + dev->Tmax - 1) / dev->Tmax;  // groups per thread
    const u32 T = (u32)((L + c - 1) / c);
    // centres: (start + (g*2H + H + t*c*2H) * stride) * G   (GStart, main.c:359-360)
    u64 k0[4], step[4];
    const u64 zero[4] = {0, 0, 0, 0};
    sc_muladd64(k0, dev->stride, g * GROUP_KEYS + ADD_H, dev->p_start);
    sc_muladd64(step, dev->stride, c * GROUP_KEYS, zero);
    SmulParams sp;
    memset(&sp, 0, sizeof sp);
    sp.k0 = to_fe(k0), sp.step = to_fe(step), sp.gtab = dev->gtab, sp.count = T, sp.mode = 1, sp.out = dev->centres;
    smul_kernel<<<(T + 127) / 128, 128, 0, dev->stream>>>(sp);
    CK(cudaGetLastError());

    AddParams ap;
    memset(&ap, 0, sizeof ap);
    ap.cx = dev->centres, ap.cy = dev->centres + (size_t)8 * T;
    ap.table = dev->add_table, ap.scratch = dev->scratch;
    ap.bloom = bloom_view(dev);
    ap.bloom_smem_words = bloom_smem ? (u32)dev->bloom_size : 0u;
    ap.sink.hits = dev->d_hits, ap.sink.count = dev->d_hit_count, ap.sink.cap = dev->hit_cap;
    ap.T = T, ap.groups_per_thread = (u32)c, ap.n_groups = L, ap.key_off0 = g * GROUP_KEYS;
    const u32 grid = (T + ADD_THREADS - 1) / ADD_THREADS;
    if (hbm) {
      ap.cand.entries = dev->cand_entries, ap.cand.counts = dev->cand_counts;
      ap.cand.overflow = dev->cand_counts + dev->sm_count;
      ap.cand.cap_per_cta = (u32)std::min<u64>(dev->cand_cap / grid, 0xffffffffu);
      CK(cudaMemsetAsync(dev->cand_counts, 0, (size_t)dev->sm_count * sizeof(u32), dev->stream));  // not the flag
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
    g += L;

    if (drain_each) {
      CK(cudaStreamSynchronize(dev->stream));
      u32 cnt = 0;
      CK(cudaMemcpy(&cnt, dev->d_hit_count, sizeof cnt, cudaMemcpyDeviceToHost));
      if (cnt > dev->hit_cap) return fail(dev, ECL_E_OVERFLOW, "hit buffer overflow in exact mode (%u > %u)", cnt, dev->hit_cap);
      const size_t old = drain_to->size();
      drain_to->resize(old + cnt);
      if (cnt) CK(cudaMemcpy(drain_to->data() + old, dev->d_hits, (size_t)cnt * sizeof(ecl_hit), cudaMemcpyDeviceToHost));
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
  CK(cudaSetDevice(dev->ordinal));
  dev->ev_used = 0, dev->launches = 0;
  dev->result.clear(), dev->result_ready = false;
  CK(cudaEventRecord(dev->ev_begin, dev->stream));
  int rc = ensure_add_resources(dev);
  if (rc) return rc;
  memcpy(dev->p_start, start_pk, 32);
  dev->p_keys = n_keys, dev->p_flags = flags;
  CK(cudaMemsetAsync(dev->d_hit_count, 0, sizeof(u32), dev->stream));
  dev->force_inline = false;
  if (dev->cand_counts) CK(cudaMemsetAsync(dev->cand_counts + dev->sm_count, 0, sizeof(u32), dev->stream));
  const u64 n_groups = n_keys / GROUP_KEYS;
  rc = launch_add(dev, 0, n_groups, (u64)dev->Tmax * dev->groups_per_thread, false, nullptr);
  if (rc) return rc;
  CK(cudaEventRecord(dev->ev_end, dev->stream));
  dev->pending = 1;
  return ECL_OK;
}

// ---------------------------------------------------------------- mul path

typedef void (*mul_kernel_fn)(const MulParams);
static mul_kernel_fn pick_mul_kernel(u32 flags) {
  const bool c = flags & ECL_A33, u = flags & ECL_A65;
  if (c && u) return mul_kernel<true, true>;
  if (c) return mul_kernel<true, false>;
  if (u) return mul_kernel<false, true>;
  return nullptr;
}

static int launch_mul(ecl_dev *dev, u32 begin, u32 end) {
  MulParams mp;
  memset(&mp, 0, sizeof mp);
  mp.scalars = dev->d_scalars + begin, mp.gtab = dev->gtab, mp.bloom = bloom_view(dev);
  mp.sink.hits = dev->d_hits, mp.sink.count = dev->d_hit_count, mp.sink.cap = dev->hit_cap;
  mp.count = end - begin;
  // keys per thread: as many as it takes to keep ~768 threads per SM busy, so that small batches still spread
  // over the whole GPU and large ones amortise the per-thread inversion
  const u32 want_threads = (u32)dev->sm_count * 768u;
  mp.B = std::min<u32>(64u, (mp.count + want_threads - 1) / want_threads);
  if (mp.B == 0) mp.B = 1;
  mp.T = (mp.count + mp.B - 1) / mp.B;
  if (mp.count > dev->mul_scratch_cap) {
    CK(cudaFree(dev->mul_scratch));
    dev->mul_scratch = nullptr, dev->mul_scratch_cap = 0;
    CK(cudaMalloc(&dev->mul_scratch, ((size_t)mp.count + 64) * 128));
    dev->mul_scratch_cap = mp.count;
  }
  mp.scratch = dev->mul_scratch;
  cudaEvent_t e0 = next_event(dev), e1 = next_event(dev);
  if (!e0 || !e1) return fail(dev, ECL_E_CUDA, "cudaEventCreate failed");
  CK(cudaEventRecord(e0, dev->stream));
  pick_mul_kernel(dev->p_flags)<<<(mp.T + 127) / 128, 128, 0, dev->stream>>>(mp);
  CK(cudaGetLastError());
  CK(cudaEventRecord(e1, dev->stream));
  dev->launches++;
  return ECL_OK;
}

extern "C" int ecl_mul_submit(ecl_dev *dev, const uint64_t (*pks)[4], uint32_t n, uint32_t flags) {
  if (!dev || !pks || n == 0) return fail(dev, ECL_E_ARG, "ecl_mul_submit: no keys");
  if (dev->pending) return fail(dev, ECL_E_STATE, "a submit is already pending; call ecl_collect first");
  if (!dev->bloom_bits) return fail(dev, ECL_E_ARG, "no filter set (ecl_set_filter)");
  if (!pick_mul_kernel(flags)) return fail(dev, ECL_E_ARG, "flags select no address type");
  CK(cudaSetDevice(dev->ordinal));
  dev->ev_used = 0, dev->launches = 0;
  dev->result.clear(), dev->result_ready = false;
  if (n > dev->scalars_cap) {
    CK(cudaFree(dev->d_scalars));
    dev->d_scalars = nullptr;
    CK(cudaMalloc(&dev->d_scalars, (size_t)n * sizeof(fe)));
    dev->scalars_cap = n;
  }
  CK(cudaEventRecord(dev->ev_begin, dev->stream));
  CK(cudaMemcpyAsync(dev->d_scalars, pks, (size_t)n * 32, cudaMemcpyHostToDevice, dev->stream));
  CK(cudaMemsetAsync(dev->d_hit_count, 0, sizeof(u32), dev->stream));
  dev->p_keys = n, dev->p_flags = flags;
  // the (hi, lo) limb image of uint64_t[4] equals fe's 8 x u32 on a little-endian host
  int rc = launch_mul(dev, 0, n);
  if (rc) return rc;
  CK(cudaEventRecord(dev->ev_end, dev->stream));
  dev->pending = 2;
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

extern "C" int ecl_collect(ecl_dev *dev, ecl_hit *hits, uint32_t cap, uint32_t *n_hits, uint64_t *keys_done) {
  if (!dev) return ECL_E_ARG;
  if (!dev->pending) return fail(dev, ECL_E_STATE, "ecl_collect without a pending submit");
  CK(cudaSetDevice(dev->ordinal));
  if (!dev->result_ready) {
    CK(cudaStreamSynchronize(dev->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, dev->ev_begin, dev->ev_end));
    dev->last_total_ms = ms;
    dev->last_hot_ms = 0;
    for (size_t i = 0; i + 1 < dev->ev_used; i += 2) {
      CK(cudaEventElapsedTime(&ms, dev->ev_pool[i], dev->ev_pool[i + 1]));
      dev->last_hot_ms += ms;
    }
    dev->last_launches = dev->launches;
    if (dev->pending == 1 && dev->cand_counts && !dev->force_inline) {
      // the asynchronous probe's candidate queue overflowed (a filter far denser than a bloom filter should be):
      // nothing may be lost, so the span is redone with the probes inline
      u32 ovf = 0;
      CK(cudaMemcpy(&ovf, dev->cand_counts + dev->sm_count, sizeof ovf, cudaMemcpyDeviceToHost));
      if (ovf) {
        dev->force_inline = true;
        CK(cudaMemsetAsync(dev->d_hit_count, 0, sizeof(u32), dev->stream));
        int rc2 = launch_add(dev, 0, dev->p_keys / GROUP_KEYS, (u64)dev->Tmax * dev->groups_per_thread, false, nullptr);
        if (rc2) {
          dev->pending = 0;
          return rc2;
        }
        CK(cudaStreamSynchronize(dev->stream));
      }
    }
    u32 cnt = 0;
    CK(cudaMemcpy(&cnt, dev->d_hit_count, sizeof cnt, cudaMemcpyDeviceToHost));
    dev->result.clear();
    if (cnt <= dev->hit_cap) {
      dev->result.resize(cnt);
      if (cnt) CK(cudaMemcpy(dev->result.data(), dev->d_hits, (size_t)cnt * sizeof(ecl_hit), cudaMemcpyDeviceToHost));
    } else {
      // Dense filter (e.g. the all-ones dump filter): redo the span in slices whose worst case fits the ring.
      CK(cudaMemsetAsync(dev->d_hit_count, 0, sizeof(u32), dev->stream));
      int rc;
      if (dev->pending == 1) {
        const u64 slice = std::max<u64>(1, dev->hit_cap / (MAX_HITS_PER_KEY * GROUP_KEYS));
        rc = launch_add(dev, 0, dev->p_keys / GROUP_KEYS, slice, true, &dev->result);
      } else {
        rc = ECL_OK;
        const u32 slice = dev->hit_cap / 2;
        for (u32 b = 0; b < (u32)dev->p_keys && rc == ECL_OK; b += slice) {
          rc = launch_mul(dev, b, std::min<u32>((u32)dev->p_keys, b + slice));
          if (rc) break;
          CK(cudaStreamSynchronize(dev->stream));
          u32 c2 = 0;
          CK(cudaMemcpy(&c2, dev->d_hit_count, sizeof c2, cudaMemcpyDeviceToHost));
          const size_t old = dev->result.size();
          dev->result.resize(old + c2);
          if (c2) CK(cudaMemcpy(dev->result.data() + old, dev->d_hits, (size_t)c2 * sizeof(ecl_hit), cudaMemcpyDeviceToHost));
          for (size_t i = old; i < dev->result.size(); ++i) dev->result[i].key_off += b;
          CK(cudaMemsetAsync(dev->d_hit_count, 0, sizeof(u32), dev->stream));
        }
      }
      if (rc) {
        dev->pending = 0;
        return rc;
      }
    }
    std::sort(dev->result.begin(), dev->result.end(), dev->pending == 1 ? hit_less_add : hit_less_mul);
    dev->result_ready = true;
  }
  if (n_hits) *n_hits = (u32)std::min<size_t>(dev->result.size(), cap);
  if (keys_done) *keys_done = dev->p_keys;
  if (dev->result.size() > cap)
    return fail(dev, ECL_E_OVERFLOW, "%zu hits do not fit the caller's buffer of %u; call ecl_collect again with a larger one", dev->result.size(), cap);
  if (hits && !dev->result.empty()) memcpy(hits, dev->result.data(), dev->result.size() * sizeof(ecl_hit));
  dev->pending = 0;
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

template <typename T>
struct DevBuf {
  T *p = nullptr;
  ~DevBuf() { cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, n * sizeof(T)); }
};

extern "C" int ecl_prim_fp(ecl_dev *dev, int op, const uint64_t (*a)[4], const uint64_t (*b)[4], uint64_t (*out)[4], uint32_t n) {
  if (!dev || !a || !out || n == 0) return fail(dev, ECL_E_ARG, "ecl_prim_fp: bad arguments");
  CK(cudaSetDevice(dev->ordinal));
  DevBuf<fe> da, db, dout;
  CK(da.alloc(n));
  CK(dout.alloc(n));
  CK(cudaMemcpy(da.p, a, (size_t)n * 32, cudaMemcpyHostToDevice));
  if (b) {
    CK(db.alloc(n));
    CK(cudaMemcpy(db.p, b, (size_t)n * 32, cudaMemcpyHostToDevice));
  }
  prim_fp_kernel<<<(n + 127) / 128, 128, 0, dev->stream>>>(op, da.p, db.p, dout.p, n);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(dev->stream));
  CK(cudaMemcpy(out, dout.p, (size_t)n * 32, cudaMemcpyDeviceToHost));
  return ECL_OK;
}

extern "C" int ecl_prim_scalar_mul(ecl_dev *dev, const uint64_t (*k)[4], uint64_t (*out_xy)[8], uint32_t n) {
  if (!dev || !k || !out_xy || n == 0) return fail(dev, ECL_E_ARG, "ecl_prim_scalar_mul: bad arguments");
  CK(cudaSetDevice(dev->ordinal));
  DevBuf<fe> dk;
  DevBuf<u32> dout;
  CK(dk.alloc(n));
  CK(dout.alloc((size_t)n * 16));
  CK(cudaMemcpy(dk.p, k, (size_t)n * 32, cudaMemcpyHostToDevice));
  SmulParams sp;
  memset(&sp, 0, sizeof sp);
  sp.scalars = dk.p, sp.gtab = dev->gtab, sp.count = n, sp.mode = 2, sp.out = dout.p;
  smul_kernel<<<(n + 127) / 128, 128, 0, dev->stream>>>(sp);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(dev->stream));
  CK(cudaMemcpy(out_xy, dout.p, (size_t)n * 64, cudaMemcpyDeviceToHost));
  return ECL_OK;
}

extern "C" int ecl_prim_hash160(ecl_dev *dev, const uint64_t (*xy)[8], uint32_t (*out33)[5], uint32_t (*out65)[5], uint32_t n) {
  if (!dev || !xy || n == 0) return fail(dev, ECL_E_ARG, "ecl_prim_hash160: bad arguments");
  CK(cudaSetDevice(dev->ordinal));
  DevBuf<u32> dxy, d33, d65;
  CK(dxy.alloc((size_t)n * 16));
  CK(cudaMemcpy(dxy.p, xy, (size_t)n * 64, cudaMemcpyHostToDevice));
  if (out33) CK(d33.alloc((size_t)n * 5));
  if (out65) CK(d65.alloc((size_t)n * 5));
  prim_hash160_kernel<<<(n + 127) / 128, 128, 0, dev->stream>>>(dxy.p, d33.p, d65.p, n);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(dev->stream));
  if (out33) CK(cudaMemcpy(out33, d33.p, (size_t)n * 20, cudaMemcpyDeviceToHost));
  if (out65) CK(cudaMemcpy(out65, d65.p, (size_t)n * 20, cudaMemcpyDeviceToHost));
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
  CK(cudaMemcpy(dh.p, h160, (size_t)n * 20, cudaMemcpyHostToDevice));
  prim_bloom_kernel<<<(n + 127) / 128, 128, 0, dev->stream>>>(bloom_view(dev), dh.p, dout.p, n);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(dev->stream));
  CK(cudaMemcpy(out, dout.p, n, cudaMemcpyDeviceToHost));
  return ECL_OK;
}

// ---------------------------------------------------------------- integer-pipe peaks

template <int KIND>
static int run_peak(ecl_dev *dev, double *gops, double *mhz) {
  DevBuf<u32> out;
  DevBuf<unsigned long long> cyc;
  CK(out.alloc(1024));
  CK(cyc.alloc(1));
  const int blocks = dev->sm_count * 8;  // 8 x 256 threads = 2048 threads per SM: full occupancy
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  unsigned long long cycles = 0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0, dev->stream));
    peak_kernel<KIND><<<blocks, 256, 0, dev->stream>>>(out.p, 0x1234567u + rep, cyc.p);
    CK(cudaEventRecord(e1, dev->stream));
    CK(cudaStreamSynchronize(dev->stream));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) {
      best = ms;
      CK(cudaMemcpy(&cycles, cyc.p, sizeof cycles, cudaMemcpyDeviceToHost));
    }
  }
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  const double insts = (double)blocks * 256.0 * PEAK_ITERS * PEAK_UNROLL * PEAK_CHAINS;
  *gops = insts / (best * 1e-3) / 1e9;
  // one block's loop time in cycles over the kernel's wall time underestimates the clock when blocks run in
  // waves; with 8 blocks/SM all resident it is one wave, so cycles/time ~ SM clock
  *mhz = (double)cycles / (best * 1e-3) / 1e6;
  return ECL_OK;
}

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

// one kind of peak.cuh by number (0..18), or a field-multiplication throughput (19..22, fp64mul.cuh); see ecl_peak_kinds in ecloop_b200/__init__.py for the names
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
  case 19: rc = run_mulbench<0, 0>(dev, gops); break;
  case 20: rc = run_mulbench<1, 0>(dev, gops); break;
  case 21: rc = run_mulbench<0, 384>(dev, gops); break;
  case 22: rc = run_mulbench<1, 384>(dev, gops); break;
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
