/* ecloop_b200.h — C-ABI of libecloop_b200.so: the B200 (sm_100a) replacement for the reference's hot path.
 *
 * The reference (vladkens/ecloop v0.5.0, one C translation unit) has no plugin or FFI layer; the seam this
 * library replaces is the pair of call sites in its worker loops:
 *
 *   main.c:430      batch_add(ctx, pk, job_size)            -> ecl_add_submit + ecl_collect
 *   main.c:531-534  ec_gtable_mul x n, ec_jacobi_grprdc,     -> ecl_mul_submit + ecl_collect
 *                   check_found_mul
 *
 * together with everything those reach down to the bloom decision (main.c:212): fe_modp_* (lib/ecc.c:269-540),
 * the point arithmetic and G table (lib/ecc.c:546-929), ctx_precompute_gpoints (main.c:219-246), prepare33/65 +
 * sha256_final + rmd160_batch (lib/addr.c, lib/sha256.c, lib/rmd160s.c) and blf_has (lib/utils.c:308-326).
 * What stays with the caller: the job dispenser (main.c:419-428), the exact bsearch second stage in list mode
 * (main.c:215), calc_priv (main.c:267-276), ctx_write_found / ctx_update. INTEGRATION.md shows the patch.
 *
 * Conventions: plain C, no CUDA or C++ types; every function returns 0 on success or a negative ECL_E_* code and
 * never calls exit(); ecl_last_error() gives the message. Field elements / scalars are `uint64_t[4]`, little-
 * endian limbs, exactly the reference's `fe` (lib/ecc.c:26). An ecl_dev is bound to one GPU and is not
 * thread-safe: one host thread per device, like one worker thread per job stream in the reference.
 * There is no CPU fallback: without a CUDA device ecl_open fails with ECL_E_NODEV.
 */
#ifndef ECLOOP_B200_H
#define ECLOOP_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECL_ABI_VERSION 2

/* flags (ctx->check_addr33 / check_addr65 / use_endo, main.c:32-34) */
#define ECL_A33 1u
#define ECL_A65 2u
#define ECL_ENDO 4u

#define ECL_OK 0
#define ECL_E_NODEV -1    /* no usable CUDA device / driver                                  */
#define ECL_E_CUDA -2     /* a CUDA call failed (message has the CUDA error string)           */
#define ECL_E_ARG -3      /* bad argument (span not a multiple of 2048, no filter set, ...)   */
#define ECL_E_OVERFLOW -4 /* more bloom-positive keys than the hit buffer holds, even after splitting */
#define ECL_E_STATE -5    /* call out of order (collect without submit, ...)                  */
#define ECL_E_DEGENERATE -6 /* the span reaches key 0 or n: a group centre equals +-m*stride*G, the point addition has
                               no inverse there (the reference asserts at this point, lib/ecc.c:666)      */

#define ECL_GROUP 2048u /* GROUP_INV_SIZE (main.c:17): spans are whole groups */

typedef struct ecl_dev ecl_dev; /* opaque: owns the stream, tables, filter copy, scratch and hit ring */

/* one bloom-positive key. key_off indexes the submitted span: key = start + key_off*stride (mod n) for add,
 * the input index for mul. endo uses calc_priv's numbering (main.c:267-276); h160 is h160_t (lib/addr.c:16). */
typedef struct ecl_hit {
  uint64_t key_off;
  uint32_t h160[5];
  uint8_t endo; /* 0..5 */
  uint8_t kind; /* 0 = addr33, 1 = addr65 */
  uint8_t pad[2];
} ecl_hit; /* 32 bytes */

int ecl_abi_version(void);
int ecl_device_count(void);
int ecl_open(ecl_dev **out, int ordinal);
void ecl_close(ecl_dev *dev);
const char *ecl_last_error(const ecl_dev *dev); /* dev may be NULL: error of the last failed ecl_open */

/* Run the hot path on an existing CUDA stream (cudaStream_t passed as void*), e.g. the caller's torch stream;
 * NULL restores the device's own stream. */
int ecl_set_stream(ecl_dev *dev, void *cuda_stream);

/* blf_t (lib/utils.c:277-280): `size_words` 64-bit words; copied to the device, caller keeps ownership.
 * Both list mode (bloom of 2*count words built by load_filter, main.c:128-130) and .blf mode use this. */
int ecl_set_filter(ecl_dev *dev, const uint64_t *bits, uint64_t size_words);

/* ctx_precompute_gpoints (main.c:219-246): builds the +-i*stride*G table on the device. Default stride is 1. */
int ecl_set_stride(ecl_dev *dev, const uint64_t stride_k[4]);

/* batch_add + check_found_add (main.c:287-403) over keys start + j*stride, j < n_keys. n_keys must be a
 * multiple of ECL_GROUP. Asynchronous: returns after the launches are queued. Any span size is tiled over the
 * whole GPU (the library picks the inversion-group size per launch); larger spans amortise the launch better. */
int ecl_add_submit(ecl_dev *dev, const uint64_t start_pk[4], uint64_t n_keys, uint32_t flags);

/* cmd_mul_worker's compute (main.c:531-534): k*G for each key, hash160, bloom. Keys are any 256-bit values
 * (as produced by fe_modn_from_hex or -raw); keys = 0 mod n are skipped. Copies pks (into pinned staging memory)
 * before returning. Up to ECL_MUL_DEPTH mul submits may be pending at once (upload and compute of one batch overlap
 * the host work on the next, main.c:549-569's producer/consumer at GPU speed); ecl_collect returns them in
 * submission order. */
#define ECL_MUL_DEPTH 2
int ecl_mul_submit(ecl_dev *dev, const uint64_t (*pks)[4], uint32_t n, uint32_t flags);
/* optional: allocate the staging (pinned host + device) buffers of all ECL_MUL_DEPTH submit slots for batches of up to n
 * keys now instead of inside the first submits (page-locking hundreds of MB takes tenths of a second) */
int ecl_mul_reserve(ecl_dev *dev, uint32_t n);

/* Wait for the submitted work and fetch its bloom-positive keys, sorted into the reference's `-t 1` emission
 * order: add -> by group of 2048, then plain before endo, then key, then endo index, then kind;
 * mul -> by key index, then kind. *n_hits = number written (<= cap); ECL_E_OVERFLOW if cap was too small
 * (the work is kept: call again with a larger buffer). *keys_done = keys covered (n_keys or n). */
int ecl_collect(ecl_dev *dev, ecl_hit *hits, uint32_t cap, uint32_t *n_hits, uint64_t *keys_done);

/* ---- bloom filters built, loaded and kept on the device (.blf tooling at scale: blf_load, blf_gen, lib/utils.c:362-475).
 * ecl_set_filter is the one-call form for filters that are already in host memory. */
/* a zeroed filter of size_words words in HBM; replaces the current filter */
int ecl_filter_alloc(ecl_dev *dev, uint64_t size_words);
/* words [offset_words, offset_words + n_words) of the filter <- bits. Asynchronous on the device's stream when bits is
 * pinned memory (ecl_host_alloc); the caller must not reuse `bits` before ecl_filter_flush returns. */
int ecl_filter_write(ecl_dev *dev, uint64_t offset_words, const uint64_t *bits, uint64_t n_words);
int ecl_filter_flush(ecl_dev *dev);
/* ends a sequence of writes / adds: measures the fill (launch planning of the asynchronous probe) */
int ecl_filter_commit(ecl_dev *dev);
/* words of the device's filter -> host (blf_save) */
int ecl_filter_read(ecl_dev *dev, uint64_t offset_words, uint64_t *bits, uint64_t n_words);
/* dst gets a copy of src's filter, device to device (NVLink peer copy when the two GPUs are peers) */
int ecl_filter_copy_peer(ecl_dev *dst, ecl_dev *src);
/* blf_gen's insert loop (lib/utils.c:453-465) for n hashes in input order: `if (blf_has) continue; blf_add; count++`.
 * *n_new is exactly the reference's count: hashes whose 20 bits were not all set by the filter as it was plus the
 * hashes before them in this call. */
int ecl_filter_add(ecl_dev *dev, const uint32_t (*h160)[5], uint32_t n, uint64_t *n_new);
/* synthetic filter for benchmarks and tests: every bit i.i.d. set with probability
 * round(fill*256)/256, from a counter-based generator (splitmix64 of seed and position): reproducible anywhere */
int ecl_filter_generate(ecl_dev *dev, uint64_t size_words, double fill, uint64_t seed);
/* fraction of set bits measured by the last ecl_set_filter / ecl_filter_commit (0.5 for small filters: not measured) */
double ecl_filter_fill(const ecl_dev *dev);
/* pinned (page-locked) host memory for staging filter chunks; plain C callers need no CUDA headers */
void *ecl_host_alloc(uint64_t bytes);
void ecl_host_free(void *p);

/* ---- device time of the last completed submit..collect, from CUDA events on the launch stream (ms) */
int ecl_last_elapsed_ms(ecl_dev *dev, float *total_ms, float *hot_kernel_ms, uint32_t *kernel_launches);

/* ---- tuning knobs (0 = library default); returns ECL_E_ARG for unsupported values.
 * groups_per_thread: upper bound on the 2048-key groups one thread walks per launch (bounds a launch to
 * 148 x 512 x groups_per_thread x 2048 keys; default 64). */
int ecl_set_tuning(ecl_dev *dev, uint32_t groups_per_thread, uint32_t hit_capacity);

/* ---- primitive entry points, one per reference routine, used by the parity tests (tests/test_gpu_*.py).
 * All take host pointers and run the same device functions the hot kernels inline. */
enum { ECL_OP_MUL = 0, ECL_OP_SQR = 1, ECL_OP_ADD = 2, ECL_OP_SUB = 3, ECL_OP_NEG = 4, ECL_OP_INV = 5 };
#ifdef ECL_EXPERIMENTAL
/* Not part of the boundary: only libecloop_b200_exp.so (built with -DECL_EXPERIMENTAL for tests/test_gpu_experimental.py)
 * knows these. FP64-pipe multiplication (csrc/fp64mul.cuh): a*b, a*b^16 chained in its own limb form, and x / y of
 * (a, b) + G by batch_add's affine formula (main.c:378-386) computed in that limb form. */
enum { ECL_OP_MUL_F64 = 6, ECL_OP_MUL_F64_CHAIN = 7, ECL_OP_AFFINE_F64_X = 8, ECL_OP_AFFINE_F64_Y = 9 };
#endif
/* fe_modp_mul/sqr/add/sub/neg/inv (lib/ecc.c:269-520) elementwise over n elements */
int ecl_prim_fp(ecl_dev *dev, int op, const uint64_t (*a)[4], const uint64_t (*b)[4], uint64_t (*out)[4], uint32_t n);
/* ec_gtable_mul + ec_jacobi_rdc (lib/ecc.c:907-929, 686-693): out_xy[i] = {x[4], y[4]} affine, zeros for k=0 */
int ecl_prim_scalar_mul(ecl_dev *dev, const uint64_t (*k)[4], uint64_t (*out_xy)[8], uint32_t n);
/* addr33 / addr65 (lib/addr.c:75-95): hash160 of affine points; either output may be NULL */
int ecl_prim_hash160(ecl_dev *dev, const uint64_t (*xy)[8], uint32_t (*out33)[5], uint32_t (*out65)[5], uint32_t n);
/* blf_has (lib/utils.c:308-326) against the filter set with ecl_set_filter */
int ecl_prim_bloom(ecl_dev *dev, const uint32_t (*h160)[5], uint8_t *out, uint32_t n);

/* ---- integer-pipe throughput microbenchmark (the roofline denominator of SURVEY §8d, measured in-process):
 * result in Gops/s (32 lanes x warp instructions / time) for: [0] LOP3, [1] IADD3, [2] SHF, [3] IMAD (lo),
 * [4] IMAD.WIDE.U32, [5] LOP3 + IMAD co-issue (sum), [6] SM clock in MHz during the run (cycle counter over
 * %globaltimer inside the kernel) */
int ecl_peak_bench(ecl_dev *dev, double out[8]);
/* one instruction kind / mix of peak.cuh by number: 0 LOP3, 1 IADD3, 2 SHF, 3 IMAD, 4 IMAD.WIDE, 5 LOP3+IMAD,
 * 6 IMAD with a constant-bank operand, 7 IMAD.HI, 8 LOP3+IMAD(const), 9 SHF+IMAD.WIDE, 10 LOP3+IMAD.HI,
 * 11 5:3 LOP3:IMAD(const), 12 two-input add, 13 LOP3+IMAD.WIDE, 14 SHF+IMAD, 15 LOP3+SHF,
 * 16 DFMA, 17 DFMA+LOP3, 18 DFMA+IMAD. Result in Gops/s (32 lanes x instructions / time). */
int ecl_peak_bench_kind(ecl_dev *dev, int kind, double *gops, double *sm_mhz);

#ifdef __cplusplus
}
#endif
#endif
