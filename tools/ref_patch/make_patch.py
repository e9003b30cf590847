#!/usr/bin/env python3
"""Regenerate tools/ref_patch/main.patch: the edit a maintainer of vladkens/ecloop would make to put
libecloop_b200.so behind the reference's own main.c. Reads /root/reference/main.c, applies the
edits below to a scratch copy and writes the unified diff (1 line of context). Run in the build container only.

The patch keeps the reference's job dispenser, filter loading, calc_priv, pk_verify_hash (which recomputes every hit
with the reference's OWN CPU field/hash code: an independent check of the GPU result), ctx_write_found and status
line; batch_add (main.c:349-403,430) and the mul worker's compute (main.c:531-534) go through the C-ABI."""
import difflib
import sys
from pathlib import Path

REF = Path("/root/reference/main.c")
OUT = Path(__file__).resolve().parent / "main.patch"

INCLUDE = '''#include "ecloop_b200.h" // B200 hot path: libecloop_b200.so (C-ABI)

// one device for the whole process; an ecl_dev is not thread-safe, so workers hold g_gpu_lock from submit to collect
static ecl_dev *g_gpu;
static pthread_mutex_t g_gpu_lock = PTHREAD_MUTEX_INITIALIZER;
'''

ADD_FN = '''// batch_add + check_found_add on the GPU: the library returns the bloom-positive keys of the job in the -t 1
// emission order; the exact list stage, calc_priv, pk_verify_hash (CPU) and ctx_write_found stay here
void batch_add_gpu(ctx_t *ctx, const fe pk, const size_t iterations) {
  const uint64_t n = (iterations + GROUP_INV_SIZE - 1) / GROUP_INV_SIZE * GROUP_INV_SIZE;
  const uint32_t flags = (ctx->check_addr33 ? ECL_A33 : 0) | (ctx->check_addr65 ? ECL_A65 : 0) | (ctx->use_endo ? ECL_ENDO : 0);
  uint32_t cap = 1 << 12, cnt = 0;
  uint64_t done = 0;
  ecl_hit *hits = malloc(cap * sizeof(ecl_hit));

  pthread_mutex_lock(&g_gpu_lock);
  int rc = ecl_add_submit(g_gpu, (const uint64_t *)pk, n, flags);
  while (rc == ECL_OK && (rc = ecl_collect(g_gpu, hits, cap, &cnt, &done)) == ECL_E_OVERFLOW) {
    cap *= 8; // the library keeps the result until the buffer is large enough
    hits = realloc(hits, cap * sizeof(ecl_hit));
  }
  if (rc != ECL_OK) {
    fprintf(stderr, "ecloop_b200: %s\\n", ecl_last_error(g_gpu));
    exit(1);
  }
  pthread_mutex_unlock(&g_gpu_lock);

  for (uint32_t i = 0; i < cnt; ++i) {
    const ecl_hit *h = &hits[i];
    if (ctx->to_find_hashes != NULL &&
        bsearch(h->h160, ctx->to_find_hashes, ctx->to_find_count, sizeof(h160_t), compare_160) == NULL)
      continue;
    fe ck;
    calc_priv(ck, pk, ctx->stride_k, h->key_off, h->endo);
    pk_verify_hash(ck, h->h160, h->kind == 0, h->endo);
    ctx_write_found(ctx, h->kind == 0 ? "addr33" : "addr65", h->h160, ck);
  }
  free(hits);
}

'''

MUL_FN = '''// ec_gtable_mul x n + ec_jacobi_grprdc + check_found_mul on the GPU
void mul_gpu(ctx_t *ctx, const fe *pk, size_t count) {
  const uint32_t flags = (ctx->check_addr33 ? ECL_A33 : 0) | (ctx->check_addr65 ? ECL_A65 : 0);
  ecl_hit hits[2 * GROUP_INV_SIZE];
  uint32_t cnt = 0;
  uint64_t done = 0;

  pthread_mutex_lock(&g_gpu_lock);
  int rc = ecl_mul_submit(g_gpu, (const uint64_t (*)[4])pk, (uint32_t)count, flags);
  if (rc == ECL_OK) rc = ecl_collect(g_gpu, hits, 2 * GROUP_INV_SIZE, &cnt, &done);
  if (rc != ECL_OK) {
    fprintf(stderr, "ecloop_b200: %s\\n", ecl_last_error(g_gpu));
    exit(1);
  }
  pthread_mutex_unlock(&g_gpu_lock);

  for (uint32_t i = 0; i < cnt; ++i) {
    const ecl_hit *h = &hits[i];
    if (ctx->to_find_hashes != NULL &&
        bsearch(h->h160, ctx->to_find_hashes, ctx->to_find_count, sizeof(h160_t), compare_160) == NULL)
      continue;
    ctx_write_found(ctx, h->kind == 0 ? "addr33" : "addr65", h->h160, pk[h->key_off]);
  }
}

'''

OPEN = '''  if (ecl_open(&g_gpu, 0) != ECL_OK) {
    fprintf(stderr, "ecloop_b200: %s\\n", ecl_last_error(NULL));
    exit(1);
  }
  ecl_set_filter(g_gpu, (const uint64_t *)ctx->blf.bits, ctx->blf.size);

'''


def edit(src: str) -> str:
    def once(s, old, new, count=1):
        assert s.count(old) == count, (old, s.count(old))
        return s.replace(old, new)

    s = src
    s = once(s, '#include "lib/utils.c"\n', '#include "lib/utils.c"\n' + INCLUDE)
    s = once(s, "void *cmd_add_worker(void *arg) {\n", ADD_FN + "void *cmd_add_worker(void *arg) {\n")
    s = once(s, "    batch_add(ctx, pk, ctx->job_size);\n", "    batch_add_gpu(ctx, pk, ctx->job_size);\n")
    s = once(s, "void *cmd_mul_worker(void *arg) {\n", MUL_FN + "void *cmd_mul_worker(void *arg) {\n")
    s = once(s, "    for (size_t i = 0; i < job->count; ++i) ec_gtable_mul(&cp[i], pk[i]);\n    ec_jacobi_grprdc(cp, job->count);\n\n"
                "    check_found_mul(ctx, pk, cp, job->count);\n",
             "    (void)cp;\n    mul_gpu(ctx, (const fe *)pk, job->count);\n")
    s = once(s, "  ctx_precompute_gpoints(ctx);\n", "  ctx_precompute_gpoints(ctx);\n  ecl_set_stride(g_gpu, (const uint64_t *)ctx->stride_k); // the +-i*stride*G table is built on the device\n", count=2)
    s = once(s, '  printf("----------------------------------------\\n");\n}\n', OPEN + '  printf("----------------------------------------\\n");\n}\n')
    return s


def main():
    src = REF.read_text()
    new = edit(src)
    diff = difflib.unified_diff(src.splitlines(keepends=True), new.splitlines(keepends=True), "a/main.c", "b/main.c", n=1)
    OUT.write_text("".join(diff))
    print(OUT, len(OUT.read_text().splitlines()), "lines")


if __name__ == "__main__":
    sys.exit(main())
