/* filter.h — the `-f` filter of the host CLI: a text list of hash160 values or a `.blf` bloom file.
 *
 * Mirrors load_filter (main.c:71-131) and the bloom container blf_t / blf_add / blf_load (lib/utils.c:274-396):
 * same file format (u32 magic 0x45434246, u32 version 1, u64 size in words, the words), same 20 bit positions
 * per hash, same list-mode sizing (2 words per unique hash). The bloom *probe* of the hot path runs on the GPU
 * (ecl_set_filter); the host keeps the sorted list for the exact second stage of ctx_check_hash (main.c:214-216).
 */
#ifndef ECL_FILTER_H
#define ECL_FILTER_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef struct h160 {
  uint32_t w[5]; /* h160_t (lib/addr.c:16): big-endian words of the digest */
} h160;

typedef struct ecl_filter {
  uint64_t *bits;  /* bloom words; NULL for a `.blf` opened with filter_open_blf (streamed, never whole in host memory) */
  uint64_t size;   /* number of 64-bit words */
  h160 *list;      /* sorted unique hashes (list mode) or NULL (bloom-only mode) */
  size_t count;
  int blf_fd;      /* streamed `.blf`: open file positioned behind the 16-byte header, else -1 */
} ecl_filter;

/* returns 0, or -1 after printing the reference's message for the failure to stderr */
int filter_load(ecl_filter *f, const char *path);
int filter_load_blf(ecl_filter *f, const char *path); /* a `.blf` file whatever its name (blf_load, lib/utils.c:362) */
/* header check of blf_load only: f->size is set, f->blf_fd stays open for filter_stream_blf. A multi-GB filter goes
 * from the page cache through pinned staging chunks straight to the GPUs instead of calloc + fread + pageable copy. */
int filter_open_blf(ecl_filter *f, const char *path);
/* next chunk of words of a streamed `.blf` into buf (up to max_words); returns the number read, 0 at the end,
 * -1 (after the reference's message) when the file is shorter than its header says */
int64_t filter_stream_blf(ecl_filter *f, uint64_t *buf, uint64_t max_words, uint64_t already);
void filter_free(ecl_filter *f);
/* second stage of ctx_check_hash: exact membership in list mode, always true in bloom-only mode */
bool filter_exact(const ecl_filter *f, const uint32_t h[5]);

void bloom_positions(const uint32_t h[5], uint64_t size_words, uint64_t pos[20]);
void bloom_add(uint64_t *bits, uint64_t size_words, const uint32_t h[5]);
bool bloom_has(const uint64_t *bits, uint64_t size_words, const uint32_t h[5]);
int bloom_save(const char *path, const uint64_t *bits, uint64_t size_words);
#endif
