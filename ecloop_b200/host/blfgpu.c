/* blfgpu.c — `ecloop blf-gen` with the filter in HBM and the insert loop on the GPU (blftool.h).
 * blf_gen (lib/utils.c:409-475) touches 20 random words of a multi-GB array per hash from one thread; here the host
 * only cuts stdin into the pieces fgets(41) would return and parses hex, 4 M hashes at a time, and
 * ecl_filter_add does the probing, the inserts and the exact count. Output file and messages are the reference's. */
#define _GNU_SOURCE
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/ecloop_b200.h"
#include "blftool.h"
#include "filter.h"

#define GEN_BLOCK_BYTES (64u << 20)
#define GEN_BATCH (1u << 22)
#define IO_CHUNK_WORDS (8u << 20) /* 64 MiB */

int blf_gen_gpu_main(int argc, const char **argv) {
  blf_gen_plan plan;
  const int rc = blf_gen_args(argc, argv, &plan);
  if (rc) return rc;
  ecl_dev *dev = NULL;
  if (ecl_open(&dev, 0) != ECL_OK) {
    fprintf(stderr, "[!] cannot open GPU 0: %s (use `blf-gen -cpu` for the host tool)\n", ecl_last_error(NULL));
    return 1;
  }
  uint64_t *chunk = ecl_host_alloc((uint64_t)IO_CHUNK_WORDS * 8);
  uint32_t(*batch)[5] = malloc((size_t)GEN_BATCH * sizeof *batch);
  char *text = malloc(GEN_BLOCK_BYTES + 64);
  if (!chunk || !batch || !text) {
    fprintf(stderr, "[!] out of memory\n");
    return 1;
  }
#define GPU_OK(call)                                                    \
  do {                                                                  \
    if ((call) != ECL_OK) {                                             \
      fprintf(stderr, "[!] GPU error: %s\n", ecl_last_error(dev));      \
      return 1;                                                         \
    }                                                                   \
  } while (0)

  if (access(plan.path, F_OK) == 0) {
    const char *todo = "delete it or choose a different file";
    printf("file %s already exists; loading...\n", plan.path);
    ecl_filter f;
    if (filter_open_blf(&f, plan.path) != 0) {
      fprintf(stderr, "[!] failed to load bloom filter: %s\n", todo);
      return 1;
    }
    if (f.size != plan.size) {
      fprintf(stderr, "[!] bloom filter size mismatch (%'zu != %'zu): %s\n", (size_t)f.size, (size_t)plan.size, todo);
      return 1;
    }
    GPU_OK(ecl_filter_alloc(dev, plan.size));
    for (uint64_t have = 0; have < f.size;) {
      const int64_t n = filter_stream_blf(&f, chunk, IO_CHUNK_WORDS, have);
      if (n <= 0) {
        fprintf(stderr, "[!] failed to load bloom filter: %s\n", todo);
        return 1;
      }
      GPU_OK(ecl_filter_write(dev, have, chunk, (uint64_t)n));
      GPU_OK(ecl_filter_flush(dev));
      have += (uint64_t)n;
    }
    filter_free(&f);
    printf("updating bloom filter...\n");
  } else {
    printf("creating bloom filter...\n");
    GPU_OK(ecl_filter_alloc(dev, plan.size));
  }
  printf("bloom filter params: n = %'llu | p = 1:%'llu | m = %'llu (%'.1f MB)\n", plan.n, plan.r, plan.m, plan.mb);

  /* stdin in the pieces fgets(line, 41) returns: up to 40 characters, ending behind a newline if one comes first;
   * a piece of exactly 40 characters is a hash (lib/utils.c:455-460, the same chunking as load_filter) */
  unsigned long long added = 0;
  uint32_t nb = 0;
  size_t have = 0;
  bool eof = false;
  while (!eof || have) {
    if (!eof) {
      const size_t got = fread(text + have, 1, GEN_BLOCK_BYTES - have, stdin);
      have += got;
      if (have < GEN_BLOCK_BYTES) eof = true;
    }
    size_t pos = 0;
    while (pos < have) {
      const size_t room = have - pos < 40 ? have - pos : 40;
      const char *nl = memchr(text + pos, '\n', room);
      if (!nl && room < 40 && !eof) break; /* the piece continues in the next block */
      const size_t take = nl ? (size_t)(nl - (text + pos)) + 1 : room;
      if (take == 40 && !memchr(text + pos, 0, 40)) {
        blf_hex40_words(batch[nb++], text + pos);
        if (nb == GEN_BATCH) {
          uint64_t fresh = 0;
          GPU_OK(ecl_filter_add(dev, (const uint32_t(*)[5])batch, nb, &fresh));
          added += fresh, nb = 0;
        }
      }
      pos += take;
    }
    memmove(text, text + pos, have - pos);
    have -= pos;
    if (eof && have == 0) break;
  }
  if (nb) {
    uint64_t fresh = 0;
    GPU_OK(ecl_filter_add(dev, (const uint32_t(*)[5])batch, nb, &fresh));
    added += fresh;
  }
  printf("added %'llu new items; saving to %s\n", added, plan.path);

  FILE *fp = fopen(plan.path, "wb"); /* blf_save (lib/utils.c:328-360) */
  if (!fp) {
    fprintf(stderr, "failed to open output file\n");
    exit(1);
  }
  const uint32_t head[2] = {0x45434246u, 1u};
  bool ok = fwrite(head, sizeof head, 1, fp) == 1 && fwrite(&plan.size, sizeof plan.size, 1, fp) == 1;
  for (uint64_t off = 0; ok && off < plan.size;) {
    const uint64_t n = plan.size - off < IO_CHUNK_WORDS ? plan.size - off : IO_CHUNK_WORDS;
    GPU_OK(ecl_filter_read(dev, off, chunk, n));
    ok = fwrite(chunk, 8, n, fp) == n;
    off += n;
  }
  if (fclose(fp) != 0) ok = false;
  if (!ok) {
    fprintf(stderr, "failed to write bloom filter bits\n");
    fprintf(stderr, "[!] failed to save bloom filter\n");
    return 1;
  }
  ecl_host_free(chunk);
  free(batch), free(text);
  ecl_close(dev);
  return 0;
}
