/* filter.c — see filter.h */
#include "filter.h"

#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define BLOOM_MAGIC 0x45434246u /* "FBCE" on disk (lib/utils.c:274) */
#define BLOOM_VERSION 1u

void bloom_positions(const uint32_t h[5], uint64_t size_words, uint64_t pos[20]) {
  /* five overlapping 64-bit lanes, four shifts; position = v mod (size*64)  (lib/utils.c:282-306) */
  const uint64_t lane[5] = {(uint64_t)h[0] << 32 | h[1], (uint64_t)h[2] << 32 | h[3], (uint64_t)h[4] << 32 | h[0],
                            (uint64_t)h[1] << 32 | h[2], (uint64_t)h[3] << 32 | h[4]};
  static const unsigned shifts[4] = {24, 28, 36, 40};
  const uint64_t nbits = size_words * 64;
  int k = 0;
  for (int s = 0; s < 4; ++s)
    for (int i = 0; i < 5; ++i) pos[k++] = ((lane[i] << shifts[s]) | (lane[(i + 1) % 5] >> shifts[s])) % nbits;
}

void bloom_add(uint64_t *bits, uint64_t size_words, const uint32_t h[5]) {
  uint64_t pos[20];
  bloom_positions(h, size_words, pos);
  for (int k = 0; k < 20; ++k) bits[pos[k] >> 6] |= 1ULL << (pos[k] & 63);
}

bool bloom_has(const uint64_t *bits, uint64_t size_words, const uint32_t h[5]) {
  uint64_t pos[20];
  bloom_positions(h, size_words, pos);
  for (int k = 0; k < 20; ++k)
    if (!((bits[pos[k] >> 6] >> (pos[k] & 63)) & 1)) return false;
  return true;
}

int bloom_save(const char *path, const uint64_t *bits, uint64_t size_words) {
  FILE *fp = fopen(path, "wb");
  if (!fp) {
    fprintf(stderr, "failed to open output file\n");
    return -1;
  }
  const uint32_t head[2] = {BLOOM_MAGIC, BLOOM_VERSION};
  int ok = fwrite(head, sizeof head, 1, fp) == 1 && fwrite(&size_words, sizeof size_words, 1, fp) == 1 &&
           fwrite(bits, sizeof(uint64_t), size_words, fp) == size_words;
  fclose(fp);
  if (!ok) fprintf(stderr, "failed to write bloom filter bits\n");
  return ok ? 0 : -1;
}

int filter_load_blf(ecl_filter *f, const char *path) {
  FILE *fp = fopen(path, "rb");
  if (!fp) {
    fprintf(stderr, "failed to open input file\n");
    return -1;
  }
  uint32_t head[2];
  uint64_t size = 0;
  if (fread(head, sizeof head, 1, fp) != 1 || fread(&size, sizeof size, 1, fp) != 1) {
    fprintf(stderr, "failed to read bloom filter header\n");
    fclose(fp);
    return -1;
  }
  if (head[0] != BLOOM_MAGIC || head[1] != BLOOM_VERSION) {
    fprintf(stderr, "invalid bloom filter version; create a new filter with blf-gen command\n");
    fclose(fp);
    return -1;
  }
  uint64_t *bits = calloc(size ? size : 1, sizeof(uint64_t));
  if (!bits || fread(bits, sizeof(uint64_t), size, fp) != size) {
    fprintf(stderr, "failed to read bloom filter bits\n");
    free(bits);
    fclose(fp);
    return -1;
  }
  fclose(fp);
  f->bits = bits, f->size = size, f->list = NULL, f->count = 0, f->blf_fd = -1;
  return 0;
}

int filter_open_blf(ecl_filter *f, const char *path) {
  memset(f, 0, sizeof *f);
  f->blf_fd = -1;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) {
    fprintf(stderr, "failed to open input file\n");
    return -1;
  }
  struct {
    uint32_t magic, version;
    uint64_t size;
  } head;
  if (read(fd, &head, sizeof head) != (ssize_t)sizeof head) {
    fprintf(stderr, "failed to read bloom filter header\n");
    close(fd);
    return -1;
  }
  if (head.magic != BLOOM_MAGIC || head.version != BLOOM_VERSION) {
    fprintf(stderr, "invalid bloom filter version; create a new filter with blf-gen command\n");
    close(fd);
    return -1;
  }
#ifdef POSIX_FADV_SEQUENTIAL
  posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
  f->size = head.size, f->blf_fd = fd;
  return 0;
}

int64_t filter_stream_blf(ecl_filter *f, uint64_t *buf, uint64_t max_words, uint64_t already) {
  const uint64_t want_words = f->size - already < max_words ? f->size - already : max_words;
  size_t want = (size_t)want_words * 8, got = 0;
  while (got < want) {
    const ssize_t r = read(f->blf_fd, (char *)buf + got, want - got);
    if (r < 0 && errno == EINTR) continue;
    if (r <= 0) {
      fprintf(stderr, "failed to read bloom filter bits\n");
      return -1;
    }
    got += (size_t)r;
  }
  return (int64_t)want_words;
}

static int cmp_h160(const void *a, const void *b) {
  const uint32_t *x = ((const h160 *)a)->w, *y = ((const h160 *)b)->w;
  for (int i = 0; i < 5; ++i)
    if (x[i] != y[i]) return x[i] > y[i] ? 1 : -1;
  return 0;
}

/* value of up to 8 leading hex digits of s, like sscanf("%8x") (stops at the first non-hex character) */
static uint32_t hex8(const char *s) {
  uint32_t v = 0;
  for (int i = 0; i < 8; ++i) {
    const char c = s[i];
    uint32_t d;
    if (c >= '0' && c <= '9') d = (uint32_t)(c - '0');
    else if (c >= 'a' && c <= 'f') d = (uint32_t)(c - 'a' + 10);
    else if (c >= 'A' && c <= 'F') d = (uint32_t)(c - 'A' + 10);
    else break;
    v = v << 4 | d;
  }
  return v;
}

int filter_load(ecl_filter *f, const char *path) {
  memset(f, 0, sizeof *f);
  f->blf_fd = -1;
  if (!path) {
    fprintf(stderr, "missing filter file\n");
    return -1;
  }
  FILE *fp = fopen(path, "rb");
  if (!fp) {
    fprintf(stderr, "failed to open filter file: %s\n", path);
    return -1;
  }
  const char *ext = strrchr(path, '.');
  if (ext && strcmp(ext, ".blf") == 0) {
    fclose(fp);
    return filter_open_blf(f, path); /* streamed to the GPUs by the caller (ecloop.c load_filter_on_devices) */
  }

  /* Text list. The reference reads with fgets into a 41-byte buffer and keeps only reads of exactly 40
   * characters, so a longer line is consumed in 40-character pieces and each full piece counts (SURVEY A.7:
   * the comment line of data/btc-bw-hash becomes one entry). Same chunking here. */
  size_t cap = 1024, n = 0;
  h160 *list = malloc(cap * sizeof *list);
  char piece[41];
  while (list && fgets(piece, sizeof piece, fp)) {
    if (strlen(piece) != 40) continue;
    if (n == cap) {
      cap *= 2;
      h160 *grown = realloc(list, cap * sizeof *list);
      if (!grown) break;
      list = grown;
    }
    for (int j = 0; j < 5; ++j) list[n].w[j] = hex8(piece + 8 * j);
    n++;
  }
  fclose(fp);
  if (!list) {
    fprintf(stderr, "out of memory while loading filter\n");
    return -1;
  }
  qsort(list, n, sizeof *list, cmp_h160);
  size_t uniq = n ? 1 : 0;
  for (size_t i = 1; i < n; ++i)
    if (cmp_h160(&list[uniq - 1], &list[i]) != 0) list[uniq++] = list[i];
  if (uniq == 0) { /* the reference reports a list of 1 uninitialised entry here; we refuse instead */
    fprintf(stderr, "filter file has no 40-digit hash lines: %s\n", path);
    free(list);
    return -1;
  }
  f->list = list, f->count = uniq;
  f->size = 2 * (uint64_t)uniq;
  f->bits = calloc(f->size, sizeof(uint64_t));
  if (!f->bits) {
    fprintf(stderr, "out of memory while loading filter\n");
    free(list);
    return -1;
  }
  for (size_t i = 0; i < uniq; ++i) bloom_add(f->bits, f->size, list[i].w);
  return 0;
}

void filter_free(ecl_filter *f) {
  free(f->bits);
  free(f->list);
  if (f->blf_fd >= 0) close(f->blf_fd);
  memset(f, 0, sizeof *f);
  f->blf_fd = -1;
}

bool filter_exact(const ecl_filter *f, const uint32_t h[5]) {
  if (!f->list) return true;
  h160 key;
  memcpy(key.w, h, sizeof key.w);
  return bsearch(&key, f->list, f->count, sizeof(h160), cmp_h160) != NULL;
}
