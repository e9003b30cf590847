/* blftool.c — see blftool.h */
#define _GNU_SOURCE
#include "blftool.h"

#include <ctype.h>
#include <math.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "filter.h"

static const char *value_of(int argc, const char **argv, const char *name) {
  for (int i = 1; i + 1 < argc; ++i)
    if (strcmp(argv[i], name) == 0) return argv[i + 1];
  return NULL;
}

/* 5 words from 40 hex characters, each group read like sscanf("%8x") */
static void words_from_hex40(uint32_t h[5], const char *s) {
  for (int j = 0; j < 5; ++j) {
    uint32_t v = 0;
    for (int i = 0; i < 8; ++i) {
      const char c = s[8 * j + i];
      uint32_t d;
      if (c >= '0' && c <= '9') d = (uint32_t)(c - '0');
      else if (c >= 'a' && c <= 'f') d = (uint32_t)(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F') d = (uint32_t)(c - 'A' + 10);
      else break;
      v = v << 4 | d;
    }
    h[j] = v;
  }
}

static int gen_usage(const char *name) {
  printf("Usage: %s blf-gen -n <count> -o <file>\n", name);
  printf("Generate a bloom filter from a list of hex-encoded hash160 values passed to stdin.\n");
  printf("\nOptions:\n");
  printf("  -n <count>      - Number of hashes to add.\n");
  printf("  -o <file>       - File to write bloom filter (must have a .blf extension).\n");
  return 1;
}

int blf_gen_args(int argc, const char **argv, blf_gen_plan *out) {
  const char *nraw = NULL;
  for (int i = 1; i < argc - 1; ++i)
    if (strcmp(argv[i], "-n") == 0) {
      nraw = argv[i + 1];
      break;
    }
  const unsigned long long n = nraw ? strtoull(nraw, NULL, 10) : 0;
  if (n == 0) {
    fprintf(stderr, "[!] missing filter size (-n <number>)\n");
    return gen_usage(argv[0]);
  }
  const char *path = value_of(argc, argv, "-o");
  if (!path) {
    fprintf(stderr, "[!] missing output file (-o <file>)\n");
    return gen_usage(argv[0]);
  }

  /* m bits for n items at p = 1e-9 with the optimal number of probes (the tool always uses 20): lib/utils.c:421-427 */
  const unsigned long long r = 1000000000ULL;
  const double p = 1.0 / (double)r;
  const unsigned long long m = (unsigned long long)((double)n * log(p) / log(1.0 / pow(2.0, log(2.0))));
  const double mb = (double)m / 8 / 1024 / 1024;
  out->n = n, out->r = r, out->m = m, out->mb = mb, out->size = (m + 63) / 64, out->path = path;
  return 0;
}

void blf_hex40_words(uint32_t h[5], const char *s) { words_from_hex40(h, s); }

int blf_gen_main(int argc, const char **argv) {
  blf_gen_plan plan;
  const int rc = blf_gen_args(argc, argv, &plan);
  if (rc) return rc;
  const unsigned long long n = plan.n, r = plan.r, m = plan.m;
  const double mb = plan.mb;
  const uint64_t size = plan.size;
  const char *path = plan.path;

  ecl_filter f = {0};
  if (access(path, F_OK) == 0) {
    const char *todo = "delete it or choose a different file";
    printf("file %s already exists; loading...\n", path);
    if (filter_load_blf(&f, path) != 0) {
      fprintf(stderr, "[!] failed to load bloom filter: %s\n", todo);
      return 1;
    }
    if (f.size != size) {
      fprintf(stderr, "[!] bloom filter size mismatch (%'zu != %'zu): %s\n", (size_t)f.size, (size_t)size, todo);
      return 1;
    }
    printf("updating bloom filter...\n");
  } else {
    printf("creating bloom filter...\n");
    f.size = size;
    f.bits = calloc(size ? size : 1, sizeof(uint64_t));
    if (!f.bits) {
      fprintf(stderr, "[!] out of memory\n");
      return 1;
    }
  }
  printf("bloom filter params: n = %'llu | p = 1:%'llu | m = %'llu (%'.1f MB)\n", n, r, m, mb);

  unsigned long long added = 0;
  char piece[41]; /* same 40-character chunking as load_filter (main.c:96-98) */
  while (fgets(piece, sizeof piece, stdin)) {
    if (strlen(piece) != 40) continue;
    uint32_t h[5];
    words_from_hex40(h, piece);
    if (bloom_has(f.bits, f.size, h)) continue;
    bloom_add(f.bits, f.size, h);
    added++;
  }
  printf("added %'llu new items; saving to %s\n", added, path);
  if (bloom_save(path, f.bits, f.size) != 0) {
    fprintf(stderr, "[!] failed to save bloom filter\n");
    return 1;
  }
  filter_free(&f);
  return 0;
}

static int check_usage(const char *name) {
  printf("Usage: %s blf-check -f <file> <hash> [hash...]\n", name);
  printf("Check if one or more hex-encoded hash160 values are in the bloom filter.\n");
  printf("\nOptions:\n");
  printf("  -f <file>       Path to the bloom filter file (required).\n");
  printf("\nArguments:\n");
  printf("  <hash>          One or more hex-encoded hash160 values to check.\n");
  printf("                  If no arguments are provided, stdin will be used as source.\n");
  return 1;
}

int blf_check_main(int argc, const char **argv) {
  const char *path = value_of(argc, argv, "-f");
  if (!path) {
    fprintf(stderr, "[!] missing input file (-f <file>)\n");
    return check_usage(argv[0]);
  }
  ecl_filter f = {0};
  if (filter_load_blf(&f, path) != 0) {
    fprintf(stderr, "[!] failed to load bloom filter\n");
    return 1;
  }
  bool from_args = false;
  for (int i = 1; i < argc; ++i) {
    if (strlen(argv[i]) != 40) continue;
    from_args = true;
    uint32_t h[5];
    words_from_hex40(h, argv[i]);
    printf("%s %s\n", argv[i], bloom_has(f.bits, f.size, h) ? "FOUND" : "NOT FOUND");
  }
  if (!from_args) {
    char line[128];
    while (fgets(line, sizeof line, stdin)) {
      char *s = line; /* strtrim (lib/utils.c:57-70) */
      while (isspace((unsigned char)*s)) ++s;
      size_t len = strlen(s);
      while (len > 0 && isspace((unsigned char)s[len - 1])) s[--len] = 0;
      if (len != 40) continue;
      uint32_t h[5];
      words_from_hex40(h, s);
      printf("%s %s\n", s, bloom_has(f.bits, f.size, h) ? "FOUND" : "NOT FOUND");
    }
  }
  filter_free(&f);
  return 0;
}
