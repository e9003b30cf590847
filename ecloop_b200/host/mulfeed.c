/* mulfeed.c — see mulfeed.h */
#include "mulfeed.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sha256_host.h"
#include "u256.h"

static void line_to_key(uint64_t key[4], char *line, size_t len, bool raw) {
  if (!raw) { /* main.c:504 */
    modn_from_hex_n(key, line, len);
    return;
  }
  uint32_t d[8]; /* -raw: key = SHA-256(line) read as a big-endian 256-bit number, not reduced (main.c:506-527) */
  sha256_bytes(d, (const uint8_t *)line, len);
  for (int i = 0; i < 4; ++i) key[i] = (uint64_t)d[6 - 2 * i] << 32 | d[7 - 2 * i];
}

uint32_t mulfeed_parse(char *text, size_t len, bool raw, uint64_t (**keys)[4], uint32_t *cap) {
  uint32_t count = 0;
  size_t pos = 0;
  while (pos < len) {
    /* the piece fgets would return: up to and including '\n', or MULFEED_LINE_MAX characters */
    const size_t room = len - pos < MULFEED_LINE_MAX ? len - pos : MULFEED_LINE_MAX;
    const char *nl = memchr(text + pos, '\n', room);
    const size_t take = nl ? (size_t)(nl - (text + pos)) + 1 : room;
    char *line = text + pos;
    size_t n = take;
    pos += take;
    if (n && line[n - 1] == '\n') --n;
    if (n && line[n - 1] == '\r') --n;
    if (!n) continue;
    if (count == *cap) {
      *cap = *cap ? *cap * 2 : 1u << 17;
      *keys = realloc(*keys, (size_t)*cap * sizeof **keys);
      if (!*keys) {
        fprintf(stderr, "out of memory\n");
        exit(1);
      }
    }
    line_to_key((*keys)[count++], line, n, raw);
  }
  return count;
}

size_t mulfeed_cut(const char *text, size_t have) {
  size_t cut = have;
  while (cut > 0 && text[cut - 1] != '\n') --cut;
  /* no newline at all: hand on whole MULFEED_LINE_MAX pieces so that piece boundaries stay where fgets puts them */
  if (cut == 0) cut = have / MULFEED_LINE_MAX * MULFEED_LINE_MAX;
  return cut;
}
