/* mulfeed.h — text side of `ecloop mul`: stdin bytes -> private keys, with the reference's line rules.
 *
 * The reference reads with fgets(line, 1025, stdin), drops one trailing '\n' and one trailing '\r', skips empty
 * lines (main.c:552-556) and parses each line with fe_modn_from_hex, or SHA-256 of the line with `-raw`
 * (main.c:503-527). These two functions do the same over a block of text so that parsing can run on several
 * threads (ecloop.c, cmd_mul). */
#ifndef ECL_MULFEED_H
#define ECL_MULFEED_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#define MULFEED_LINE_MAX 1024 /* characters fgets(buf, MAX_LINE_SIZE = 1025) returns at most (main.c:18) */

/* Parse `len` bytes of whole lines into *keys (grown with realloc, *cap entries); returns the number of keys.
 * text must have one writable spare byte after text[len - 1]. */
uint32_t mulfeed_parse(char *text, size_t len, bool raw, uint64_t (**keys)[4], uint32_t *cap);
/* Where to cut a block of `have` bytes so that only whole lines go to the parser; the rest is carried over. */
size_t mulfeed_cut(const char *text, size_t have);
#endif
