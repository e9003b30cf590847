/* host_sanitize.c — the C host's bookkeeping under AddressSanitizer + UBSan ("run host glue under
 * -fsanitize=address,undefined in CPU-only tests"). Built and run by tests/test_cli_host.py; exits 0 when every check
 * holds and the sanitizers stayed silent. Usage: host_sanitize <hash-list> <blf-out> */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "filter.h"
#include "jobplan.h"
#include "mulfeed.h"
#include "sha256_host.h"
#include "u256.h"

#define CHECK(c)                                                          \
  do {                                                                    \
    if (!(c)) {                                                           \
      fprintf(stderr, "check failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); \
      return 1;                                                           \
    }                                                                     \
  } while (0)

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  /* scalars */
  u256 a, b, r;
  modn_from_hex(a, "fffffffffffffffffffffffffffffffebaaedce6af48a03bbfd25e8cd0364140"); /* n - 1 */
  modn_mul(r, a, a);
  CHECK(r[0] == 1 && !r[1] && !r[2] && !r[3]); /* (n-1)^2 = 1 */
  modn_neg(b, a);
  CHECK(b[0] == 1 && !b[1]);
  modn_add_stride(r, a, b, 5); /* (n-1) + 5*1 = 4 mod n ... with the carry convention: no 2^256 overflow -> stays n+4 */
  u256_from_hex(a, "");
  CHECK(u256_is_zero(a));
  char longhex[200];
  memset(longhex, 'f', sizeof longhex - 1);
  longhex[sizeof longhex - 1] = 0;
  u256_from_hex(a, longhex); /* digits beyond 64 are dropped, nothing is written past the array */
  CHECK(a[3] == ~0ULL && u256_bitlen(a) == 256);

  /* filter: list mode, exact lookup, save / load round trip */
  ecl_filter f;
  CHECK(filter_load(&f, argv[1]) == 0);
  CHECK(f.list && f.count > 0 && f.size == 2 * f.count);
  CHECK(filter_exact(&f, f.list[0].w) && bloom_has(f.bits, f.size, f.list[f.count - 1].w));
  CHECK(bloom_save(argv[2], f.bits, f.size) == 0);
  ecl_filter g;
  CHECK(filter_load_blf(&g, argv[2]) == 0);
  CHECK(!g.list && g.size == f.size && memcmp(g.bits, f.bits, f.size * 8) == 0);
  filter_free(&g);
  CHECK(filter_load(&g, argv[2]) == 0); /* name ends in .blf: header only, words streamed in chunks */
  CHECK(!g.list && !g.bits && g.size == f.size && g.blf_fd >= 0);
  {
    uint64_t *chunk = malloc(f.size * 8);
    uint64_t have = 0;
    while (have < g.size) {
      const int64_t n = filter_stream_blf(&g, chunk + have, 5, have);
      CHECK(n > 0);
      have += (uint64_t)n;
    }
    CHECK(memcmp(chunk, f.bits, f.size * 8) == 0);
    free(chunk);
  }
  filter_free(&g);
  filter_free(&f);
  CHECK(filter_load(&f, "/nonexistent/x") == -1);

  /* mul feeder: ragged text, long lines, no trailing newline, -raw */
  const char *text = "1\n\r\nc936\r\n  zz  \n0x10\n";
  size_t len = strlen(text);
  char *buf = malloc(len + 4000 + 2);
  memcpy(buf, text, len);
  memset(buf + len, '7', 3000); /* one 3000-character line without newline: pieces of 1024 */
  len += 3000;
  uint64_t(*keys)[4] = NULL;
  uint32_t cap = 0;
  uint32_t n = mulfeed_parse(buf, len, false, &keys, &cap);
  CHECK(n == 4 + 3 && keys[0][0] == 1 && keys[1][0] == 0xc936 && keys[2][0] == 0 && keys[3][0] == 0x10);
  n = mulfeed_parse(buf, len, true, &keys, &cap);
  CHECK(n == 7);
  CHECK(mulfeed_cut(buf, len) == strlen(text));
  free(keys);
  free(buf);
  uint32_t d[8];
  sha256_bytes(d, (const uint8_t *)"abc", 3);
  CHECK(d[0] == 0xba7816bfu && d[7] == 0xf20015adu);

  /* job plan: ragged single job, fused spans, wrap guard */
  job_plan jp;
  u256 rs, re, start;
  u256_set64(rs, 0x8000);
  u256_set64(re, 0xffff);
  jobplan_init(&jp, rs, re, 0, false);
  jobplan_choose_span(&jp, 2048, 8);
  CHECK(jp.job_keys == 0x7fff && jp.visit_keys == 0x8000 && jobplan_take(&jp, start) == 1 && jobplan_take(&jp, start) == 0);
  modn_from_hex(rs, "400000000000000000");
  modn_from_hex(re, "40000000ffffffffff");
  jobplan_init(&jp, rs, re, 0, false);
  jobplan_choose_span(&jp, 2048, 8);
  uint64_t jobs = 0, spans = 0, k;
  while ((k = jobplan_take(&jp, start)) != 0) jobs += k, spans++;
  CHECK(jobs == (1ull << 19) && spans == 256);
  return 0;
}
