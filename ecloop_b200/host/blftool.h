/* blftool.h — `ecloop blf-gen` / `ecloop blf-check`: the reference's offline bloom-filter tools (lib/utils.c:400-529)
 * for the drop-in binary: same sizing formula, same file bytes, same messages.
 *   blf_gen_main      host-only insert loop, like the reference (`blf-gen -cpu`, and the unit tests);
 *   blf_gen_gpu_main  the insert loop on the GPU (blfgpu.c): the filter lives in HBM, hashes are parsed
 *                     on the host in blocks and handed to ecl_filter_add, which keeps the reference's
 *                     `if (blf_has) continue; blf_add; count++` semantics exactly (same "added N new items");
 *   blf_check_main    host-only, like the reference. */
#ifndef ECL_BLFTOOL_H
#define ECL_BLFTOOL_H
#include <stdint.h>

typedef struct blf_gen_plan { /* lib/utils.c:421-427 */
  unsigned long long n, r, m;
  double mb;
  uint64_t size; /* words */
  const char *path;
} blf_gen_plan;

int blf_gen_args(int argc, const char **argv, blf_gen_plan *out); /* 0, or the exit code after printing the usage */
void blf_hex40_words(uint32_t h[5], const char *s);               /* 40 hex characters read like 5 x sscanf("%8x") */
int blf_gen_main(int argc, const char **argv);                    /* returns the process exit code */
int blf_gen_gpu_main(int argc, const char **argv);
int blf_check_main(int argc, const char **argv);
#endif
