/* blftool.h — `ecloop blf-gen` / `ecloop blf-check`: the reference's offline bloom-filter tools (lib/utils.c:400-529)
 * for the drop-in binary. Host-only like in the reference (they run once, before a search); same sizing formula,
 * same file bytes, same messages. */
#ifndef ECL_BLFTOOL_H
#define ECL_BLFTOOL_H
int blf_gen_main(int argc, const char **argv);   /* returns the process exit code */
int blf_check_main(int argc, const char **argv);
#endif
