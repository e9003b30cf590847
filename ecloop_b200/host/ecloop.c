/* ecloop.c — the `ecloop add|mul|rnd` command line on top of libecloop_b200.so (the B200 hot path).
 *
 * Drop-in for the reference binary's process contract: same argv grammar
 * (main.c:750-865), same stdin protocol for `mul` (main.c:552-556), same banner, found lines, `-o` file lines,
 * status line and error messages, same keys visited per `-r`/`-d` (Appendix A.1/A.2). What changed is who does the
 * work: the reference starts `-t` CPU worker threads that each run batch_add / ec_gtable_mul on their jobs
 * (main.c:405-435, 486-540); here one host thread per GPU ("rank thread") takes spans of consecutive jobs from
 * the same dispenser and hands them to ecl_add_submit / ecl_mul_submit, then does what the reference's worker
 * does after the bloom decision: exact list lookup, calc_priv, verification, ctx_write_found.
 *
 * There is no CPU compute path: without a usable GPU the program exits with an error.
 * Extra knobs (do not exist in the reference): `-gpus N` / env ECLOOP_GPUS (default: all visible devices),
 * env ECLOOP_SPAN_JOBS (jobs of 2^21 keys fused into one submit).
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <locale.h>
#include <pthread.h>
#include <signal.h>
#include <stdatomic.h>
#include <stdarg.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/select.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <termios.h>
#include <unistd.h>

#include "../../include/ecloop_b200.h"
#include "blftool.h"
#include "filter.h"
#include "jobplan.h"
#include "mulfeed.h"
#include "sha256_host.h"
#include "u256.h"

#define ECLOOP_VERSION "0.5.0"           /* the reference version this CLI mirrors (main.c:15) */
#define MUL_BATCH_KEYS (1u << 22)       /* keys per ecl_mul_submit: 56 keys per GPU thread share one inversion */

enum command { CMD_NONE, CMD_ADD, CMD_MUL, CMD_RND };

typedef struct app {
  enum command cmd;
  int argc;
  const char **argv;

  /* devices */
  int n_gpus;
  ecl_dev **dev;

  /* options */
  size_t threads_shown; /* `-t`: printed in the banner like the reference; compute runs on the GPUs */
  uint32_t flags;       /* ECL_A33 | ECL_A65 | ECL_ENDO */
  bool quiet, color, raw_text, has_seed;
  FILE *outfile;
  ecl_filter filter;

  /* search range (add, rnd) */
  u256 range_s, range_e, stride;
  unsigned ord_offs, ord_size;

  /* job dispenser + progress, all under `mu` */
  pthread_mutex_t mu;
  job_plan plan;
  uint64_t k_checked, k_found;
  uint64_t t_start, t_update, t_print, t_pause_at, paused_ms;
  volatile bool paused;
  bool finished;
  atomic_int fatal; /* a rank thread hit a library error (set from any thread, polled without the lock) */

  struct mul_pipe *mul; /* mul: stdin reader -> parser threads -> rank threads */
} app;

static uint64_t now_us(void) {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return (uint64_t)tv.tv_sec * 1000000 + (uint64_t)tv.tv_usec;
}

static uint64_t now_ms(void) {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return (uint64_t)tv.tv_sec * 1000 + (uint64_t)tv.tv_usec / 1000;
}

static void clear_status_line(void) {
  fputs("\033[2K\r", stderr);
  fflush(stderr);
}

static void die(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
  exit(1);
}

/* ------------------------------------------------------------------ argv helpers (lib/utils.c:162-185) */

static bool opt_flag(const app *a, const char *name) {
  for (int i = 1; i < a->argc; ++i)
    if (strcmp(a->argv[i], name) == 0) return true;
  return false;
}

static const char *opt_value(const app *a, const char *name) {
  for (int i = 1; i + 1 < a->argc; ++i)
    if (strcmp(a->argv[i], name) == 0) return a->argv[i + 1];
  return NULL;
}

/* ------------------------------------------------------------------ status line + found output */

static void print_status_locked(app *a) { /* ctx_print_unlocked (main.c:134-144) */
  const char *hint = a->finished ? "" : (a->paused ? " ('r' – resume)" : " ('p' – pause)");
  int64_t eff = (int64_t)(a->t_update - a->t_start) - (int64_t)a->paused_ms;
  if (eff < 1) eff = 1;
  const double secs = (double)eff / 1000.0;
  const double rate = (double)a->k_checked / secs / 1e6;
  clear_status_line();
  fprintf(stderr, "%.2fs ~ %.2f Mkeys/s ~ %'zu / %'zu%s%c", secs, rate, (size_t)a->k_found, (size_t)a->k_checked, hint,
          a->finished ? '\n' : '\r');
  fflush(stderr);
}

static void progress_add(app *a, uint64_t keys) { /* ctx_update (main.c:158-172) */
  const uint64_t ts = now_ms();
  pthread_mutex_lock(&a->mu);
  a->k_checked += keys;
  a->t_update = ts;
  if (ts - a->t_print >= 100) {
    a->t_print = ts;
    print_status_locked(a);
  }
  pthread_mutex_unlock(&a->mu);
  while (a->paused) usleep(100000);
}

static void finish(app *a) { /* ctx_finish (main.c:174-180) */
  pthread_mutex_lock(&a->mu);
  a->finished = true;
  print_status_locked(a);
  if (a->outfile) fclose(a->outfile), a->outfile = NULL;
  pthread_mutex_unlock(&a->mu);
}

static void write_found(app *a, int kind, const uint32_t h[5], const u256 pk) { /* ctx_write_found (main.c:182-203) */
  const char *label = kind == 0 ? "addr33" : "addr65";
  pthread_mutex_lock(&a->mu);
  if (!a->quiet) {
    clear_status_line();
    printf("%s: %08x%08x%08x%08x%08x <- %016llx%016llx%016llx%016llx\n", label, h[0], h[1], h[2], h[3], h[4],
           (unsigned long long)pk[3], (unsigned long long)pk[2], (unsigned long long)pk[1], (unsigned long long)pk[0]);
  }
  if (a->outfile) {
    fprintf(a->outfile, "%s\t%08x%08x%08x%08x%08x\t%016llx%016llx%016llx%016llx\n", label, h[0], h[1], h[2], h[3], h[4],
            (unsigned long long)pk[3], (unsigned long long)pk[2], (unsigned long long)pk[1], (unsigned long long)pk[0]);
    fflush(a->outfile);
  }
  a->k_found += 1;
  const uint64_t ts = now_ms(); /* the reference redraws the status after every hit; dense filters make that the
                                   bottleneck, so the redraw (stderr only) is limited to the usual 10 Hz */
  if (ts - a->t_print >= 100) {
    a->t_print = ts;
    print_status_locked(a);
  }
  pthread_mutex_unlock(&a->mu);
}

/* ------------------------------------------------------------------ hits -> private keys (main.c:248-285) */

static void recover_key(u256 pk, const u256 start, const u256 stride, uint64_t off, unsigned endo) { /* calc_priv */
  modn_add_stride(pk, start, stride, off);
  if (endo == 2 || endo == 3) modn_mul(pk, pk, SECP_LAMBDA);
  if (endo == 4 || endo == 5) modn_mul(pk, pk, SECP_LAMBDA2);
  if (endo == 1 || endo == 3 || endo == 5) modn_neg(pk, pk);
}

typedef struct hit_buf {
  ecl_hit *hits;
  uint32_t cap;
  /* verification scratch */
  uint64_t (*pks)[4];
  uint64_t (*xy)[8];
  uint32_t (*h33)[5], (*h65)[5];
  uint32_t vcap;
} hit_buf;

static int collect_hits(app *a, ecl_dev *dev, hit_buf *hb, uint32_t *n) {
  for (;;) {
    if (!hb->hits) {
      hb->cap = hb->cap ? hb->cap : 4096;
      hb->hits = malloc((size_t)hb->cap * sizeof(ecl_hit));
      if (!hb->hits) die("out of memory");
    }
    uint64_t done = 0;
    const int rc = ecl_collect(dev, hb->hits, hb->cap, n, &done);
    if (rc == ECL_OK) return 0;
    if (rc != ECL_E_OVERFLOW || hb->cap >= (1u << 30)) {
      fprintf(stderr, "ecloop: GPU error: %s\n", ecl_last_error(dev));
      a->fatal = 1;
      return -1;
    }
    free(hb->hits); /* caller buffer too small: the library keeps the result, ask again with more room */
    hb->hits = NULL;
    hb->cap *= 8;
  }
}

static void ensure_verify_scratch(hit_buf *hb, uint32_t n) {
  if (n <= hb->vcap) return;
  free(hb->pks), free(hb->xy), free(hb->h33), free(hb->h65);
  hb->vcap = n < 1024 ? 1024 : n;
  hb->pks = malloc((size_t)hb->vcap * sizeof *hb->pks);
  hb->xy = malloc((size_t)hb->vcap * sizeof *hb->xy);
  hb->h33 = malloc((size_t)hb->vcap * sizeof *hb->h33);
  hb->h65 = malloc((size_t)hb->vcap * sizeof *hb->h65);
  if (!hb->pks || !hb->xy || !hb->h33 || !hb->h65) die("out of memory");
}

/* One span of the add path after the GPU returned its bloom-positive keys: exact list stage, key recovery,
 * pk_verify_hash (recomputed from the recovered key, on the GPU through the primitive entry points; a mismatch
 * is fatal like in the reference, main.c:255-262), then the found lines in the reference's `-t 1` order. */
static int report_add_hits(app *a, ecl_dev *dev, hit_buf *hb, uint32_t n, const u256 start) {
  uint32_t m = 0;
  for (uint32_t i = 0; i < n; ++i)
    if (filter_exact(&a->filter, hb->hits[i].h160)) hb->hits[m++] = hb->hits[i];
  if (!m) return 0;
  ensure_verify_scratch(hb, m);
  for (uint32_t i = 0; i < m; ++i) recover_key(hb->pks[i], start, a->stride, hb->hits[i].key_off, hb->hits[i].endo);
  if (ecl_prim_scalar_mul(dev, (const uint64_t(*)[4])hb->pks, hb->xy, m) != ECL_OK ||
      ecl_prim_hash160(dev, (const uint64_t(*)[8])hb->xy, hb->h33, hb->h65, m) != ECL_OK) {
    fprintf(stderr, "ecloop: GPU error: %s\n", ecl_last_error(dev));
    a->fatal = 1;
    return -1;
  }
  for (uint32_t i = 0; i < m; ++i) {
    const ecl_hit *h = &hb->hits[i];
    uint32_t *again = h->kind == 0 ? hb->h33[i] : hb->h65[i];
    if (getenv("ECLOOP_TEST_CORRUPT_VERIFY")) again[4] ^= 1u; /* test hook: drive the mismatch branch (main.c:255-262) */
    if (memcmp(again, h->h160, 20) != 0) {
      const uint64_t *pk = hb->pks[i];
      fprintf(stderr, "[!] error: hash mismatch (compressed: %d endo: %zu)\n", h->kind == 0, (size_t)h->endo);
      fprintf(stderr, "pk: %016llx%016llx%016llx%016llx\n", (unsigned long long)pk[3], (unsigned long long)pk[2],
              (unsigned long long)pk[1], (unsigned long long)pk[0]);
      fprintf(stderr, "lh: %08x%08x%08x%08x%08x\n", h->h160[0], h->h160[1], h->h160[2], h->h160[3], h->h160[4]);
      fprintf(stderr, "rh: %08x%08x%08x%08x%08x\n", again[0], again[1], again[2], again[3], again[4]);
      exit(1);
    }
    write_found(a, h->kind, h->h160, hb->pks[i]);
  }
  return 0;
}

/* ------------------------------------------------------------------ add / rnd: dispenser + rank threads */

/* next span of consecutive jobs from the shared plan (cmd_add_worker's critical section, main.c:419-428) */
static uint64_t take_span(app *a, u256 start) {
  pthread_mutex_lock(&a->mu);
  const uint64_t jobs = a->fatal ? 0 : jobplan_take(&a->plan, start);
  pthread_mutex_unlock(&a->mu);
  return jobs;
}

typedef struct rank_arg {
  app *a;
  int rank;
} rank_arg;

static void *add_rank_main(void *p) {
  app *a = ((rank_arg *)p)->a;
  ecl_dev *dev = a->dev[((rank_arg *)p)->rank];
  hit_buf hb = {0};
  const uint64_t visit = a->plan.visit_keys;
  const uint64_t per_key = (a->flags & ECL_ENDO) ? 6 : 1;
  u256 start;
  uint64_t jobs;
  while ((jobs = take_span(a, start)) != 0) {
    uint32_t n = 0;
    if (ecl_add_submit(dev, start, jobs * visit, a->flags) != ECL_OK) {
      fprintf(stderr, "ecloop: GPU error: %s\n", ecl_last_error(dev));
      a->fatal = 1;
      break;
    }
    if (collect_hits(a, dev, &hb, &n) != 0) break;
    if (report_add_hits(a, dev, &hb, n, start) != 0) break;
    progress_add(a, jobs * a->plan.job_keys * per_key); /* main.c:431 */
  }
  free(hb.hits), free(hb.pks), free(hb.xy), free(hb.h33), free(hb.h65);
  return NULL;
}

static void run_add_ranks(app *a, bool fixed_job) {
  pthread_t th[64];
  rank_arg arg[64];
  const char *env = getenv("ECLOOP_SPAN_JOBS"); /* default 2048 jobs = 2^32 keys per submit */
  jobplan_init(&a->plan, a->range_s, a->range_e, a->ord_offs, fixed_job);
  jobplan_choose_span(&a->plan, env ? strtoull(env, NULL, 10) : 2048, (unsigned)a->n_gpus);
  for (int r = 0; r < a->n_gpus; ++r) {
    arg[r].a = a, arg[r].rank = r;
    pthread_create(&th[r], NULL, add_rank_main, &arg[r]);
  }
  for (int r = 0; r < a->n_gpus; ++r) pthread_join(th[r], NULL);
  if (a->fatal) exit(1);
}

static void set_stride_all(app *a) { /* ctx_precompute_gpoints (main.c:219-246) now runs on each device */
  memset(a->stride, 0, sizeof(u256));
  a->stride[a->ord_offs / 64] = 1ULL << (a->ord_offs % 64);
  for (int r = 0; r < a->n_gpus; ++r)
    if (ecl_set_stride(a->dev[r], a->stride) != ECL_OK) die("ecloop: GPU error: %s", ecl_last_error(a->dev[r]));
}

static void cmd_add(app *a) { /* main.c:437-454 */
  set_stride_all(a);
  a->t_start = now_ms();
  run_add_ranks(a, false);
  finish(a);
}

/* ---- rnd (main.c:580-662) */

static uint64_t random64(app *a) { /* rand64 (lib/utils.c:83-106) */
  if (a->has_seed) return (uint64_t)rand() << 32 | (uint64_t)rand();
  static FILE *ur;
  if (!ur && !(ur = fopen("/dev/urandom", "rb"))) die("failed to open /dev/urandom");
  uint64_t v;
  if (fread(&v, sizeof v, 1, ur) != 1) die("failed to read from /dev/urandom");
  return v;
}

static void random_in_range(app *a, u256 out, const u256 lo, const u256 hi) { /* fe_rand_range (lib/utils.c:129-153) */
  u256 span, one = {1, 0, 0, 0}, x;
  modn_sub(span, hi, lo);
  u256_add_raw(span, span, one);
  const unsigned bits = u256_bitlen(span);
  do {
    for (int i = 0; i < 4; ++i) x[i] = random64(a);
    x[3] &= 0xfffffffefffffc2fULL; /* the reference masks the top limb like this (lib/utils.c:121,126) */
    const unsigned top = (bits - 1) / 64;
    for (unsigned i = top + 1; i < 4; ++i) x[i] = 0;
    if (bits % 64) x[top] &= (1ULL << (bits % 64)) - 1;
  } while (u256_cmp(x, span) >= 0);
  modn_add(out, x, lo);
}

static void print_window_line(const u256 v, unsigned size, unsigned offs, bool color) { /* print_range_mask (main.c:593-617) */
  const int hi = 255 - (int)offs, lo = hi - (int)size + 1;
  for (int i = 0; i < 64; ++i) {
    if (i && i % 16 == 0) putchar(' ');
    const int b0 = 4 * i, b1 = b0 + 3, bit = 255 - b1;
    const char c = "0123456789abcdef"[(v[bit / 64] >> (bit % 64)) & 0xF];
    const bool dyn = (b0 >= lo && b0 <= hi) || (b1 >= lo && b1 <= hi);
    if (dyn && color) fputs("\033[33m", stdout);
    putchar(c);
    if (dyn && color) fputs("\033[0m", stdout);
  }
  putchar('\n');
}

static void cmd_rnd(app *a) {
  if (a->ord_offs > 255 - a->ord_size) a->ord_offs = 255 - a->ord_size;
  printf("[RANDOM MODE] offs: %d ~ bits: %d\n\n", (int)a->ord_offs, (int)a->ord_size);
  set_stride_all(a);
  a->t_start = now_ms();
  u256 lo, hi;
  u256_copy(lo, a->range_s);
  u256_copy(hi, a->range_e);
  const char *env = getenv("ECLOOP_RND_WINDOWS"); /* test hook: stop after this many windows */
  uint64_t max_windows = env ? strtoull(env, NULL, 10) : 0, windows = 0;
  for (;;) {
    const uint64_t c0 = a->k_checked, f0 = a->k_found, t0 = now_ms();
    random_in_range(a, a->range_s, lo, hi);
    u256_copy(a->range_e, a->range_s);
    for (unsigned i = a->ord_offs; i < a->ord_offs + a->ord_size; ++i) {
      a->range_s[i / 64] &= ~(1ULL << (i % 64));
      a->range_e[i / 64] |= 1ULL << (i % 64);
    }
    if (u256_cmp(a->range_s, lo) <= 0) u256_copy(a->range_s, lo);
    if (u256_cmp(a->range_e, hi) >= 0) u256_copy(a->range_e, hi);
    print_window_line(a->range_s, a->ord_size, a->ord_offs, a->color);
    print_window_line(a->range_e, a->ord_size, a->ord_offs, a->color);
    fflush(stdout);
    pthread_mutex_lock(&a->mu);
    print_status_locked(a);
    pthread_mutex_unlock(&a->mu);
    const bool whole = u256_cmp(a->range_s, lo) == 0 && u256_cmp(a->range_e, hi) == 0;
    run_add_ranks(a, true);
    uint64_t dt = now_ms() - t0;
    if (dt < 1) dt = 1;
    clear_status_line();
    printf("%'zu / %'zu ~ %.1fs\n\n", (size_t)(a->k_found - f0), (size_t)(a->k_checked - c0), (double)dt / 1000.0);
    fflush(stdout);
    if (whole || (max_windows && ++windows >= max_windows)) break;
  }
  finish(a);
}

/* ------------------------------------------------------------------ mul (main.c:458-578) */

/* The reference reads stdin with fgets on one thread and parses on the workers (main.c:549-569, 503-527). At GPU
 * speed the text side is the bottleneck (10 M keys = 650 MB of hex), so the feeder is a three-stage pipeline:
 *   reader thread          read()s large blocks, cuts them at a line boundary, numbers them;
 *   parser threads (-t)    split a block into lines with the reference's rules and turn each into a key;
 *   rank threads (per GPU) take parsed blocks IN SEQUENCE ORDER, fuse them into submits of up to 2^22 keys and keep
 *                          TWO submits in flight (ECL_MUL_DEPTH): the upload and kernels of one batch run while the
 *                          previous batch's hits are reported and the next one is gathered.
 * Reader and parsers start BEFORE the GPUs are opened (setup), so context creation and the window-table build overlap
 * with reading and parsing. With one GPU the found lines come out in input order, like `-t 1`. */

#define MUL_BLOCK_BYTES (8u << 20)
#define MUL_RING 64 /* blocks in flight: up to 512 MB of text may be parsed ahead while the devices come up */

typedef struct mul_block {
  char *text;          /* raw bytes, whole lines (the last block may lack the final newline) */
  char *own;           /* the block's own buffer (pipes); text points into the mapped input when stdin is a file */
  size_t len;
  uint64_t (*keys)[4]; /* parsed keys */
  uint32_t count, cap;
  int state;           /* 0 free, 1 text ready, 2 being parsed, 3 parsed */
  uint64_t seq;
} mul_block;

typedef struct mul_pipe {
  app *a;
  mul_block ring[MUL_RING];
  uint64_t seq_read, seq_parse, seq_take; /* next sequence number to fill / to parse / to hand to a GPU */
  uint32_t take_off;                      /* keys of block seq_take already handed out */
  bool eof;
  pthread_cond_t cv;
  pthread_t reader_th, parse_th[64];
  int n_parsers;
  uint64_t us_read, us_parse, us_gpu, us_report; /* stage totals for ECLOOP_VERBOSE */
} mul_pipe;

static void *mul_parser_main(void *p) {
  mul_pipe *mp = p;
  app *a = mp->a;
  for (;;) {
    pthread_mutex_lock(&a->mu);
    while (!(mp->seq_parse < mp->seq_read) && !mp->eof && !a->fatal) pthread_cond_wait(&mp->cv, &a->mu);
    if (!(mp->seq_parse < mp->seq_read) || a->fatal) {
      pthread_mutex_unlock(&a->mu);
      return NULL;
    }
    mul_block *b = &mp->ring[mp->seq_parse++ % MUL_RING];
    b->state = 2;
    pthread_mutex_unlock(&a->mu);
    const uint64_t t0 = now_us();
    b->count = mulfeed_parse(b->text, b->len, a->raw_text, &b->keys, &b->cap);
    const uint64_t dt = now_us() - t0;
    pthread_mutex_lock(&a->mu);
    mp->us_parse += dt;
    b->state = 3;
    pthread_cond_broadcast(&mp->cv);
    pthread_mutex_unlock(&a->mu);
  }
}

static void mul_publish_block(mul_pipe *mp, mul_block *b, bool last) {
  app *a = mp->a;
  pthread_mutex_lock(&a->mu);
  b->seq = mp->seq_read, b->state = 1;
  mp->seq_read++;
  if (last) mp->eof = true;
  pthread_cond_broadcast(&mp->cv);
  pthread_mutex_unlock(&a->mu);
}

/* stdin is a regular file (`ecloop mul ... < keys.txt`): no copy at all — the file is mapped and the blocks are slices of
 * the mapping cut at line boundaries; the parser threads fault the pages in, in parallel. read() into private buffers
 * tops out near 2.4 GB/s on one thread (36 Mkeys/s of hex keys), below what parsers and GPU can take. */
static bool mul_reader_mapped(mul_pipe *mp) {
  app *a = mp->a;
  struct stat st;
  if (fstat(STDIN_FILENO, &st) != 0 || !S_ISREG(st.st_mode)) return false;
  const off_t at = lseek(STDIN_FILENO, 0, SEEK_CUR);
  if (at < 0 || at >= st.st_size) return false;
  const size_t pg = (size_t)sysconf(_SC_PAGESIZE), skip = (size_t)at % pg, len = (size_t)(st.st_size - at);
  char *base = mmap(NULL, len + skip, PROT_READ | PROT_WRITE, MAP_PRIVATE, STDIN_FILENO, at - (off_t)skip);
  if (base == MAP_FAILED) return false;
  madvise(base, len + skip, MADV_SEQUENTIAL);
  char *text = base + skip;
  const uint64_t tr = now_us();
  for (size_t pos = 0; pos < len && !a->fatal;) {
    pthread_mutex_lock(&a->mu);
    while (mp->seq_read - mp->seq_take >= MUL_RING && !a->fatal) pthread_cond_wait(&mp->cv, &a->mu);
    mul_block *b = &mp->ring[mp->seq_read % MUL_RING];
    pthread_mutex_unlock(&a->mu);
    if (a->fatal) break;
    size_t take = len - pos;
    const bool last = take <= MUL_BLOCK_BYTES;
    if (!last) take = mulfeed_cut(text + pos, MUL_BLOCK_BYTES);
    b->text = text + pos, b->len = take;
    pos += take;
    mul_publish_block(mp, b, last);
  }
  mp->us_read += now_us() - tr;
  return true; /* the mapping lives until exit: the rank threads still read keys parsed from it */
}

/* reader: blocks of whole lines; the tail after the last newline moves to the front of the next block */
static void *mul_reader_main(void *p) {
  mul_pipe *mp = p;
  app *a = mp->a;
  if (mul_reader_mapped(mp)) {
    pthread_mutex_lock(&a->mu);
    mp->eof = true;
    pthread_cond_broadcast(&mp->cv);
    pthread_mutex_unlock(&a->mu);
    return NULL;
  }
#ifdef F_SETPIPE_SZ
  fcntl(STDIN_FILENO, F_SETPIPE_SZ, 1 << 20); /* a pipe's default 64 KB costs a wake-up per 1000 keys; ignored for files */
#endif
  char *carry = malloc(MUL_BLOCK_BYTES + 2);
  size_t carry_len = 0;
  if (!carry) die("out of memory");
  for (bool more = true; more && !a->fatal;) {
    pthread_mutex_lock(&a->mu);
    while (mp->seq_read - mp->seq_take >= MUL_RING && !a->fatal) pthread_cond_wait(&mp->cv, &a->mu);
    mul_block *b = &mp->ring[mp->seq_read % MUL_RING];
    pthread_mutex_unlock(&a->mu);
    if (a->fatal) break;
    if (!b->own && !(b->own = malloc(MUL_BLOCK_BYTES + 2))) die("out of memory");
    b->text = b->own;
    memcpy(b->text, carry, carry_len);
    size_t have = carry_len;
    carry_len = 0;
    const uint64_t tr = now_us();
    while (have < MUL_BLOCK_BYTES) {
      const ssize_t got = read(STDIN_FILENO, b->text + have, MUL_BLOCK_BYTES - have);
      if (got < 0 && errno == EINTR) continue;
      if (got <= 0) { /* EOF (or error): this is the last block */
        more = false;
        break;
      }
      have += (size_t)got;
    }
    mp->us_read += now_us() - tr;
    size_t cut = have;
    if (more) { /* keep whole lines; a block without any newline is passed on as is (pieces of 1024 characters) */
      cut = mulfeed_cut(b->text, have);
      carry_len = have - cut;
      memcpy(carry, b->text + cut, carry_len);
    }
    b->len = cut;
    mul_publish_block(mp, b, !more);
  }
  pthread_mutex_lock(&a->mu);
  mp->eof = true;
  pthread_cond_broadcast(&mp->cv);
  pthread_mutex_unlock(&a->mu);
  free(carry);
  return NULL;
}

static void mul_start_feeder(app *a) { /* before open_devices: the text side runs while the GPUs come up */
  static mul_pipe pipe;
  mul_pipe *mp = &pipe;
  mp->a = a, a->mul = mp;
  pthread_cond_init(&mp->cv, NULL);
  mp->n_parsers = (int)(a->threads_shown < 32 ? a->threads_shown : 32);
  for (int i = 0; i < mp->n_parsers; ++i) pthread_create(&mp->parse_th[i], NULL, mul_parser_main, mp);
  pthread_create(&mp->reader_th, NULL, mul_reader_main, mp);
}

/* next batch for a GPU: parsed blocks in sequence order until `room` keys are gathered. may_wait = false returns 0
 * instead of sleeping when nothing is parsed yet (the caller has a submit in flight whose hits it can report first).
 * *drained is set when the input is exhausted. */
static uint32_t mul_gather(mul_pipe *mp, uint64_t (*keys)[4], uint32_t room, bool may_wait, bool *drained) {
  app *a = mp->a;
  uint32_t n = 0;
  pthread_mutex_lock(&a->mu);
  for (;;) {
    mul_block *b = &mp->ring[mp->seq_take % MUL_RING];
    const bool ready = mp->seq_take < mp->seq_read && b->state == 3;
    if (ready) { /* a block of very short lines can hold more keys than one submit: take it in parts */
      const uint32_t avail = b->count - mp->take_off, left = room - n, k = avail < left ? avail : left;
      memcpy(keys + n, b->keys + mp->take_off, (size_t)k * sizeof *keys);
      n += k, mp->take_off += k;
      if (mp->take_off == b->count) {
        b->state = 0, mp->take_off = 0;
        mp->seq_take++;
        pthread_cond_broadcast(&mp->cv);
      }
      if (n == room) break;
      continue;
    }
    if (n || a->fatal) break; /* something to do (or giving up) */
    if (mp->eof && mp->seq_take == mp->seq_read) {
      *drained = true;
      break;
    }
    if (!may_wait) break;
    pthread_cond_wait(&mp->cv, &a->mu);
  }
  pthread_mutex_unlock(&a->mu);
  return n;
}

static void *mul_rank_main(void *p) {
  mul_pipe *mp = ((rank_arg *)p)->a->mul;
  app *a = mp->a;
  ecl_dev *dev = a->dev[((rank_arg *)p)->rank];
  hit_buf hb = {0};
  struct {
    uint64_t (*keys)[4];
    uint32_t n;
    uint64_t t_submit;
  } bt[ECL_MUL_DEPTH];
  for (int i = 0; i < ECL_MUL_DEPTH; ++i)
    if (!(bt[i].keys = malloc((size_t)MUL_BATCH_KEYS * sizeof *bt[i].keys))) die("out of memory");
  int head = 0, inflight = 0;
  bool drained = false;
  while (!a->fatal) {
    if (!drained && inflight < ECL_MUL_DEPTH) {
      const int slot = (head + inflight) % ECL_MUL_DEPTH;
      bt[slot].n = mul_gather(mp, bt[slot].keys, MUL_BATCH_KEYS, inflight == 0, &drained);
      if (bt[slot].n) {
        bt[slot].t_submit = now_us();
        if (ecl_mul_submit(dev, (const uint64_t(*)[4])bt[slot].keys, bt[slot].n, a->flags & (ECL_A33 | ECL_A65)) != ECL_OK) {
          fprintf(stderr, "ecloop: GPU error: %s\n", ecl_last_error(dev));
          a->fatal = 1;
          break;
        }
        inflight++;
        if (inflight < ECL_MUL_DEPTH) continue; /* try to queue a second batch before waiting for the first */
      }
    }
    if (!inflight) {
      if (drained) break;
      continue;
    }
    uint32_t nh = 0; /* oldest submit: hits -> found lines (check_found_mul, main.c:458-479: no verification here) */
    if (collect_hits(a, dev, &hb, &nh) != 0) break;
    const uint64_t t1 = now_us();
    for (uint32_t i = 0; i < nh; ++i)
      if (filter_exact(&a->filter, hb.hits[i].h160)) write_found(a, hb.hits[i].kind, hb.hits[i].h160, bt[head].keys[hb.hits[i].key_off]);
    progress_add(a, bt[head].n);
    pthread_mutex_lock(&a->mu);
    mp->us_gpu += t1 - bt[head].t_submit, mp->us_report += now_us() - t1;
    pthread_mutex_unlock(&a->mu);
    head = (head + 1) % ECL_MUL_DEPTH, inflight--;
  }
  if (a->fatal) {
    pthread_mutex_lock(&a->mu);
    pthread_cond_broadcast(&mp->cv);
    pthread_mutex_unlock(&a->mu);
  }
  for (int i = 0; i < ECL_MUL_DEPTH; ++i) free(bt[i].keys);
  free(hb.hits), free(hb.pks), free(hb.xy), free(hb.h33), free(hb.h65);
  return NULL;
}

static void cmd_mul(app *a) {
  mul_pipe *mp = a->mul; /* reader and parsers have been running since setup() */
  pthread_t rank_th[64];
  rank_arg arg[64];
  for (int r = 0; r < a->n_gpus; ++r) /* staging buffers of both submit slots: page-locking them is device set-up */
    if (ecl_mul_reserve(a->dev[r], MUL_BATCH_KEYS) != ECL_OK) die("ecloop: GPU error: %s", ecl_last_error(a->dev[r]));
  a->t_start = now_ms(); /* the clock of the status line starts when the devices are ready, like cmd_add */
  for (int r = 0; r < a->n_gpus; ++r) {
    arg[r].a = a, arg[r].rank = r;
    pthread_create(&rank_th[r], NULL, mul_rank_main, &arg[r]);
  }
  pthread_join(mp->reader_th, NULL);
  for (int i = 0; i < mp->n_parsers; ++i) pthread_join(mp->parse_th[i], NULL);
  for (int r = 0; r < a->n_gpus; ++r) pthread_join(rank_th[r], NULL);
  if (a->fatal) exit(1);
  if (getenv("ECLOOP_VERBOSE"))
    fprintf(stderr, "\nmul stages: read %.3f s, parse %.3f s (sum over %d threads), gpu submit..collect %.3f s (overlapping), report %.3f s\n",
            mp->us_read / 1e6, mp->us_parse / 1e6, mp->n_parsers, mp->us_gpu / 1e6, mp->us_report / 1e6);
  finish(a);
}

/* ------------------------------------------------------------------ option parsing (main.c:666-865) */

static void parse_range(app *a) { /* arg_search_range */
  const char *raw = opt_value(a, "-r");
  if (!raw) {
    u256_set64(a->range_s, ECL_GROUP);
    u256_copy(a->range_e, SECP_P);
    return;
  }
  char *copy = strdup(raw), *sep = copy ? strchr(copy, ':') : NULL;
  if (!sep) die("invalid search range, use format: -r 8000:ffff");
  *sep = 0;
  modn_from_hex(a->range_s, copy);
  modn_from_hex(a->range_e, sep + 1);
  free(copy);
  const bool tiny = !a->range_s[1] && !a->range_s[2] && !a->range_s[3] && a->range_s[0] <= ECL_GROUP;
  if (tiny) die("invalid search range, start <= %#lx", (unsigned long)ECL_GROUP);
  if (u256_cmp(a->range_e, SECP_P) > 0) die("invalid search range, end > FE_P");
  if (u256_cmp(a->range_s, a->range_e) >= 0) die("invalid search range, start >= end");
}

static void parse_offs_size(app *a) { /* load_offs_size, SURVEY A.2 */
  const unsigned min_size = 20, max_size = 64;
  const unsigned range_bits = u256_bitlen(a->range_e);
  const unsigned floor_bits = range_bits > min_size ? range_bits : min_size;
  const unsigned default_bits = range_bits < 32 ? floor_bits : 32;
  const unsigned max_offs = floor_bits - default_bits > 1 ? floor_bits - default_bits : 1;
  const char *raw = opt_value(a, "-d");
  if (!raw) {
    a->ord_offs = a->cmd == CMD_RND ? (unsigned)(random64(a) % max_offs) : 0;
    a->ord_size = default_bits;
    return;
  }
  const char *sep = strchr(raw, ':');
  if (!sep) die("invalid offset:size format, use format: -d 128:32");
  const unsigned offs = (unsigned)atoi(raw), size = (unsigned)atoi(sep + 1);
  if (offs > 255) die("invalid offset, max is 255");
  if (size < min_size || size > max_size) die("invalid size, min is %d and max is %d", (int)min_size, (int)max_size);
  a->ord_offs = offs < max_offs ? offs : max_offs;
  a->ord_size = size;
}

static void usage(const char *name) { /* main.c:750-772; the reference's self-benchmark commands are not part of this build */
  printf("Usage: %s <cmd> [-t <threads>] [-f <file>] [-a <addr_type>] [-r <range>]\n", name);
  printf("v%s ~ https://github.com/vladkens/ecloop\n", ECLOOP_VERSION);
  printf("\nCompute commands:\n");
  printf("  add             - search in given range with batch addition\n");
  printf("  mul             - search hex encoded private keys (from stdin)\n");
  printf("  rnd             - search random range of bits in given range\n");
  printf("\nCompute options:\n");
  printf("  -f <file>       - filter file to search (list of hashes or bloom fitler)\n");
  printf("  -o <file>       - output file to write found keys (default: stdout)\n");
  printf("  -t <threads>    - number of threads to run (default: 1)\n");
  printf("  -a <addr_type>  - address type to search: c - addr33, u - addr65 (default: c)\n");
  printf("  -r <range>      - search range in hex format (example: 8000:ffff, default all)\n");
  printf("  -d <offs:size>  - bit offset and size for search (example: 128:32, default: 0:32)\n");
  printf("  -q              - quiet mode (no output to stdout; -o required)\n");
  printf("  -endo           - use endomorphism (default: false)\n");
  printf("\nOther commands:\n");
  printf("  blf-gen         - create bloom filter from list of hex-encoded hash160\n");
  printf("  blf-check       - check bloom filter for given hex-encoded hash160\n");
  printf("\nB200 build:\n");
  printf("  -gpus <n>       - number of GPUs to use (default: all visible; env ECLOOP_GPUS)\n");
  printf("\n");
}

static uint32_t seed_from_text(const char *s) { /* encode_seed (lib/utils.c:108-116) */
  uint32_t h = 0;
  while (*s) h = (h << 5) - h + (unsigned char)*s++;
  return h;
}

typedef struct open_arg {
  app *a;
  int rank;
  bool failed;
  char msg[512];
} open_arg;

static void *open_one_device(void *p) {
  open_arg *o = p;
  if (ecl_open(&o->a->dev[o->rank], o->rank) != ECL_OK) {
    o->failed = true; /* ecl_last_error(NULL) is one buffer for all threads: good enough for a message */
    snprintf(o->msg, sizeof o->msg, "%s", ecl_last_error(NULL));
  }
  return NULL;
}

/* The `-f` filter goes to every GPU. List mode: a few KB, one ecl_set_filter each. `.blf` (blf_load, lib/utils.c:362-396):
 * the file is never held in host memory; it is read in 64 MiB pieces into two pinned staging buffers and every piece
 * is sent to all GPUs at once (asynchronous copies on each GPU's own PCIe link), so a multi-GB filter is resident
 * everywhere after ONE pass over the file. ECLOOP_BLF_PEER=1 sends it to GPU 0 only and fans out over NVLink. */
#define BLF_CHUNK_WORDS (8u << 20)
static void load_filter_on_devices(app *a) {
  ecl_filter *f = &a->filter;
  if (f->bits) {
    for (int r = 0; r < a->n_gpus; ++r)
      if (ecl_set_filter(a->dev[r], f->bits, f->size) != ECL_OK) die("ecloop: GPU error: %s", ecl_last_error(a->dev[r]));
    return;
  }
  const uint64_t t0 = now_us();
  const bool peer = getenv("ECLOOP_BLF_PEER") != NULL && a->n_gpus > 1;
  const int direct = peer ? 1 : a->n_gpus;
  uint64_t *buf[2] = {ecl_host_alloc((uint64_t)BLF_CHUNK_WORDS * 8), ecl_host_alloc((uint64_t)BLF_CHUNK_WORDS * 8)};
  if (!buf[0] || !buf[1]) die("ecloop: cannot allocate pinned staging memory");
  for (int r = 0; r < direct; ++r)
    if (ecl_filter_alloc(a->dev[r], f->size) != ECL_OK) die("ecloop: GPU error: %s", ecl_last_error(a->dev[r]));
  uint64_t have = 0;
  for (int c = 0; have < f->size; ++c) {
    uint64_t *b = buf[c & 1];
    if (c >= 2) /* the copies that read this buffer two pieces ago must have landed */
      for (int r = 0; r < direct; ++r)
        if (ecl_filter_flush(a->dev[r]) != ECL_OK) die("ecloop: GPU error: %s", ecl_last_error(a->dev[r]));
    const int64_t n = filter_stream_blf(f, b, BLF_CHUNK_WORDS, have);
    if (n <= 0) exit(1);
    for (int r = 0; r < direct; ++r)
      if (ecl_filter_write(a->dev[r], have, b, (uint64_t)n) != ECL_OK) die("ecloop: GPU error: %s", ecl_last_error(a->dev[r]));
    have += (uint64_t)n;
  }
  for (int r = 0; r < direct; ++r)
    if (ecl_filter_commit(a->dev[r]) != ECL_OK) die("ecloop: GPU error: %s", ecl_last_error(a->dev[r]));
  for (int r = direct; r < a->n_gpus; ++r)
    if (ecl_filter_copy_peer(a->dev[r], a->dev[0]) != ECL_OK) die("ecloop: GPU error: %s", ecl_last_error(a->dev[r]));
  ecl_host_free(buf[0]), ecl_host_free(buf[1]);
  if (getenv("ECLOOP_VERBOSE"))
    fprintf(stderr, "filter: %.2f GiB resident on %d GPU(s) in %.2f s (%s)\n", (double)f->size * 8 / (1u << 30), a->n_gpus,
            (double)(now_us() - t0) / 1e6, peer ? "GPU 0 + NVLink peer copies" : "one pass, all GPUs");
}

static void open_devices(app *a) {
  const int have = ecl_device_count();
  if (have <= 0) die("ecloop: no CUDA device found; this build has no CPU compute path (%s)", ecl_last_error(NULL));
  int want = have;
  const char *g = opt_value(a, "-gpus");
  if (!g) g = getenv("ECLOOP_GPUS");
  if (g) want = atoi(g);
  if (want < 1) want = 1;
  if (want > have) want = have;
  if (want > 64) want = 64;
  a->dev = calloc((size_t)want, sizeof *a->dev);
  a->n_gpus = want;
  pthread_t th[64];
  open_arg arg[64];
  for (int r = 0; r < want; ++r) { /* contexts, window tables: every GPU at the same time */
    arg[r].a = a, arg[r].rank = r, arg[r].failed = false;
    pthread_create(&th[r], NULL, open_one_device, &arg[r]);
  }
  for (int r = 0; r < want; ++r) pthread_join(th[r], NULL);
  for (int r = 0; r < want; ++r)
    if (arg[r].failed) die("ecloop: cannot open GPU %d: %s", r, arg[r].msg);
  load_filter_on_devices(a);
}

static void setup(app *a) { /* init (main.c:774-865) */
  if (a->argc > 1) { /* the offline bloom tools come first, like main.c:776-778 */
    if (!strcmp(a->argv[1], "blf-gen")) /* on the GPU when there is one (blfgpu.c); `-cpu` keeps the host loop */
      exit(!opt_flag(a, "-cpu") && ecl_device_count() > 0 ? blf_gen_gpu_main(a->argc, a->argv) : blf_gen_main(a->argc, a->argv));
    if (!strcmp(a->argv[1], "blf-check")) exit(blf_check_main(a->argc, a->argv));
  }
  a->color = isatty(fileno(stdout));
  a->cmd = CMD_NONE;
  if (a->argc > 1) {
    if (!strcmp(a->argv[1], "add")) a->cmd = CMD_ADD;
    if (!strcmp(a->argv[1], "mul")) a->cmd = CMD_MUL;
    if (!strcmp(a->argv[1], "rnd")) a->cmd = CMD_RND;
  }
  if (a->cmd == CMD_NONE) {
    if (opt_flag(a, "-v")) printf("ecloop v%s\n", ECLOOP_VERSION);
    else usage(a->argv[0]);
    exit(0);
  }
  const char *seed = opt_value(a, "-seed");
  if (seed) { /* the reference frees an argv pointer here and aborts on glibc; we just seed */
    a->has_seed = true;
    srand(seed_from_text(seed));
  }
  if (filter_load(&a->filter, opt_value(a, "-f")) != 0) exit(1);

  a->quiet = opt_flag(a, "-q");
  const char *out = opt_value(a, "-o");
  if (out) a->outfile = fopen(out, "a");
  if (!out && a->quiet) die("quiet mode chosen without output file");

  const char *addr = opt_value(a, "-a");
  if (addr && strchr(addr, 'c')) a->flags |= ECL_A33;
  if (addr && strchr(addr, 'u')) a->flags |= ECL_A65;
  if (!(a->flags & (ECL_A33 | ECL_A65))) a->flags |= ECL_A33;
  if (opt_flag(a, "-endo") && a->cmd != CMD_MUL) a->flags |= ECL_ENDO;

  pthread_mutex_init(&a->mu, NULL);
  long cpus = sysconf(_SC_NPROCESSORS_ONLN);
  if (cpus < 1) cpus = 1;
  const char *t = opt_value(a, "-t");
  unsigned long long threads = t ? strtoull(t, NULL, 10) : (unsigned long long)cpus;
  if (threads < 1) threads = 1;
  if (threads > 320) threads = 320;
  a->threads_shown = (size_t)threads;
  a->t_start = a->t_update = now_ms();
  a->t_print = a->t_start - 5000;

  parse_range(a);
  parse_offs_size(a);
  if (a->cmd == CMD_MUL) {
    a->raw_text = opt_flag(a, "-raw");
    mul_start_feeder(a);
  }
  open_devices(a);

  printf("threads: %zu ~ addr33: %d ~ addr65: %d ~ endo: %d | filter: ", a->threads_shown, (a->flags & ECL_A33) != 0,
         (a->flags & ECL_A65) != 0, (a->flags & ECL_ENDO) != 0);
  if (a->filter.list) printf("list (%'zu)\n", a->filter.count);
  else printf("bloom\n");
  if (a->cmd == CMD_ADD) {
    printf("range_s: %016llx %016llx %016llx %016llx\n", (unsigned long long)a->range_s[3], (unsigned long long)a->range_s[2],
           (unsigned long long)a->range_s[1], (unsigned long long)a->range_s[0]);
    printf("range_e: %016llx %016llx %016llx %016llx\n", (unsigned long long)a->range_e[3], (unsigned long long)a->range_e[2],
           (unsigned long long)a->range_e[1], (unsigned long long)a->range_e[0]);
  }
  printf("----------------------------------------\n");
  fflush(stdout);
  if (getenv("ECLOOP_VERBOSE")) fprintf(stderr, "ecloop_b200: %d GPU(s), ABI %d\n", a->n_gpus, ecl_abi_version());
}

/* ------------------------------------------------------------------ pause / resume on the controlling tty */

static struct termios g_tty_saved;
static int g_tty_fd = -1;
static app *g_app;

static void tty_restore(void) {
  if (g_tty_fd < 0) return;
  tcsetattr(g_tty_fd, TCSANOW, &g_tty_saved);
  close(g_tty_fd);
  g_tty_fd = -1;
}

static void *tty_main(void *unused) { /* tty_cb (main.c:874-888): 'p' pauses between spans, 'r' resumes */
  (void)unused;
  for (;;) {
    const int fd = g_tty_fd;
    if (fd < 0) break;
    fd_set set;
    FD_ZERO(&set);
    FD_SET(fd, &set);
    if (select(fd + 1, &set, NULL, NULL, NULL) < 0) break;
    char ch;
    if (read(fd, &ch, 1) <= 0) continue;
    app *a = g_app;
    pthread_mutex_lock(&a->mu);
    if (ch == 'p' && !a->paused) a->t_pause_at = now_ms(), a->paused = true, print_status_locked(a);
    else if (ch == 'r' && a->paused) a->paused_ms += now_ms() - a->t_pause_at, a->paused = false, print_status_locked(a);
    pthread_mutex_unlock(&a->mu);
  }
  return NULL;
}

static void tty_start(app *a) {
  atexit(tty_restore);
  g_app = a;
  g_tty_fd = open("/dev/tty", O_RDONLY | O_NONBLOCK);
  if (g_tty_fd < 0) return;
  tcgetattr(g_tty_fd, &g_tty_saved);
  struct termios raw = g_tty_saved;
  raw.c_lflag &= ~(tcflag_t)(ICANON | ECHO);
  tcsetattr(g_tty_fd, TCSANOW, &raw);
  pthread_t th;
  pthread_create(&th, NULL, tty_main, NULL);
}

static void on_sigint(int sig) { /* main.c:867-872 */
  fflush(stderr);
  fflush(stdout);
  printf("\n");
  exit(sig);
}

int main(int argc, const char **argv) {
  setlocale(LC_NUMERIC, ""); /* thousands separators in %'zu, like main.c:892 */
  static app a;
  a.argc = argc, a.argv = argv;
  setup(&a);
  signal(SIGINT, on_sigint);
  tty_start(&a);
  if (a.cmd == CMD_ADD) cmd_add(&a);
  if (a.cmd == CMD_MUL) cmd_mul(&a);
  if (a.cmd == CMD_RND) cmd_rnd(&a);
  for (int r = 0; r < a.n_gpus; ++r) ecl_close(a.dev[r]);
  return 0;
}
