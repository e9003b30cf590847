/* ecl_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, gcc) of the reference's hot path, used as the parity checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs. Nothing in the product
 * (ecloop_b200/, include/) may include, link or call this file.
 *
 * Parity status: PINNED — checked in tests/test_oracle.py against (i) the reference's own known-answer
 * fixtures (9 / 13 puzzle keys, 1080 brain-wallet keys, SURVEY Appendix B vectors) and (ii) full dumps
 * produced by the unmodified reference binary built from /root/reference (oracle/_ref, tools/gen_golden.py),
 * committed under tests/golden/.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 */
#ifndef ECL_ORACLE_H
#define ECL_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t orc_fe[4]; /* little-endian 64-bit limbs, same as `fe` (lib/ecc.c:26) */

typedef struct {
  uint64_t key_off;  /* index j of the key inside the span: key = start + j*stride (mod n)        */
  uint32_t h160[5];  /* h160_t word order (lib/addr.c:16)                                          */
  uint8_t endo;      /* 0..5, calc_priv numbering (main.c:267-276)                                 */
  uint8_t kind;      /* 0 = addr33, 1 = addr65                                                     */
  uint8_t pad[2];
  uint64_t pk[4];    /* recovered private key (calc_priv) — for mul: the input key                */
} orc_hit;

#define ORC_A33 1u
#define ORC_A65 2u
#define ORC_ENDO 4u

/* filter = bloom (bits,size) + optional sorted unique list (main.c:205-217) */
typedef struct {
  const uint64_t *bits;
  uint64_t size; /* words */
  const uint32_t *list; /* count*5 words sorted by compare_160, or NULL = bloom-only mode */
  uint64_t count;
} orc_filter;

/* Fp (lib/ecc.c:269-520) */
void orc_fp_add(orc_fe r, const orc_fe a, const orc_fe b);
void orc_fp_sub(orc_fe r, const orc_fe a, const orc_fe b);
void orc_fp_neg(orc_fe r, const orc_fe a);
void orc_fp_mul(orc_fe r, const orc_fe a, const orc_fe b);
void orc_fp_sqr(orc_fe r, const orc_fe a);
void orc_fp_inv(orc_fe r, const orc_fe a);
void orc_fp_grpinv(orc_fe *r, uint32_t n);
/* Fn (lib/ecc.c:166-265) */
void orc_fn_add(orc_fe r, const orc_fe a, const orc_fe b);
void orc_fn_sub(orc_fe r, const orc_fe a, const orc_fe b);
void orc_fn_neg(orc_fe r, const orc_fe a);
void orc_fn_mul(orc_fe r, const orc_fe a, const orc_fe b);
void orc_fn_add_stride(orc_fe r, const orc_fe base, const orc_fe stride, uint64_t offset);
void orc_fn_from_hex(orc_fe r, const char *hex);
/* group (lib/ecc.c:546-929): returns 0 and affine x,y; returns 1 for the point at infinity */
int orc_ec_mul_g(orc_fe x, orc_fe y, const orc_fe k);
int orc_ec_add(orc_fe rx, orc_fe ry, const orc_fe px, const orc_fe py, const orc_fe qx, const orc_fe qy);
/* hashes (lib/sha256.c:399-453, lib/rmd160s.c:122-336, lib/addr.c:33-131) */
void orc_sha256_blocks(uint32_t state[8], const uint8_t *data, size_t nblocks);
void orc_rmd160_block(uint32_t state[5], const uint32_t w[16]);
void orc_hash160_33(uint32_t h[5], const orc_fe x, const orc_fe y);
void orc_hash160_65(uint32_t h[5], const orc_fe x, const orc_fe y);
/* bloom (lib/utils.c:282-326) */
void orc_blf_positions(uint64_t pos[20], const uint32_t h[5], uint64_t size_words);
void orc_blf_add(uint64_t *bits, uint64_t size_words, const uint32_t h[5]);
int orc_blf_has(const uint64_t *bits, uint64_t size_words, const uint32_t h[5]);
int orc_filter_check(const orc_filter *f, const uint32_t h[5]); /* main.c:205-217 */
/* key recovery (main.c:267-276) */
void orc_calc_priv(orc_fe pk, const orc_fe start, const orc_fe stride, uint64_t off, uint8_t endo);

/* batch_add + check_found_add over a span of n_keys (multiple of 2048) keys start + j*stride
 * (main.c:287-403). Hits are emitted in the reference's `-t 1` order. Returns number of hits found
 * (may exceed cap; only the first cap are stored). */
uint64_t orc_add_span(const orc_fe start, const orc_fe stride, uint64_t n_keys, uint32_t flags,
                      const orc_filter *f, orc_hit *hits, uint64_t cap);

/* cmd_add + cmd_add_worker at -t 1 (main.c:405-454): job plan of SURVEY A.1 on top of orc_add_span.
 * key_off in the hits is relative to range_s. *k_checked gets the status-line counter. */
uint64_t orc_add_range(const orc_fe range_s, const orc_fe range_e, uint32_t ord_offs, uint32_t flags,
                       const orc_filter *f, orc_hit *hits, uint64_t cap, uint64_t *k_checked);

/* cmd_mul_worker compute part (main.c:531-534): k*G per key, hash160, filter. pks already parsed. Keys
 * that are 0 mod n are skipped (documented divergence). key_off = index into pks. */
uint64_t orc_mul_batch(const orc_fe *pks, uint64_t n, uint32_t flags, const orc_filter *f, orc_hit *hits,
                       uint64_t cap);

/* hash160 of k*G for each key (dump helper for parity tests): out33/out65 may be NULL */
void orc_pubkey_hashes(const orc_fe *pks, uint64_t n, uint32_t *out33, uint32_t *out65, uint64_t *outxy);

#ifdef __cplusplus
}
#endif
#endif
