/* u256.h — 256-bit scalar bookkeeping for the host CLI (private keys mod n, range arithmetic, hex parsing).
 *
 * Host-side key bookkeeping of the reference: fe_modn_add/sub/neg/mul, fe_modn_add_stride, fe_modn_from_hex,
 * fe_cmp, fe_bitlen (lib/ecc.c:45-265). Written on unsigned __int128 limb products with a plain
 * 512 -> 256 bit fold by 2^256 - n, not the reference's Montgomery form; values are identical.
 * A u256 is uint64_t[4], little-endian limbs, the memory image of the reference's `fe` and of the C-ABI.
 */
#ifndef ECL_U256_H
#define ECL_U256_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef uint64_t u256[4];

extern const u256 SECP_N;       /* group order */
extern const u256 SECP_P;       /* field prime (only used as the default upper range bound) */
extern const u256 SECP_LAMBDA;  /* A1 (lib/ecc.c:36): k -> k*lambda matches x -> beta*x */
extern const u256 SECP_LAMBDA2; /* A2 = lambda^2 */

void u256_set64(u256 r, uint64_t v);
void u256_copy(u256 r, const u256 a);
int u256_cmp(const u256 a, const u256 b);
bool u256_is_zero(const u256 a);
unsigned u256_bitlen(const u256 a);
uint64_t u256_add_raw(u256 r, const u256 a, const u256 b); /* returns carry */
uint64_t u256_sub_raw(u256 r, const u256 a, const u256 b); /* returns borrow */

/* mod n. add/sub follow the reference's conventions exactly (they matter for range bookkeeping):
 * add subtracts n only when the 256-bit sum overflowed, sub adds n only on borrow (lib/ecc.c:174-200). */
void modn_add(u256 r, const u256 a, const u256 b);
void modn_sub(u256 r, const u256 a, const u256 b);
void modn_neg(u256 r, const u256 a);
void modn_mul(u256 r, const u256 a, const u256 b);            /* canonical result */
void modn_add_stride(u256 r, const u256 base, const u256 stride, uint64_t off); /* base + off*stride */

/* right-to-left hex parse that skips non-hex characters (lib/ecc.c:81-95); digits beyond 64 are dropped */
void u256_from_hex(u256 r, const char *hex);
void u256_from_hex_n(u256 r, const char *hex, size_t len); /* same, explicit length (no terminator needed) */
void modn_from_hex(u256 r, const char *hex); /* + one conditional subtraction of n (lib/ecc.c:262-265) */
void modn_from_hex_n(u256 r, const char *hex, size_t len);
#endif
