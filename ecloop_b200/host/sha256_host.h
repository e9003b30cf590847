/* sha256_host.h — SHA-256 of a short byte string on the host, used only to turn a `-raw` stdin line into a
 * private key (cmd_mul_worker's parse step, main.c:506-527, which stays on the host). The hot-path SHA-256
 * (of public keys) is in csrc/hash160.cuh. */
#ifndef ECL_SHA256_HOST_H
#define ECL_SHA256_HOST_H
#include <stddef.h>
#include <stdint.h>
void sha256_bytes(uint32_t digest_words[8], const uint8_t *msg, size_t len);
#endif
