/* jobplan.c — see jobplan.h */
#include "jobplan.h"

#include <string.h>

void jobplan_init(job_plan *jp, const u256 range_s, const u256 range_e, unsigned ord_offs, bool fixed_job) {
  memset(jp, 0, sizeof *jp);
  u256_copy(jp->next, range_s);
  u256_copy(jp->first, range_s);
  u256_copy(jp->range_e, range_e);
  jp->stride[ord_offs / 64] = 1ULL << (ord_offs % 64); /* stride_k = 2^offs (main.c:222-223) */
  u256 r;
  modn_sub(r, range_e, range_s);
  const bool small = !r[1] && !r[2] && !r[3] && r[0] < JOBPLAN_MAX_JOB;
  jp->job_keys = (small && !fixed_job) ? r[0] : JOBPLAN_MAX_JOB;
  jp->visit_keys = (jp->job_keys + JOBPLAN_GROUP - 1) / JOBPLAN_GROUP * JOBPLAN_GROUP;
  u256 jk;
  u256_set64(jk, jp->job_keys);
  modn_mul(jp->job_inc, jk, jp->stride); /* main.c:413-415 */
  jp->span_jobs = 1;
}

void jobplan_choose_span(job_plan *jp, uint64_t max_span, unsigned n_ranks) {
  uint64_t span = max_span ? max_span : 1;
  if (jp->visit_keys != jp->job_keys) span = 1; /* a ragged job cannot be fused with its neighbour */
  u256 r;
  modn_sub(r, jp->range_e, jp->first);
  const unsigned shift = u256_bitlen(jp->job_inc) ? u256_bitlen(jp->job_inc) - 1 : 0; /* job_inc = J * 2^offs */
  if ((jp->job_keys & (jp->job_keys - 1)) == 0 && u256_bitlen(r) <= shift + 40) {
    uint64_t jobs = 0; /* ceil(r / job_inc), fits 41 bits */
    bool rest = false;
    for (unsigned i = 0; i < shift && i < 256; ++i) rest |= (r[i / 64] >> (i % 64)) & 1ULL;
    for (unsigned i = shift; i < 256 && i - shift < 64; ++i) jobs |= ((r[i / 64] >> (i % 64)) & 1ULL) << (i - shift);
    jobs += rest;
    if (n_ranks == 0) n_ranks = 1;
    const uint64_t per_rank = (jobs + n_ranks - 1) / n_ranks;
    if (per_rank < span) span = per_rank ? per_rank : 1;
  }
  jp->span_jobs = span;
}

uint64_t jobplan_take(job_plan *jp, u256 start) {
  uint64_t jobs = 0;
  while (jobs < jp->span_jobs) {
    if (u256_cmp(jp->next, jp->range_e) >= 0 || u256_cmp(jp->next, jp->first) < 0) break; /* main.c:420-424 */
    if (jobs == 0) u256_copy(start, jp->next);
    u256 before;
    u256_copy(before, jp->next);
    modn_add(jp->next, jp->next, jp->job_inc); /* main.c:427 */
    jobs++;
    if (u256_cmp(jp->next, before) < 0) break; /* wrapped: the next job is not contiguous with this span */
  }
  return jobs;
}
