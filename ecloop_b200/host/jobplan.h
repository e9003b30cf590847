/* jobplan.h — which keys an `add` / `rnd` run visits, and in which spans they go to the GPUs.
 *
 * The arithmetic of cmd_add + cmd_add_worker's dispenser (main.c:405-454): job size
 * J = min(range_e - range_s, 2^21) keys, job i starts at range_s + i*J*stride (mod-n addition with the reference's
 * carry convention), the loop stops when the next start is >= range_e or has wrapped below the first start, and a
 * job visits ceil(J/2048)*2048 keys. Consecutive jobs are contiguous in units of the stride, so up to span_jobs of
 * them are handed out as ONE span (one ecl_add_submit); a job whose size is not a multiple of 2048 cannot be fused.
 * Not thread-safe: the caller serialises jobplan_take (ecloop.c holds its mutex). */
#ifndef ECL_JOBPLAN_H
#define ECL_JOBPLAN_H
#include <stdbool.h>
#include <stdint.h>

#include "u256.h"

#define JOBPLAN_MAX_JOB (2u * 1024 * 1024) /* MAX_JOB_SIZE (main.c:16) */
#define JOBPLAN_GROUP 2048u                /* GROUP_INV_SIZE (main.c:17) */

typedef struct job_plan {
  u256 next, first, range_e, stride, job_inc;
  uint64_t job_keys;   /* ctx->job_size: what the status counter advances by per job */
  uint64_t visit_keys; /* keys a job really visits */
  uint64_t span_jobs;  /* jobs fused per span */
} job_plan;

/* fixed_job: rnd mode always uses 2^21-key jobs (main.c:625); add mode derives the size from the range (main.c:442) */
void jobplan_init(job_plan *jp, const u256 range_s, const u256 range_e, unsigned ord_offs, bool fixed_job);
/* spans of at most max_span jobs, shrunk so that a short range still spreads over n_ranks */
void jobplan_choose_span(job_plan *jp, uint64_t max_span, unsigned n_ranks);
/* next span: returns the number of jobs (0 = exhausted) and its first key */
uint64_t jobplan_take(job_plan *jp, u256 start);
#endif
