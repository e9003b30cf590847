// CPU check of the add kernels' launch planner (ecloop_b200/csrc/launch_plan.h): for every span size the geometry covers
// the span, stays within the GPU's resident threads, and wastes less than one group per thread.
#include <stdio.h>
#include <stdlib.h>

#include "../../ecloop_b200/csrc/launch_plan.h"

static int fails = 0;
#define CHECK(cond_)                                                                                             \
  do {                                                                                                         \
    if (!(cond_)) {                                                                                              \
      if (fails++ < 10) printf("FAIL %s (keys=%llu Tmax=%u -> T=%u c=%u Hr=%u)\n", #cond_, (unsigned long long)keys, Tmax, lp.T, lp.c, lp.Hr); \
    }                                                                                                          \
  } while (0)

static void one(uint64_t keys, uint32_t Tmax) {
  const launch_plan lp = plan_launch(keys, Tmax);
  const uint64_t per_thread = (uint64_t)lp.c * 2 * lp.Hr, covered = per_thread * lp.T;
  CHECK(lp.T >= 1 && lp.T <= Tmax);
  CHECK(lp.Hr >= 64 && lp.Hr <= 1024 && lp.c >= 1);
  CHECK(covered >= keys);                  /* every key has an owner */
  CHECK(covered - keys < per_thread);      /* only the last thread's groups overhang the span */
  /* idle lanes: the overhang of the last thread and, for large spans, the rounding of Hr: below 1 % from 2^28 keys on */
  if (keys >= (1ull << 28) && Tmax == 75776) {
    const double used = (double)keys / ((double)Tmax * per_thread);
    CHECK(used > 0.99);
  }
}

int main(void) {
  const uint32_t tmaxes[] = {75776, 512, 200, 32, 1};
  for (uint32_t Tmax : tmaxes) {
    for (uint64_t g = 1; g <= 5000; ++g) one(g * 2048, Tmax);
    for (int b = 11; b <= 40; ++b) {
      one(1ull << b, Tmax);
      one((1ull << b) + 2048, Tmax);
      one((1ull << b) - 2048 > 0 ? (1ull << b) - 2048 + 2048 * (b == 11) : 2048, Tmax);
    }
    srand(7);
    for (int i = 0; i < 200000; ++i) one(((uint64_t)rand() * 2147483648ull + (uint64_t)rand()) % (1ull << 36) / 2048 * 2048 + 2048, Tmax);
  }
  /* the shapes DESIGN.md quotes */
  {
    const launch_plan lp = plan_launch(1ull << 32, 75776);
    uint64_t keys = 1ull << 32;
    uint32_t Tmax = 75776;
    CHECK(lp.c == 28 && lp.Hr == 1013 && (lp.T + 511) / 512 == 148);
  }
  {
    const launch_plan lp = plan_launch(1ull << 29, 75776);
    uint64_t keys = 1ull << 29;
    uint32_t Tmax = 75776;
    CHECK(lp.c == 4 && lp.Hr == 886);
  }
  printf(fails ? "%d failures\n" : "ok\n", fails);
  return fails != 0;
}
