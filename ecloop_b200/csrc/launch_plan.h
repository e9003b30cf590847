// launch_plan.h - launch geometry of the add kernels (host-only arithmetic, no CUDA: tests/csrc/plan_test.cpp checks its
// invariants on the CPU).
#pragma once
#include <stdint.h>

#include <algorithm>

// Launch geometry for `keys` consecutive keys: T threads walk c groups of 2*Hr keys each. The half group Hr is a
// run-time quantity (table prefix (i+1)*s*G, i < Hr, plus the step 2*Hr*s*G), so instead of rounding the span up to
// whole rounds of 148 x 512 threads x 2048 keys (a 2^32-key span left 1.2 % of the lanes idle in its tail launch,
// a 2^29-key span 13 %), Hr is chosen so that T * c * 2*Hr covers the span within one group per thread.
struct launch_plan {
  uint32_t T, c, Hr;
};
// H = the table's half group (ADD_H), hr_min = the smallest half group worth launching
static inline launch_plan plan_launch(uint64_t keys, uint32_t Tmax, uint32_t H = 1024, uint32_t hr_min = 64) {
  const uint64_t full = (uint64_t)Tmax * 2 * H;
  const uint64_t c = (keys + full - 1) / full;
  const uint64_t per_thread = (keys + Tmax - 1) / Tmax;
  uint64_t Hr = (per_thread + 2 * c - 1) / (2 * c);
  Hr = std::min<uint64_t>(H, std::max<uint64_t>(hr_min, Hr));
  launch_plan lp;
  lp.c = (uint32_t)c, lp.Hr = (uint32_t)Hr;
  lp.T = (uint32_t)((keys + c * 2 * Hr - 1) / (c * 2 * Hr));
  return lp;
}

