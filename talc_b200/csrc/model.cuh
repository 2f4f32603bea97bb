// model.cuh -- count prediction intervals and successor tagging.
//
// Decisions are made in IEEE double with separate multiply/add (the reference's x86-64 build has
// no FMA: SURVEY F7), so this file must be compiled with -fmad=false; std::pow(x,2) is x*x.
#pragma once
#include <math.h>

#include "defs.cuh"

namespace talc {

enum Tag : u8 { kExpected = 0, kUnexpected = 1, kBreakpoint = 2 };

// Explorer.cpp:1185-1198.  upper==true is the reference's `classe == UNEXPECTED` branch (noise upper
// bound), upper==false the `EXPECTED` branch (lower bound).  sqrt of a negative value is NaN and
// makes the comparison false (cc = 0 lower bound).
TALC_HD bool expected_by_model(u32 nextc, u32 cc, double alpha, bool upper) {
  if (cc <= 3) {
    if (upper) return ((double)nextc <= ((double)(cc + 0.5) + alpha * sqrt((double)(cc + 0.5))));
    return ((double)nextc >= ((double)(cc - 0.5) + (1 - alpha) * sqrt((double)(cc - 0.5))));
  }
  if (upper) {
    const double x = (alpha / 2 + sqrt((double)(cc + 0.96)));
    return ((double)nextc <= x * x);
  }
  const double y = (alpha / 2 - sqrt((double)(cc + 0.02)));
  return ((double)nextc >= y * y);
}

// Explorer.cpp:1200-1217
TALC_HD bool expected_by_last_node(u32 nextc, u32 cc, double alpha) {
  return expected_by_model(nextc, cc, alpha, true) & expected_by_model(nextc, cc, alpha, false);
}

// Explorer.cpp:1226-1298.  Returns the number of tags (0 on a dead end, else 4).
TALC_HDN int tag_next_nodes(const u32 cnt[4], const u32 col[4], u32 count, const Params& P, bool complex_, u8 tag[4],
                           double dist[4]) {
  int counter = 0;
  u32 lambda_noise = 0;
  u32 nbExpected = 0, nbBreakpoints = 0, nbUnexpected = 0;
  for (int i = 0; i < 4; ++i)
    if (cnt[i] >= P.min_count) counter++;
  if (counter == 0) return 0;
  lambda_noise = (u32)(i32)((double)count * P.sr_error);
  const double sq = sqrt((double)count);
  for (int b = 0; b < 4; ++b) {
    const u32 nextc = cnt[b];
    dist[b] = fabs((double)count - (double)nextc) / sq;
    if (nextc >= P.min_count) {
      if (expected_by_model(nextc, count, P.alpha, false) || (counter == 1)) {
        tag[b] = kExpected;
        ++nbExpected;
      } else if (lambda_noise >= P.min_count) {
        if (!expected_by_model(nextc, lambda_noise, P.alpha, true) || (col[b] > 0)) {
          tag[b] = kBreakpoint;
          ++nbBreakpoints;
        } else {
          tag[b] = kUnexpected;
          ++nbUnexpected;
        }
      } else {
        tag[b] = kBreakpoint;
        ++nbBreakpoints;
      }
    } else
      tag[b] = kUnexpected;  // not counted in nbUnexpected (:1275)
  }
  if ((nbExpected == 0) & (nbBreakpoints == 1)) {
    for (int t = 0; t < 4; ++t)
      if (tag[t] == kBreakpoint) tag[t] = kExpected;
  }
  if ((nbExpected == 1) & (nbUnexpected > 0) & !complex_) {
    int sum = 0;
    u32 index = 0;
    for (u32 i = 0; i < 4; ++i) {
      if (tag[i] == kUnexpected) {
        if (sum == 0) index = i;
        sum += (int)cnt[i];
        if (cnt[index] < cnt[i]) index = i;
      }
    }
    if (!expected_by_model((u32)sum, lambda_noise, P.alpha, true)) tag[index] = kBreakpoint;
  }
  return 4;
}

}  // namespace talc
