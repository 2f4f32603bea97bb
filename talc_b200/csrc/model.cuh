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
TALC_HDN bool expected_by_model(u32 nextc, u32 cc, double alpha, bool upper) {
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
TALC_HDN bool expected_by_last_node(u32 nextc, u32 cc, double alpha) {
  return expected_by_model(nextc, cc, alpha, true) & expected_by_model(nextc, cc, alpha, false);
}

// The two interval bounds a step needs, each evaluated once instead of once per successor: the lower
// bound of the "expected" interval around the current count (Explorer.cpp:1191,1196) and the upper
// bound of the noise interval around lambda = (int)(count * SR_ERROR_RATE) (Explorer.cpp:1190,1195).
// Same expressions, same operation order as expected_by_model, so the comparisons are bit-identical.
TALC_HD double model_lower_bound(u32 cc, double alpha) {
  if (cc <= 3) return ((double)(cc - 0.5) + (1 - alpha) * sqrt((double)(cc - 0.5)));  // NaN for cc = 0
  const double y = (alpha / 2 - sqrt((double)(cc + 0.02)));
  return y * y;
}
TALC_HD double model_upper_bound(u32 cc, double alpha) {
  if (cc <= 3) return ((double)(cc + 0.5) + alpha * sqrt((double)(cc + 0.5)));
  const double x = (alpha / 2 + sqrt((double)(cc + 0.96)));
  return x * x;
}

// Per-step constants of the tagging rules: they depend on the current count only, so the caller can
// evaluate them while the four table probes are still in flight.
struct StepBounds {
  double lower;       // lower bound of the "expected" interval around count
  double upperNoise;  // upper bound of the noise interval around lambda (valid iff noiseModel)
  double sq;          // sqrt(count), the denominator of dist = |count - next| / sqrt(count)  (Explorer.cpp:1247)
  u32 lambda;         // (int)(count * SR_ERROR_RATE)
  bool noiseModel;    // lambda >= MIN_COUNT
};
TALC_HDN StepBounds step_bounds(u32 count, const Params& P) {
  StepBounds b;
  b.lambda = (u32)(i32)((double)count * P.sr_error);
  b.noiseModel = b.lambda >= P.min_count;
  b.lower = model_lower_bound(count, P.alpha);
  b.upperNoise = b.noiseModel ? model_upper_bound(b.lambda, P.alpha) : 0.0;
  b.sq = sqrt((double)count);
  return b;
}
// The same constants from per-context lookup tables (filled on the device by model_tabs_kernel with the very
// functions above, so the values are bit-identical); counts beyond the table are computed.
struct ModelTabs {
  const double* lower;  // model_lower_bound(c, alpha)
  const double* upper;  // model_upper_bound(c, alpha)
  const double* sq;     // sqrt((double)c)
  u32 n;
};
TALC_HD StepBounds step_bounds_tab(u32 count, const Params& P, const ModelTabs& M) {
  if (count >= M.n) return step_bounds(count, P);
  StepBounds b;
  b.lambda = (u32)(i32)((double)count * P.sr_error);
  b.noiseModel = b.lambda >= P.min_count;
  b.lower = M.lower[count];
  b.upperNoise = b.noiseModel ? M.upper[b.lambda] : 0.0;  // lambda <= count < n
  b.sq = M.sq[count];
  return b;
}
TALC_HD double step_dist(u32 count, u32 nextc, const StepBounds& b) {
  return fabs((double)count - (double)nextc) / b.sq;
}

// Explorer.cpp:1226-1298.  Returns the number of tags (0 on a dead end, else 4).  The distance of a
// successor is step_dist(); the reference only ever reads it for successors that are not UNEXPECTED.
TALC_HDN int tag_next_nodes(const u32 cnt[4], const u32 col[4], const StepBounds& B, const Params& P, bool complex_,
                           u8 tag[4]) {
  int counter = 0;
  u32 nbExpected = 0, nbBreakpoints = 0, nbUnexpected = 0;
  for (int i = 0; i < 4; ++i)
    if (cnt[i] >= P.min_count) counter++;
  if (counter == 0) return 0;
  for (int b = 0; b < 4; ++b) {
    const u32 nextc = cnt[b];
    if (nextc >= P.min_count) {
      if (((double)nextc >= B.lower) || (counter == 1)) {
        tag[b] = kExpected;
        ++nbExpected;
      } else if (B.noiseModel) {
        if (!((double)nextc <= B.upperNoise) || (col[b] > 0)) {
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
    // nbUnexpected > 0 implies the noise model is active, so upperNoise is valid
    if (!((double)(u32)sum <= B.upperNoise)) tag[index] = kBreakpoint;
  }
  return 4;
}

}  // namespace talc
