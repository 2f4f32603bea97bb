// xdrop.cuh -- register-resident form of SeqAn's gapped X-drop seed extension (Trail.cpp:373,391 ->
// extendSeed(.., GappedXDrop); restated cell for cell in xdrop_extend_scalar, align.cuh).
//
// With Score(0,-1,-1) the running best is 0 for ever, so a cell survives iff its score >= -X, i.e. iff it lies
// within |diagonal| <= X.  The band is therefore fixed to the diagonals around the seed and needs no memory at
// all: lane j of the warp that owns the read keeps S adjacent diagonals in registers,
//     index g = S*j + i  (0 <= g < 32*S),   diagonal k = col - row = 2*g - 32*S + (d & 1)
// for anti-diagonal d.  Cells of even and odd anti-diagonals alternate (vE / vO), the diagonal predecessor of
// a cell is the lane's own previous value, and the two gap predecessors are own/neighbour values of the other
// parity: ONE shuffle per anti-diagonal.  Going from an even to the next odd anti-diagonal every cell moves one
// column on (query characters shift down one index), from odd to even one row on (database characters shift
// up): one more shuffle, and a single new character per anti-diagonal for the whole warp.  The window
// [minCol, maxCol) of the original is kept as two uniform scalars and updated from two warp reductions.
// Window updates and the end-position rules are shared with the host mirror in tests/hostemu (xd_next_window,
// xd_finish), which is fuzzed against xdrop_extend_scalar; the device transcription is checked on the GPU.
#pragma once
#include "defs.cuh"

namespace talc {

static const i32 kXdU = -(1 << 29);  // "undefined": below every reachable score, and stays below after +gap

struct XdHist {  // pre-trim windows [min, max) of the last three anti-diagonals (3 = the newest)
  i32 min1, max1, min2, max2, min3, max3;
  TALC_HD void push(i32 mn, i32 mx) {
    min1 = min2; max1 = max2;
    min2 = min3; max2 = max3;
    min3 = mn; max3 = mx;
  }
};
TALC_HD XdHist xd_hist_init() {
  XdHist h;
  h.min1 = 1; h.max1 = -1;  // nothing (len 0)
  h.min2 = 1; h.max2 = 0;   // anti-diagonal 0: the origin, len 1
  h.min3 = 1; h.max3 = 1;   // anti-diagonal 1: the two boundary cells, len 2
  return h;
}
// first column index of the register band on anti-diagonal d
TALC_HD i32 xd_col_base(i32 d, i32 S) { return (d - 32 * S + (d & 1)) >> 1; }

// The two trimming loops + clamps of _extendSeedGappedXDropOneDirection after anti-diagonal d was computed with
// the window [minCol, maxCol):
//   loC = lowest column c in [minCol, maxCol] with cell(d, c) or cell(d-1, c-1) defined   (INT32_MAX: none)
//   hiC = highest column c in [minCol-1, maxCol-1] with cell(d, c) or cell(d-1, c) defined (INT32_MIN: none)
TALC_HD void xd_next_window(i32 d, i32 rows, i32 cols, i32 loC, i32 hiC, i32& minCol, i32& maxCol) {
  const i32 off3 = minCol - 1, maxPre = maxCol;
  i32 nmin = (loC == INT32_MAX) ? maxPre + 1 : loC;
  i32 nmax = (hiC == INT32_MIN) ? off3 : hiC + 1;
  ++nmax;
  const i32 lo = d + 2 - rows;  // end of the database segment reached?
  if (lo > nmin) nmin = lo;
  if (cols < nmax) nmax = cols;  // end of the query segment reached?
  minCol = nmin;
  maxCol = nmax;
}

// End-position rules.  at(which, col): value of anti-diagonal d (which = 3) or d-1 (which = 2) at column col,
// kXdU when absent; argmax1(lo, hi, col): greatest value of anti-diagonal d-2 over columns [lo, hi], lowest
// column among equals.
template <class At, class ArgMax>
TALC_HD void xd_finish(i32 d, const XdHist& h, At at, ArgMax argmax1, u32& ext_rows, u32& ext_cols, i32& end_score) {
  i32 col = h.max3 - 1;
  i32 row = d - col;
  i32 sc = at(3, col);
  if (sc == kXdU) {
    const i32 len2 = h.max2 - h.min2 + 2;
    const i32 v = at(2, h.max2 - 1);
    if (v != kXdU) {  // reached the end of the query segment
      col = h.max2 - 1;
      row = d - 1 - col;
      sc = v;
    } else if (len2 > 2) {
      const i32 v2 = at(2, h.max2 - 2);
      if (v2 != kXdU) {  // reached the end of the database segment
        col = h.max2 - 2;
        row = d - 1 - col;
        sc = v2;
      }
    }
  }
  if (sc == kXdU) {  // general case: first strictly greatest entry two anti-diagonals back
    i32 c = 0;
    const i32 v = argmax1(h.min1 - 1, h.max1, c);
    if (v > kXdU) {
      sc = v;
      col = c;
      row = d - 2 - col;
    }
  }
  if (sc != kXdU) {
    ext_rows = (u32)row;
    ext_cols = (u32)col;
    end_score = sc;  // = -(edit distance of the two extended prefixes): every surviving cell holds its exact distance
  }
}

// The part of the cell rule that is the same for every cell of anti-diagonal d computed with the window
// [minCol, maxCol) (minCol < maxCol inside the loop).
struct XdRow {
  i32 W;      // maxCol - minCol > 0
  i32 negX;   // survival threshold
  i32 negd;   // boundary value
  bool col0;  // column 0 of the matrix is still alive on this anti-diagonal (it is the cell left of the window)
  bool row0;  // row 0 of the matrix is still alive (it is the cell right of the window)
};
TALC_HD XdRow xd_row(i32 d, i32 X, i32 minCol, i32 maxCol) {
  XdRow r;
  r.W = maxCol - minCol;
  r.negX = -X;
  r.negd = -d;
  r.col0 = (minCol == 1) & (d < X);
  r.row0 = (d == maxCol) & (d < X);
  return r;
}
// One cell: value, and its contribution to the two trimming reductions.  t = column - minCol, so that every range
// test of the original (col in [minCol, maxCol), [minCol, maxCol], [minCol-1, maxCol-1], == minCol-1, == maxCol) is
// one unsigned comparison against W; stored values are kXdU or >= -X > kXdU, so "a or b is defined" is max(a, b) > kXdU.
// loT / hiT are the reductions in the same shifted coordinate (xd_window_bounds turns them back into columns).
TALC_HD i32 xd_cell(i32 a, i32 b, i32 dg, bool match, i32 t, const XdRow& R, i32& loT, i32& hiT) {
  const i32 g = (a > b ? a : b) - 1;
  const i32 s = dg - (match ? 0 : 1);
  const i32 tmp = g > s ? g : s;
  i32 val = (((u32)t < (u32)R.W) & (tmp >= R.negX)) ? tmp : kXdU;
  if ((t == -1) & R.col0) val = R.negd;   // column 0 of the matrix
  if ((t == R.W) & R.row0) val = R.negd;  // row 0 of the matrix
  const i32 va = val > a ? val : a, vb = val > b ? val : b;
  const i32 tl = (((u32)t <= (u32)R.W) & (va > kXdU)) ? t : INT32_MAX;
  const i32 th = (((u32)(t + 1) <= (u32)R.W) & (vb > kXdU)) ? t : INT32_MIN;
  loT = tl < loT ? tl : loT;
  hiT = th > hiT ? th : hiT;
  return val;
}
// reductions in window coordinates -> the column sentinels xd_next_window expects
TALC_HD void xd_window_bounds(i32 loT, i32 hiT, i32 minCol, i32& loC, i32& hiC) {
  loC = (loT == INT32_MAX) ? INT32_MAX : loT + minCol;
  hiC = (hiT == INT32_MIN) ? INT32_MIN : hiT + minCol;
}

#if defined(__CUDA_ARCH__)
template <int S>
__device__ __noinline__ void xdrop_extend_reg(const SeqView& queryArg, u32 qoff, u32 qlen, const SeqView& databaseArg,
                                              u32 doff, u32 dlen, int X, u32& ext_rows, u32& ext_cols, i32& end_score,
                                              DpStats* st) {
  const SeqView query = queryArg, database = databaseArg;  // by value: the fields stay in registers
  const i32 lane = (i32)(threadIdx.x & 31u);
  const i32 cols = (i32)qlen + 1, rows = (i32)dlen + 1;
  ext_rows = 0;
  ext_cols = 0;
  end_score = 0;
  if (rows == 1 || cols == 1) return;
  const i32 g0 = S * lane;
  i32 vE[S], vO[S], vOld[S];
  u32 qc[S], tc[S];
#pragma unroll
  for (int i = 0; i < S; ++i) {
    const i32 g = g0 + i;
    vE[i] = (g == 16 * S) ? 0 : kXdU;
    vO[i] = ((g == 16 * S - 1) | (g == 16 * S)) ? (X >= 1 ? -1 : kXdU) : kXdU;
    vOld[i] = kXdU;
    const i32 col = 1 - 16 * S + g, row = 1 + 16 * S - g;  // anti-diagonal 2
    qc[i] = (col >= 1 && col <= (i32)qlen) ? query.code(qoff + (u32)(col - 1)) : 6u;
    tc[i] = (row >= 1 && row <= (i32)dlen) ? database.code(doff + (u32)(row - 1)) : 7u;
  }
  i32 minCol = 1, maxCol = 2, d = 1;
  XdHist h = xd_hist_init();
  u32 cells = 0;
  bool lastOdd;
#pragma unroll 1
  for (;;) {
    // ---------------- even anti-diagonal d = 2m
    ++d;
    const i32 m = d >> 1;
    h.push(minCol, maxCol);
    {
      const i32 cb = m - 16 * S;
      const u32 qi = (u32)(m + 16 * S - 1);  // query character lane 31 needs on the next (odd) anti-diagonal
      const u32 nq = (qi < qlen) ? query.code(qoff + qi) : 6u;
      i32 aEdge = __shfl_up_sync(0xffffffffu, vO[S - 1], 1);
      if (lane == 0) aEdge = kXdU;
      i32 loT = INT32_MAX, hiT = INT32_MIN, loC, hiC;
      const XdRow R = xd_row(d, X, minCol, maxCol);
      const i32 t0 = cb + g0 - minCol;
#pragma unroll
      for (int i = 0; i < S; ++i) {
        const i32 a = i ? vO[i - 1] : aEdge, b = vO[i], dg = vE[i];
        vOld[i] = dg;
        vE[i] = xd_cell(a, b, dg, qc[i] == tc[i], t0 + i, R, loT, hiT);
      }
      loT = __reduce_min_sync(0xffffffffu, loT);
      hiT = __reduce_max_sync(0xffffffffu, hiT);
      xd_window_bounds(loT, hiT, minCol, loC, hiC);
      cells += (u32)(maxCol - minCol);
      xd_next_window(d, rows, cols, loC, hiC, minCol, maxCol);
      const u32 e = __shfl_down_sync(0xffffffffu, qc[0], 1);
#pragma unroll
      for (int i = 0; i + 1 < S; ++i) qc[i] = qc[i + 1];
      qc[S - 1] = (lane == 31) ? nq : e;
    }
    if (!(minCol < maxCol)) { lastOdd = false; break; }
    // ---------------- odd anti-diagonal d = 2m + 1
    ++d;
    h.push(minCol, maxCol);
    {
      const i32 cb = m + 1 - 16 * S;
      const u32 ti = (u32)(m + 16 * S);  // database character lane 0 needs on the next (even) anti-diagonal
      const u32 nt = (ti < dlen) ? database.code(doff + ti) : 7u;
      i32 bEdge = __shfl_down_sync(0xffffffffu, vE[0], 1);
      if (lane == 31) bEdge = kXdU;
      i32 loT = INT32_MAX, hiT = INT32_MIN, loC, hiC;
      const XdRow R = xd_row(d, X, minCol, maxCol);
      const i32 t0 = cb + g0 - minCol;
#pragma unroll
      for (int i = 0; i < S; ++i) {
        const i32 a = vE[i], b = (i + 1 < S) ? vE[i + 1] : bEdge, dg = vO[i];
        vOld[i] = dg;
        vO[i] = xd_cell(a, b, dg, qc[i] == tc[i], t0 + i, R, loT, hiT);
      }
      loT = __reduce_min_sync(0xffffffffu, loT);
      hiT = __reduce_max_sync(0xffffffffu, hiT);
      xd_window_bounds(loT, hiT, minCol, loC, hiC);
      cells += (u32)(maxCol - minCol);
      xd_next_window(d, rows, cols, loC, hiC, minCol, maxCol);
      const u32 e = __shfl_up_sync(0xffffffffu, tc[S - 1], 1);
#pragma unroll
      for (int i = S - 1; i > 0; --i) tc[i] = tc[i - 1];
      tc[0] = (lane == 0) ? nt : e;
    }
    if (!(minCol < maxCol)) { lastOdd = true; break; }
  }
  if (st) st->cells_xdrop += cells;
  // ---------------- end position
  const i32 cb3 = xd_col_base(d, S), cb2 = xd_col_base(d - 1, S), cb1 = xd_col_base(d - 2, S);
  auto pick = [&](const i32* v, i32 cb, i32 col) -> i32 {
    const i32 g = col - cb;
    const bool in = (g >= 0) & (g < 32 * S);
    const i32 gi = in ? g : 0;
    const i32 i = gi % S;
    i32 x = v[0];
#pragma unroll
    for (int q = 1; q < S; ++q) x = (i == q) ? v[q] : x;
    x = __shfl_sync(0xffffffffu, x, gi / S);
    return in ? x : kXdU;
  };
  auto at = [&](int which, i32 col) -> i32 {
    if (which == 3) return lastOdd ? pick(vO, cb3, col) : pick(vE, cb3, col);
    return lastOdd ? pick(vE, cb2, col) : pick(vO, cb2, col);
  };
  auto argmax1 = [&](i32 lo, i32 hi, i32& col) -> i32 {
    i32 bv = kXdU, bc = INT32_MAX;
#pragma unroll
    for (int i = 0; i < S; ++i) {
      const i32 c = cb1 + g0 + i;
      if ((c >= lo) & (c <= hi) & (vOld[i] > bv)) { bv = vOld[i]; bc = c; }
    }
    const i32 mx = __reduce_max_sync(0xffffffffu, bv);
    col = __reduce_min_sync(0xffffffffu, (bv == mx && bv > kXdU) ? bc : INT32_MAX);
    return mx;
  };
  xd_finish(d, h, at, argmax1, ext_rows, ext_cols, end_score);
}
#endif

}  // namespace talc
