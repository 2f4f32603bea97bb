// stdsort.cuh -- libstdc++ (GCC 13) std::sort, reproduced step for step.
//
// The reference breaks ties by whatever permutation std::sort happens to produce
// (Explorer.cpp:409 anchors, :744-769 gardening).  std::sort is an unstable
// introsort, so to emit the same bytes the device has to run the same algorithm:
// bits/stl_algo.h  __sort -> __introsort_loop (threshold 16, depth limit 2*floor(lg n),
// __move_median_to_first(first, first+1, mid, last-1), __unguarded_partition, heap-sort
// fallback) -> __final_insertion_sort; bits/stl_heap.h for the fallback.
// Only the comparator outcomes matter for the permutation, not the element type.
#pragma once
#include "defs.cuh"

namespace talc {
namespace stdsort_detail {

template <class T>
TALC_HD void swap_(T& a, T& b) { T t = a; a = b; b = t; }

template <class T, class Less>
TALC_HD void move_median_to_first(T* result, T* a, T* b, T* c, Less less) {
  if (less(*a, *b)) {
    if (less(*b, *c)) swap_(*result, *b);
    else if (less(*a, *c)) swap_(*result, *c);
    else swap_(*result, *a);
  } else if (less(*a, *c)) swap_(*result, *a);
  else if (less(*b, *c)) swap_(*result, *c);
  else swap_(*result, *b);
}

template <class T, class Less>
TALC_HD T* unguarded_partition(T* first, T* last, T* pivot, Less less) {
  TALC_ROLLED
  for (;;) {
    while (less(*first, *pivot)) ++first;
    --last;
    while (less(*pivot, *last)) --last;
    if (!(first < last)) return first;
    swap_(*first, *last);
    ++first;
  }
}

template <class T, class Less>
TALC_HD void push_heap_(T* first, i64 holeIndex, i64 topIndex, T value, Less less) {
  i64 parent = (holeIndex - 1) / 2;
  while (holeIndex > topIndex && less(first[parent], value)) {
    first[holeIndex] = first[parent];
    holeIndex = parent;
    parent = (holeIndex - 1) / 2;
  }
  first[holeIndex] = value;
}

template <class T, class Less>
TALC_HD void adjust_heap(T* first, i64 holeIndex, i64 len, T value, Less less) {
  const i64 topIndex = holeIndex;
  i64 secondChild = holeIndex;
  while (secondChild < (len - 1) / 2) {
    secondChild = 2 * (secondChild + 1);
    if (less(first[secondChild], first[secondChild - 1])) secondChild--;
    first[holeIndex] = first[secondChild];
    holeIndex = secondChild;
  }
  if ((len & 1) == 0 && secondChild == (len - 2) / 2) {
    secondChild = 2 * (secondChild + 1);
    first[holeIndex] = first[secondChild - 1];
    holeIndex = secondChild - 1;
  }
  push_heap_(first, holeIndex, topIndex, value, less);
}

template <class T, class Less>
TALC_HD void heap_sort_all(T* first, T* last, Less less) {  // __partial_sort(first, last, last)
  const i64 len = last - first;
  if (len >= 2) {  // __make_heap
    i64 parent = (len - 2) / 2;
    TALC_ROLLED
    for (;;) {
      T value = first[parent];
      adjust_heap(first, parent, len, value, less);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) {  // __sort_heap / __pop_heap
    --last;
    T value = *last;
    *last = *first;
    adjust_heap(first, (i64)0, (i64)(last - first), value, less);
  }
}

template <class T, class Less>
TALC_HD void unguarded_linear_insert(T* last, Less less) {
  T val = *last;
  T* next = last;
  --next;
  while (less(val, *next)) {
    *last = *next;
    last = next;
    --next;
  }
  *last = val;
}

template <class T, class Less>
TALC_HD void insertion_sort(T* first, T* last, Less less) {
  if (first == last) return;
  TALC_ROLLED
  for (T* i = first + 1; i != last; ++i) {
    if (less(*i, *first)) {
      T val = *i;
      TALC_ROLLED
      for (T* p = i; p != first; --p) *p = *(p - 1);  // move_backward(first, i, i + 1)
      *first = val;
    } else
      unguarded_linear_insert(i, less);
  }
}

}  // namespace stdsort_detail

template <class T, class Less>
TALC_HD void std_sort(T* first, T* last, Less less) {
  using namespace stdsort_detail;
  if (first == last) return;
  const i64 n = last - first;
  // __introsort_loop with the recursion unrolled onto an explicit stack; the two halves are
  // disjoint, so the order in which they are processed does not change the result
  struct Frame { T* first; T* last; int depth; };
  Frame stack[40];  // depth limit 2*floor(lg n) <= 38 for any n < 2^19
  int sp = 0;
  int lg = 0;
  TALC_ROLLED
  for (i64 v = n; v > 1; v >>= 1) ++lg;  // std::__lg
  stack[sp++] = Frame{first, last, 2 * lg};
  while (sp > 0) {
    Frame f = stack[--sp];
    while (f.last - f.first > 16) {
      if (f.depth == 0) {
        heap_sort_all(f.first, f.last, less);
        break;
      }
      --f.depth;
      T* mid = f.first + (f.last - f.first) / 2;
      move_median_to_first(f.first, f.first + 1, mid, f.last - 1, less);
      T* cut = unguarded_partition(f.first + 1, f.last, f.first, less);
      stack[sp++] = Frame{cut, f.last, f.depth};
      f.last = cut;
    }
  }
  // __final_insertion_sort
  if (n > 16) {
    insertion_sort(first, first + 16, less);
    TALC_ROLLED
    for (T* i = first + 16; i != last; ++i) unguarded_linear_insert(i, less);
  } else
    insertion_sort(first, last, less);
}

// The only instantiation the device code uses: every comparator of the reference is `key(l) < key(r)` on a
// scalar key, so all sorts run through one routine on (key, original index) pairs -- the permutation only
// depends on the comparator outcomes -- and the callers then gather their records by `idx`.
struct SortKey {
  i64 key;
  u32 idx;
};
struct SortKeyLess {
  TALC_HD bool operator()(const SortKey& a, const SortKey& b) const { return a.key < b.key; }
};
TALC_HDN void std_sort_keys_large(SortKey* first, u32 n) { std_sort(first, first + n, SortKeyLess()); }
// up to 16 elements std::sort is its final insertion sort alone (the introsort loop does nothing): the anchor
// lists of every gap attempt take this short path and never touch the introsort code
TALC_HDN void std_sort_keys(SortKey* first, u32 n) {
  if (n > 16) { std_sort_keys_large(first, n); return; }
  if (n > 1) stdsort_detail::insertion_sort(first, first + n, SortKeyLess());
}
// non-negative doubles order like their bit patterns
TALC_HD i64 sort_key_of_nonneg_double(double d) {
  i64 k;
  memcpy(&k, &d, sizeof k);
  return k;
}

}  // namespace talc
