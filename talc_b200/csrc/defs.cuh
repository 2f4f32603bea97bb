// defs.cuh -- shared definitions for the TALC correction path on B200.
//
// Everything marked TALC_HD is plain C++ without warp intrinsics so that the same
// source compiles (a) with nvcc for sm_100a, where it is the product, and (b) with
// g++ inside tests/hostemu, where it is only a debugging aid that lets the logic be
// diffed against the oracle without a GPU.  The shipped library has no CPU path.
#pragma once
#include <stdint.h>
#include <string.h>

// TALC_HD: small helpers, always inlined.  TALC_HDN: the big routines; they stay real calls on the device
// (forcing the whole per-read pipeline into one function body makes the NVVM optimiser run for an hour).
#if defined(__CUDACC__)
#define TALC_HD __host__ __device__ __forceinline__
#define TALC_HDN __host__ __device__ __noinline__
#else
#define TALC_HD inline
#define TALC_HDN inline
#endif

namespace talc {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;

// Tunables of the reference: Settings.cpp:33-69 (CLI-settable) and the file-static
// constants of Explorer.cpp:85-102 / Jellyfish.cpp:64 / Read.cpp:361,368.
struct Params {
  u32 K;
  u32 min_count;        // gp_MIN_COUNT
  u32 window;           // gp_WINDOW_SIZE
  u32 max_branches;     // gp_MAX_NB_COMPETING_PATHS
  double alpha;         // gp_ALPHA
  double sr_error;      // gp_SR_ERROR_RATE
  double min_inner;     // gp_MIN_INNER_SCORE
  double min_border;    // gp_MIN_BORDER_SCORE
  i32 cycle_mode;       // 0: SeqAn2 Horspool over mixed alphabets (stride-K windows); 1: exact first occurrence
  i32 q11_zero;         // Explorer.cpp:705 UB loop counter modelled as j = 0
};

enum : u32 {
  kMinStartAnchors = 3,      // Explorer.cpp:85
  kMaxStartAnchors = 5,      // :86
  kMaxInCount = 100000,      // :88
  kMaxBorderPaths = 75,      // :90
  kMaxInnerPaths = 50,       // :91
  kCheckInterval = 6,        // :93
  kMaxBorderFailures = 3,    // :97
  kBorderMaxLen = 500,       // Read.cpp:361,368
  kColouredCountThr = 10000  // Jellyfish.cpp:64
};

// per-read outcome (main.cpp:262-296)
enum ReadStatus : u8 {
  kReadOk = 0,
  kReadNoSolid = 1,      // "No solid kmer could be found."
  kReadNoStructure = 2,  // "Unable to define convenient structure."
  kReadShort = 3,        // len <= K : untouched, not logged
  kReadOverflow = 250    // internal: scratch arena too small, re-run with a larger arena
};

// Dna5 code of an input character (SeqAn char -> Dna5): ACGT either case, U -> T, else N(4)
TALC_HD u32 base_code(u8 c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: return 4;
  }
}
TALC_HD u8 code_char(u32 code) { return (u8)("ACGTN"[code]); }

TALC_HD u64 kmer_mask(u32 K) { return (K >= 32) ? ~0ull : ((1ull << (2 * K)) - 1ull); }

// k-mers are 2-bit packed with the first base in the most significant position.
// Successors (utils.cpp:370-387): RIGHT x[1..]+b ; LEFT b+x[..K-2]
TALC_HD u64 kmer_next(u64 kmer, u32 b, bool right, u32 K) {
  return right ? (((kmer << 2) | (u64)b) & kmer_mask(K)) : ((kmer >> 2) | ((u64)b << (2 * (K - 1))));
}

// ---- bump arena over a per-thread slice of global scratch --------------------
struct Arena {
  u8* base;
  u32 cap;
  u32 top;
  u32 overflow;
  TALC_HD void init(u8* b, u32 c) { base = b; cap = c; top = 0; overflow = 0; }
  TALC_HD u32 mark() const { return top; }
  TALC_HD void release(u32 m) { top = m; }
  // returns nullptr (and latches overflow) when the slice is exhausted
  TALC_HD void* alloc(u32 bytes) {
    u32 t = (top + 7u) & ~7u;
    if (bytes > cap || t > cap - bytes) { overflow = 1; return nullptr; }
    top = t + bytes;
    return base + t;
  }
};

// ---- a read as the kernels see it: ASCII bytes in HBM --------------------------
struct ReadView {
  const u8* s;  // ASCII
  u32 len;
  TALC_HD u32 code(u32 i) const { return base_code(s[i]); }
  // packed k-mer starting at base i; ok=false if it contains N
  TALC_HD u64 kmer_at(u32 i, u32 K, bool& ok) const {
    u64 v = 0;
    ok = true;
    for (u32 j = 0; j < K; ++j) {
      u32 c = code(i + j);
      if (c > 3) { ok = false; c = 0; }
      v = (v << 2) | c;
    }
    return v;
  }
};

// The reference string of a search in *walk order*: element i is read[start + i] for a
// RIGHT-ward walk and read[start - i] for a LEFT-ward walk.  Every reference string the
// reference builds (anchor+gap+target, target+gap+anchor, anchor+border, border+anchor:
// Explorer.cpp:925-938,1042-1053) is a contiguous slice of the raw read, so no copy is made.
struct RefView {
  const u8* s;
  i32 start;
  i32 step;  // +1 or -1
  u32 len;
  TALC_HD u32 code(u32 i) const { return base_code(s[start + (i32)i * step]); }
};

// A trail's sequence in walk order, 2-bit packed (32 bases per u64, first base most significant).
// Paths only ever contain A,C,G,T (anchors come from the table, successors from kDict).
struct PathView {
  const u64* w;
  u32 len;
  TALC_HD u32 code(u32 i) const { return (u32)((w[i >> 5] >> (62 - 2 * (i & 31))) & 3ull); }
};
TALC_HD void path_set(u64* w, u32 i, u32 c) {
  const u32 sh = 62 - 2 * (i & 31);
  w[i >> 5] = (w[i >> 5] & ~(3ull << sh)) | ((u64)c << sh);
}
// k-mer (first base most significant) whose first base is walk index i
TALC_HD u64 path_kmer_fwd(const u64* w, u32 i, u32 K) {
  const u32 wi = i >> 5, off = 2 * (i & 31);
  u64 hi = w[wi] << off;
  if (off && (off + 2 * K > 64)) hi |= w[wi + 1] >> (64 - off);
  return hi >> (64 - 2 * K);
}

}  // namespace talc
