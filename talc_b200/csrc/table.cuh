// table.cuh -- the k-mer count table in HBM (replaces the reference's
// std::map<Dna5String, pair<uint,uint>> : Settings.cpp:26,50; queries Jellyfish.cpp:308-321,
// 383-393,485-496).  Only equality look-ups are ever made, so an open-addressed hash is legal.
//
// Layout: power-of-two array of 16-byte slots {u64 key, u32 count, u32 colour}; two slots
// share one 32-byte DRAM sector and probing walks sector by sector (pair-aligned linear
// probing), so a hit or a miss in the home sector costs exactly one sector fetch.
// Load factor <= 0.5.  Empty slots hold key = ~0 (keys use at most 60 bits).
#pragma once
#include "defs.cuh"

namespace talc {

struct __attribute__((aligned(16))) Slot {
  u64 key;
  u32 count;
  u32 colour;
};

static const u64 kEmptyKey = ~0ull;

struct TableView {
  const Slot* slots;
  u64 mask;  // capacity - 1 (capacity is a power of two >= 2)
};

TALC_HD u64 hash_kmer(u64 k) {  // murmur3 fmix64
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ Slot load_slot(const Slot* p) {
  // one 16-byte read-only load; the two slots of a sector are fetched by two such loads
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  Slot s;
  s.key = ((u64)v.y << 32) | v.x;
  s.count = v.z;
  s.colour = v.w;
  return s;
}
// one 32-byte sector = the two slots of a probe step, fetched by ONE 256-bit read-only load (sm_100a);
// volatile keeps the load where it is written: the callers issue all their sector loads before they look at any
__device__ __forceinline__ void load_sector(const Slot* p, Slot& a, Slot& b) {
  u64 x0, x1, x2, x3;
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x0), "=l"(x1), "=l"(x2), "=l"(x3) : "l"(p));
  a.key = x0;
  a.count = (u32)x1;
  a.colour = (u32)(x1 >> 32);
  b.key = x2;
  b.count = (u32)x3;
  b.colour = (u32)(x3 >> 32);
}
#else
inline Slot load_slot(const Slot* p) { return *p; }
inline void load_sector(const Slot* p, Slot& a, Slot& b) { a = p[0]; b = p[1]; }
#endif

// resolve one key against its already-fetched home sector; returns 1 hit, 0 definite miss, -1 keep probing
TALC_HD int sector_resolve(const Slot& s0, const Slot& s1, u64 key, u32& count, u32& colour) {
  if (s0.key == key) { count = s0.count; colour = s0.colour; return 1; }
  if (s0.key == kEmptyKey) { count = 0; colour = 0; return 0; }
  if (s1.key == key) { count = s1.count; colour = s1.colour; return 1; }
  if (s1.key == kEmptyKey) { count = 0; colour = 0; return 0; }
  return -1;
}

// continue a probe sequence past sector b
// (bounded by the capacity: a table without a free slot -- refused at seal -- cannot hang the GPU)
TALC_HDN bool table_probe_from(const TableView& t, u64 b, u64 key, u32& count, u32& colour) {
  for (u64 n = 0; n <= t.mask; n += 2) {
    b = (b + 2) & t.mask;
    Slot s0, s1;
    load_sector(t.slots + b, s0, s1);
    const int r = sector_resolve(s0, s1, key, count, colour);
    if (r >= 0) return r == 1;
  }
  count = 0;
  colour = 0;
  return false;
}
// the same by value: (colour << 32) | count, 0 when absent (no reference arguments: nothing is forced into local memory)
TALC_HDN u64 table_probe_from_v(const TableView& t, u64 b, u64 key) {
  u32 count = 0, colour = 0;
  table_probe_from(t, b, key, count, colour);
  return ((u64)colour << 32) | (u64)count;
}
// Two consecutive sectors of a probe sequence resolved without branches.  A key cannot sit behind an empty slot of
// its probe sequence (no deletions), so: found anywhere -> that slot; else any empty slot -> absent; else go on.
// Returns false when the probe has to continue past the second sector.
TALC_HD bool resolve4(const Slot& s0, const Slot& s1, const Slot& s2, const Slot& s3, u64 key, u32& count, u32& colour) {
  const bool h0 = s0.key == key, h1 = s1.key == key, h2 = s2.key == key, h3 = s3.key == key;
  const bool e = (s0.key == kEmptyKey) | (s1.key == kEmptyKey) | (s2.key == kEmptyKey) | (s3.key == kEmptyKey);
  count = h0 ? s0.count : h1 ? s1.count : h2 ? s2.count : h3 ? s3.count : 0u;
  colour = h0 ? s0.colour : h1 ? s1.colour : h2 ? s2.colour : h3 ? s3.colour : 0u;
  return h0 | h1 | h2 | h3 | e;
}

// point look-up: (count, colour) or (0,0) when absent (Jellyfish.cpp:317-318,492-493)
TALC_HD bool table_lookup(const TableView& t, u64 key, u32& count, u32& colour) {
  const u64 b = hash_kmer(key) & t.mask & ~1ull;
  Slot s0, s1;
  load_sector(t.slots + b, s0, s1);
  const int r = sector_resolve(s0, s1, key, count, colour);
  if (r >= 0) return r == 1;
  return table_probe_from(t, b, key, count, colour);
}

// the four successor counts in A,C,G,T order (Jellyfish.cpp:308-321).  Device: lane l probes successor l & 3 (the
// warp runs one look-up instead of four, lanes 4..31 mirror lanes 0..3) and four shuffles hand every lane all
// four results.  Host form: the four home sectors one after the other.
#if defined(__CUDA_ARCH__)
__device__ __noinline__ void table_next_counts(const TableView& t, u64 kmer, bool right, u32 K, u32 cnt[4], u32 col[4]) {
  const u32 b = threadIdx.x & 3u;
  u32 c, l;
  table_lookup(t, kmer_next(kmer, b, right, K), c, l);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    cnt[i] = __shfl_sync(0xffffffffu, c, i);
    col[i] = __shfl_sync(0xffffffffu, l, i);
  }
}
#else
TALC_HDN void table_next_counts(const TableView& t, u64 kmer, bool right, u32 K, u32 cnt[4], u32 col[4]) {
  for (u32 i = 0; i < 4; ++i) table_lookup(t, kmer_next(kmer, i, right, K), cnt[i], col[i]);
}
#endif

// the same in two halves, so that a caller can do arithmetic while the eight loads are in flight
struct NextProbe {
  u64 key[4], b[4];
  Slot s0[4], s1[4];
};
TALC_HD void table_next_issue(const TableView& t, u64 kmer, bool right, u32 K, NextProbe& q) {
#pragma unroll
  for (u32 i = 0; i < 4; ++i) {
    q.key[i] = kmer_next(kmer, i, right, K);
    q.b[i] = hash_kmer(q.key[i]) & t.mask & ~1ull;
  }
#pragma unroll
  for (u32 i = 0; i < 4; ++i) load_sector(t.slots + q.b[i], q.s0[i], q.s1[i]);
}
TALC_HD void table_next_resolve(const TableView& t, const NextProbe& q, u32 cnt[4], u32 col[4]) {
#pragma unroll
  for (u32 i = 0; i < 4; ++i) {
    const int r = sector_resolve(q.s0[i], q.s1[i], q.key[i], cnt[i], col[i]);
    if (r < 0) table_probe_from(t, q.b[i], q.key[i], cnt[i], col[i]);
  }
}

// Jellyfish.cpp:383-393
#if defined(__CUDA_ARCH__)
__device__ __noinline__ int table_out_degree(const TableView& t, u64 kmer, bool right, u32 K, u32 min_count) {
  u32 c, l;
  table_lookup(t, kmer_next(kmer, threadIdx.x & 3u, right, K), c, l);
  return __popc(__ballot_sync(0xffffffffu, c >= min_count) & 0xFu);
}
#else
TALC_HDN int table_out_degree(const TableView& t, u64 kmer, bool right, u32 K, u32 min_count) {
  u32 cnt[4], col[4];
  table_next_counts(t, kmer, right, K, cnt, col);
  int d = 0;
  for (u32 b = 0; b < 4; ++b) d += (cnt[b] >= min_count) ? 1 : 0;
  return d;
}
#endif

// ---------------------------------------------------------------------------------------------------------
// Successor tables (device only): the walk never asks for ONE k-mer, it asks for the four successors of a k-mer
// (getNextCountsFromDBG / getOutDegree, Jellyfish.cpp:308-321,383-393).  RIGHT successors x[1..]+b share the
// (K-1)-mer context x[1..], LEFT successors b+x[..K-2] share x[..K-2]; so two more tables, keyed by context --
// one holding, per (K-1)-mer c, the counts of the four k-mers c+b (RIGHT walks), the other those of b+c (LEFT
// walks) -- answer such a question with ONE 32-byte sector instead of four.  They are derived on the device
// from the k-mer table above whenever that one is sealed (built, imported or copied), so the exchange format
// between GPUs stays the k-mer table.  Only "colour > 0" is ever tested on successors (Explorer.cpp:1258): one
// bit per successor.
struct __attribute__((aligned(32))) CtxBucket {
  u64 ctx;      // (K-1)-mer context, kEmptyKey when free
  u32 cnt[4];   // counts of the four successors in A,C,G,T order (0 = absent)
  u32 colmask;  // bit b: successor b carries a junction colour
  u32 pad;
};
struct CtxView {
  const CtxBucket* b;
  u64 mask;  // capacity - 1 (power of two)
};
TALC_HD u64 ctx_of(u64 kmer, bool right, u32 K) { return right ? (kmer & kmer_mask(K - 1)) : (kmer >> 2); }
// bucket index of a context: one multiply and the high half (the walk hashes once per trail and step; a full
// 64-bit finaliser is five dependent multiplies and shifts on its critical path)
TALC_HD u64 hash_ctx(u64 c) {
  c *= 0x9E3779B97F4A7C15ull;
  return c ^ (c >> 32);
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void load_bucket(const CtxBucket* p, u64& ctx, u32 cnt[4], u32& colmask) {
  u64 x0, x1, x2, x3;
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x0), "=l"(x1), "=l"(x2), "=l"(x3) : "l"(p));
  ctx = x0;
  cnt[0] = (u32)x1;
  cnt[1] = (u32)(x1 >> 32);
  cnt[2] = (u32)x2;
  cnt[3] = (u32)(x2 >> 32);
  colmask = (u32)x3;
}
__device__ __noinline__ void ctx_probe_from(const CtxView& v, u64 i, u64 ctx, u32 cnt[4], u32& colmask) {
  for (u64 n = 0; n <= v.mask; ++n) {
    i = (i + 1) & v.mask;
    u64 c;
    load_bucket(v.b + i, c, cnt, colmask);
    if (c == ctx) return;
    if (c == kEmptyKey) break;
  }
  cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0;
  colmask = 0;
}
// the four successor counts of a context: home bucket and the next one are fetched together
__device__ __forceinline__ void ctx_lookup(const CtxView& v, u64 ctx, u32 cnt[4], u32& colmask) {
  const u64 i = hash_ctx(ctx) & v.mask, i2 = (i + 1) & v.mask;
  u64 c0, c1;
  u32 n1[4], m1;
  load_bucket(v.b + i, c0, cnt, colmask);
  load_bucket(v.b + i2, c1, n1, m1);
  if (c0 == ctx) return;
  const bool second = (c0 != kEmptyKey) & (c1 == ctx);
  const bool none = (c0 == kEmptyKey) | (c1 == kEmptyKey);
#pragma unroll
  for (int q = 0; q < 4; ++q) cnt[q] = second ? n1[q] : 0u;
  colmask = second ? m1 : 0u;
  if (!second && !none) ctx_probe_from(v, i2, ctx, cnt, colmask);
}
#endif

}  // namespace talc
