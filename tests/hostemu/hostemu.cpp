// hostemu.cpp -- compiles the product's TALC_HD device code with g++ so that its logic can be
// diffed against the oracle on a machine without a GPU.  TEST AID ONLY: nothing in the shipped
// library links this file, and the product has no CPU path.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../talc_b200/csrc/correct.cuh"

using namespace talc;

struct EmuTable {
  std::vector<Slot> slots;
  u64 mask;
  u32 K;
};

static Slot* emu_find(EmuTable* t, u64 key, bool insert) {
  u64 b = hash_kmer(key) & t->mask & ~1ull;
  for (;;) {
    for (int i = 0; i < 2; ++i) {
      Slot& s = t->slots[b + i];
      if (s.key == key) return &s;
      if (s.key == kEmptyKey) {
        if (!insert) return nullptr;
        s.key = key;
        s.count = 0;
        s.colour = 0;
        return &s;
      }
    }
    b = (b + 2) & t->mask;
  }
}

static u64 revcomp(u64 k, u32 K) {
  u64 r = 0;
  for (u32 i = 0; i < K; ++i) {
    r = (r << 2) | (3 - (k & 3));
    k >>= 2;
  }
  return r;
}

extern "C" {

struct emu_params {
  uint32_t K, MIN_COUNT, WINDOW_SIZE, MAX_NB_COMPETING_PATHS;
  double ALPHA, SR_ERROR_RATE, MIN_INNER_SCORE, MIN_BORDER_SCORE;
  int32_t cycle_mode, q11_zero_init;
};
static Params to_params(const emu_params* q) {
  Params p;
  p.K = q->K; p.min_count = q->MIN_COUNT; p.window = q->WINDOW_SIZE; p.max_branches = q->MAX_NB_COMPETING_PATHS;
  p.alpha = q->ALPHA; p.sr_error = q->SR_ERROR_RATE; p.min_inner = q->MIN_INNER_SCORE; p.min_border = q->MIN_BORDER_SCORE;
  p.cycle_mode = q->cycle_mode; p.q11_zero = q->q11_zero_init;
  return p;
}

// sequential build with the reference's semantics (Jellyfish.cpp:236-295, utils.cpp:658-669)
void* emu_table_build(const emu_params* q, const uint64_t* keys, const int64_t* counts, uint64_t n, const uint64_t* jkeys,
                      const int64_t* jcounts, uint64_t nj, int use_junctions) {
  EmuTable* t = new EmuTable;
  t->K = q->K;
  u64 cap = 2;
  while (cap < 2 * n + 2) cap <<= 1;
  t->slots.assign(cap, Slot{kEmptyKey, 0, 0});
  t->mask = cap - 1;
  for (u64 i = 0; i < n; ++i) {
    if ((u32)(int)counts[i] >= q->MIN_COUNT) {
      if (!emu_find(t, keys[i], false)) emu_find(t, keys[i], true)->count = (u32)(int)counts[i];
    }
  }
  if (use_junctions) {
    for (u64 i = 0; i < nj; ++i) {
      if ((u32)(int)jcounts[i] < kColouredCountThr) {
        if (Slot* s = emu_find(t, jkeys[i], false)) s->colour = (u32)(int)jcounts[i];
        if (Slot* s = emu_find(t, revcomp(jkeys[i], q->K), false)) s->colour = (u32)(int)jcounts[i];
      }
    }
  }
  for (u32 b = 0; b < 4; ++b) {
    u64 homo = 0;
    for (u32 i = 0; i < q->K; ++i) homo = (homo << 2) | b;
    if (Slot* s = emu_find(t, homo, false)) s->colour = 0;
  }
  return t;
}
void emu_table_free(void* t) { delete (EmuTable*)t; }

int emu_correct_reads(void* tab, const emu_params* q, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads,
                      uint32_t arena_bytes, int wide, uint8_t* out, uint64_t out_capacity, uint64_t* out_offsets,
                      uint8_t* status, uint64_t* counters /* kNumCounters */) {
  EmuTable* t = (EmuTable*)tab;
  Params P = to_params(q);
  TableView tv;
  tv.slots = t->slots.data();
  tv.mask = t->mask;
  Counters total;
  memset(&total, 0, sizeof(total));
  std::vector<u8> arena(arena_bytes);
  std::vector<u32> cov;
  u64 pos = 0;
  out_offsets[0] = 0;
  for (u32 r = 0; r < n_reads; ++r) {
    ReadView rd;
    rd.s = bases + offsets[r];
    rd.len = (u32)(offsets[r + 1] - offsets[r]);
    // coverage (what the coverage kernel produces)
    cov.clear();
    if (rd.len >= P.K) {
      for (u32 i = 0; i + P.K <= rd.len; ++i) {
        bool ok;
        u64 km = rd.kmer_at(i, P.K, ok);
        u32 c = 0, cl = 0;
        if (ok) table_lookup(tv, km, c, cl);
        cov.push_back(c);
      }
    }
    Corrector* cx = new Corrector;
    cx->T = tv;
    cx->P = P;
    cx->tabs.n = 0;
    cx->tabs.lower = cx->tabs.upper = cx->tabs.sq = nullptr;
    Counters mine;
    memset(&mine, 0, sizeof(mine));
    cx->ctr = &mine;
    ReadJob job;
    job.rd = rd;
    job.cov = cov.data();
    job.arena = arena.data();
    job.arena_bytes = arena_bytes;
    job.wide = wide != 0;
    u8 st = cx->run(job);
    mine.cells_nw += cx->dps.cells_nw;
    mine.cells_lcs += cx->dps.cells_lcs;
    mine.cells_ovl += cx->dps.cells_ovl;
    mine.cells_xdrop += cx->dps.cells_xdrop;
    if (st == kReadOverflow) {
      total.reads_overflow++;
    } else {
      const u64* src = (const u64*)&mine;
      u64* dst = (u64*)&total;
      for (int i = 0; i < kNumCounters; ++i) dst[i] += src[i];
      if (st == kReadOk) total.reads_ok++;
    }
    status[r] = st;
    u32 olen;
    if (st == kReadOk) olen = cx->corrected_length();
    else olen = rd.len;
    if (pos + olen > out_capacity) { delete cx; return -1; }
    if (st == kReadOk) cx->emit(out + pos, 0, 1);
    else for (u32 i = 0; i < rd.len; ++i) out[pos + i] = code_char(rd.code(i));
    pos += olen;
    if (st != kReadOverflow) total.bases_out += olen;
    out_offsets[r + 1] = pos;
    delete cx;
  }
  memcpy(counters, &total, sizeof(total));
  return 0;
}

int emu_num_counters() { return kNumCounters; }

// ---- primitives
static SeqView bytes_view(const char* s, u32 n) {
  SeqView v;
  v.s = (const u8*)s; v.start = 0; v.step = 1; v.w = nullptr; v.len = n;
  return v;
}
static std::vector<u64> pack(const char* s, u32 n) {
  std::vector<u64> w((n + 31) / 32 + 2, 0);
  for (u32 i = 0; i < n; ++i) path_set(w.data(), i, base_code((u8)s[i]) & 3);
  return w;
}
// packed != 0: second sequence is given to the kernel as a packed trail (must be ACGT only)
int emu_nw(const char* a, const char* b, int packed) {
  std::vector<u8> ar(1 << 20);
  Arena A; A.init(ar.data(), (u32)ar.size());
  u32 an = strlen(a), bn = strlen(b);
  std::vector<u64> w = pack(b, bn);
  SeqView vb = packed ? view_of_path(w.data(), bn) : bytes_view(b, bn);
  return -nw_distance(bytes_view(a, an), an, vb, bn, A, nullptr);
}
int emu_lcs(const char* a, const char* b, int packed) {
  std::vector<u8> ar(1 << 20);
  Arena A; A.init(ar.data(), (u32)ar.size());
  u32 an = strlen(a), bn = strlen(b);
  std::vector<u64> w = pack(b, bn);
  SeqView vb = packed ? view_of_path(w.data(), bn) : bytes_view(b, bn);
  return lcs_length(bytes_view(a, an), an, vb, bn, A, nullptr);
}
// walk-order overlap score == the oracle's RIGHT mode; LEFT mode is the same on reversed strings
int emu_overlap(const char* ref, const char* cand) {
  std::vector<u8> ar(1 << 22);
  Arena A; A.init(ar.data(), (u32)ar.size());
  u32 an = strlen(ref), bn = strlen(cand);
  return overlap_score(bytes_view(ref, an), an, bytes_view(cand, bn), bn, A, nullptr);
}
void emu_xdrop(const char* query_seg, const char* database_seg, int xdrop, int wide, uint64_t* ext_rows, uint64_t* ext_cols,
               int* overflow) {
  std::vector<u8> ar(1 << 22);
  Arena A; A.init(ar.data(), (u32)ar.size());
  u32 qn = strlen(query_seg), dn = strlen(database_seg);
  u32 er = 0, ec = 0;
  xdrop_extend(bytes_view(query_seg, qn), 0, qn, bytes_view(database_seg, dn), 0, dn, xdrop, er, ec, A, wide != 0, nullptr);
  *ext_rows = er; *ext_cols = ec; *overflow = (int)A.overflow;
}
// getSeedAndExtension on walk-order strings (RIGHT as is; for LEFT pass reversed strings and right=0)
void emu_seed_extend(const char* reference, const char* candidate, int xdrop, int right, uint32_t K, int32_t* ref_ext,
                     int32_t* cand_ext, int32_t* score, int32_t* stop) {
  std::vector<u8> ar(1 << 22);
  Arena A; A.init(ar.data(), (u32)ar.size());
  u32 rn = strlen(reference), cn = strlen(candidate);
  SeedExt e = seed_and_extension(bytes_view(reference, rn), bytes_view(candidate, cn), xdrop, right != 0, K, A, true, nullptr);
  *ref_ext = e.ref_ext; *cand_ext = e.cand_ext; *score = e.score; *stop = e.stop ? 1 : 0;
}
void emu_std_sort_perm(const int64_t* keys, uint32_t n, uint32_t* perm) {
  for (u32 i = 0; i < n; ++i) perm[i] = i;
  std_sort(perm, perm + n, [keys](const u32& a, const u32& b) { return keys[a] < keys[b]; });
}
int emu_tag_next_nodes(const emu_params* q, const uint32_t* counts4, const uint32_t* colours4, uint32_t count, int complex_,
                       int32_t* tags4, double* dist4) {
  Params P = to_params(q);
  u8 tag[4] = {0, 0, 0, 0};
  const StepBounds sb = step_bounds(count, P);
  int n = tag_next_nodes(counts4, colours4, sb, P, complex_ != 0, tag);
  for (int i = 0; i < n; ++i) { tags4[i] = tag[i] == kExpected ? 0 : tag[i] == kUnexpected ? 1 : 7; dist4[i] = step_dist(count, counts4[i], sb); }
  return n;
}
}
