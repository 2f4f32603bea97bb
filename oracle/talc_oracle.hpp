// =============================================================================
// talc_oracle.hpp -- CPU restatement of TALC's correction hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, link, import or execute it, and only as the checker / reported
// CPU baseline.  The product (talc_b200/) never includes or links this file.
//
// PARITY UNPINNED: the reference (lbroseus/TALC, /root/reference) cannot be built
// here (it needs SeqAn2 headers that are neither vendored nor installed, and it
// ships no tests, fixtures or golden vectors).  This file restates src/*.cpp of
// the reference function by function (each function cites the file:line it
// follows) and restates the published algorithms of the SeqAn 2.x routines the
// reference calls (globalAlignment, localAlignment, extendSeed/GappedXDrop,
// Finder/Pattern<Horspool>, Dna5 conversion, reverseComplement) from their
// documented behaviour.  See DESIGN.md "Oracle" for the list of recalled-not-
// verified points and the switches that cover them.
//
// Strings are Dna5 sequences stored as upper-case chars over {A,C,G,T,N}.
// =============================================================================
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <tuple>
#include <unordered_map>
#include <utility>
#include <vector>

namespace talc_oracle {

typedef std::string Seq;
typedef std::pair<unsigned int, unsigned int> CCount;  // (SR count, junction colour)

// utils.hpp:51-57 (only the values that are live on the hot path matter; the
// numeric order is kept because Status values are compared for (in)equality only)
enum Status { EXPECTED, UNEXPECTED, LOWCOUNT, SUPPORTED, CORRECTED, UNCORRECTED, ABSENT, BREAKPOINT };
enum Direction { LEFT, RIGHT };
enum Location { HEAD, INNER, TAIL, UNKNOWN };

// Settings.cpp:33-69 + file-static tunables Explorer.cpp:85-102, Jellyfish.cpp:64
struct Params {
  unsigned int K = 21;
  unsigned int MIN_COUNT = 2;            // gp_MIN_COUNT
  double ALPHA = 2.57;                   // gp_ALPHA
  unsigned int WINDOW_SIZE = 9;          // gp_WINDOW_SIZE
  double SR_ERROR_RATE = 0.025;          // gp_SR_ERROR_RATE
  double MIN_INNER_SCORE = 0.7;          // gp_MIN_INNER_SCORE
  double MIN_BORDER_SCORE = 0.7;         // gp_MIN_BORDER_SCORE
  unsigned int MAX_NB_COMPETING_PATHS = 7;  // gp_MAX_NB_COMPETING_PATHS
  // ---- switches for points that cannot be verified without SeqAn / the binary
  // 0: SeqAn-2.x Pattern<CharString,Horspool> over a Dna5 haystack: the bad-
  //    character table is indexed by the needle's *char* ordinals while the
  //    haystack is looked up by its *Dna5* ordinal, so every shift is the full
  //    needle length: windows 0,K,2K,... only.  1: exact first occurrence.
  int cycle_mode = 0;
  // Explorer.cpp:705 uninitialised loop counter; g++ -O3 behaves as j=0 (true).
  bool q11_zero_init = true;
};

// ---- instrumentation: properties of the algorithm on a given input
struct Counters {
  uint64_t reads = 0, bases_in = 0, bases_out = 0;
  uint64_t reads_short = 0, reads_nosolid = 0, reads_nostruct = 0, reads_corrected = 0;
  uint64_t lookups_seg = 0;   // W_seg  : L-K+1 per read           (Jellyfish.cpp:490)
  uint64_t lookups_deg = 0;   // W_deg  : 4 per getOutDegree call  (Jellyfish.cpp:383)
  uint64_t lookups_walk = 0;  // W_walk : 4 per whatsNext call     (Explorer.cpp:564,635)
  uint64_t steps_inner = 0, steps_border = 0;
  uint64_t frontier_sum = 0, frontier_max = 0;
  uint64_t cells_nw = 0, cells_lcs = 0, cells_ovl = 0, cells_xdrop = 0;
  uint64_t calls_nw = 0, calls_lcs = 0, calls_ovl = 0, calls_xdrop = 0;
  uint64_t gaps = 0, gaps_bridged = 0, gap_attempts = 0;
  uint64_t borders = 0, borders_corrected = 0, border_cutoff_500 = 0;
  uint64_t ev_frontier_over50 = 0, ev_maxlength = 0, ev_gardening = 0, ev_garden_ties = 0,
           ev_garden_q16 = 0, ev_cycle = 0, ev_bridge = 0, ev_edge = 0, ev_q9 = 0, ev_anchor_overlap = 0,
           ev_bug_alignment = 0, ev_sort_gt16 = 0;
  void add(const Counters& o);
  std::string json() const;
};

// ---- the k-mer table (Settings.cpp:26,50: std::map<Dna5String, pair<uint,uint>>)
// Only equality lookups are used by the reference, so the hashed flavour is a
// legal (faster) stand-in for tests; the ordered flavour mirrors the reference's
// cost and is what the CPU baseline is timed with.
class Table {
 public:
  explicit Table(bool ordered = false) : ordered_(ordered) {}
  bool ordered() const { return ordered_; }
  size_t size() const { return ordered_ ? m_.size() : h_.size(); }
  bool insert_first_wins(const Seq& k, CCount v);
  CCount* find(const Seq& k);
  const CCount* find(const Seq& k) const { return const_cast<Table*>(this)->find(k); }
  // enumerate (for dumping the table to the device-side tests)
  template <class F> void for_each(F f) const {
    if (ordered_) for (auto& kv : m_) f(kv.first, kv.second);
    else for (auto& kv : h_) f(kv.first, kv.second);
  }
 private:
  bool ordered_;
  std::map<Seq, CCount> m_;
  std::unordered_map<Seq, CCount> h_;
};

// Dna5 conversion (SeqAn TranslateTableCharToDna5_): A,C,G,T (either case), U->T, else N
char to_dna5(char c);
Seq to_dna5(const std::string& s);
Seq reverse_complement(const Seq& s);

// Jellyfish.cpp:236-295 + utils.cpp:658-669.  Returns #lines evaluated; n_kept via table.size().
struct BuildStats { uint64_t lines = 0, kept = 0, bad_lines = 0, junction_lines = 0; };
BuildStats build_cdbg(Table& t, const Params& p, const std::string& dump_path, const std::string& junction_path,
                      bool use_junctions);
// same semantics on in-memory (kmer,count) lists, in list order
void build_cdbg_from_lists(Table& t, const Params& p, const std::vector<std::pair<std::string, long long>>& kmers,
                           const std::vector<std::pair<std::string, long long>>& junctions, bool use_junctions);
void decolour_repeats(Table& t, unsigned int K);  // utils.cpp:658-669

// ---- alignment primitives
int nw_score(const Seq& a, const Seq& b, Counters* c = nullptr);          // globalAlignment Score(0,-1,-1)
int lcs_score(const Seq& a, const Seq& b, Counters* c = nullptr);         // localAlignment  Score(1,0,0)
int overlap_score(const Seq& ref, const Seq& cand, Direction d, Counters* c = nullptr);  // Trail.cpp:145-174
// SeqAn extendSeed(..., GappedXDrop) for Score(0,-1,-1) on the given segments.
// Returns extension along the database/H segment (rows) and the query/V segment (cols).
void xdrop_extend(const Seq& query_seg, const Seq& database_seg, bool extend_left, int xdrop, size_t& ext_rows,
                  size_t& ext_cols, Counters* c = nullptr);
// Trail.cpp:289-302 (Finder + Pattern<CharString,Horspool>): first match position or -1
long horspool_first(const Seq& haystack, const Seq& needle, int cycle_mode);

// Explorer.cpp:1185-1217
bool isExpectedbyMyModel(unsigned int nextc, unsigned int cc, const Params& p, Status classe);
bool isExpectedbyMyLastNode(unsigned int nextc, unsigned int cc, const Params& p);
// Explorer.cpp:1226-1298
void tagNextNodes(std::vector<std::pair<Status, double>>& tags, const std::vector<CCount>& nextCounts, unsigned int count,
                  const Params& p, bool complex_);

// Trail.cpp:341-437
std::tuple<Seq, Seq, int, double, bool> getSeedAndExtension(const Seq& reference, const Seq& candidate, int xdrop,
                                                           Direction d, unsigned int seedSize, Counters* c = nullptr);

// ---- per-read driver (main.cpp:258-296)
enum ReadStatus { READ_OK = 0, READ_NO_SOLID = 1, READ_NO_STRUCTURE = 2, READ_SHORT = 3 };
struct ReadResult {
  ReadStatus status = READ_OK;
  Seq corrected;  // == raw sequence unless status == READ_OK
  // what outputBasicReadStats (Read.cpp:418-433; call disabled at main.cpp:305) would print for this read:
  // span and number of m_InKmersPositions as defineStructure2 / correct2 left them
  unsigned stat_span = 0, stat_regions = 0;
};
// optional stage dump for differential debugging of the CUDA path
struct StageDump {
  std::vector<CCount> coverage;
  double threshold = 0;
  std::vector<std::tuple<unsigned, unsigned, int>> regions_found, regions_final;
  std::string trace;
};
ReadResult correct_read(const Seq& raw_dna5, const Table& t, const Params& p, Counters& c, StageDump* dump = nullptr);

// std::sort permutation probe (libstdc++ introsort) used by the device sort-replica test:
// sorts indices 0..n-1 by keys[idx] with operator< on the key and returns the permutation.
std::vector<unsigned> std_sort_permutation(const std::vector<long long>& keys);

}  // namespace talc_oracle
