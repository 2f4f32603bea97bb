// =============================================================================
// talc_oracle.cpp -- CPU restatement of TALC's correction hot path.
// TEST INFRASTRUCTURE ONLY (see talc_oracle.hpp header).  PARITY UNPINNED.
//
// All file:line citations are relative to /root/reference/src/.
// Build flags mirror the reference Makefile:3 (-O3, no -march, no -ffast-math) and
// add -ffp-contract=off so double arithmetic matches the reference's x86-64 SSE2
// code (mulsd/addsd/sqrtsd, no FMA) -- decisions depend on it.
// =============================================================================
#include "talc_oracle.hpp"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

namespace talc_oracle {

// ----------------------------------------------------------------------------
// Counters
// ----------------------------------------------------------------------------
#define TALC_COUNTER_FIELDS(X)                                                                                       \
  X(reads) X(bases_in) X(bases_out) X(reads_short) X(reads_nosolid) X(reads_nostruct) X(reads_corrected)            \
  X(lookups_seg) X(lookups_deg) X(lookups_walk) X(steps_inner) X(steps_border) X(frontier_sum)                      \
  X(cells_nw) X(cells_lcs) X(cells_ovl) X(cells_xdrop) X(calls_nw) X(calls_lcs) X(calls_ovl) X(calls_xdrop)         \
  X(gaps) X(gaps_bridged) X(gap_attempts) X(borders) X(borders_corrected) X(border_cutoff_500)                      \
  X(ev_frontier_over50) X(ev_maxlength) X(ev_gardening) X(ev_garden_ties) X(ev_garden_q16) X(ev_cycle)              \
  X(ev_bridge) X(ev_edge) X(ev_q9) X(ev_anchor_overlap) X(ev_bug_alignment) X(ev_sort_gt16)

void Counters::add(const Counters& o) {
#define X(f) f += o.f;
  TALC_COUNTER_FIELDS(X)
#undef X
  frontier_max = std::max(frontier_max, o.frontier_max);
}

std::string Counters::json() const {
  std::ostringstream os;
  os << "{";
#define X(f) os << "\"" #f "\": " << f << ", ";
  TALC_COUNTER_FIELDS(X)
#undef X
  os << "\"frontier_max\": " << frontier_max << "}";
  return os.str();
}

// ----------------------------------------------------------------------------
// Table
// ----------------------------------------------------------------------------
bool Table::insert_first_wins(const Seq& k, CCount v) {
  if (ordered_) return m_.insert(std::make_pair(k, v)).second;
  return h_.insert(std::make_pair(k, v)).second;
}
CCount* Table::find(const Seq& k) {
  if (ordered_) {
    auto it = m_.find(k);
    return it == m_.end() ? nullptr : &it->second;
  }
  auto it = h_.find(k);
  return it == h_.end() ? nullptr : &it->second;
}

// ----------------------------------------------------------------------------
// Dna5 helpers (SeqAn alphabet conversion)
// ----------------------------------------------------------------------------
char to_dna5(char c) {
  switch (c) {
    case 'A': case 'a': return 'A';
    case 'C': case 'c': return 'C';
    case 'G': case 'g': return 'G';
    case 'T': case 't': case 'U': case 'u': return 'T';
    default: return 'N';
  }
}
Seq to_dna5(const std::string& s) {
  Seq r(s.size(), 'N');
  for (size_t i = 0; i < s.size(); ++i) r[i] = to_dna5(s[i]);
  return r;
}
Seq reverse_complement(const Seq& s) {
  Seq r(s.size(), 'N');
  for (size_t i = 0; i < s.size(); ++i) {
    char c = s[s.size() - 1 - i];
    r[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
  }
  return r;
}

// ----------------------------------------------------------------------------
// Table build: Jellyfish.cpp:236-295
// ----------------------------------------------------------------------------
static const unsigned int kColouredCountThr = 10000;  // Jellyfish.cpp:64
static const char kDict[4] = {'A', 'C', 'G', 'T'};     // Jellyfish.cpp:63, Explorer.cpp:104

// one dump line: Jellyfish.cpp:257-264
static inline void ingest_count(Table& t, const Params& p, const std::string& kmer0, long long count, BuildStats* st) {
  Seq kmer = to_dna5(kmer0);                             // :259 TSeq kmer = kmer0
  // :260  std::stoi(count) >= gp_MIN_COUNT  (int vs unsigned: compared as unsigned)
  if ((unsigned int)(int)count >= p.MIN_COUNT) {
    t.insert_first_wins(kmer, std::make_pair((unsigned int)(int)count, 0u));  // :262 map::insert -> first wins
    if (st) st->kept++;                                  // :263 (counts attempts, as the reference does)
  }
  if (st) st->lines++;
}
// one junction line: Jellyfish.cpp:283-286
static inline void ingest_junction(Table& t, const std::string& jmer0, long long jcount) {
  Seq jmer = to_dna5(jmer0);
  bool below = (unsigned int)(int)jcount < kColouredCountThr;  // int vs unsigned compare
  if (below) {
    if (CCount* e = t.find(jmer)) e->second = (unsigned int)(int)jcount;
  }
  Seq rc = reverse_complement(jmer);                     // :285
  if (below) {
    if (CCount* e = t.find(rc)) e->second = (unsigned int)(int)jcount;
  }
}

void decolour_repeats(Table& t, unsigned int K) {        // utils.cpp:658-669
  for (int b = 0; b < 4; ++b) {
    Seq homo(K, kDict[b]);
    if (CCount* e = t.find(homo)) e->second = 0;
  }
}

BuildStats build_cdbg(Table& t, const Params& p, const std::string& dump_path, const std::string& junction_path,
                      bool use_junctions) {
  BuildStats st;
  std::ifstream in(dump_path);
  std::string line;
  while (std::getline(in, line)) {                       // :251
    std::istringstream iss(line);
    std::string kmer0, count;
    if (iss >> kmer0 >> count) {                         // :257
      long long c = 0;
      try { c = std::stoi(count); } catch (...) { st.bad_lines++; continue; }  // reference would terminate
      ingest_count(t, p, kmer0, c, &st);
    } else {
      st.bad_lines++;                                    // :268 prints and skips
    }
  }
  in.close();
  if (use_junctions) {                                   // :273
    std::ifstream jn(junction_path);
    std::string jmer0, jcount;
    while (std::getline(jn, line)) {
      std::istringstream iss2(line);
      if (iss2 >> jmer0 >> jcount) {                     // :281
        long long c = 0;
        try { c = std::stoi(jcount); } catch (...) { st.bad_lines++; continue; }
        ingest_junction(t, jmer0, c);
        st.junction_lines++;
      } else {
        st.bad_lines++;
      }
    }
  }
  decolour_repeats(t, p.K);                              // main.cpp:232
  return st;
}

void build_cdbg_from_lists(Table& t, const Params& p, const std::vector<std::pair<std::string, long long>>& kmers,
                           const std::vector<std::pair<std::string, long long>>& junctions, bool use_junctions) {
  for (auto& kc : kmers) ingest_count(t, p, kc.first, kc.second, nullptr);
  if (use_junctions)
    for (auto& jc : junctions) ingest_junction(t, jc.first, jc.second);
  decolour_repeats(t, p.K);
}

// ----------------------------------------------------------------------------
// Table queries: utils.cpp:370-387, Jellyfish.cpp:116-126,308-321,383-393,485-496
// ----------------------------------------------------------------------------
static Seq formNextKmer(const Seq& kmer, char new_base, Direction direction) {  // utils.cpp:370-387
  if (direction == RIGHT) return kmer.substr(1) + new_base;
  return std::string(1, new_base) + kmer.substr(0, kmer.size() - 1);
}

static std::vector<CCount> getNextCounts(const Seq& kmer, Direction direction, const Table& t) {  // Jellyfish.cpp:308-321
  std::vector<CCount> cc;
  for (int b = 0; b < 4; ++b) {
    const CCount* e = t.find(formNextKmer(kmer, kDict[b], direction));
    cc.push_back(e ? *e : std::make_pair(0u, 0u));
  }
  return cc;
}

static int getOutDegree(const Seq& kmer, Direction direction, const Table& t, const Params& p, Counters& c) {  // :383-393
  std::vector<CCount> next = getNextCounts(kmer, direction, t);
  c.lookups_deg += 4;
  int deg = 0;
  for (size_t i = 0; i < next.size(); ++i)
    if (next[i].first >= p.MIN_COUNT) ++deg;
  return deg;
}

static std::vector<CCount> getLRCountsInSR(const Seq& seq, unsigned int K, const Table& t, Counters& c) {  // :69-82,485-496
  std::vector<CCount> counts;
  unsigned readLength = (unsigned)seq.size();
  for (unsigned start = 0; start < readLength - K + 1; ++start) {
    const CCount* e = t.find(seq.substr(start, K));
    counts.push_back(e ? *e : std::make_pair(0u, 0u));
  }
  c.lookups_seg += counts.size();
  return counts;
}

// ----------------------------------------------------------------------------
// Alignment primitives.  N is a 5th symbol equal to itself.
// ----------------------------------------------------------------------------
int nw_score(const Seq& a, const Seq& b, Counters* c) {  // globalAlignment(.., Score(0,-1,-1)) : -(edit distance)
  const size_t n = a.size(), m = b.size();
  if (c) { c->cells_nw += (uint64_t)n * m; c->calls_nw++; }
  std::vector<int> row(m + 1);
  for (size_t j = 0; j <= m; ++j) row[j] = (int)j;
  for (size_t i = 1; i <= n; ++i) {
    int diag = row[0];
    row[0] = (int)i;
    for (size_t j = 1; j <= m; ++j) {
      int up = row[j];
      int v = std::min(std::min(up + 1, row[j - 1] + 1), diag + (a[i - 1] == b[j - 1] ? 0 : 1));
      diag = up;
      row[j] = v;
    }
  }
  return -row[m];
}

int lcs_score(const Seq& a, const Seq& b, Counters* c) {  // localAlignment(.., Score(1,0,0)) : LCS length
  const size_t n = a.size(), m = b.size();
  if (c) { c->cells_lcs += (uint64_t)n * m; c->calls_lcs++; }
  std::vector<int> row(m + 1, 0);
  for (size_t i = 1; i <= n; ++i) {
    int diag = 0;
    for (size_t j = 1; j <= m; ++j) {
      int up = row[j];
      int v = std::max(std::max(up, row[j - 1]), diag + (a[i - 1] == b[j - 1] ? 1 : 0));
      diag = up;
      row[j] = v;
    }
  }
  return row[m];
}

// Trail.cpp:145-174: globalAlignment(alignG, Score(4,-3,-2), AlignConfig<..>, LinearGaps)
//   RIGHT: AlignConfig<true,true,false,false>  -> leading gaps free in both sequences
//   LEFT : AlignConfig<false,false,true,true>  -> trailing gaps free in both sequences
int overlap_score(const Seq& ref, const Seq& cand, Direction d, Counters* c) {
  const int MATCH = 4, MISMATCH = -3, GAP = -2;
  const size_t n = ref.size(), m = cand.size();
  if (c) { c->cells_ovl += (uint64_t)n * m; c->calls_ovl++; }
  const bool free_begin = (d == RIGHT);
  std::vector<int> prev(m + 1), cur(m + 1);
  for (size_t j = 0; j <= m; ++j) prev[j] = free_begin ? 0 : (int)j * GAP;
  int best_last_col = prev[m];  // max over last column (j == m), rows 0..n
  for (size_t i = 1; i <= n; ++i) {
    cur[0] = free_begin ? 0 : (int)i * GAP;
    for (size_t j = 1; j <= m; ++j) {
      int v = prev[j - 1] + (ref[i - 1] == cand[j - 1] ? MATCH : MISMATCH);
      v = std::max(v, prev[j] + GAP);
      v = std::max(v, cur[j - 1] + GAP);
      cur[j] = v;
    }
    best_last_col = std::max(best_last_col, cur[m]);
    std::swap(prev, cur);
  }
  if (free_begin) return prev[m];
  int best = best_last_col;
  for (size_t j = 0; j <= m; ++j) best = std::max(best, prev[j]);  // last row
  return best;
}

// SeqAn 2.x seeds_extension.h _extendSeedGappedXDropOneDirection, Seed<Simple>, Score(0,-1,-1)
// (SURVEY B.4).  query_seg = part of V (seq2) to extend into, database_seg = part of H (seq1).
// For extend_left the segments are the prefixes and are consumed from their ends.
void xdrop_extend(const Seq& querySeg, const Seq& databaseSeg, bool extend_left, int scoreDropOff, size_t& ext_rows,
                  size_t& ext_cols, Counters* c) {
  ext_rows = 0;
  ext_cols = 0;
  if (c) c->calls_xdrop++;
  typedef size_t TSize;
  const int scoreMatch = 0, scoreMismatch = -1, scoreGap = -1;
  TSize cols = querySeg.size() + 1;
  TSize rows = databaseSeg.size() + 1;
  if (rows == 1 || cols == 1) return;

  int len = 2 * (int)std::max(cols, rows);
  int minErrScore = INT_MIN / len;
  int gapCost = std::max(scoreGap, minErrScore);
  int undefined = INT_MIN - gapCost;

  std::vector<int> antiDiag1, antiDiag2, antiDiag3;
  TSize minCol = 1, maxCol = 2;
  TSize offset1 = 0, offset2 = 0, offset3 = 0;

  // _initAntiDiags
  antiDiag2.assign(1, 0);
  antiDiag3.assign(2, 0);
  if (-gapCost > scoreDropOff) {
    antiDiag3[0] = undefined;
    antiDiag3[1] = undefined;
  } else {
    antiDiag3[0] = gapCost;
    antiDiag3[1] = gapCost;
  }
  TSize antiDiagNo = 1;
  int best = 0;

  while (minCol < maxCol) {
    ++antiDiagNo;
    // _swapAntiDiags
    {
      std::vector<int> temp;
      temp.swap(antiDiag1);
      antiDiag1.swap(antiDiag2);
      antiDiag2.swap(antiDiag3);
      antiDiag3.swap(temp);
    }
    offset1 = offset2;
    offset2 = offset3;
    offset3 = minCol - 1;
    // _initAntiDiag3
    {
      int minScore = best - scoreDropOff;
      antiDiag3.resize(maxCol + 1 - offset3);
      antiDiag3[0] = undefined;
      antiDiag3[maxCol - offset3] = undefined;
      if ((int)antiDiagNo * gapCost > minScore) {
        if (offset3 == 0) antiDiag3[0] = (int)antiDiagNo * gapCost;
        if (antiDiagNo - maxCol == 0) antiDiag3[maxCol - offset3] = (int)antiDiagNo * gapCost;
      }
    }

    int antiDiagBest = (int)antiDiagNo * gapCost;
    for (TSize col = minCol; col < maxCol; ++col) {
      TSize i3 = col - offset3, i2 = col - offset2, i1 = col - offset1;
      TSize queryPos, dbPos;
      if (!extend_left) {
        queryPos = col - 1;
        dbPos = antiDiagNo - col - 1;
      } else {
        queryPos = cols - 1 - col;
        dbPos = rows - 1 + col - antiDiagNo;
      }
      int tmp = std::max(antiDiag2[i2 - 1], antiDiag2[i2]) + gapCost;
      tmp = std::max(tmp, antiDiag1[i1 - 1] + (querySeg[queryPos] == databaseSeg[dbPos] ? scoreMatch : scoreMismatch));
      if (tmp < best - scoreDropOff) {
        antiDiag3[i3] = undefined;
      } else {
        antiDiag3[i3] = tmp;
        antiDiagBest = std::max(antiDiagBest, tmp);
      }
    }
    if (c) c->cells_xdrop += (maxCol - minCol);
    best = std::max(best, antiDiagBest);

    while (minCol - offset3 < antiDiag3.size() && antiDiag3[minCol - offset3] == undefined &&
           minCol - offset2 - 1 < antiDiag2.size() && antiDiag2[minCol - offset2 - 1] == undefined) {
      ++minCol;
    }
    while (maxCol - offset3 > 0 && (antiDiag3[maxCol - offset3 - 1] == undefined) &&
           (antiDiag2[maxCol - offset2 - 1] == undefined)) {
      --maxCol;
    }
    ++maxCol;

    minCol = (TSize)std::max((int)minCol, (int)antiDiagNo + 2 - (int)rows);  // end of databaseSeg reached?
    maxCol = std::min(maxCol, cols);                                         // end of querySeg reached?
  }

  // positions of the longest extension
  TSize longestExtensionCol = antiDiag3.size() + offset3 - 2;
  TSize longestExtensionRow = antiDiagNo - longestExtensionCol;
  int longestExtensionScore = antiDiag3[longestExtensionCol - offset3];

  if (longestExtensionScore == undefined) {
    if (antiDiag2[antiDiag2.size() - 2] != undefined) {  // reached end of query segment
      longestExtensionCol = antiDiag2.size() + offset2 - 2;
      longestExtensionRow = antiDiagNo - 1 - longestExtensionCol;
      longestExtensionScore = antiDiag2[longestExtensionCol - offset2];
    } else if (antiDiag2.size() > 2 && antiDiag2[antiDiag2.size() - 3] != undefined) {  // end of database segment
      longestExtensionCol = antiDiag2.size() + offset2 - 3;
      longestExtensionRow = antiDiagNo - 1 - longestExtensionCol;
      longestExtensionScore = antiDiag2[longestExtensionCol - offset2];
    }
  }
  if (longestExtensionScore == undefined) {  // general case: first strictly greatest on antiDiag1
    for (TSize i = 0; i < antiDiag1.size(); ++i) {
      if (antiDiag1[i] > longestExtensionScore) {
        longestExtensionScore = antiDiag1[i];
        longestExtensionCol = i + offset1;
        longestExtensionRow = antiDiagNo - 2 - longestExtensionCol;
      }
    }
  }
  if (longestExtensionScore != undefined) {  // _updateExtendedSeed
    ext_rows = longestExtensionRow;
    ext_cols = longestExtensionCol;
  }
}

// Trail.cpp:295-298: Finder<Dna5String> + Pattern<CharString,Horspool>, first match (or -1)
long horspool_first(const Seq& haystack, const Seq& needle, int cycle_mode) {
  const size_t n = haystack.size(), k = needle.size();
  if (n < k || k == 0) return -1;
  if (cycle_mode == 1) {
    size_t p = haystack.find(needle);
    return p == std::string::npos ? -1 : (long)p;
  }
  // cycle_mode 0: bad-character table built from the needle's char ordinals (65..84) but
  // indexed with the haystack's Dna5 ordinals (0..4) -> every shift equals |needle|.
  for (size_t p = 0; p + k <= n; p += k)
    if (haystack.compare(p, k, needle) == 0) return (long)p;
  return -1;
}

// ----------------------------------------------------------------------------
// Statistical model: Explorer.cpp:1185-1217 (double, no FMA; pow(x,2) == x*x)
// ----------------------------------------------------------------------------
bool isExpectedbyMyModel(unsigned int nextc, unsigned int cc, const Params& p, Status classe) {
  const double A = p.ALPHA;
  if ((cc <= 3) & (classe == UNEXPECTED)) return ((double)nextc <= ((double)(cc + 0.5) + A * sqrt((double)(cc + 0.5))));
  else if ((cc <= 3) & (classe == EXPECTED))
    return ((double)nextc >= ((double)(cc - 0.5) + (1 - A) * sqrt((double)(cc - 0.5))));
  else if ((cc > 3) & (classe == UNEXPECTED)) {
    double x = (A / 2 + sqrt((double)(cc + 0.96)));
    return ((double)nextc <= x * x);
  } else {
    double x = (A / 2 - sqrt((double)(cc + 0.02)));
    return ((double)nextc >= x * x);
  }
}

bool isExpectedbyMyLastNode(unsigned int nextc, unsigned int cc, const Params& p) {
  const double A = p.ALPHA;
  bool isExpected = true;
  if (cc <= 3) {
    isExpected &= ((double)nextc <= ((double)(cc + 0.5) + A * sqrt((double)(cc + 0.5))));
    isExpected &= ((double)nextc >= ((double)(cc - 0.5) + (1 - A) * sqrt((double)(cc - 0.5))));
  }
  if (cc > 3) {
    double x = (A / 2 + sqrt((double)(cc + 0.96)));
    isExpected &= ((double)nextc <= x * x);
    double y = (A / 2 - sqrt((double)(cc + 0.02)));
    isExpected &= ((double)nextc >= y * y);
  }
  return isExpected;
}

// Explorer.cpp:1226-1298
void tagNextNodes(std::vector<std::pair<Status, double>>& nodeTags, const std::vector<CCount>& nextCounts,
                  unsigned int count, const Params& p, bool complex_) {
  nodeTags.clear();
  int counter = 0;
  double dist = 0;
  unsigned int nextc = 0;
  unsigned int lambda_noise = 0;
  unsigned int nbExpected = 0, nbBreakpoints = 0, nbUnexpected = 0;

  for (size_t i = 0; i < nextCounts.size(); ++i)
    if ((unsigned int)(int)nextCounts[i].first >= p.MIN_COUNT) counter++;  // :1238
  if (counter > 0) {
    lambda_noise = (unsigned int)(int)((double)count * p.SR_ERROR_RATE);  // :1242
    for (size_t b = 0; b < nextCounts.size(); ++b) {
      nextc = nextCounts[b].first;
      dist = std::fabs((double)count - (double)nextc) / sqrt((double)count);  // :1247 (integer-valued |.|)
      if (nextc >= p.MIN_COUNT) {
        if (isExpectedbyMyModel(nextc, count, p, EXPECTED) || (counter == 1)) {
          nodeTags.push_back(std::make_pair(EXPECTED, dist));
          ++nbExpected;
        } else if (lambda_noise >= p.MIN_COUNT) {
          if (!isExpectedbyMyModel(nextc, lambda_noise, p, UNEXPECTED) || (nextCounts[b].second > 0)) {
            nodeTags.push_back(std::make_pair(BREAKPOINT, dist));
            ++nbBreakpoints;
          } else {
            nodeTags.push_back(std::make_pair(UNEXPECTED, dist));
            ++nbUnexpected;
          }
        } else {
          nodeTags.push_back(std::make_pair(BREAKPOINT, dist));
          ++nbBreakpoints;
        }
      } else
        nodeTags.push_back(std::make_pair(UNEXPECTED, dist));  // :1275 (does not count in nbUnexpected)
    }
  }
  if ((nbExpected == 0) & (nbBreakpoints == 1)) {  // :1278
    for (size_t t = 0; t < nodeTags.size(); ++t)
      if (nodeTags[t].first == BREAKPOINT) nodeTags[t].first = EXPECTED;
  }
  if ((nbExpected == 1) & (nbUnexpected > 0) & !complex_) {  // :1282
    counter = 0;
    unsigned int index = 0;
    for (size_t i = 0; i < nextCounts.size(); ++i) {
      if (nodeTags[i].first == UNEXPECTED) {
        if (counter == 0) index = (unsigned)i;
        counter += nextCounts[i].first;
        if (nextCounts[index].first < nextCounts[i].first) index = (unsigned)i;
      }
    }
    if (!isExpectedbyMyModel((unsigned int)counter, lambda_noise, p, UNEXPECTED)) nodeTags[index].first = BREAKPOINT;
  }
}

// ----------------------------------------------------------------------------
// Trail (Trail.hpp:29-109, Trail.cpp)
// ----------------------------------------------------------------------------
struct Trail {
  Seq seq;                 // m_sequence
  Seq lastKmer;            // m_lastStep.first
  CCount lastCount;        // m_lastStep.second
  double lastScore = 0;    // m_lastScore
  unsigned int failures = 0;  // m_nbFailuresInARow
  int nbBreakpoints = 0;
  double distance = 0;     // m_distance
  int leftAnchor = -1, rightAnchor = -1;

  Trail() {}
  Trail(const Seq& kmer, const CCount& cc) : seq(kmer), lastKmer(kmer), lastCount(cc) {}  // Trail.cpp:57-65
  // Trail.cpp:76-84 (+ addNewBase :313-330, makeTipNode :332-338: colour dropped)
  Trail(const Trail& path, char newBase, Direction d, unsigned int count)
      : seq(d == RIGHT ? path.seq + newBase : std::string(1, newBase) + path.seq),
        lastKmer(formNextKmer(path.lastKmer, newBase, d)),
        lastCount(std::make_pair(count, 0u)),
        lastScore(path.lastScore),
        failures(path.failures),
        nbBreakpoints(path.nbBreakpoints),
        distance(path.distance),
        leftAnchor(path.leftAnchor),
        rightAnchor(path.rightAnchor) {}
  unsigned int length() const { return (unsigned int)seq.size(); }
};

struct Anchor {  // anchorTuple (kmer, position, count)
  Seq kmer;
  unsigned int pos;
  unsigned int count;
};

// Trail.cpp:273-285
static bool checkAims(Trail& t, const std::vector<Anchor>& aims, Direction d) {
  bool isEqual = false;
  unsigned int i = 0;
  while ((!isEqual) & (i < aims.size())) {
    isEqual = (t.lastKmer == aims[i].kmer);
    ++i;
  }
  if (isEqual) {  // setReachedAim :263-267
    if (d == RIGHT) t.rightAnchor = (int)aims[i - 1].pos;
    else t.leftAnchor = (int)aims[i - 1].pos;
  }
  return isEqual;
}

// Trail.cpp:289-302
static bool thinkIveAlreadyGotThere(const Trail& child, const Seq& history, const Params& p) {
  long alreadyOccurred = -1;
  if (history.size() > child.lastKmer.size()) alreadyOccurred = horspool_first(history, child.lastKmer, p.cycle_mode);
  return alreadyOccurred > 0;
}

// Trail.cpp:341-437
std::tuple<Seq, Seq, int, double, bool> getSeedAndExtension(const Seq& reference, const Seq& candidate, int xdrop,
                                                           Direction direction, unsigned int seedSize, Counters* c) {
  Seq refExtension, histExtension;
  int posOnRef = -1;
  bool state = true;
  bool stopThere = false;
  double score = 0;
  const Seq* seq1;
  const Seq* seq2;
  if (reference.size() < candidate.size()) {
    seq1 = &candidate;
    seq2 = &reference;
    state = false;
  } else {
    seq1 = &reference;
    seq2 = &candidate;
  }
  if (direction == RIGHT) {
    // Seed<Simple> seedR(0, 0, seedSize-1, seedSize-1); extendSeed(seedR, seq1 /*H*/, seq2 /*V*/, EXTEND_RIGHT, ..)
    size_t endH = seedSize - 1, endV = seedSize - 1;
    size_t er = 0, ec = 0;
    xdrop_extend(seq2->substr(endV), seq1->substr(endH), false, xdrop, er, ec, c);
    endH += er;
    endV += ec;
    if (state) {
      histExtension = candidate.substr(0, endV);
      refExtension = reference.substr(0, endH);
      posOnRef = (int)endH;
    } else {
      histExtension = candidate.substr(0, endH);
      refExtension = reference.substr(0, endV);
      posOnRef = (int)endV;
    }
  } else {
    // Seed<Simple> seedL(|seq1|-seedSize, |seq2|-seedSize, |seq1|-1, |seq2|-1); EXTEND_LEFT
    size_t beginH = seq1->size() - seedSize, beginV = seq2->size() - seedSize;
    size_t er = 0, ec = 0;
    xdrop_extend(seq2->substr(0, beginV), seq1->substr(0, beginH), true, xdrop, er, ec, c);
    beginH -= er;
    beginV -= ec;
    if (state) {
      histExtension = candidate.substr(beginV);
      refExtension = reference.substr(beginH);
      posOnRef = (int)beginH;
    } else {
      histExtension = candidate.substr(beginH);
      refExtension = reference.substr(beginV);
      posOnRef = (int)beginV;
    }
  }
  if (std::max(refExtension.size(), histExtension.size()) >= seedSize) {  // :408
    score = nw_score(refExtension, histExtension, c);                      // :422 (orientation-free)
  } else {
    if (c) c->ev_bug_alignment++;
    score = (-1) * xdrop;  // :432
    stopThere = true;
  }
  return std::make_tuple(refExtension, histExtension, posOnRef, score, stopThere);
}

// Trail.cpp:193-216
static bool seedAndExtend(Trail& t, const Seq& reference, Direction direction, int xdrop, unsigned int MAX_FAILURES,
                          const Params& p, Counters& c) {
  auto res = getSeedAndExtension(reference, t.seq, xdrop, direction, p.K, &c);
  bool ok = (std::get<1>(res).size() == t.seq.size());
  if (!ok) t.failures++;
  else t.failures = 0;
  t.lastScore = std::get<3>(res);
  if (direction == RIGHT) t.rightAnchor = std::get<2>(res);
  else t.leftAnchor = std::get<2>(res);
  ok = (t.failures <= MAX_FAILURES);
  ok &= !std::get<4>(res);
  return ok;
}

// ----------------------------------------------------------------------------
// Trajectory (Trajectory.hpp:33-104, Trajectory.cpp)
// ----------------------------------------------------------------------------
struct Trajectory {
  Seq seq;
  unsigned int leftAnchor = 0, rightAnchor = 0;
  double score = -200000;
  double idScore = -1;
  int nbBreakpoints = 0;
  double lastScore = 0;
  double meanDistance = 0;
  Trajectory() {}
  explicit Trajectory(const Trail& t)  // Trajectory.cpp:43-48 (Q22)
      : seq(t.seq),
        leftAnchor((unsigned int)t.leftAnchor),
        rightAnchor((unsigned int)t.rightAnchor),
        score(-200000),
        idScore(0),
        nbBreakpoints(t.nbBreakpoints),
        lastScore(t.lastScore),
        meanDistance(t.distance / (t.length() + 0.01)) {}
  unsigned int length() const { return (unsigned int)seq.size(); }
};

static double computeEditDistance(const Seq& reference, const Seq& history, Counters& c) {  // Trajectory.cpp:386-428
  if ((history.size() > 0) & (reference.size() > 0)) return nw_score(reference, history, &c);
  return -100000;
}
static double computeIDScore(const Seq& gap, const Seq& history, Counters& c) {  // Trajectory.cpp:337-384
  if ((gap.size() > 0) & (history.size() > 0)) return lcs_score(gap, history, &c);
  return -1;
}
static double computePercentID(const Seq& s1, const Seq& s2, Counters& c) {  // Trajectory.cpp:505-528
  double len = (s2.size() <= s1.size()) ? (double)s1.size() : (double)s2.size();
  return lcs_score(s1, s2, &c) / len;
}
static void scoreSequence(Trajectory& tr, const Seq& reference, Counters& c) {  // Trajectory.cpp:239-243
  tr.score = computeEditDistance(reference, tr.seq, c);
  tr.idScore = computeIDScore(reference, tr.seq, c) / std::max(reference.size(), tr.seq.size());
}

// Trajectory.cpp:89-112
static void trajTrim(Trajectory& tr, unsigned int minSize, unsigned int intervalLength, unsigned int nbFailuresInARow,
                     Direction d) {
  unsigned int nbBases = nbFailuresInARow * intervalLength;
  if (tr.length() >= nbBases + minSize) {
    if (d == RIGHT) tr.seq = tr.seq.substr(0, tr.length() - nbBases);
    else tr.seq = tr.seq.substr(nbBases);
  }
}

// Trajectory.cpp:482-503
static std::tuple<Seq, Seq, int, double> findStopPosition(const Seq& reference, const Seq& shorterPath, int xdrop,
                                                          Direction d, unsigned int K, Counters& c) {
  int xdrop1 = xdrop;
  bool goFurther = true;
  std::tuple<Seq, Seq, int, double, bool> ext, next;
  next = getSeedAndExtension(reference, shorterPath, xdrop1, d, K, &c);
  do {
    --xdrop1;
    ext = next;
    next = getSeedAndExtension(reference, shorterPath, xdrop1, d, K, &c);
    if (std::get<1>(next).size() < std::get<1>(ext).size()) goFurther = false;
  } while (goFurther & (xdrop1 > 0));
  return std::make_tuple(std::get<0>(ext), std::get<1>(ext), std::get<2>(ext), std::get<3>(ext));
}

// Trajectory.cpp:114-155
static void trajReshape(Trajectory& tr, const Seq& reference, unsigned int K, Direction d, bool shorter, Counters& c) {
  Seq newSeq;
  Seq tmp = tr.seq;
  int xdrop1 = (int)tr.lastScore * (-1);
  std::tuple<Seq, Seq, int, double> er;
  if (!shorter) {
    er = findStopPosition(tmp, reference, xdrop1, d, K, c);
    if (d == LEFT) newSeq = tmp.substr(std::get<2>(er));
    else newSeq = tmp.substr(0, std::get<2>(er));
  } else {
    er = findStopPosition(reference, tmp, xdrop1, d, K, c);
    if (d == LEFT) {
      newSeq = reference.substr(0, std::get<2>(er));
      newSeq += tmp;
    } else {
      newSeq = tr.seq;
      newSeq += reference.substr(std::get<2>(er));
    }
  }
  tr.idScore = computePercentID(std::get<0>(er), std::get<1>(er), c);
  tr.score = std::get<3>(er);
  tr.seq = newSeq;
}

// Trajectory.cpp:157-211
static bool trajCutAnchors(Trajectory& tr, Location location, unsigned int limit, unsigned int K, Counters& c) {
  Seq truncSeq;
  bool isOK = true;
  unsigned int len = tr.length();
  switch (location) {
    case HEAD:
      if (len > K) truncSeq = tr.seq.substr(0, len - K);
      break;
    case TAIL:
      if (len > K) truncSeq = tr.seq.substr(K);
      break;
    case INNER:
      if (len >= 2 * K) {
        truncSeq = tr.seq.substr(K, len - 2 * K);
      } else if ((len < 2 * K) & (len > K)) {
        if (tr.rightAnchor + 2 * K - len <= limit) {
          tr.rightAnchor = tr.rightAnchor + 2 * K - len;
          c.ev_anchor_overlap++;
        } else
          isOK = false;
      } else
        isOK = false;
      break;
    default: break;
  }
  tr.seq = truncSeq;
  return isOK;
}

// Trajectory.cpp:282-303 (Q21)
static unsigned int findBestBridge(const std::vector<Trajectory>& tr) {
  unsigned int index = 0;
  std::vector<unsigned int> exAequo;
  for (unsigned int i = 1; i < tr.size(); i++)
    if (tr[i].score > tr[index].score) index = i;
  for (unsigned int i = index + 1; i < tr.size(); i++)
    if (tr[i].score == tr[index].score) exAequo.push_back(i);
  for (unsigned int i = 0; i < exAequo.size(); i++)
    if (tr[exAequo[i]].meanDistance > tr[index].meanDistance) index = exAequo[i];
  return index;
}
// Trajectory.cpp:306-334
static std::pair<bool, unsigned int> findBestBORDER(const std::vector<Trajectory>& tr) {
  if (tr.empty()) return std::make_pair(false, 0u);
  return std::make_pair(true, findBestBridge(tr));  // identical selection rule
}

// ----------------------------------------------------------------------------
// Explorer (Explorer.hpp:47-165, Explorer.cpp)
// ----------------------------------------------------------------------------
struct Region {
  unsigned int start = 0, end = 0;
  Status status = EXPECTED;
};

static const unsigned int p_MIN_START_ANCHORS = 3;        // Explorer.cpp:85
static const unsigned int p_MAX_START_ANCHORS = 5;        // :86
static const unsigned int p_MAX_IN_COUNT = 100000;        // :88
static const unsigned int p_MAX_NB_OF_BORDER_PATHS = 75;  // :90
static const unsigned int p_MAX_NB_OF_INNER_PATHS = 50;   // :91
static const unsigned int p_CHECK_INTERVAL = 6;           // :93
static const double p_ALLOWED_FAILURE_RATE = 0.3;         // :94
static const int p_MAX_NB_BORDER_FAILURES = 3;            // :97

// Explorer.cpp:402-411
static void sortAnchorsByNearest(double cc, std::vector<Anchor>& anchors, Counters& c) {
  if (anchors.size() > 16) c.ev_sort_gt16++;
  std::sort(anchors.begin(), anchors.end(), [cc](const Anchor& l, const Anchor& r) {
    return std::abs((int)cc - (int)l.count) < std::abs((int)cc - (int)r.count);
  });
}

struct GardenRank {  // tuple<index, score rank, dist rank, rank sum>
  unsigned int idx, r1, r2, sum;
};

// Explorer.cpp:773-865
static bool doABitOfGardening(std::vector<unsigned int>& indexOfKeptPaths, const std::vector<Trail>& paths,
                              const Params& p, Counters& c) {
  indexOfKeptPaths.clear();
  c.ev_gardening++;
  struct IdxSD { unsigned int idx; double score, dist; };
  std::vector<IdxSD> t1, t2;
  std::vector<unsigned int> rankWithTies1, rankWithTies2;
  std::vector<GardenRank> rankings, newrankings;
  bool ties = true;
  bool isComplex = false;
  const unsigned int MAXP = p.MAX_NB_COMPETING_PATHS;
  unsigned int nb = (unsigned int)paths.size();
  nb = std::min(nb, MAXP);
  unsigned int s = 0;
  for (unsigned t = 0; t < paths.size(); t++) {
    rankings.push_back(GardenRank{t, 0, 0, 0});
    rankWithTies1.push_back(t);
    rankWithTies2.push_back(t);
    t1.push_back(IdxSD{t, paths[t].lastScore, paths[t].distance});
    t2.push_back(IdxSD{t, paths[t].lastScore, paths[t].distance});
  }
  if (paths.size() > 16) c.ev_sort_gt16++;
  std::sort(t1.begin(), t1.end(), [](const IdxSD& l, const IdxSD& r) { return l.score > r.score; });  // sortByScore
  rankings[t1[0].idx].r1 = rankWithTies1[0];
  for (unsigned int t = 1; t < rankWithTies1.size(); t++) {
    if (t1[t].score == t1[t - 1].score) rankWithTies1[t] = rankWithTies1[t - 1];
    else rankWithTies1[t] = rankWithTies1[t - 1] + 1;
    rankings[t1[t].idx].r1 = rankWithTies1[t];
  }
  std::sort(t2.begin(), t2.end(), [](const IdxSD& l, const IdxSD& r) { return l.dist < r.dist; });  // sortByLikelihood
  rankings[t2[0].idx].r2 = rankWithTies2[0];
  for (unsigned int t = 1; t < rankWithTies2.size(); t++) {
    if (t2[t].dist == t2[t - 1].dist) rankWithTies2[t] = rankWithTies2[t - 1];
    else rankWithTies2[t] = rankWithTies2[t - 1] + 1;
    rankings[t2[t].idx].r2 = rankWithTies2[t];
  }
  for (unsigned int t = 0; t < rankings.size(); t++) {
    rankings[t].sum = rankings[t].r1 + rankings[t].r2;
    if ((rankings[t].sum == 0) || (paths.size() <= MAXP)) newrankings.push_back(rankings[t]);
  }
  if (newrankings.empty()) {
    c.ev_garden_ties++;
    std::sort(rankings.begin(), rankings.end(),
              [](const GardenRank& l, const GardenRank& r) { return l.r1 < r.r1; });  // sortByMaxScore
    s = 0;
    ties = false;
    do {
      if ((s <= nb) || ties) newrankings.push_back(rankings[s]);
      if (s < rankings.size() - 1) ties = (rankings[s + 1].r1 == rankings[s].r1);
      ++s;
    } while (((s <= nb) || ties) & (s < rankings.size()));

    if (newrankings.size() > MAXP) {
      // NB: newrankings[MAXP] may be read after pops below; pops never release storage.
      GardenRank atMax = newrankings[MAXP];
      if (newrankings[0].r1 != atMax.r1) {
        newrankings.pop_back();
        ties = true;
        while ((newrankings.size() >= MAXP) & ties) {
          ties = (newrankings.back().r1 == newrankings[newrankings.size() - 2].r1);
          ties |= (newrankings.size() >= MAXP);
          if (ties) newrankings.pop_back();
        }
      }
      // :852  (size > MAX) & (r1[0] == r1[MAX]); element [MAX] is the stale-but-intact slot
      if ((newrankings.size() > MAXP) & (newrankings[0].r1 == atMax.r1)) {
        isComplex = true;
        c.ev_garden_q16++;
        std::sort(newrankings.begin(), newrankings.end(),
                  [](const GardenRank& l, const GardenRank& r) { return l.r2 < r.r2; });  // sortByMinDist
        for (unsigned int t = 0; t < MAXP; t++) indexOfKeptPaths.push_back(newrankings[t].idx);
      }
    }
    for (unsigned int t = 0; t < newrankings.size(); t++) indexOfKeptPaths.push_back(newrankings[t].idx);  // :860
  } else
    for (unsigned int t = 0; t < newrankings.size(); t++) indexOfKeptPaths.push_back(newrankings[t].idx);
  return isComplex;
}

class Explorer {
 public:
  Explorer(const Seq& refSequence, const std::vector<CCount>& coverage, double lambda, const Table& t, const Params& p,
           Counters& c, StageDump* dump)
      : m_sequence(refSequence), m_coverage(coverage), m_priorLambda_noise(lambda), T(t), P(p), C(c), D(dump) {}

  void reset() {  // Explorer.cpp:155-174 (m_complexRegion is NOT reset: Q12)
    m_weak = Seq();
    m_weakStatus = EXPECTED;
    m_LEFT = Region();
    m_RIGHT = Region();
    m_LEFT_anchors.clear();
    m_RIGHT_anchors.clear();
    m_location = UNKNOWN;
    m_direction = RIGHT;
    m_fullPaths.clear();
    m_longPaths.clear();
    m_shortPaths.clear();
  }

  void setWeakSequence() {  // :218-226
    const unsigned K = P.K;
    if (m_location == INNER) m_weak = m_sequence.substr(m_LEFT.end + K, m_RIGHT.start - (m_LEFT.end + K));
    else if (m_location == HEAD) m_weak = m_sequence.substr(0, m_RIGHT.start);
    else m_weak = m_sequence.substr(m_LEFT.end + K);
    m_weakStatus = UNCORRECTED;
  }

  void initializeINNER(const Region& l, const Region& r, Direction d) {  // :228-243
    reset();
    m_location = INNER;
    m_direction = d;
    m_LEFT = l;
    m_RIGHT = r;
    setWeakSequence();
    anchorLEFTHandSide();
    anchorRIGHTHandSide();
  }
  void initializeHEAD(const Region& r) {  // :245-257
    reset();
    m_location = HEAD;
    m_direction = LEFT;
    m_RIGHT = r;
    setWeakSequence();
    anchorRIGHTHandSide();
  }
  void initializeTAIL(const Region& l) {  // :259-271
    reset();
    m_location = TAIL;
    m_direction = RIGHT;
    m_LEFT = l;
    setWeakSequence();
    anchorLEFTHandSide();
  }

  Seq kmerAt(unsigned int pos) const { return m_sequence.substr(pos, P.K); }  // utils.cpp:625

  // Explorer.cpp:413-478
  void anchorLEFTHandSide() {
    bool goFurther = true;
    unsigned int nbKmers = m_LEFT.end - m_LEFT.start + 1;
    unsigned int pivot = m_LEFT.end;
    unsigned int limit = m_LEFT.start;
    int degree = 0;
    std::vector<unsigned int> anchorPos;
    double current_count = (double)m_coverage[pivot].first;
    double next_count = 0;
    unsigned int j = pivot;
    anchorPos.push_back(pivot);
    while (goFurther & (j >= limit + 1)) {
      next_count = m_coverage[j - 1].first;
      if ((next_count >= P.MIN_COUNT) & (next_count < p_MAX_IN_COUNT))
        goFurther = isExpectedbyMyLastNode((unsigned int)next_count, (unsigned int)current_count, P);
      else goFurther = false;
      if (!goFurther & (current_count >= P.MIN_COUNT) & (next_count >= P.MIN_COUNT) & (next_count < p_MAX_IN_COUNT)) {
        anchorPos.push_back(j - 1);
        goFurther = true;
        current_count = next_count;
      }
      --j;
    }
    for (int anc = 0; anc < (int)anchorPos.size(); anc++) {
      Seq anchor = kmerAt(anchorPos[anc]);
      degree = getOutDegree(anchor, RIGHT, T, P, C);
      if ((anc == 0) || ((anc != 0) & (degree > 1)))
        m_LEFT_anchors.push_back(Anchor{anchor, anchorPos[anc], m_coverage[anc].first});  // Q6: cov[anc]
    }
    if (m_LEFT_anchors.size() < std::min(p_MIN_START_ANCHORS, nbKmers)) {
      j = pivot;
      goFurther = true;
      while ((j >= limit + 1) & (m_LEFT_anchors.size() < std::min(p_MIN_START_ANCHORS, nbKmers))) {
        for (unsigned int i = 0; i < m_LEFT_anchors.size(); --i) goFurther &= (m_LEFT_anchors[i].pos != (j - 1));  // Q7
        if (goFurther) {
          Seq anchor = kmerAt(j - 1);
          degree = getOutDegree(anchor, RIGHT, T, P, C);
          if (degree > 1) m_LEFT_anchors.push_back(Anchor{anchor, j - 1, m_coverage[j - 1].first});
        }
        --j;
      }
    }
    sortAnchorsByNearest(m_priorLambda_noise / P.SR_ERROR_RATE, m_LEFT_anchors, C);
  }

  // Explorer.cpp:480-543
  void anchorRIGHTHandSide() {
    bool goFurther = true;
    unsigned int nbKmers = m_RIGHT.end - m_RIGHT.start + 1;
    unsigned int pivot = m_RIGHT.start;
    unsigned int limit = m_RIGHT.end;
    std::vector<unsigned int> anchorPos;
    double current_count = (double)m_coverage[pivot].first;
    double next_count = 0;
    unsigned int j = pivot;
    anchorPos.push_back(pivot);
    while (goFurther & ((j + 1) <= limit)) {
      next_count = m_coverage[j + 1].first;
      if ((next_count >= P.MIN_COUNT) & (next_count < p_MAX_IN_COUNT))
        goFurther = isExpectedbyMyLastNode((unsigned int)next_count, (unsigned int)current_count, P);
      else goFurther = false;
      if (!goFurther & (current_count >= P.MIN_COUNT) & (next_count >= P.MIN_COUNT) & (next_count < p_MAX_IN_COUNT)) {
        anchorPos.push_back(j + 1);
        goFurther = true;
        current_count = next_count;
      }
      ++j;
    }
    for (int anc = 0; anc < (int)anchorPos.size(); anc++) {
      Seq anchor = kmerAt(anchorPos[anc]);
      int degree = getOutDegree(anchor, LEFT, T, P, C);
      if ((anc == 0) || ((anc != 0) & (degree > 1)))
        m_RIGHT_anchors.push_back(Anchor{anchor, anchorPos[anc], m_coverage[anc].first});  // Q6
    }
    if (m_RIGHT_anchors.size() < std::min(p_MIN_START_ANCHORS, nbKmers)) {
      j = pivot;
      goFurther = true;
      while (((j + 1) <= limit) & (m_RIGHT_anchors.size() < std::min(p_MIN_START_ANCHORS, nbKmers))) {
        for (unsigned int i = 0; i < m_RIGHT_anchors.size(); --i) goFurther &= (m_RIGHT_anchors[i].pos != (j + 1));  // Q7
        if (goFurther) {
          Seq anchor = kmerAt(j + 1);
          int degree = getOutDegree(anchor, LEFT, T, P, C);
          if (degree > 1) m_RIGHT_anchors.push_back(Anchor{anchor, j + 1, m_coverage[j + 1].first});
        }
        --j;  // Q8 (sic): wraps below zero, loop ends when j+1 > limit
      }
    }
    sortAnchorsByNearest(m_priorLambda_noise / P.SR_ERROR_RATE, m_RIGHT_anchors, C);
  }

  // Explorer.cpp:689-706
  void scoreBridges(std::vector<Trail>& paths, unsigned int stepCounter, const Seq& reference, Direction d) {
    Seq truncatedReference;
    int bound = 0;
    if (d == RIGHT) {
      bound = P.K + stepCounter + P.WINDOW_SIZE;
      if ((size_t)bound >= reference.size()) truncatedReference = reference;
      else truncatedReference = reference.substr(0, bound);
    } else {
      bound = (int)reference.size() - (int)P.K - (int)stepCounter - (int)P.WINDOW_SIZE;
      if (bound < 0) truncatedReference = reference;
      else truncatedReference = reference.substr(bound);
    }
    if (!P.q11_zero_init) return;  // Q11: UB counter; modelled as j = 0
    for (unsigned int j = 0; j < paths.size(); j++)
      paths[j].lastScore = overlap_score(truncatedReference, paths[j].seq, d, &C);  // Trail.cpp:145-174
  }

  // Explorer.cpp:1103-1118
  void recordEdge(const Trail& trail, const Seq& reference) {
    C.ev_edge++;
    Trajectory myTip(trail);
    trajTrim(myTip, P.K, p_CHECK_INTERVAL, trail.failures, m_direction);
    bool shorter = (myTip.length() <= reference.size());
    trajReshape(myTip, reference, P.K, m_direction, shorter, C);
    if (trajCutAnchors(myTip, m_location, 0, P.K, C)) {
      if (shorter) m_shortPaths.push_back(myTip);
      else m_longPaths.push_back(myTip);
    }
  }

  // Explorer.cpp:546-612
  void oneMoreStep(const Seq& reference, std::vector<Trail>& competingPaths, unsigned int& stepCounter) {
    std::vector<Trail> newCompetingPaths;
    std::vector<CCount> nextCounts;
    std::vector<std::pair<Status, double>> nodeTags;
    std::vector<unsigned int> indexOfKeptPaths;
    bool cycle = false, aimReached = false, complex_ = false;
    C.steps_inner++;
    C.frontier_sum += competingPaths.size();
    C.frontier_max = std::max<uint64_t>(C.frontier_max, competingPaths.size());
    for (unsigned int t = 0; t < competingPaths.size(); t++) {
      complex_ = (competingPaths.size() > P.MAX_NB_COMPETING_PATHS);
      nextCounts = getNextCounts(competingPaths[t].lastKmer, m_direction, T);  // Trail::whatsNext
      C.lookups_walk += 4;
      tagNextNodes(nodeTags, nextCounts, competingPaths[t].lastCount.first, P, complex_);
      for (unsigned int i = 0; i < nodeTags.size(); i++) {
        if (nodeTags[i].first != UNEXPECTED) {
          newCompetingPaths.push_back(Trail(competingPaths[t], kDict[i], m_direction, nextCounts[i].first));
          Trail& nb = newCompetingPaths.back();
          if (nodeTags[i].first == BREAKPOINT) nb.nbBreakpoints++;
          nb.distance += nodeTags[i].second;
          if (m_direction == RIGHT) aimReached = checkAims(nb, m_RIGHT_anchors, m_direction);
          else aimReached = checkAims(nb, m_LEFT_anchors, m_direction);
          if (aimReached) {
            C.ev_bridge++;
            m_fullPaths.push_back(Trajectory(nb));  // recordBridge :1097
            if (nb.length() > reference.size()) newCompetingPaths.pop_back();  // Q10
          } else {
            cycle = thinkIveAlreadyGotThere(nb, competingPaths[t].seq, P);
            if (cycle) {
              C.ev_cycle++;
              newCompetingPaths.pop_back();
            }
          }
        }
      }
    }
    complex_ = (newCompetingPaths.size() > P.MAX_NB_COMPETING_PATHS);
    ++stepCounter;
    if (complex_ & (stepCounter % p_CHECK_INTERVAL == 0)) {
      scoreBridges(newCompetingPaths, stepCounter, reference, m_direction);
      m_complexRegion |= doABitOfGardening(indexOfKeptPaths, newCompetingPaths, P, C);
      competingPaths.clear();
      for (unsigned int t = 0; t < indexOfKeptPaths.size(); t++)
        competingPaths.push_back(newCompetingPaths[indexOfKeptPaths[t]]);
    } else
      competingPaths = newCompetingPaths;
  }

  // Explorer.cpp:709-740
  void scoreEdges(int& xdrop, std::vector<Trail>& newCompetingPaths, unsigned int stepCounter, const Seq& reference,
                  Direction d) {
    (void)stepCounter;
    std::vector<Trail> newSelectedPaths, trashPaths;
    int new_xdrop = 0, current_xdrop = 0;
    if (!newCompetingPaths.empty()) {
      xdrop += 2;
      for (unsigned int t = 0; t < newCompetingPaths.size(); t++) {
        bool ok = seedAndExtend(newCompetingPaths[t], reference, d, xdrop, p_MAX_NB_BORDER_FAILURES, P, C);
        if (!ok) trashPaths.push_back(newCompetingPaths[t]);
        else {
          newSelectedPaths.push_back(newCompetingPaths[t]);
          current_xdrop = (int)(newCompetingPaths[t].lastScore * (-1));
          if ((new_xdrop > current_xdrop) || (new_xdrop == 0)) new_xdrop = current_xdrop;  // Q14
        }
      }
      xdrop = new_xdrop;
      newCompetingPaths = newSelectedPaths;
      if (newCompetingPaths.empty())
        for (unsigned int t = 0; t < trashPaths.size(); t++) recordEdge(trashPaths[t], reference);
    }
  }

  // Explorer.cpp:615-687
  void oneMoreStepInTheDark(int& xdrop, const Seq& reference, std::vector<Trail>& competingPaths,
                            unsigned int& stepCounter, unsigned int PATH_MAXLENGTH) {
    std::vector<Trail> newCompetingPaths;
    std::vector<CCount> nextCounts;
    std::vector<std::pair<Status, double>> nodeTags;
    std::vector<unsigned int> indexOfKeptPaths;
    bool complex_ = false, cycle = false;
    unsigned int counter = 0;
    C.steps_border++;
    C.frontier_sum += competingPaths.size();
    C.frontier_max = std::max<uint64_t>(C.frontier_max, competingPaths.size());
    for (unsigned int t = 0; t < competingPaths.size(); t++) {
      counter = 0;
      complex_ = (competingPaths.size() > 7);
      nextCounts = getNextCounts(competingPaths[t].lastKmer, m_direction, T);
      C.lookups_walk += 4;
      tagNextNodes(nodeTags, nextCounts, competingPaths[t].lastCount.first, P, complex_);
      for (unsigned int i = 0; i < nodeTags.size(); i++) {
        if (nodeTags[i].first != UNEXPECTED) {
          ++counter;
          newCompetingPaths.push_back(Trail(competingPaths[t], kDict[i], m_direction, nextCounts[i].first));
          Trail& nb = newCompetingPaths.back();
          if (nodeTags[i].first == BREAKPOINT) nb.nbBreakpoints++;
          nb.distance += nodeTags[i].second;
          cycle = thinkIveAlreadyGotThere(nb, competingPaths[t].seq, P);
          if (cycle || (stepCounter + 1 > PATH_MAXLENGTH)) {
            if (cycle) C.ev_cycle++;
            seedAndExtend(nb, reference, m_direction, xdrop, p_MAX_NB_BORDER_FAILURES, P, C);
            recordEdge(nb, reference);
            newCompetingPaths.pop_back();
          }
        }
      }
      if (counter == 0) {  // dead end
        seedAndExtend(competingPaths[t], reference, m_direction, xdrop, p_MAX_NB_BORDER_FAILURES, P, C);
        recordEdge(competingPaths[t], reference);
      }
    }
    ++stepCounter;
    if ((stepCounter % p_CHECK_INTERVAL == 0) || (newCompetingPaths.size() >= p_MAX_NB_OF_BORDER_PATHS)) {
      scoreEdges(xdrop, newCompetingPaths, stepCounter, reference, m_direction);
      if (newCompetingPaths.size() > 5) {
        m_complexRegion |= doABitOfGardening(indexOfKeptPaths, newCompetingPaths, P, C);
        competingPaths.clear();
        for (unsigned int t = 0; t < indexOfKeptPaths.size(); t++)
          competingPaths.push_back(newCompetingPaths[indexOfKeptPaths[t]]);
      } else
        competingPaths = newCompetingPaths;
    } else
      competingPaths = newCompetingPaths;
  }

  // Explorer.cpp:868-989
  bool searchBridge() {
    const unsigned K = P.K;
    bool pathHasBeenFound = false;
    unsigned int PATH_MAXLENGTH = 0;
    unsigned int stepCounter = 0;
    std::vector<Anchor> anchors = (m_direction == LEFT) ? m_RIGHT_anchors : m_LEFT_anchors;
    unsigned int limit = (unsigned int)anchors.size();
    limit = std::min(limit, p_MAX_START_ANCHORS);
    std::vector<Trail> competingPaths;
    for (int s = 0; s < (int)limit; s++) {
      if (pathHasBeenFound) continue;
      C.gap_attempts++;
      competingPaths.clear();
      m_fullPaths.clear();
      stepCounter = 0;
      unsigned int whichStart = anchors[s].pos;
      Seq currentTarget = (m_direction == RIGHT) ? m_sequence.substr(m_RIGHT.start, m_RIGHT.end + K - m_RIGHT.start)
                                                 : m_sequence.substr(m_LEFT.start, m_LEFT.end + K - m_LEFT.start);
      Seq currentGap;
      if ((m_direction == RIGHT) & (whichStart + K < m_RIGHT.start))
        currentGap = m_sequence.substr(whichStart + K, m_RIGHT.start - (whichStart + K));
      else if ((m_direction == LEFT) & (m_LEFT.end + K < whichStart))
        currentGap = m_sequence.substr(m_LEFT.end + K, whichStart - (m_LEFT.end + K));
      PATH_MAXLENGTH = (unsigned int)(int)(1.2 * currentGap.size() + 3 * K);
      const Seq& currentAnchor = anchors[s].kmer;
      CCount currentCount = m_coverage[whichStart];
      competingPaths.push_back(Trail(currentAnchor, currentCount));
      Seq currentRefSeq;
      if (m_direction == RIGHT) {
        currentRefSeq = currentAnchor + currentGap + currentTarget;
        competingPaths[0].leftAnchor = (int)whichStart;
      } else {
        currentRefSeq = currentTarget + currentGap + currentAnchor;
        competingPaths[0].rightAnchor = (int)whichStart;
      }
      while ((!competingPaths.empty()) & (competingPaths.size() <= p_MAX_NB_OF_INNER_PATHS) &
             (stepCounter < PATH_MAXLENGTH)) {
        oneMoreStep(currentRefSeq, competingPaths, stepCounter);
      }
      if (competingPaths.size() > p_MAX_NB_OF_INNER_PATHS) C.ev_frontier_over50++;
      else if (!competingPaths.empty()) C.ev_maxlength++;

      if (!m_fullPaths.empty() & !pathHasBeenFound) {
        std::vector<unsigned int> index;
        unsigned int lim = m_RIGHT.end;
        for (unsigned int t = 0; t < m_fullPaths.size(); t++) {
          scoreSequence(m_fullPaths[t], currentRefSeq, C);
          if (trajCutAnchors(m_fullPaths[t], m_location, lim, K, C)) index.push_back(t);
        }
        if (m_fullPaths.size() != index.size()) {  // Q9
          C.ev_q9++;
          m_shortPaths = m_fullPaths;
          m_fullPaths.clear();
          for (unsigned int t = 0; t < index.size(); t++) m_fullPaths.push_back(m_shortPaths[t]);
        }
        if (!m_fullPaths.empty()) {
          unsigned int bestOne = findBestBridge(m_fullPaths);
          const Seq& bestPath = m_fullPaths[bestOne].seq;
          unsigned int weakLen = (unsigned int)m_weak.size();
          double diff = (double)weakLen - (double)bestPath.size();
          if ((diff < weakLen * 0.05 || ((weakLen < 6) & (bestPath.size() < 6))) &
              (m_fullPaths[bestOne].idScore >= P.MIN_INNER_SCORE)) {  // Q13
            m_LEFT.end = m_fullPaths[bestOne].leftAnchor;
            m_RIGHT.start = m_fullPaths[bestOne].rightAnchor;
            m_weak = m_fullPaths[bestOne].seq;
            m_weakStatus = CORRECTED;
            pathHasBeenFound = true;
          }
        }
      }
    }
    return pathHasBeenFound;
  }

  // Explorer.cpp:310-329
  bool sortOutBestBorder(Trajectory& out) {
    auto r = findBestBORDER(m_longPaths);
    if (r.first) { out = m_longPaths[r.second]; return true; }
    r = findBestBORDER(m_shortPaths);
    if (r.first) { out = m_shortPaths[r.second]; return true; }
    out = Trajectory();
    return false;
  }

  // Explorer.cpp:992-1081
  bool searchEdge() {
    const unsigned K = P.K;
    bool pathHasBeenFound = false;
    unsigned int PATH_MAXLENGTH = 0;
    unsigned int stepCounter = 0;
    std::vector<Anchor> anchors = (m_direction == LEFT) ? m_RIGHT_anchors : m_LEFT_anchors;
    unsigned int limit = (unsigned int)anchors.size();
    limit = std::min(limit, p_MAX_START_ANCHORS);
    std::vector<Trail> competingPaths;
    int xdrop;
    for (int s = 0; s < (int)limit; s++) {
      stepCounter = 0;
      competingPaths.clear();
      xdrop = (int)((int)p_CHECK_INTERVAL * p_ALLOWED_FAILURE_RATE + 1);  // Q15: 6*0.3+1 = 2.8 -> 2
      unsigned int whichStart = anchors[s].pos;
      Seq currentGap = (m_location == HEAD) ? m_sequence.substr(0, whichStart) : m_sequence.substr(whichStart + K);
      PATH_MAXLENGTH = (unsigned int)(int)(1.2 * currentGap.size() + 2 * K);
      const Seq& currentAnchor = anchors[s].kmer;
      CCount currentCount = m_coverage[whichStart];
      competingPaths.push_back(Trail(currentAnchor, currentCount));
      Seq currentRefSeq;
      if (m_direction == RIGHT) {
        currentRefSeq = currentAnchor + currentGap;
        competingPaths[0].leftAnchor = (int)whichStart;
      } else {
        currentRefSeq = currentGap + currentAnchor;
        competingPaths[0].rightAnchor = (int)whichStart;
      }
      while ((!competingPaths.empty()) & (competingPaths.size() <= p_MAX_NB_OF_INNER_PATHS) &
             (stepCounter < PATH_MAXLENGTH)) {
        oneMoreStepInTheDark(xdrop, currentRefSeq, competingPaths, stepCounter, PATH_MAXLENGTH);
      }
      if (competingPaths.size() > p_MAX_NB_OF_INNER_PATHS) C.ev_frontier_over50++;
      else if (!competingPaths.empty()) C.ev_maxlength++;
    }
    if (!m_shortPaths.empty() || !m_longPaths.empty()) {
      Trajectory winner;
      sortOutBestBorder(winner);
      unsigned int weakLen = (unsigned int)m_weak.size();
      double diff = (double)weakLen - (double)winner.length();
      double minScore;
      if ((weakLen >= 300) || m_complexRegion) minScore = std::max(0.75, P.MIN_BORDER_SCORE);
      else minScore = P.MIN_BORDER_SCORE;
      if ((diff < weakLen * 0.05 || ((weakLen < 6) & (winner.length() < 6))) & (winner.idScore >= minScore)) {
        pathHasBeenFound = true;
        m_weak = winner.seq;
        m_weakStatus = CORRECTED;
        if (m_location == TAIL) m_LEFT.end = winner.leftAnchor;
        else m_RIGHT.start = winner.rightAnchor;
      }
    }
    return pathHasBeenFound;
  }

  Seq m_sequence;
  Seq m_weak;
  Status m_weakStatus = EXPECTED;
  std::vector<CCount> m_coverage;
  double m_priorLambda_noise;
  Region m_LEFT, m_RIGHT;
  std::vector<Anchor> m_LEFT_anchors, m_RIGHT_anchors;
  Location m_location = UNKNOWN;
  Direction m_direction = RIGHT;
  bool m_complexRegion = false;
  std::vector<Trajectory> m_fullPaths, m_longPaths, m_shortPaths;
  const Table& T;
  const Params& P;
  Counters& C;
  StageDump* D;
};

// ----------------------------------------------------------------------------
// Read (Read.cpp)
// ----------------------------------------------------------------------------
// Read.cpp:440-489
static bool findINRegions(std::vector<Region>& out, const std::vector<CCount>& counts, const Params& p) {
  unsigned int pos = 0, current_start = 0, current_count = 0;
  bool state = false;
  if (counts.size() > 1) {
    while (pos < counts.size()) {
      current_count = counts[pos].first;
      if ((current_count >= p.MIN_COUNT) & (state == false)) {
        current_start = pos;
        state = true;
      } else if ((current_count < p.MIN_COUNT) & (state == true)) {
        Region r; r.start = current_start; r.end = pos - 1; r.status = UNEXPECTED;
        out.push_back(r);
        state = false;
      }
      ++pos;
    }
    if (state == true) {
      Region r; r.start = current_start; r.end = pos - 1; r.status = UNEXPECTED;
      out.push_back(r);
    }
  }
  return !out.empty();
}

// Read.cpp:493-518
static double computeSeqErrorThreshold(const std::vector<CCount>& counts, const Params& p) {
  std::vector<unsigned int> INcounts;
  double robMean = p.MIN_COUNT;  // Q2
  unsigned int first, last;
  for (unsigned int pos = 0; pos < counts.size(); pos++)
    if (counts[pos].first >= p.MIN_COUNT) INcounts.push_back(counts[pos].first);
  std::sort(INcounts.begin(), INcounts.end());
  INcounts.size() > 10 ? first = (unsigned int)(0.15 * (double)INcounts.size()) : first = 0;
  INcounts.size() > 10 ? last = (unsigned int)(0.90 * (double)INcounts.size()) : last = (unsigned int)INcounts.size();
  for (unsigned int i = first; i < last; i++) robMean += INcounts[i];
  robMean /= last - first;
  return robMean * p.SR_ERROR_RATE;
}

// Read.cpp:524-600
static void analyzeINRegions(std::vector<Region>& regs, const Seq& ref, const std::vector<CCount>& counts,
                             double solidityThr, const Table& T, const Params& p, Counters& C) {
  std::vector<Region> kept;
  const unsigned K = p.K;
  bool OK = true;
  unsigned int new_start_pos = 0, new_end_pos = 0, c = 0;
  int span = 0;
  for (unsigned int reg = 0; reg < regs.size(); reg++) {
    span = 0;
    OK = true;
    new_start_pos = regs[reg].start;
    Seq kmer = ref.substr(new_start_pos, K);
    const bool lastAndNone = ((regs.size() == reg + 1) & kept.empty());
    // non-short-circuit '&': getOutDegree is always evaluated (:548)
    int degL = getOutDegree(kmer, LEFT, T, p, C);
    if (!lastAndNone & (degL == 0) & (new_start_pos != 0)) {
      OK = false;
      while ((new_start_pos < regs[reg].end) & !OK) {
        ++new_start_pos;
        kmer = ref.substr(new_start_pos, K);
        if (getOutDegree(kmer, LEFT, T, p, C) > 1) OK = true;  // Q3
      }
    }
    new_end_pos = regs[reg].end;
    if (OK & !lastAndNone) {
      kmer = ref.substr(new_end_pos, K);
      int degR = getOutDegree(kmer, RIGHT, T, p, C);
      if ((degR == 0) & (new_end_pos != counts.size() - 1)) {
        OK = false;
        while ((new_end_pos > regs[reg].start) & !OK) {  // Q4: original start
          --new_end_pos;
          kmer = ref.substr(new_end_pos, K);
          if (getOutDegree(kmer, RIGHT, T, p, C) > 1) OK = true;
        }
      }
    }
    if (OK) {
      if (reg + 1 < regs.size()) span = ((int)regs[reg + 1].start - (int)(new_end_pos + K));
      if (span < 0) {
        if ((int)regs[reg + 1].end + span >= (int)regs[reg + 1].start) regs[reg + 1].start -= span;  // Q5
        else {
          regs[reg + 1].start = new_start_pos;
          OK = false;
        }
      }
      if (OK) {
        c = 0;
        for (unsigned int i = new_start_pos; i <= new_end_pos; i++) c < counts[i].first ? c = counts[i].first : c += 0;
        if (!isExpectedbyMyModel(c, (unsigned int)solidityThr, p, UNEXPECTED)) {
          Region r; r.start = new_start_pos; r.end = new_end_pos; r.status = EXPECTED;
          kept.push_back(r);
        }
      }
    }
  }
  if (!kept.empty()) regs = kept;
}

ReadResult correct_read(const Seq& raw, const Table& T, const Params& p, Counters& C, StageDump* dump) {
  ReadResult res;
  res.corrected = raw;
  const unsigned K = p.K;
  C.reads++;
  C.bases_in += raw.size();
  if (!((int)raw.size() > (int)K)) {  // main.cpp:262
    res.status = READ_SHORT;
    C.reads_short++;
    C.bases_out += raw.size();
    return res;
  }
  // Read::reCoverage (Read.cpp:174-195)
  std::vector<CCount> coverage = getLRCountsInSR(raw, K, T, C);
  int nbIn = 0;
  for (size_t i = 0; i < coverage.size(); ++i)
    if (coverage[i].first > p.MIN_COUNT) nbIn++;  // Q1: strict
  if (dump) dump->coverage = coverage;
  if (!(nbIn > 0)) {
    res.status = READ_NO_SOLID;
    C.reads_nosolid++;
    C.bases_out += raw.size();
    return res;
  }
  // Read::defineStructure2 (Read.cpp:260-276)
  std::vector<Region> regs;
  bool checok = findINRegions(regs, coverage, p);
  if (dump) for (auto& r : regs) dump->regions_found.push_back(std::make_tuple(r.start, r.end, (int)r.status));
  double thr = computeSeqErrorThreshold(coverage, p);
  if (dump) dump->threshold = thr;
  analyzeINRegions(regs, raw, coverage, thr, T, p, C);
  if (dump) for (auto& r : regs) dump->regions_final.push_back(std::make_tuple(r.start, r.end, (int)r.status));

  // Read::setInitialStructure (Read.cpp:214-258)
  Seq head, tail;
  bool headPresent = false, tailPresent = false;
  std::vector<Seq> inner;  // m_newInnerStructure (sequence part)
  if (!regs.empty()) {
    unsigned int len = 0;
    if (regs[0].start > 0) {
      head = raw.substr(0, regs[0].start);
      headPresent = true;
      len += (unsigned)head.size();
    }
    if (regs.back().end + 1 < coverage.size()) {
      tail = raw.substr(regs.back().end + K);
      tailPresent = true;
      len += (unsigned)tail.size();
    }
    for (unsigned int i = 0; i + 1 < regs.size(); i++) {
      // extractSolidSequence: infix(start, end+K).  A start beyond end+K (possible after Q4/Q5
      // mutations) is undefined in the reference; the structure check below fails then.
      Seq solid = (regs[i].end + K >= regs[i].start) ? raw.substr(regs[i].start, regs[i].end + K - regs[i].start) : Seq();
      inner.push_back(solid);
      len += (unsigned)solid.size();
      Seq weak;
      if (regs[i + 1].start > regs[i].end + K) weak = raw.substr(regs[i].end + K, regs[i + 1].start - (regs[i].end + K));
      inner.push_back(weak);
      len += (unsigned)weak.size();
    }
    Seq lastSolid = (regs.back().end + K >= regs.back().start)
                        ? raw.substr(regs.back().start, regs.back().end + K - regs.back().start)
                        : Seq();
    inner.push_back(lastSolid);
    len += (unsigned)lastSolid.size();
    checok &= (len == raw.size());
  }
  auto fill_stats = [&]() {  // Read.cpp:421-424
    res.stat_regions = (unsigned)regs.size();
    res.stat_span = 0;
    for (auto& r : regs) res.stat_span += r.end - r.start + 1;
  };
  if (!checok) {
    fill_stats();
    res.status = READ_NO_STRUCTURE;
    C.reads_nostruct++;
    C.bases_out += raw.size();
    return res;
  }

  // Read::correct2 (Read.cpp:336-386)
  Explorer ex(raw, coverage, thr, T, p, C, dump);
  for (int reg = 0; reg < (int)regs.size() - 1; reg++) {
    C.gaps++;
    ex.initializeINNER(regs[reg], regs[reg + 1], RIGHT);
    bool success = ex.searchBridge();
    if (!success) {
      ex.initializeINNER(regs[reg], regs[reg + 1], LEFT);
      success = ex.searchBridge();
    }
    if (success) C.gaps_bridged++;
    // updateINNER (Read.cpp:294-303)
    regs[reg] = ex.m_LEFT;
    regs[reg + 1] = ex.m_RIGHT;
    inner[2 * reg + 1] = ex.m_weak;
    inner[2 * reg] = raw.substr(ex.m_LEFT.start, ex.m_LEFT.end + K - ex.m_LEFT.start);
    inner[2 * (reg + 1)] = raw.substr(ex.m_RIGHT.start, ex.m_RIGHT.end + K - ex.m_RIGHT.start);
    if (dump) {
      std::ostringstream os;
      os << "gap " << reg << " ok=" << success << " L=" << ex.m_LEFT.start << "-" << ex.m_LEFT.end << " R=" << ex.m_RIGHT.start
         << "-" << ex.m_RIGHT.end << " weak=" << ex.m_weak << "\n";
      dump->trace += os.str();
    }
  }
  if (headPresent) {
    if (head.size() <= 500) {
      C.borders++;
      ex.initializeHEAD(regs[0]);
      if (ex.searchEdge()) {  // updateHEAD (Read.cpp:305-311)
        C.borders_corrected++;
        regs[0] = ex.m_RIGHT;
        inner[0] = raw.substr(ex.m_RIGHT.start, ex.m_RIGHT.end + K - ex.m_RIGHT.start);
        head = ex.m_weak;
      }
      if (dump) dump->trace += "head -> " + head + "\n";
    } else
      C.border_cutoff_500++;
  }
  if (tailPresent) {
    if (tail.size() <= 500) {
      C.borders++;
      ex.initializeTAIL(regs.back());
      if (ex.searchEdge()) {  // updateTAIL (Read.cpp:313-318)
        C.borders_corrected++;
        regs.back() = ex.m_LEFT;
        inner.back() = raw.substr(ex.m_LEFT.start, ex.m_LEFT.end + K - ex.m_LEFT.start);
        tail = ex.m_weak;
      }
      if (dump) dump->trace += "tail -> " + tail + "\n";
    } else
      C.border_cutoff_500++;
  }
  // updateCorrSeq (Read.cpp:320-326)
  Seq corr = head;
  for (size_t i = 0; i < inner.size(); ++i) corr += inner[i];
  corr += tail;
  res.corrected = corr;
  res.status = READ_OK;
  fill_stats();
  C.reads_corrected++;
  C.bases_out += corr.size();
  return res;
}

std::vector<unsigned> std_sort_permutation(const std::vector<long long>& keys) {
  std::vector<unsigned> idx(keys.size());
  for (unsigned i = 0; i < idx.size(); ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&keys](unsigned a, unsigned b) { return keys[a] < keys[b]; });
  return idx;
}

}  // namespace talc_oracle
