// dump_parse.hpp -- host-side reader of Jellyfish `dump -c` text (KMER<ws>COUNT per line), the input of
// buildCDBG (Jellyfish.cpp:251-269 for the counts, :278-289 for the junction list).
//
// Line semantics follow `iss >> kmer0 >> count` + std::stoi: two whitespace-separated tokens, anything
// after them is ignored, lines with fewer tokens are skipped (the reference prints "Error when building
// dBG..." and continues).  A count token without leading digits, or beyond int range, makes the
// reference's std::stoi throw and the program die; here such lines are counted as bad and skipped.
// k-mers that are not exactly K letters of ACGT (either case) can never be matched by a K-mer of a read
// in the 2-bit table, and Jellyfish never emits them; they are dropped.
//
// The file is mapped and split into per-thread chunks at line boundaries; the entries come back in file
// order, which is all "first line wins" needs.
#pragma once
#include <fcntl.h>
#include <string.h>
#include <stdint.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <thread>
#include <vector>

struct DumpEntries {
  std::vector<uint64_t> keys;
  std::vector<int64_t> counts;  // std::stoi value
  uint64_t lines = 0, bad_lines = 0, dropped_kmers = 0;
};

namespace dump_detail {

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

struct Chunk {
  std::vector<uint64_t> keys;
  std::vector<int64_t> counts;
  uint64_t lines = 0, bad = 0, dropped = 0;
};

inline void parse_range(const char* p, const char* end, uint32_t K, uint32_t min_count, bool filter, Chunk& out) {
  while (p < end) {
    const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
    if (!eol) eol = end;
    const char* q = p;
    while (q < eol && is_ws(*q)) ++q;
    const char* k0 = q;
    while (q < eol && !is_ws(*q)) ++q;
    const char* k1 = q;
    while (q < eol && is_ws(*q)) ++q;
    const char* c0 = q;
    while (q < eol && !is_ws(*q)) ++q;
    const char* c1 = q;
    if (k1 > k0 && c1 > c0) {
      // std::stoi
      const char* d = c0;
      bool neg = false;
      if (*d == '+' || *d == '-') { neg = (*d == '-'); ++d; }
      if (d < c1 && *d >= '0' && *d <= '9') {
        int64_t v = 0;
        bool range_ok = true;
        while (d < c1 && *d >= '0' && *d <= '9') {
          v = v * 10 + (*d - '0');
          if (v > 2147483648LL) { range_ok = false; break; }
          ++d;
        }
        if (neg) v = -v;
        if (range_ok && v <= 2147483647LL) {
          out.lines++;
          const bool keep = !filter || ((uint32_t)(int32_t)v >= min_count);
          if (keep) {
            uint64_t key = 0;
            bool ok = ((uint32_t)(k1 - k0) == K);
            for (const char* s = k0; ok && s < k1; ++s) {
              uint64_t c;
              switch (*s) {
                case 'A': case 'a': c = 0; break;
                case 'C': case 'c': c = 1; break;
                case 'G': case 'g': c = 2; break;
                case 'T': case 't': case 'U': case 'u': c = 3; break;
                default: ok = false; c = 0; break;
              }
              key = (key << 2) | c;
            }
            if (ok) {
              out.keys.push_back(key);
              out.counts.push_back(v);
            } else {
              out.dropped++;
            }
          }
        } else {
          out.bad++;
        }
      } else {
        out.bad++;
      }
    } else if (eol > p || eol < end) {
      out.bad++;  // fewer than two tokens (an empty line included)
    }
    p = eol + 1;
  }
}

}  // namespace dump_detail

inline bool parse_dump_file(const char* path, uint32_t K, uint32_t min_count, bool filter, DumpEntries& out, std::string& err) {
  using namespace dump_detail;
  int fd = open(path, O_RDONLY);
  if (fd < 0) {
    err = std::string("cannot open ") + path;
    return false;
  }
  struct stat st;
  if (fstat(fd, &st) != 0) {
    close(fd);
    err = std::string("cannot stat ") + path;
    return false;
  }
  const size_t size = (size_t)st.st_size;
  if (size == 0) {
    close(fd);
    return true;
  }
  const char* data = (const char*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (data == MAP_FAILED) {
    err = std::string("cannot map ") + path;
    return false;
  }
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 4;
  if (nt > 64) nt = 64;
  if (size < (1u << 20)) nt = 1;
  std::vector<size_t> cut(nt + 1, size);
  cut[0] = 0;
  for (unsigned t = 1; t < nt; ++t) {
    size_t pos = size / nt * t;
    const char* nl = (const char*)memchr(data + pos, '\n', size - pos);
    cut[t] = nl ? (size_t)(nl - data) + 1 : size;
    if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
  }
  std::vector<Chunk> chunks(nt);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t)
    th.emplace_back([&, t]() { parse_range(data + cut[t], data + cut[t + 1], K, min_count, filter, chunks[t]); });
  for (auto& x : th) x.join();
  size_t total = 0;
  for (auto& c : chunks) total += c.keys.size();
  out.keys.reserve(total);
  out.counts.reserve(total);
  for (auto& c : chunks) {
    out.keys.insert(out.keys.end(), c.keys.begin(), c.keys.end());
    out.counts.insert(out.counts.end(), c.counts.begin(), c.counts.end());
    out.lines += c.lines;
    out.bad_lines += c.bad;
    out.dropped_kmers += c.dropped;
  }
  munmap((void*)data, size);
  return true;
}
