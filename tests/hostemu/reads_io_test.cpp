// reads_io_test.cpp -- C entry points around talc_b200/csrc/host/reads_io.hpp for tests/test_reads_io.py (no GPU):
// parse a FASTA / FASTQ file in batches exactly as talc_main.cpp does, and format a batch as 70-column FASTA.
#include "../../talc_b200/csrc/host/reads_io.hpp"

extern "C" {

// Parses `path` with the given batch limits.  Returns the number of reads (or -1 if the file cannot be opened, -2 on
// a parse error) and fills: ids_out (all ids joined with '\n'), bases_out (all bases), offs_out (n+1 offsets into
// bases_out), batch_sizes_out (reads per batch, up to max_batches), *n_batches.  Capacities are the caller's promise.
long rio_parse(const char* path, long batch_reads, long batch_bases, char* ids_out, long ids_cap, unsigned char* bases_out,
               long bases_cap, unsigned long long* offs_out, long offs_cap, long* batch_sizes_out, long max_batches, long* n_batches) {
  ReadParser parser(path);
  if (!parser.is_open()) return -1;
  long n = 0, nb = 0, idpos = 0;
  unsigned long long base = 0;
  offs_out[0] = 0;
  for (;;) {
    Batch b;
    if (!parser.next_batch(b, (size_t)batch_reads, (size_t)batch_bases)) return -2;
    if (b.ids.empty()) break;
    if (b.offs.size() != b.ids.size() + 1) return -3;
    if (nb < max_batches) batch_sizes_out[nb] = (long)b.ids.size();
    ++nb;
    for (size_t r = 0; r < b.ids.size(); ++r) {
      const std::string& id = b.ids[r];
      if (idpos + (long)id.size() + 1 > ids_cap || n + 2 > offs_cap) return -4;
      memcpy(ids_out + idpos, id.data(), id.size());
      idpos += (long)id.size();
      ids_out[idpos++] = '\n';
      const unsigned long long len = b.offs[r + 1] - b.offs[r];
      if ((long)(base + len) > bases_cap) return -4;
      memcpy(bases_out + base, b.bases.data() + b.offs[r], len);
      base += len;
      offs_out[++n] = base;
    }
  }
  if (idpos < ids_cap) ids_out[idpos] = 0;
  *n_batches = nb;
  return n;
}

// Formats n records (ids joined with '\n', sequences out[ooffs[r] .. ooffs[r+1])) with `threads` workers; returns the
// number of bytes written to dst (or -1 if dst_cap is too small).
long rio_format(const char* ids, long n, const unsigned char* out, const unsigned long long* ooffs, int threads, char* dst, long dst_cap) {
  Batch b;
  const char* p = ids;
  for (long r = 0; r < n; ++r) {
    const char* e = strchr(p, '\n');
    b.ids.emplace_back(p, e ? (size_t)(e - p) : strlen(p));
    p = e ? e + 1 : p + strlen(p);
  }
  std::vector<std::string> parts;
  std::vector<uint64_t> offs(ooffs, ooffs + n + 1);
  format_fasta(b, out, offs.data(), threads, parts);
  long pos = 0;
  for (const auto& s : parts) {
    if (pos + (long)s.size() > dst_cap) return -1;
    memcpy(dst + pos, s.data(), s.size());
    pos += (long)s.size();
  }
  return pos;
}

}  // extern "C"
