// talc_main.cpp -- the drop-in `talc` command line on top of libtalc_b200.so.
//
// Same interface and the same output files as the reference program (main.cpp:83-325):
//   talc <reads.fa|fq> --SRCounts <dump> [--junctions <dump>] -k <K> [-o <prefix>] [-t <N>] [tunables]
//   <prefix>.config.txt (Settings.cpp:160-185)   <prefix>.stats_basics.txt header (Read.cpp:394-415)
//   <prefix>.log, appended, one line per failed read (io.cpp:105-111; messages main.cpp:290,294)
//   <prefix>.fa, every read in input order, 70 columns (main.cpp:310, io.cpp:50-75)
// Exit codes follow main.cpp: 1 on a parse error (:199) or an empty table (:320), 0 otherwise -- including
// unreadable input (:323).  The work is done on the GPU through the C ABI; there is no CPU path.
// -t is accepted for compatibility and used for host-side formatting only.  Extension: --gpus N shards
// the reads over N devices of the box (table replicated device to device).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "talc_b200.h"

struct Cli {
  std::string reads, dump, junctions, out = "out", queryMode = "memory", tableCache;
  bool useJunctions = false, reverse = false;
  int threads = 1, gpus = 1;
  talc_params p;
  bool haveK = false, haveSR = false;
};

static bool to_int(const std::string& s, long& v) {
  char* e = nullptr;
  v = strtol(s.c_str(), &e, 10);
  return e && *e == 0 && !s.empty();
}
static bool to_dbl(const std::string& s, double& v) {
  char* e = nullptr;
  v = strtod(s.c_str(), &e);
  return e && *e == 0 && !s.empty();
}

// mirrors the option table of main.cpp:101-195 (SeqAn ArgumentParser: "-x" short and "--long" names)
static int parse(int argc, const char** argv, Cli& c) {
  talc_params_default(&c.p, 21);
  std::vector<std::string> pos;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto val = [&](std::string& dst) { if (i + 1 >= argc) return false; dst = argv[++i]; return true; };
    std::string v;
    long n;
    double d;
    if (a == "-h" || a == "--help" || a == "--version") return 2;
    else if (a == "-o" || a == "--output") { if (!val(c.out)) return 1; }
    else if (a == "-k" || a == "--kmerSize") { if (!val(v) || !to_int(v, n) || n < 18 || n > 30) return 1; c.p.K = (uint32_t)n; c.haveK = true; }
    else if (a == "-qm" || a == "--query-mode") { if (!val(c.queryMode) || (c.queryMode != "memory" && c.queryMode != "jellyfish2")) return 1; }
    else if (a == "-SR" || a == "--SRCounts") { if (!val(c.dump)) return 1; c.haveSR = true; }
    else if (a == "-j" || a == "--junctions") { if (!val(c.junctions)) return 1; c.useJunctions = true; }
    else if (a == "--tableCache") { if (!val(c.tableCache)) return 1; }  // extension: binary cache of the built table
    else if (a == "-jf2" || a == "--pathToJF2") { if (!val(v)) return 1; }
    else if (a == "-MIN_INNER_SCORE" || a == "--MIN_INNER_SCORE") { if (!val(v) || !to_dbl(v, d) || d < 0.3 || d > 0.9) return 1; c.p.min_inner_score = d; }
    else if (a == "-MIN_BORDER_SCORE" || a == "--MIN_BORDER_SCORE") { if (!val(v) || !to_dbl(v, d) || d < 0.5 || d > 0.9) return 1; c.p.min_border_score = d; }
    else if (a == "-MIN_COUNT" || a == "--MIN_COUNT") { if (!val(v) || !to_int(v, n) || n < 2) return 1; c.p.min_count = (uint32_t)n; }
    else if (a == "-SR_ERROR_RATE" || a == "--SR_ERROR_RATE") { if (!val(v) || !to_dbl(v, d) || d < 0.01 || d > 0.1) return 1; c.p.sr_error_rate = d; }
    else if (a == "-WINDOW_SIZE" || a == "--WINDOW_SIZE") { if (!val(v) || !to_int(v, n) || n < 6) return 1; c.p.window_size = (uint32_t)n; }
    else if (a == "-MAX_NB_BRANCHES" || a == "--MAX_NB_BRANCHES") { if (!val(v) || !to_int(v, n) || n < 5) return 1; c.p.max_nb_branches = (uint32_t)n; }
    else if (a == "-ALPHA_FOR_PRED" || a == "--ALPHA_FOR_PRED") { if (!val(v) || !to_dbl(v, d) || d < 0.67) return 1; c.p.alpha = d; }
    else if (a == "-t" || a == "--num_threads") { if (!val(v) || !to_int(v, n) || n < 1) return 1; c.threads = (int)n; }
    else if (a == "-DEBUG_MODE" || a == "--DEBUG_MODE") { if (!val(v)) return 1; }
    else if (a == "-rev" || a == "--reverse") { c.reverse = true; }
    else if (a == "--gpus") { if (!val(v) || !to_int(v, n) || n < 1) return 1; c.gpus = (int)n; }
    else if (a == "--cycle-mode") { if (!val(v) || !to_int(v, n)) return 1; c.p.cycle_mode = (int32_t)n; }
    else if (a.size() > 1 && a[0] == '-') return 1;
    else pos.push_back(a);
  }
  if (pos.size() != 1 || !c.haveK || !c.haveSR) return 1;
  if (c.reverse) { std::cerr << "talc: --reverse is not supported by the GPU build\n"; return 1; }
  if (c.queryMode != "memory") { std::cerr << "talc: only query mode 'memory' is supported\n"; return 1; }
  c.reads = pos[0];
  return 0;
}

static void write_config(const Cli& c) {  // Settings.cpp:160-185
  std::ofstream f(c.out + ".config.txt", std::ios_base::trunc);
  f << "TALC: Parameters used for sample: " << c.out << "\n"
    << "****************************\n"
    << "INPUT=" << c.reads << "\n"
    << "OUTPUT=" << c.out << "\n"
    << "STATS=" << c.out << ".stats_basics.txt" << "\n"
    << "****************************\n"
    << "KmerSize=" << c.p.K << "\n"
    << "Junction mode activated? " << c.useJunctions << "\n"
    << "queryMode=" << c.queryMode << "\n"
    << "****************************\n"
    << "MIN_INNER_SCORE=" << c.p.min_inner_score << "\n"
    << "MIN_BORDER_SCORE=" << c.p.min_border_score << "\n"
    << "MAX_NB_BRANCHES=" << c.p.max_nb_branches << "\n"
    << "ALPHA=" << c.p.alpha << "\n"
    << "MIN_SR_COUNT=" << c.p.min_count << "\n"
    << "WINDOW_SIZE=" << c.p.window_size << "\n"
    << "****************************" << std::endl;
}
static void write_stats_header(const std::string& path) {  // Read.cpp:394-415
  std::ofstream f(path, std::ios_base::trunc);
  f << "read_name\traw_length\twhead_length\twtail_length\tnbInKmers\tnbSolidKmers\tnbSolidReg\tnbInWeakReg\t"
       "nbInCorrReg\tCorrHead?\tCorrHeadLen\tCorrTail?\tCorrTailLen\tCorrlength\tnbInKmers2\n";
}

// io.cpp:26-48 (SeqAn readRecords): FASTA or FASTQ by the first byte; ids = header without marker;
// sequence letters must be ACGTN (either case), anything else is a parse error.
static bool load_reads(const std::string& path, std::vector<std::string>& ids, std::vector<uint8_t>& bases,
                       std::vector<uint64_t>& offs) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { std::cerr << "ERROR: Could not open file " << path << "\n"; return false; }
  std::string data;
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, n);
  fclose(f);
  offs.assign(1, 0);
  size_t p = 0;
  const size_t N = data.size();
  auto line = [&](size_t& b, size_t& e) {  // [b,e) without the terminator; false at end of file
    if (p >= N) return false;
    b = p;
    const void* nl = memchr(data.data() + p, '\n', N - p);
    e = nl ? (size_t)((const char*)nl - data.data()) : N;
    p = e + 1;
    while (e > b && data[e - 1] == '\r') --e;
    return true;
  };
  auto push_seq = [&](size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) {
      const char ch = data[i];
      switch (ch) {
        case 'A': case 'C': case 'G': case 'T': case 'N': case 'a': case 'c': case 'g': case 't': case 'n':
          bases.push_back((uint8_t)ch);
          break;
        case ' ': case '\t': break;
        default: return false;
      }
    }
    return true;
  };
  size_t b, e;
  bool first = true, fastq = false, open = false;
  while (line(b, e)) {
    if (first) {
      if (e == b) continue;
      if (data[b] == '>') fastq = false;
      else if (data[b] == '@') fastq = true;
      else return false;
      first = false;
    }
    if (!fastq) {
      if (e > b && data[b] == '>') {
        if (open) offs.push_back(bases.size());
        ids.emplace_back(data, b + 1, e - b - 1);
        open = true;
      } else if (!push_seq(b, e)) {
        std::cout << "ERROR: Unexpected character found" << std::endl;
        return false;
      }
    } else {
      if (e == b) continue;
      if (data[b] != '@') return false;
      ids.emplace_back(data, b + 1, e - b - 1);
      size_t sb, se, xb, xe;
      if (!line(sb, se) || !push_seq(sb, se)) { std::cout << "ERROR: Unexpected character found" << std::endl; return false; }
      if (!line(xb, xe) || !line(xb, xe)) return false;
      offs.push_back(bases.size());
    }
  }
  if (!fastq && open) offs.push_back(bases.size());
  return true;
}

int main(int argc, const char** argv) {
  std::cout << "******************************************************\n"
            << "* TALC : Transcriptome-Aware Long Read Correction    *\n"
            << "*        B200 build (libtalc_b200)                   *\n"
            << "******************************************************" << std::endl;
  Cli cli;
  const int pr = parse(argc, argv, cli);
  if (pr == 2) {
    std::cout << "talc <reads> --SRCounts <dump> [--junctions <dump>] -k <K> [-o <prefix>] [-t <N>] [--gpus <N>] [--tableCache <file>]\n";
    return 0;
  }
  if (pr != 0) { std::cerr << "talc: PARSE_ERROR\n"; return 1; }
  write_config(cli);
  write_stats_header(cli.out + ".stats_basics.txt");

  std::vector<std::string> ids;
  std::vector<uint8_t> bases;
  std::vector<uint64_t> offs;
  std::cout << "[TALC]: Attempting to load sequences." << std::endl;
  if (!load_reads(cli.reads, ids, bases, offs) || offs.size() != ids.size() + 1) {
    std::cout << "[TALC]: ISSUE WITH INPUT FILES" << std::endl;
    return 0;  // main.cpp:323 falls off main
  }
  std::cout << "[TALC]: " << ids.size() << " long read(s) loaded" << std::endl;

  std::vector<talc_ctx*> ctx(cli.gpus, nullptr);
  for (int g = 0; g < cli.gpus; ++g) {
    if (talc_ctx_create(&cli.p, g, &ctx[g]) != 0) {
      std::cerr << "talc: cannot create a GPU context on device " << g << ": " << talc_last_error(nullptr) << "\n";
      return 2;
    }
  }
  uint64_t nLines = 0, nKept = 0;
  int rc = -1;
  // --tableCache <file> (extension): reuse the built table if the file exists, else build it from the
  // dump and write the file; a cache made from other --SRCounts / --junctions files (size, mtime), another -k or
  // MIN_COUNT is refused by the library and rebuilt
  if (!cli.tableCache.empty())
    rc = talc_table_load_cache_for(ctx[0], cli.tableCache.c_str(), cli.dump.c_str(),
                                   cli.useJunctions ? cli.junctions.c_str() : nullptr, &nKept);
  if (rc == TALC_ERR_STALE) std::cout << "[TALC]: " << talc_last_error(ctx[0]) << " -> rebuilding it." << std::endl;
  if (rc != 0) {
    rc = talc_table_load_dump(ctx[0], cli.dump.c_str(), cli.useJunctions ? cli.junctions.c_str() : nullptr, &nLines, &nKept);
    if (rc == 0 && nKept > 0 && !cli.tableCache.empty() && talc_table_save(ctx[0], cli.tableCache.c_str()) != 0)
      std::cerr << "talc: " << talc_last_error(ctx[0]) << "\n";
  }
  if (rc != 0) { std::cerr << "talc: " << talc_last_error(ctx[0]) << "\n"; nKept = 0; }
  std::cout << "[TALC]: SR-dBG contains " << nKept << " nodes." << std::endl;
  if (nKept == 0) {
    std::cout << "[TALC]: The de Bruijn Graph is empty...Correction aborted." << std::endl;
    return 1;  // main.cpp:320
  }
  for (int g = 1; g < cli.gpus; ++g) {
    if (talc_table_copy(ctx[g], ctx[0]) != 0) { std::cerr << "talc: " << talc_last_error(ctx[g]) << "\n"; return 2; }
  }

  // shard contiguous blocks of reads, balanced by bases, over the devices
  const size_t R = ids.size();
  std::vector<size_t> cut(cli.gpus + 1, R);
  cut[0] = 0;
  {
    const uint64_t total = offs[R];
    size_t r = 0;
    for (int g = 1; g < cli.gpus; ++g) {
      const uint64_t want = total / cli.gpus * g;
      while (r < R && offs[r] < want) ++r;
      cut[g] = r;
    }
  }
  std::vector<std::vector<uint8_t>> out(cli.gpus), status(cli.gpus);
  std::vector<std::vector<uint64_t>> ooffs(cli.gpus);
  std::vector<int> rcs(cli.gpus, 0);
  std::vector<std::thread> th;
  for (int g = 0; g < cli.gpus; ++g) {
    th.emplace_back([&, g]() {
      const size_t r0 = cut[g], r1 = cut[g + 1];
      const uint32_t n = (uint32_t)(r1 - r0);
      std::vector<uint64_t> lo(n + 1);
      for (uint32_t i = 0; i <= n; ++i) lo[i] = offs[r0 + i] - offs[r0];
      out[g].resize(2 * lo[n] + 64ull * n + 4096);
      ooffs[g].assign(n + 1, 0);
      status[g].assign(n + 1, 0);
      rcs[g] = talc_correct_batch(ctx[g], bases.data() + offs[r0], lo.data(), n, out[g].data(), out[g].size(), ooffs[g].data(),
                                  status[g].data(), nullptr);
    });
  }
  for (auto& t : th) t.join();
  for (int g = 0; g < cli.gpus; ++g) {
    if (rcs[g] != 0) { std::cerr << "talc: correction failed on device " << g << ": " << talc_last_error(ctx[g]) << "\n"; return 2; }
  }

  // failed-read log in input order (the reference's order under -t 1), then the FASTA
  {
    std::ofstream lg;
    bool opened = false;
    for (int g = 0; g < cli.gpus; ++g) {
      for (size_t r = cut[g]; r < cut[g + 1]; ++r) {
        const uint8_t st = status[g][r - cut[g]];
        const char* msg = st == TALC_READ_NO_STRUCTURE ? "Unable to define convenient structure."
                          : st == TALC_READ_NO_SOLID   ? "No solid kmer could be found." : nullptr;
        if (msg) {
          if (!opened) { lg.open(cli.out + ".log", std::ios_base::app); opened = true; }
          lg << "[Read: " << ids[r] << " ]: " << msg << std::endl;
        }
      }
    }
  }
  FILE* fo = fopen((cli.out + ".fa").c_str(), "wb");
  if (!fo) { std::cerr << "ERROR: Could not open the file " << cli.out << ".fa\n"; return 0; }
  std::string buf;
  for (int g = 0; g < cli.gpus; ++g) {
    for (size_t r = cut[g]; r < cut[g + 1]; ++r) {
      const size_t i = r - cut[g];
      const uint8_t* s = out[g].data() + ooffs[g][i];
      const size_t len = ooffs[g][i + 1] - ooffs[g][i];
      buf.clear();
      buf += '>';
      buf += ids[r];
      buf += '\n';
      if (len == 0) buf += '\n';
      for (size_t j = 0; j < len; j += 70) {
        buf.append((const char*)s + j, std::min<size_t>(70, len - j));
        buf += '\n';
      }
      fwrite(buf.data(), 1, buf.size(), fo);
    }
  }
  fclose(fo);
  for (auto* c : ctx) talc_ctx_destroy(c);
  std::cout << "[TALC]: Looks like we are done now." << std::endl;
  return 0;
}
