// =============================================================================
// talc_oracle_main.cpp -- `talc`-compatible CLI around the CPU restatement.
// TEST INFRASTRUCTURE ONLY (checker + reported CPU baseline).  PARITY UNPINNED.
//
// Follows main.cpp:83-325 (driver), Settings.cpp:74-185 (.config.txt),
// io.cpp:26-75,105-111 (FASTA/FASTQ in, 70-column FASTA out, .log append),
// Read.cpp:394-415 (.stats_basics.txt header).  The stdout debug flood of the
// reference is intentionally not reproduced: stdout is not graded.
//
// Extra, oracle-only options (not part of the reference CLI):
//   --oracle-table map|hash   ordered std::map (reference cost, default) or hash
//   --cycle-mode 0|1          see talc_oracle.hpp Params::cycle_mode
//   --stats-json FILE         write counters + wall-clock of the correction loop
//   --max-reads N             correct only the first N reads (bounded CPU sample)
// =============================================================================
#include <omp.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "talc_oracle.hpp"

using namespace talc_oracle;

struct Cli {
  std::string reads, dump, junctions, out = "out", queryMode = "memory", jf2;
  bool useJunctions = false, reverse = false;
  int threads = 1;
  Params p;
  bool have_k = false, have_sr = false;
  // oracle-only
  bool ordered = true;
  std::string statsJson;
  long maxReads = -1;
};

static bool parse_cli(int argc, const char** argv, Cli& c, bool& help) {
  help = false;
  std::vector<std::string> pos;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto need = [&](std::string& dst) {
      if (i + 1 >= argc) return false;
      dst = argv[++i];
      return true;
    };
    std::string v;
    if (a == "-h" || a == "--help" || a == "--version") { help = true; return false; }
    else if (a == "-o" || a == "--output") { if (!need(c.out)) return false; }
    else if (a == "-k" || a == "--kmerSize") {
      if (!need(v)) return false;
      int k = atoi(v.c_str());
      if (k < 18 || k > 30) return false;  // main.cpp:115-116
      c.p.K = k; c.have_k = true;
    }
    else if (a == "-qm" || a == "--query-mode") {
      if (!need(c.queryMode)) return false;
      if (c.queryMode != "memory" && c.queryMode != "jellyfish2") return false;
    }
    else if (a == "-SR" || a == "--SRCounts") { if (!need(c.dump)) return false; c.have_sr = true; }
    else if (a == "-j" || a == "--junctions") { if (!need(c.junctions)) return false; c.useJunctions = true; }
    else if (a == "-jf2" || a == "--pathToJF2") { if (!need(c.jf2)) return false; }
    else if (a == "-MIN_INNER_SCORE" || a == "--MIN_INNER_SCORE") {
      if (!need(v)) return false; c.p.MIN_INNER_SCORE = atof(v.c_str());
      if (c.p.MIN_INNER_SCORE < 0.3 || c.p.MIN_INNER_SCORE > 0.9) return false;
    }
    else if (a == "-MIN_BORDER_SCORE" || a == "--MIN_BORDER_SCORE") {
      if (!need(v)) return false; c.p.MIN_BORDER_SCORE = atof(v.c_str());
      if (c.p.MIN_BORDER_SCORE < 0.5 || c.p.MIN_BORDER_SCORE > 0.9) return false;
    }
    else if (a == "-MIN_COUNT" || a == "--MIN_COUNT") {
      if (!need(v)) return false; int x = atoi(v.c_str()); if (x < 2) return false; c.p.MIN_COUNT = x;
    }
    else if (a == "-SR_ERROR_RATE" || a == "--SR_ERROR_RATE") {
      if (!need(v)) return false; c.p.SR_ERROR_RATE = atof(v.c_str());
      if (c.p.SR_ERROR_RATE < 0.01 || c.p.SR_ERROR_RATE > 0.1) return false;
    }
    else if (a == "-WINDOW_SIZE" || a == "--WINDOW_SIZE") {
      if (!need(v)) return false; int x = atoi(v.c_str()); if (x < 6) return false; c.p.WINDOW_SIZE = x;
    }
    else if (a == "-MAX_NB_BRANCHES" || a == "--MAX_NB_BRANCHES") {
      if (!need(v)) return false; int x = atoi(v.c_str()); if (x < 5) return false; c.p.MAX_NB_COMPETING_PATHS = x;
    }
    else if (a == "-ALPHA_FOR_PRED" || a == "--ALPHA_FOR_PRED") {
      if (!need(v)) return false; c.p.ALPHA = atof(v.c_str()); if (c.p.ALPHA < 0.67) return false;
    }
    else if (a == "-t" || a == "--num_threads") {
      if (!need(v)) return false; c.threads = atoi(v.c_str()); if (c.threads < 1) return false;
    }
    else if (a == "-DEBUG_MODE" || a == "--DEBUG_MODE") { if (!need(v)) return false; }
    else if (a == "-rev" || a == "--reverse") { c.reverse = true; }
    else if (a == "--oracle-table") { if (!need(v)) return false; c.ordered = (v != "hash"); }
    else if (a == "--cycle-mode") { if (!need(v)) return false; c.p.cycle_mode = atoi(v.c_str()); }
    else if (a == "--stats-json") { if (!need(c.statsJson)) return false; }
    else if (a == "--max-reads") { if (!need(v)) return false; c.maxReads = atol(v.c_str()); }
    else if (!a.empty() && a[0] == '-' && a.size() > 1) return false;  // unknown option -> PARSE_ERROR
    else pos.push_back(a);
  }
  if (pos.size() != 1 || !c.have_k || !c.have_sr) return false;
  c.reads = pos[0];
  return true;
}

// Settings.cpp:160-185
static void outputConfig(const Cli& c) {
  std::ofstream f(c.out + ".config.txt", std::ios_base::trunc);
  f << "TALC: Parameters used for sample: " << c.out << "\n"
    << "****************************" << "\n"
    << "INPUT=" << c.reads << "\n"
    << "OUTPUT=" << c.out << "\n"
    << "STATS=" << c.out + ".stats_basics.txt" << "\n"
    << "****************************" << "\n"
    << "KmerSize=" << c.p.K << "\n"
    << "Junction mode activated? " << c.useJunctions << "\n"
    << "queryMode=" << c.queryMode << "\n"
    << "****************************" << "\n"
    << "MIN_INNER_SCORE=" << c.p.MIN_INNER_SCORE << "\n"
    << "MIN_BORDER_SCORE=" << c.p.MIN_BORDER_SCORE << "\n"
    << "MAX_NB_BRANCHES=" << c.p.MAX_NB_COMPETING_PATHS << "\n"
    << "ALPHA=" << c.p.ALPHA << "\n"
    << "MIN_SR_COUNT=" << c.p.MIN_COUNT << "\n"
    << "WINDOW_SIZE=" << c.p.WINDOW_SIZE << "\n"
    << "****************************" << std::endl;
}

// Read.cpp:394-415
static void statsHeader(const std::string& path) {
  std::ofstream f(path, std::ios_base::trunc);
  f << "read_name\traw_length\twhead_length\twtail_length\tnbInKmers\tnbSolidKmers\tnbSolidReg\tnbInWeakReg\t"
       "nbInCorrReg\tCorrHead?\tCorrHeadLen\tCorrTail?\tCorrTailLen\tCorrlength\tnbInKmers2\n";
}

// io.cpp:26-48 via SeqAn readRecords: FASTA or FASTQ by first byte.
static int loadSeqData(std::vector<std::string>& ids, std::vector<Seq>& seqs, const std::string& path) {
  std::ifstream in(path);
  if (!in) { std::cerr << "ERROR: Could not open file " << path << "\n"; return 1; }
  std::string line;
  auto chomp = [](std::string& s) { while (!s.empty() && (s.back() == '\r' || s.back() == '\n')) s.pop_back(); };
  bool have = false;
  bool fastq = false;
  std::string seq;
  auto valid = [](const std::string& s) {
    for (char ch : s) {
      switch (ch) {
        case 'A': case 'C': case 'G': case 'T': case 'N': case 'a': case 'c': case 'g': case 't': case 'n': break;
        default: return false;
      }
    }
    return true;
  };
  while (std::getline(in, line)) {
    chomp(line);
    if (!have) {
      if (line.empty()) continue;
      if (line[0] == '>') fastq = false;
      else if (line[0] == '@') fastq = true;
      else return 1;
      have = true;
    }
    if (!fastq) {
      if (!line.empty() && line[0] == '>') {
        if (ids.size() > seqs.size()) seqs.push_back(to_dna5(seq));
        ids.push_back(line.substr(1));
        seq.clear();
      } else {
        std::string s;
        for (char ch : line) if (ch != ' ' && ch != '\t') s += ch;
        if (!valid(s)) { std::cout << "ERROR: Unexpected character found" << std::endl; return 1; }
        seq += s;
      }
    } else {
      if (line.empty()) continue;
      if (line[0] != '@') return 1;
      ids.push_back(line.substr(1));
      std::string s, plus, qual;
      if (!std::getline(in, s)) return 1;
      chomp(s);
      if (!valid(s)) { std::cout << "ERROR: Unexpected character found" << std::endl; return 1; }
      if (!std::getline(in, plus)) return 1;
      if (!std::getline(in, qual)) return 1;
      seqs.push_back(to_dna5(s));
    }
  }
  if (!fastq && ids.size() > seqs.size()) seqs.push_back(to_dna5(seq));
  return 0;
}

// io.cpp:50-75 via SeqAn writeRecords: 70 columns
static int outputSeqData(const std::vector<std::string>& ids, const std::vector<Seq>& seqs, const std::string& path) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { std::cerr << "ERROR: Could not open the file " << path << "\n"; return 1; }
  std::string buf;
  for (size_t r = 0; r < ids.size(); ++r) {
    buf.clear();
    buf += '>'; buf += ids[r]; buf += '\n';
    const Seq& s = seqs[r];
    if (s.empty()) buf += '\n';
    for (size_t i = 0; i < s.size(); i += 70) { buf.append(s, i, 70); buf += '\n'; }
    fwrite(buf.data(), 1, buf.size(), f);
  }
  fclose(f);
  return 0;
}

// io.cpp:105-111
static void throwToLog(const std::string& name, const std::string& msg, const std::string& path) {
  std::ofstream f(path, std::ios_base::app);
  f << "[Read: " << name << " ]: " << msg << std::endl;
}

int main(int argc, const char** argv) {
  Cli cli;
  bool help = false;
  if (!parse_cli(argc, argv, cli, help)) {
    if (help) { std::cout << "talc (oracle restatement) <reads> --SRCounts F [--junctions F] -k K [-o P] [-t N]\n"; return 0; }
    std::cerr << "talc: PARSE_ERROR\n";
    return 1;  // main.cpp:199
  }
  outputConfig(cli);                           // Settings.cpp:122
  statsHeader(cli.out + ".stats_basics.txt");  // main.cpp:204
  const std::string logFile = cli.out + ".log";

  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::string> ids;
  std::vector<Seq> seqs;
  if (loadSeqData(ids, seqs, cli.reads) != 0) {  // main.cpp:219,323
    std::cout << "[TALC]: ISSUE WITH INPUT FILES" << std::endl;
    return 0;
  }
  auto t1 = std::chrono::steady_clock::now();
  Table table(cli.ordered);
  BuildStats bs = build_cdbg(table, cli.p, cli.dump, cli.junctions, cli.useJunctions);  // main.cpp:231-232
  auto t2 = std::chrono::steady_clock::now();
  if (table.size() == 0) {
    std::cout << "[TALC]: The de Bruijn Graph is empty...Correction aborted." << std::endl;
    return 1;  // main.cpp:320
  }
  omp_set_num_threads(cli.threads);  // main.cpp:242
  size_t nreads = ids.size();
  if (cli.maxReads >= 0 && (size_t)cli.maxReads < nreads) nreads = (size_t)cli.maxReads;
  std::vector<Counters> perThread(cli.threads);
  std::vector<unsigned char> status(ids.size(), 255);
  uint64_t basesIn = 0;
#pragma omp parallel for schedule(dynamic)  // main.cpp:247
  for (size_t r = 0; r < nreads; r++) {
    Counters& C = perThread[omp_get_thread_num()];
    Seq s = seqs[r];
    if (cli.reverse) s = reverse_complement(s);  // main.cpp:253
    ReadResult res = correct_read(s, table, cli.p, C);
    status[r] = (unsigned char)res.status;
    if (res.status == READ_OK) {
      seqs[r] = cli.reverse ? reverse_complement(res.corrected) : res.corrected;  // main.cpp:285-286
    } else {
      seqs[r] = s;  // Q24: failed reads stay reverse-complemented under --reverse
      if (cli.threads > 1) {  // log order is completion order (F11)
        if (res.status == READ_NO_STRUCTURE) {
#pragma omp critical
          throwToLog(ids[r], "Unable to define convenient structure.", logFile);
        } else if (res.status == READ_NO_SOLID) {
#pragma omp critical
          throwToLog(ids[r], "No solid kmer could be found.", logFile);
        }
      }
    }
  }
  if (cli.threads == 1) {
    for (size_t r = 0; r < nreads; r++) {
      if (status[r] == READ_NO_STRUCTURE) throwToLog(ids[r], "Unable to define convenient structure.", logFile);
      else if (status[r] == READ_NO_SOLID) throwToLog(ids[r], "No solid kmer could be found.", logFile);
    }
  }
  auto t3 = std::chrono::steady_clock::now();
  outputSeqData(ids, seqs, cli.out + ".fa");  // main.cpp:310
  auto t4 = std::chrono::steady_clock::now();

  Counters total;
  for (auto& c : perThread) total.add(c);
  basesIn = total.bases_in;
  auto secs = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
  if (!cli.statsJson.empty()) {
    std::ofstream f(cli.statsJson);
    f << "{\"threads\": " << cli.threads << ", \"table_ordered\": " << (cli.ordered ? 1 : 0)
      << ", \"table_entries\": " << table.size() << ", \"dump_lines\": " << bs.lines << ", \"reads\": " << nreads
      << ", \"bases_in\": " << basesIn << ", \"t_load_reads_s\": " << secs(t0, t1) << ", \"t_build_table_s\": " << secs(t1, t2)
      << ", \"t_correct_s\": " << secs(t2, t3) << ", \"t_write_s\": " << secs(t3, t4)
      << ", \"mbp_per_s_correct\": " << (basesIn / 1e6) / std::max(1e-9, secs(t2, t3)) << ", \"counters\": " << total.json()
      << "}\n";
  }
  std::cout << "[TALC]: Looks like we are done now." << std::endl;
  return 0;
}
