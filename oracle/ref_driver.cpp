// oracle/_ref/dicey_ref -- TEST INFRASTRUCTURE, not product code.
//
// The reference's own arithmetic, compiled from the sources where they lie under
// /root/reference (nothing is copied): SDSL csa_wt<> construct/count/locate/extract,
// src/neighbors.h, src/needle.h + src/align.h, src/version.h and nlohmann json are
// #included verbatim.  Only the driver loops are restated here, because hunter.h /
// silica.h / index.h cannot be compiled without Boost + htslib (neither is installed):
//
//   cmd_index   restates  src/index.h:96-123     (dump text -> construct -> store_to_checked_file)
//   run_hunt    restates  src/hunter.h:289-444   (per-query loop) + :53-97 (DnaHit, helpers)
//   json_hunt   restates  src/hunter.h:99-160    (JSON envelope, adjacent (chr,start) dedupe)
//   run_seed    restates  src/silica.h:449-573   FM/NW part of `search` (thal gate left out:
//                                                 every candidate is reported)
//   cmd_padcount restates src/padlock.h:381-427  (exact + neighbourhood count totals)
//   cmd_thal    calls     src/thal.h verbatim    (get_thermodynamic_values + thal(), thal_end1, temponly:
//                                                 one temperature per (oligo, site) line; dumps the tables)
//   cmd_search  restates  src/silica.h:208-660   (`search` end to end around the verbatim SDSL / neighbors.h /
//                                                 needle.h / thal.h / nlohmann code: Tm gate, de-duplication,
//                                                 amplicon pairing, penalties, writeJsonPrimerOut)
//   cmd_jsonfloat prints  nlohmann::json(double).dump() for 64-bit patterns (the number format the
//                                                 product's jsonnum.hpp must reproduce)
//
// Built by oracle/Makefile into oracle/_ref/ (git-ignored, travels to the GPU box).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// execute it.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <boost/dynamic_bitset.hpp>  // oracle/shim
#include <boost/multi_array.hpp>     // oracle/shim

#include <sdsl/suffix_arrays.hpp>  // verbatim reference (src/xxsds/include)
#include <nlohmann/json.hpp>       // verbatim reference (src/jlib)

#include "neighbors.h"  // verbatim reference (src/neighbors.h)
#include "needle.h"     // verbatim reference (src/needle.h, src/align.h)
#include "version.h"    // verbatim reference (src/version.h)
#include "thal.h"       // verbatim reference (src/thal.h: primer3 thal, the Tm gate of silica.h:508-519)

using namespace sdsl;

namespace refdrv {

typedef csa_wt<> TIndex;  // hunter.h:252, silica.h:339, index.h:79

// ---------------------------------------------------------------- helpers (restated)

// util.h:54-114 (complement / revcomplement / reverseComplement; input is upper-cased first)
static char complement(char n) {
  switch (n) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    case 'U': return 'A';
    case 'R': return 'Y';
    case 'Y': return 'R';
    case 'S': return 'S';
    case 'W': return 'W';
    case 'K': return 'M';
    case 'M': return 'K';
    case 'B': return 'V';
    case 'V': return 'B';
    case 'D': return 'H';
    case 'H': return 'D';
    case 'N': return 'N';
  }
  return 'N';
}
static void reverseComplement(std::string& s) {
  for (auto& ch : s) ch = (char)std::toupper((unsigned char)ch);
  for (auto& ch : s) ch = complement(ch);
  std::reverse(s.begin(), s.end());
}
// util.h:208-219
static std::string replaceNonDna(std::string const& str, std::vector<std::string>& msg) {
  std::string out;
  for (uint32_t i = 0; i < str.size(); ++i) {
    if ((str[i] == 'A') || (str[i] == 'C') || (str[i] == 'G') || (str[i] == 'T')) out.append(str, i, 1);
    else {
      msg.push_back("Warning: Non-DNA character in nucleotide sequence detected and replaced by 'N'!");
      out.append("N");
    }
  }
  return out;
}

// hunter.h:53-66
struct DnaHit {
  int32_t score;
  uint32_t chr;
  uint32_t start;
  char strand;
  std::string refalign;
  std::string queryalign;
  DnaHit(int32_t sc, uint32_t const refIndex, uint32_t const s, char const orient, std::string const& ra,
         std::string const& qa)
      : score(sc), chr(refIndex), start(s), strand(orient), refalign(ra), queryalign(qa) {}
  bool operator<(const DnaHit& b) const {
    return ((score > b.score) || ((score == b.score) && (chr < b.chr)) ||
            ((score == b.score) && (chr == b.chr) && (start < b.start)));
  }
};

// hunter.h:69-77
template <typename TAlign>
static uint32_t trailGap(TAlign const& align) {
  uint32_t lastAlignedPos = align.shape()[1] - 1;
  for (uint32_t j = 0; j < align.shape()[1]; ++j) {
    if (align[1][j] != '-') lastAlignedPos = j;
  }
  return (align.shape()[1] - lastAlignedPos - 1);
}
// hunter.h:79-88
template <typename TScore>
static int32_t hammingScore(std::string const& a, std::string const& b, TScore const& sc) {
  int32_t score = 0;
  for (uint32_t i = 0; ((i < a.size()) && (i < b.size())); ++i) {
    if (a[i] == b[i]) score += sc.match;
    else score += sc.mismatch;
  }
  return score;
}
// hunter.h:90-97
static uint32_t nucleotideLength(std::string const& seq) {
  uint32_t s = 0;
  for (uint32_t i = 0; i < seq.size(); ++i)
    if (seq[i] != '-') ++s;
  return s;
}

struct HuntParams {
  bool indel = true;
  bool reverse = true;
  uint32_t distance = 1;
  uint32_t maxNeighborhood = 10000;
  std::size_t max_locations = 1000;
  bool counters = false;
};

// Work counters of SURVEY.md section 8(d): R rank queries, L locate LF steps, H hits,
// X context bytes, plus neighbour strings searched and backward-search iterations.
struct Work {
  uint64_t R = 0, L = 0, H = 0, X = 0, strings = 0, steps = 0;
  void add(Work const& o) { R += o.R; L += o.L; H += o.H; X += o.X; strings += o.strings; steps += o.steps; }
};

struct QueryResult {
  std::string name;
  std::string sequence;      // after upper-casing / N replacement (meta "sequence")
  uint32_t distance = 0;     // possibly clamped
  bool tooShort = false;
  std::vector<std::string> msg;
  std::vector<DnaHit> push;  // in push order
  std::vector<DnaHit> sorted;  // after std::sort
  Work work;
};

// Number of backward_search iterations sdsl::count executes for `s`
// (suffix_array_algorithm.hpp:207-226): loop runs while the interval is non-empty.
static uint32_t executedSteps(TIndex const& fm, std::string const& s) {
  uint64_t l = 0, r = fm.size() - 1;
  uint32_t e = 0;
  auto it = s.end();
  while (s.begin() < it && r + 1 - l > 0) {
    --it;
    backward_search(fm, l, r, (TIndex::char_type)*it, l, r);
    ++e;
  }
  return e;
}

// hunter.h:289-444 for one query.
static void run_hunt(TIndex const& fm_index, std::vector<uint32_t> const& seqlen, std::string const& qname,
                     std::string const& qseq, HuntParams const& p, QueryResult& out) {
  out.name = qname;
  out.sequence = qseq;
  out.distance = p.distance;
  std::vector<DnaHit>& ht = out.push;
  std::vector<std::string>& msg = out.msg;

  // hunter.h:299-303
  if (out.sequence.size() < 10) {
    msg.push_back("Error: Input sequence is shorter than 10 nucleotides!");
    out.tooShort = true;
    return;
  }
  // hunter.h:306-309
  for (auto& ch : out.sequence) ch = (char)std::toupper((unsigned char)ch);
  out.sequence = replaceNonDna(out.sequence, msg);
  std::string revSequence = out.sequence;
  reverseComplement(revSequence);
  // hunter.h:312-315
  if (out.distance >= out.sequence.size()) {
    out.distance = out.sequence.size() - 1;
    msg.push_back("Warning: Distance was adjusted to sequence length!");
  }
  // hunter.h:318-323
  std::size_t pre_context = 0, post_context = 0;
  if (p.indel) {
    pre_context += out.distance;
    post_context += out.distance;
  }
  // hunter.h:326-339
  typedef std::set<char> TAlphabet;
  char tmp[] = {'A', 'C', 'G', 'T'};
  TAlphabet alphabet(tmp, tmp + sizeof(tmp) / sizeof(tmp[0]));
  typedef std::set<std::string> TStringSet;
  std::vector<TStringSet> fwrv(2, TStringSet());
  dicey::neighbors(out.sequence, alphabet, out.distance, p.indel, p.maxNeighborhood, fwrv[0]);
  if (p.reverse) dicey::neighbors(revSequence, alphabet, out.distance, p.indel, p.maxNeighborhood, fwrv[1]);
  // hunter.h:342-345
  if ((fwrv[0].size() >= p.maxNeighborhood) || (fwrv[1].size() >= p.maxNeighborhood)) {
    std::string m = "Warning: Neighborhood size exceeds " + std::to_string(p.maxNeighborhood) +
                    " candidates. Only first " + std::to_string(p.maxNeighborhood) +
                    " neighbors are searched, results are likely incomplete!";
    msg.push_back(m);
  }
  // hunter.h:348-433
  uint32_t hits = 0;
  for (uint32_t fwrvidx = 0; fwrvidx < fwrv.size(); ++fwrvidx) {
    for (TStringSet::const_iterator it = fwrv[fwrvidx].begin();
         ((it != fwrv[fwrvidx].end()) && (hits < p.max_locations)); ++it) {
      std::string query = *it;
      std::size_t m = query.size();
      std::size_t occs = sdsl::count(fm_index, query.begin(), query.end());
      if (p.counters) {
        uint32_t e = executedSteps(fm_index, query);
        out.work.strings += 1;
        out.work.steps += e;
        out.work.R += 2 * (uint64_t)(e > 0 ? e - 1 : 0);
      }
      if (occs > 0) {
        auto locations = locate(fm_index, query.begin(), query.begin() + m);
        std::sort(locations.begin(), locations.end());
        for (std::size_t i = 0; ((i < std::min(occs, p.max_locations)) && (hits < p.max_locations)); ++i) {
          int64_t bestPos = locations[i];
          int64_t cumsum = 0;
          uint32_t refIndex = 0;
          for (; (refIndex + 1 < seqlen.size()) && (bestPos >= cumsum + seqlen[refIndex]); ++refIndex)
            cumsum += seqlen[refIndex];
          uint32_t chrpos = bestPos - cumsum;
          std::size_t pre_extract = pre_context;
          std::size_t post_extract = post_context;
          if (pre_extract > locations[i]) pre_extract = locations[i];
          if (locations[i] + m + post_extract > fm_index.size()) post_extract = fm_index.size() - locations[i] - m;
          auto s = extract(fm_index, locations[i] - pre_extract, locations[i] + m + post_extract - 1);
          if (p.counters) {
            out.work.H += 1;
            out.work.X += m + pre_extract + post_extract;
            // csa_wt.hpp:340-354: LF steps until the SA index is a multiple of the sampling rate.
            // Recover the SA index from the text position through the ISA (test-side only).
            uint64_t idx = fm_index.isa[locations[i]];
            while (idx % TIndex::sa_sample_dens != 0) { idx = fm_index.lf[idx]; out.work.L += 1; }
          }
          std::string pre = s.substr(0, pre_extract);
          s = s.substr(pre_extract);
          if (pre.find_last_of('\n') != std::string::npos) pre = pre.substr(pre.find_last_of('\n') + 1);
          std::string post = s.substr(m);
          post = post.substr(0, post.find_first_of('\n'));

          std::string genomicseq = pre + s.substr(0, m) + post;
          if (pre.size() < chrpos) chrpos -= pre.size();
          dicey::DnaScore<int32_t> sc(0, -1, -1, -1);
          typedef boost::multi_array<char, 2> TAlign;
          dicey::AlignConfig<false, true> global;
          std::string const& qq = (fwrvidx == 0) ? out.sequence : revSequence;
          char strand = (fwrvidx == 0) ? '+' : '-';
          if (p.indel) {
            TAlign align;
            int32_t score = dicey::needle(genomicseq, qq, align, global, sc);
            std::string refalign = "";
            std::string queryalign = "";
            bool leadGap = true;
            for (uint32_t j = 0; (j < (align.shape()[1] - trailGap(align))); ++j) {
              if (align[1][j] != '-') leadGap = false;
              if (!leadGap) {
                refalign += align[0][j];
                queryalign += align[1][j];
              } else {
                ++chrpos;
              }
            }
            ht.push_back(DnaHit(score, refIndex, chrpos + 1, strand, refalign, queryalign));
          } else {
            int32_t score = hammingScore(genomicseq, qq, sc);
            ht.push_back(DnaHit(score, refIndex, chrpos + 1, strand, genomicseq, qq));
          }
          ++hits;
        }
      }
    }
  }
  // hunter.h:434-437
  if (hits >= p.max_locations) {
    std::string m = "Warning: More than " + std::to_string(p.max_locations) + " matches found. Only first " +
                    std::to_string(p.max_locations) + " matches are reported, results are likely incomplete!";
    msg.push_back(m);
  }
  // hunter.h:440
  out.sorted = ht;
  std::sort(out.sorted.begin(), out.sorted.end());
}

// hunter.h:99-160; `genome`/`outfile` strings are what the CLI was given.
static std::string json_hunt(QueryResult const& q, HuntParams const& p, std::vector<std::string> const& qn,
                             std::string const& genome, std::string const& outfile) {
  std::ostringstream rcfile;
  bool errors = false;
  rcfile << "{";
  rcfile << "\"errors\": [";
  for (uint32_t i = 0; i < q.msg.size(); ++i) {
    std::string msgtype = "warning";
    if (q.msg[i].rfind("Error", 0) == 0) {
      errors = true;
      msgtype = "error";
    }
    nlohmann::json err;
    err["type"] = msgtype;
    err["title"] = q.msg[i];
    if (i > 0) rcfile << ',';
    rcfile << err.dump();
  }
  rcfile << "]";
  if (!errors) {
    rcfile << ",\"meta\":";
    nlohmann::json meta;
    meta["version"] = dicey::diceyVersionNumber;
    meta["subcommand"] = "hunt";
    meta["distance"] = q.distance;
    meta["sequence"] = q.sequence;
    if (!q.name.empty()) meta["name"] = q.name;
    meta["genome"] = genome;
    meta["outfile"] = outfile;
    meta["maxmatches"] = p.max_locations;
    meta["hamming"] = (!p.indel);
    meta["forwardonly"] = (!p.reverse);
    rcfile << meta.dump() << ',';
    uint32_t oldchr = 999999;
    uint32_t oldstart = 0;
    bool firstData = true;
    rcfile << "\"data\":[";
    for (uint32_t i = 0; i < q.sorted.size(); ++i) {
      if ((oldchr != q.sorted[i].chr) || (oldstart != q.sorted[i].start)) {
        if (!firstData) rcfile << ',';
        firstData = false;
        nlohmann::json j;
        j["distance"] = std::abs(q.sorted[i].score);
        j["chr"] = qn[q.sorted[i].chr];
        j["start"] = q.sorted[i].start;
        j["end"] = q.sorted[i].start + nucleotideLength(q.sorted[i].refalign) - 1;
        j["strand"] = std::string(1, q.sorted[i].strand);
        j["refalign"] = q.sorted[i].refalign;
        j["queryalign"] = q.sorted[i].queryalign;
        rcfile << j.dump();
      }
      oldchr = q.sorted[i].chr;
      oldstart = q.sorted[i].start;
    }
    rcfile << ']';
  }
  rcfile << '}';
  return rcfile.str();
}

// silica.h:449-573 without the thal gate: every candidate (seed neighbour x occurrence)
// is reported with its extracted genomic context, chromosome position, alignpos and the
// dedupe decision the reference would take if every candidate passed the Tm cut.
struct SeedCand {
  uint32_t strand, refIndex, chrpos, alignpos;
  std::string genomicseq;  // before the |primer| cut (input to thal in the reference)
  std::string nbr;
};
static void run_seed(TIndex const& fm_index, std::vector<uint32_t> const& seqlen, std::string const& primer,
                     uint32_t kmer, uint32_t distance, bool indel, uint32_t maxNeighborhood,
                     std::size_t max_locations, std::vector<SeedCand>& out, uint32_t& hitsOut) {
  typedef std::set<char> TAlphabet;
  char tmp[] = {'A', 'C', 'G', 'T'};
  TAlphabet alphabet(tmp, tmp + sizeof(tmp) / sizeof(tmp[0]));
  typedef std::set<std::string> TStringSet;
  std::vector<TStringSet> fwrv(2, TStringSet());
  std::string sequence = primer;
  uint32_t koffset = sequence.size() - kmer;
  sequence = sequence.substr(sequence.size() - kmer);
  dicey::neighbors(sequence, alphabet, distance, indel, maxNeighborhood, fwrv[0]);
  std::string revSequence = sequence;
  reverseComplement(revSequence);
  dicey::neighbors(revSequence, alphabet, distance, indel, maxNeighborhood, fwrv[1]);
  std::size_t pre_context = 0, post_context = 0;  // silica.h:412-418
  if (indel) {
    pre_context += distance;
    post_context += distance;
  }
  uint32_t hits = 0;
  for (uint32_t fwrvidx = 0; fwrvidx < fwrv.size(); ++fwrvidx) {
    for (TStringSet::const_iterator it = fwrv[fwrvidx].begin();
         ((it != fwrv[fwrvidx].end()) && (hits < max_locations)); ++it) {
      std::string query = *it;
      std::size_t m = query.size();
      std::size_t occs = sdsl::count(fm_index, query.begin(), query.end());
      if (occs > 0) {
        auto locations = locate(fm_index, query.begin(), query.begin() + m);
        std::sort(locations.begin(), locations.end());
        for (std::size_t i = 0; ((i < std::min(occs, max_locations)) && (hits < max_locations)); ++i) {
          int64_t bestPos = locations[i];
          int64_t cumsum = 0;
          uint32_t refIndex = 0;
          for (; (refIndex + 1 < seqlen.size()) && (bestPos >= cumsum + seqlen[refIndex]); ++refIndex)
            cumsum += seqlen[refIndex];
          uint32_t chrpos = bestPos - cumsum;
          std::size_t pre_extract = pre_context;
          std::size_t post_extract = post_context;
          if (fwrvidx) post_extract += koffset;
          else pre_extract += koffset;
          if (pre_extract > locations[i]) pre_extract = locations[i];
          if (locations[i] + m + post_extract > fm_index.size()) post_extract = fm_index.size() - locations[i] - m;
          auto s = extract(fm_index, locations[i] - pre_extract, locations[i] + m + post_extract - 1);
          std::string pre = s.substr(0, pre_extract);
          s = s.substr(pre_extract);
          if (pre.find_last_of('\n') != std::string::npos) pre = pre.substr(pre.find_last_of('\n') + 1);
          std::string post = s.substr(m);
          post = post.substr(0, post.find_first_of('\n'));
          std::string genomicseq = pre + s.substr(0, m) + post;
          if (pre.size() <= chrpos) chrpos -= pre.size();
          std::string searchSeq = fwrvidx ? revSequence : sequence;
          uint32_t alignpos = chrpos;
          dicey::DnaScore<int32_t> sc(0, -1, -1, -1);
          typedef boost::multi_array<char, 2> TAlign;
          dicey::AlignConfig<false, true> global;
          TAlign align;
          dicey::needle(genomicseq, searchSeq, align, global, sc);
          bool leadGap = true;
          for (uint32_t j = 0; (j < (align.shape()[1] - trailGap(align))); ++j) {
            if (align[1][j] != '-') leadGap = false;
            if (leadGap) ++alignpos;
          }
          SeedCand c;
          c.strand = fwrvidx;
          c.refIndex = refIndex;
          c.chrpos = chrpos;
          c.alignpos = alignpos;
          c.genomicseq = genomicseq;
          c.nbr = query;
          out.push_back(c);
          ++hits;
        }
      }
    }
  }
  hitsOut = hits;
}

// ---------------------------------------------------------------- I/O helpers

static bool read_lines(std::string const& path, std::vector<std::string>& lines) {
  std::ifstream f(path.c_str());
  if (!f.is_open()) return false;
  std::string line;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (!line.empty()) lines.push_back(line);
  }
  return true;
}
// "name<TAB>length" per record (the first two columns of a .fai); seqlen = length + 1 (util.h:201)
static bool read_records(std::string const& path, std::vector<std::string>& names, std::vector<uint32_t>& seqlen) {
  std::vector<std::string> lines;
  if (!read_lines(path, lines)) return false;
  for (auto const& l : lines) {
    std::istringstream is(l);
    std::string n;
    uint64_t len;
    if (!(is >> n >> len)) return false;
    names.push_back(n);
    seqlen.push_back((uint32_t)(len + 1));
  }
  return !names.empty();
}
static void split_query(std::string const& line, std::string& name, std::string& seq) {
  std::size_t t = line.find('\t');
  if (t == std::string::npos) {
    name.clear();
    seq = line;
  } else {
    name = line.substr(0, t);
    seq = line.substr(t + 1);
  }
}

static int usage() {
  std::cerr
      << "dicey_ref (oracle built from the reference sources; test infrastructure)\n"
         "  index <dump.txt> <out.fm9> [tmpdir]\n"
         "  hunt  <in.fm9> <records.tsv> <queries.txt> [-d D] [-n] [-f] [-m MAXMATCH] [-x MAXNBR]\n"
         "        [--records out.tsv] [--json out.jsonl] [--genome STR] [--outfile STR]\n"
         "        [--counters] [--threads P] [--limit N]\n"
         "  seed  <in.fm9> <records.tsv> <primers.txt> [-k K] [-d D] [-n] [-m MAXLOC] [-x MAXNBR]\n"
         "  neighbors <query> [-d D] [-n] [-x MAXNBR]\n"
         "  needle <genomic> <query>\n"
         "  count <in.fm9> <patterns.txt>       (l r occ per pattern)\n"
         "  locate <in.fm9> <patterns.txt>      (sorted positions per pattern)\n"
         "  extract <in.fm9> <lo> <hi>\n"
         "  padcount <in.fm9> <arms.txt> [-d D] [-n]\n"
         "  dump <in.fm9>                       (n, sigma, C, BWT, SA, text; small indexes only)\n";
  return 2;
}

struct Args {
  std::vector<std::string> pos;
  uint32_t d = 1, x = 10000, k = 15, threads = 1;
  std::size_t m = 1000;
  bool mset = false;
  bool hamming = false, forward = false, counters = false;
  std::string records, json, genome = "genome.fa.gz", outfile = "";
  std::size_t limit = 0;
  double cutTemp = 45.0;       // silica.h:232
  uint32_t maxProd = 15000;    // silica.h:233
  bool prune = false;          // silica.h:226 (-q)
  uint32_t pruneCount = 0;
};
static bool parse(int argc, char** argv, Args& a) {
  for (int i = 2; i < argc; ++i) {
    std::string s = argv[i];
    auto need = [&](const char* what) -> const char* {
      if (i + 1 >= argc) { std::cerr << "missing value for " << what << "\n"; std::exit(2); }
      return argv[++i];
    };
    if (s == "-d") a.d = std::strtoul(need("-d"), 0, 10);
    else if (s == "-x") a.x = std::strtoul(need("-x"), 0, 10);
    else if (s == "-k") a.k = std::strtoul(need("-k"), 0, 10);
    else if (s == "-m") { a.m = std::strtoull(need("-m"), 0, 10); a.mset = true; }
    else if (s == "-n") a.hamming = true;
    else if (s == "-c") a.cutTemp = std::strtod(need("-c"), 0);
    else if (s == "-l") a.maxProd = std::strtoul(need("-l"), 0, 10);
    else if (s == "-q") { a.prune = true; a.pruneCount = std::strtoul(need("-q"), 0, 10); }
    else if (s == "-f") a.forward = true;
    else if (s == "--counters") a.counters = true;
    else if (s == "--threads") a.threads = std::strtoul(need("--threads"), 0, 10);
    else if (s == "--records") a.records = need("--records");
    else if (s == "--json") a.json = need("--json");
    else if (s == "--genome") a.genome = need("--genome");
    else if (s == "--outfile") a.outfile = need("--outfile");
    else if (s == "--limit") a.limit = std::strtoull(need("--limit"), 0, 10);
    else a.pos.push_back(s);
  }
  return true;
}

// index.h:96-123 (the FASTA -> dump step is done by the caller; the dump format is
// "records upper-cased, joined by '\n', trailing '\n'").
static int cmd_index(Args const& a) {
  if (a.pos.size() < 2) return usage();
  std::string tmpdir = a.pos.size() > 2 ? a.pos[2] : std::string("/tmp");
  TIndex fm_index;
  cache_config config(true, tmpdir, "dicey_ref_" + std::to_string((unsigned long)getpid()));
  construct(fm_index, a.pos[0], config, 1);
  if (!store_to_checked_file(fm_index, a.pos[1])) {
    std::cerr << "cannot store " << a.pos[1] << "\n";
    return 1;
  }
  std::cout << "{\"n\": " << fm_index.size() << ", \"sigma\": " << (uint32_t)fm_index.sigma << "}" << std::endl;
  return 0;
}

static bool load_index(std::string const& path, TIndex& fm) {
  if (!load_from_checked_file(fm, path)) {  // hunter.h:256
    std::cerr << "Error: FM-Index cannot be loaded!\n";
    return false;
  }
  return true;
}

static void write_hit(std::ostream& os, char tag, DnaHit const& h) {
  os << tag << '\t' << h.score << '\t' << h.chr << '\t' << h.start << '\t' << h.strand << '\t' << h.refalign << '\t'
     << h.queryalign << '\n';
}

static int cmd_hunt(Args const& a) {
  if (a.pos.size() < 3) return usage();
  TIndex fm;
  auto t0 = std::chrono::steady_clock::now();
  if (!load_index(a.pos[0], fm)) return 1;
  auto t1 = std::chrono::steady_clock::now();
  std::vector<std::string> names;
  std::vector<uint32_t> seqlen;
  if (!read_records(a.pos[1], names, seqlen)) { std::cerr << "bad records file\n"; return 1; }
  std::vector<std::string> qlines;
  if (!read_lines(a.pos[2], qlines)) { std::cerr << "bad query file\n"; return 1; }
  if (a.limit && qlines.size() > a.limit) qlines.resize(a.limit);
  HuntParams p;
  p.indel = !a.hamming;
  p.reverse = !a.forward;
  p.distance = a.d;
  p.maxNeighborhood = a.x;
  p.max_locations = a.mset ? a.m : 1000;
  p.counters = a.counters;
  std::vector<QueryResult> res(qlines.size());
  uint32_t P = std::max<uint32_t>(1, a.threads);
  auto t2 = std::chrono::steady_clock::now();
  {
    // The reference is single-threaded; with --threads P the query list is split into
    // P contiguous shards run by independent threads over the shared read-only index.
    std::vector<std::thread> th;
    std::atomic<std::size_t> next(0);
    for (uint32_t t = 0; t < P; ++t) {
      th.emplace_back([&]() {
        const std::size_t chunk = 16;
        for (;;) {
          std::size_t b = next.fetch_add(chunk);
          if (b >= qlines.size()) break;
          std::size_t e = std::min(qlines.size(), b + chunk);
          for (std::size_t i = b; i < e; ++i) {
            std::string name, seq;
            split_query(qlines[i], name, seq);
            run_hunt(fm, seqlen, name, seq, p, res[i]);
          }
        }
      });
    }
    for (auto& t : th) t.join();
  }
  auto t3 = std::chrono::steady_clock::now();
  Work total;
  uint64_t nhits = 0;
  for (auto const& r : res) { total.add(r.work); nhits += r.push.size(); }
  if (!a.records.empty()) {
    std::ofstream os(a.records.c_str());
    for (std::size_t i = 0; i < res.size(); ++i) {
      QueryResult const& r = res[i];
      os << "Q\t" << i << '\t' << r.sequence << '\t' << r.distance << '\t' << r.msg.size() << '\t' << r.push.size()
         << '\n';
      for (auto const& m : r.msg) os << "M\t" << m << '\n';
      for (auto const& h : r.push) write_hit(os, 'P', h);
      for (auto const& h : r.sorted) write_hit(os, 'S', h);
      if (a.counters)
        os << "W\t" << r.work.strings << '\t' << r.work.steps << '\t' << r.work.R << '\t' << r.work.L << '\t'
           << r.work.H << '\t' << r.work.X << '\n';
    }
  }
  if (!a.json.empty()) {
    std::ofstream os(a.json.c_str());
    for (auto const& r : res) os << json_hunt(r, p, names, a.genome, a.outfile) << std::endl;  // hunter.h:159
  }
  double load_s = std::chrono::duration<double>(t1 - t0).count();
  double loop_s = std::chrono::duration<double>(t3 - t2).count();
  std::cout << "{\"queries\": " << res.size() << ", \"threads\": " << P << ", \"load_s\": " << load_s
            << ", \"loop_s\": " << loop_s << ", \"queries_per_s\": " << (loop_s > 0 ? res.size() / loop_s : 0.0)
            << ", \"hits\": " << nhits << ", \"n\": " << fm.size();
  if (a.counters)
    std::cout << ", \"strings\": " << total.strings << ", \"steps\": " << total.steps << ", \"R\": " << total.R
              << ", \"L\": " << total.L << ", \"H\": " << total.H << ", \"X\": " << total.X;
  std::cout << "}" << std::endl;
  return 0;
}

static int cmd_seed(Args const& a) {
  if (a.pos.size() < 3) return usage();
  TIndex fm;
  if (!load_index(a.pos[0], fm)) return 1;
  std::vector<std::string> names;
  std::vector<uint32_t> seqlen;
  if (!read_records(a.pos[1], names, seqlen)) return 1;
  std::vector<std::string> qlines;
  if (!read_lines(a.pos[2], qlines)) return 1;
  std::size_t maxloc = a.mset ? a.m : 10000;  // silica.h:223
  for (std::size_t i = 0; i < qlines.size(); ++i) {
    std::string name, seq;
    split_query(qlines[i], name, seq);
    for (auto& ch : seq) ch = (char)std::toupper((unsigned char)ch);
    if (seq.size() <= a.k) {  // silica.h:371 keeps only primers longer than kmer... reported as skipped
      std::cout << "Q\t" << i << "\tskipped\n";
      continue;
    }
    std::vector<SeedCand> cands;
    uint32_t hits = 0;
    run_seed(fm, seqlen, seq, a.k, a.d, !a.hamming, a.x, maxloc, cands, hits);
    std::cout << "Q\t" << i << '\t' << seq << '\t' << hits << '\n';
    for (auto const& c : cands)
      std::cout << "C\t" << c.strand << '\t' << c.refIndex << '\t' << c.chrpos << '\t' << c.alignpos << '\t'
                << c.genomicseq << '\t' << c.nbr << '\n';
  }
  return 0;
}

static int cmd_neighbors(Args const& a) {
  if (a.pos.size() < 1) return usage();
  typedef std::set<char> TAlphabet;
  char tmp[] = {'A', 'C', 'G', 'T'};
  TAlphabet alphabet(tmp, tmp + 4);
  // queries: either one literal or, if the argument names a readable file, one per line
  std::vector<std::string> qs;
  if (!read_lines(a.pos[0], qs)) qs.push_back(a.pos[0]);
  for (auto const& q : qs) {
    std::set<std::string> st;
    dicey::neighbors(q, alphabet, a.d, !a.hamming, a.x, st);
    std::cout << "Q\t" << q << '\t' << st.size() << '\n';
    for (auto const& s : st) std::cout << s << '\n';
  }
  return 0;
}

static int cmd_needle(Args const& a) {
  if (a.pos.size() < 2) return usage();
  // pairs: literal "<genomic> <query>" or a file with "genomic<TAB>query" lines
  std::vector<std::pair<std::string, std::string>> pairs;
  std::vector<std::string> lines;
  if (a.pos[1] == "-" && read_lines(a.pos[0], lines)) {
    for (auto const& l : lines) {
      std::size_t t = l.find('\t');
      pairs.push_back(std::make_pair(l.substr(0, t), l.substr(t + 1)));
    }
  } else {
    pairs.push_back(std::make_pair(a.pos[0], a.pos[1]));
  }
  for (auto const& pr : pairs) {
    dicey::DnaScore<int32_t> sc(0, -1, -1, -1);
    typedef boost::multi_array<char, 2> TAlign;
    dicey::AlignConfig<false, true> global;
    TAlign align;
    int32_t score = dicey::needle(pr.first, pr.second, align, global, sc);
    std::string r0, r1;
    for (std::size_t j = 0; j < align.shape()[1]; ++j) {
      r0 += align[0][j];
      r1 += align[1][j];
    }
    std::cout << score << '\t' << r0 << '\t' << r1 << '\n';
  }
  return 0;
}

static int cmd_count(Args const& a, bool doLocate) {
  if (a.pos.size() < 2) return usage();
  TIndex fm;
  if (!load_index(a.pos[0], fm)) return 1;
  std::vector<std::string> pats;
  if (!read_lines(a.pos[1], pats)) return 1;
  for (auto const& s : pats) {
    uint64_t l = 0, r = 0;
    uint64_t occ = backward_search(fm, 0, fm.size() - 1, s.begin(), s.end(), l, r);
    uint64_t cnt = sdsl::count(fm, s.begin(), s.end());
    if (cnt != occ) { std::cerr << "count/backward_search disagree\n"; return 1; }
    if (!doLocate) {
      std::cout << l << '\t' << r << '\t' << occ << '\n';
    } else {
      auto loc = locate(fm, s.begin(), s.end());
      std::sort(loc.begin(), loc.end());
      std::cout << occ;
      for (auto v : loc) std::cout << '\t' << v;
      std::cout << '\n';
    }
  }
  return 0;
}

static int cmd_extract(Args const& a) {
  if (a.pos.size() < 3) return usage();
  TIndex fm;
  if (!load_index(a.pos[0], fm)) return 1;
  uint64_t lo = std::strtoull(a.pos[1].c_str(), 0, 10), hi = std::strtoull(a.pos[2].c_str(), 0, 10);
  auto s = extract(fm, lo, hi);
  std::cout.write(s.data(), s.size());
  return 0;
}

// padlock.h:381-427: exact counts of the four arm strings and the neighbourhood totals with
// early exit once the sum exceeds maxNeighborHits (padlock.h:153-155).
static int cmd_padcount(Args const& a) {
  if (a.pos.size() < 2) return usage();
  TIndex fm;
  if (!load_index(a.pos[0], fm)) return 1;
  std::vector<std::string> arms;
  if (!read_lines(a.pos[1], arms)) return 1;
  typedef std::set<char> TAlphabet;
  char tmp[] = {'A', 'C', 'G', 'T'};
  TAlphabet alphabet(tmp, tmp + 4);
  bool indel = !a.hamming;
  for (auto const& arm : arms) {
    std::string rarm = arm;
    reverseComplement(rarm);
    uint64_t exact = sdsl::count(fm, arm.begin(), arm.end()) + sdsl::count(fm, rarm.begin(), rarm.end());
    std::set<std::string> fw, rv;
    dicey::neighbors(arm, alphabet, a.d, indel, 10000, fw);  // padlock.h:396
    dicey::neighbors(rarm, alphabet, a.d, indel, 10000, rv);
    uint64_t total = 0;
    for (auto const& s : fw) total += sdsl::count(fm, s.begin(), s.end());
    for (auto const& s : rv) total += sdsl::count(fm, s.begin(), s.end());
    std::cout << arm << '\t' << exact << '\t' << total << '\n';
  }
  return 0;
}

// thal <primer3_config_dir/> <pairs.tsv> [params_out.tsv]: the Tm gate of `dicey search`
// (silica.h:316-329, 508-512): thal_end1, dicey's default conditions (37 C, 50 mM monovalent,
// 1.5 mM divalent, 0.6 mM dNTP, 50 nM oligo).  One line per pair: success flag, Tm as %.17g and as
// the raw 64-bit pattern.  With a third argument the thermodynamic tables the reference loaded
// from its parameter files are dumped (bit patterns), for tests that cannot read /root/reference.
static void dump_table(std::ostream& os, const char* name, const double* p, size_t n) {
  os << name << '\t' << n;
  for (size_t i = 0; i < n; ++i) {
    uint64_t u;
    memcpy(&u, p + i, 8);
    os << '\t' << std::hex << u << std::dec;
  }
  os << '\n';
}
static int cmd_thal(Args const& a) {
  if (a.pos.size() < 2) return usage();
  std::string cfg = a.pos[0];
  if (cfg.empty() || cfg.back() != '/') cfg += '/';
  primer3thal::thal_args ta;
  primer3thal::set_thal_default_args(&ta);
  ta.temponly = 1;
  ta.type = primer3thal::thal_end1;
  primer3thal::get_thermodynamic_values(cfg.c_str());
  ta.temp = 37.0; ta.mv = 50.0; ta.dv = 1.5; ta.dna_conc = 50.0; ta.dntp = 0.6;   // silica.h:242-246
  ta.temp += primer3thal::ABSOLUTE_ZERO;
  std::vector<std::string> lines;
  if (!read_lines(a.pos[1], lines)) return 1;
  for (auto const& line : lines) {
    size_t t = line.find('\t');
    if (t == std::string::npos) continue;
    std::string o1 = line.substr(0, t), o2 = line.substr(t + 1);
    primer3thal::thal_results o;
    bool ok = primer3thal::thal((const unsigned char*)o1.c_str(), (const unsigned char*)o2.c_str(), &ta, &o);
    uint64_t u;
    double tm = o.temp;
    memcpy(&u, &tm, 8);
    char buf[64];
    snprintf(buf, sizeof(buf), "%.17g", tm);
    std::cout << (ok ? 1 : 0) << '\t' << buf << '\t' << std::hex << u << std::dec << '\n';
  }
  if (a.pos.size() >= 3) {
    using namespace primer3thal;
    std::ofstream os(a.pos[2].c_str());
    dump_table(os, "stackEntropies", &stackEntropies[0][0][0][0], 625);
    dump_table(os, "stackEnthalpies", &stackEnthalpies[0][0][0][0], 625);
    dump_table(os, "stackint2Entropies", &stackint2Entropies[0][0][0][0], 625);
    dump_table(os, "stackint2Enthalpies", &stackint2Enthalpies[0][0][0][0], 625);
    dump_table(os, "dangleEntropies3", &dangleEntropies3[0][0][0], 125);
    dump_table(os, "dangleEnthalpies3", &dangleEnthalpies3[0][0][0], 125);
    dump_table(os, "dangleEntropies5", &dangleEntropies5[0][0][0], 125);
    dump_table(os, "dangleEnthalpies5", &dangleEnthalpies5[0][0][0], 125);
    dump_table(os, "interiorLoopEntropies", interiorLoopEntropies, 30);
    dump_table(os, "bulgeLoopEntropies", bulgeLoopEntropies, 30);
    dump_table(os, "interiorLoopEnthalpies", interiorLoopEnthalpies, 30);
    dump_table(os, "bulgeLoopEnthalpies", bulgeLoopEnthalpies, 30);
    dump_table(os, "tstackEntropies", &tstackEntropies[0][0][0][0], 625);
    dump_table(os, "tstackEnthalpies", &tstackEnthalpies[0][0][0][0], 625);
    dump_table(os, "tstack2Entropies", &tstack2Entropies[0][0][0][0], 625);
    dump_table(os, "tstack2Enthalpies", &tstack2Enthalpies[0][0][0][0], 625);
    dump_table(os, "atpS", &atpS[0][0], 25);
    dump_table(os, "atpH", &atpH[0][0], 25);
    double sc = saltCorrectS(ta.mv, ta.dv, ta.dntp);
    double rc[2] = {R * log(ta.dna_conc / 1000000000.0), R * log(ta.dna_conc / 4000000000.0)};
    dump_table(os, "saltCorrection", &sc, 1);
    dump_table(os, "RC_symmetric_asymmetric", rc, 2);
  }
  primer3thal::destroy_thal_structures();
  return 0;
}

// search <in.fm9> <records.tsv> <primers.fa> <primer3_config/> [-k K] [-d D] [-n] [-m MAXLOC] [-x MAXNBR]
//        [-c CUTTEMP] [-l MAXPRODSIZE] [-q PRUNE]: restates silica() (silica.h:340-641) and
// writeJsonPrimerOut (silica.h:100-187) around the verbatim SDSL / neighbors.h / needle.h / thal.h /
// nlohmann code; the FM / NW part is run_seed above.  Amplicon sequences come from the index text
// (the reference reads them back from the FASTA with faidx_fetch_seq and upper-cases them).
struct PrimerBind {  // silica.h:69-82
  uint32_t refIndex, pos, primerId;
  bool onFor;
  double temp, perfTemp;
  std::string genome;
  bool operator<(const PrimerBind& b) const { return (temp > b.temp); }
};
struct PcrProduct {  // silica.h:84-98
  uint32_t refIndex, leng, forPos, revPos, forId, revId;
  double forTemp, revTemp, penalty;
  bool operator<(const PcrProduct& b) const { return (penalty < b.penalty); }
};
static int cmd_search(Args const& a) {
  if (a.pos.size() < 4) return usage();
  TIndex fm_index;
  if (!load_index(a.pos[0], fm_index)) return 1;
  std::vector<std::string> seqname;
  std::vector<uint32_t> seqlen;
  if (!read_records(a.pos[1], seqname, seqlen)) return 1;
  uint32_t nseq = seqlen.size();
  std::string cfg = a.pos[3];
  if (cfg.empty() || cfg.back() != '/') cfg += '/';
  uint32_t kmer = a.k, distance = a.d;
  bool indel = !a.hamming;
  std::size_t max_locations = a.mset ? a.m : 10000;
  double cutTemp = a.cutTemp, penDiff = 0.6, penMis = 0.4, penLen = 0.001, cutofPen = -1.0;
  uint32_t maxProdSize = a.maxProd;
  primer3thal::thal_args ta;
  primer3thal::set_thal_default_args(&ta);
  ta.temponly = 1;
  ta.type = primer3thal::thal_end1;
  primer3thal::get_thermodynamic_values(cfg.c_str());
  ta.temp = 37.0; ta.mv = 50.0; ta.dv = 1.5; ta.dna_conc = 50.0; ta.dntp = 0.6;
  ta.temp += primer3thal::ABSOLUTE_ZERO;
  std::vector<std::string> msg, pName, pSeq;
  std::vector<PrimerBind> allp;
  std::vector<PcrProduct> pcrColl;
  bool fatal = false;
  // primers (silica.h:355-408)
  {
    std::ifstream fafile(a.pos[2].c_str());
    std::string fan, tmpfasta, line;
    auto flush = [&]() {
      if ((!fan.empty()) && (!tmpfasta.empty()) && (tmpfasta.size() > kmer)) {
        std::string qr = tmpfasta.substr(tmpfasta.size() - kmer);
        if ((!a.prune) || (sdsl::count(fm_index, qr.begin(), qr.end()) <= a.pruneCount)) {
          reverseComplement(qr);
          if ((!a.prune) || (sdsl::count(fm_index, qr.begin(), qr.end()) <= a.pruneCount)) {
            std::string inseq = replaceNonDna(tmpfasta, msg);
            if ((inseq.size() < 10) || (inseq.size() < kmer)) {
              msg.push_back("Error: Input sequence is shorter than 10 nucleotides or shorter than the selected k-mer length!");
              fatal = true;
              return;
            }
            if (distance >= inseq.size()) {
              distance = inseq.size() - 1;
              msg.push_back("Warning: Distance was adjusted to sequence length!");
            }
            pName.push_back(fan);
            pSeq.push_back(inseq);
          }
        }
      }
    };
    while (!fatal && std::getline(fafile, line)) {
      if (line.empty()) continue;
      if (line[0] == '>') {
        if ((!fan.empty()) && (!tmpfasta.empty()) && (tmpfasta.size() > kmer)) { flush(); tmpfasta = ""; }
        fan = line.substr(1);
      } else {
        std::string up = line;
        for (auto& ch : up) ch = (char)std::toupper((unsigned char)ch);
        tmpfasta += up;
      }
    }
    if (!fatal) flush();
  }
  std::vector<std::vector<PrimerBind>> forBind(nseq), revBind(nseq);
  for (uint32_t primerId = 0; !fatal && primerId < pSeq.size(); ++primerId) {
    std::string forQuery = pSeq[primerId], revQuery = pSeq[primerId];
    reverseComplement(revQuery);
    primer3thal::thal_results oi;
    bool ok1 = primer3thal::thal((const unsigned char*)forQuery.c_str(), (const unsigned char*)revQuery.c_str(), &ta, &oi);
    if ((!ok1) || (oi.temp == primer3thal::THAL_ERROR_SCORE)) { msg.push_back("Error: Thermodynamical calculation failed!"); fatal = true; break; }
    double matchTemp = oi.temp;
    uint32_t koffset = pSeq[primerId].size() - kmer;
    // neighbourhood cap warning (silica.h:456-459)
    {
      typedef std::set<char> TAlphabet;
      char tmp[] = {'A', 'C', 'G', 'T'};
      TAlphabet alphabet(tmp, tmp + 4);
      std::string sequence = pSeq[primerId].substr(pSeq[primerId].size() - kmer), rev = sequence;
      reverseComplement(rev);
      std::set<std::string> f0, f1;
      dicey::neighbors(sequence, alphabet, distance, indel, a.x, f0);
      dicey::neighbors(rev, alphabet, distance, indel, a.x, f1);
      if ((f0.size() >= a.x) || (f1.size() >= a.x))
        msg.push_back("Warning: Neighborhood size exceeds " + std::to_string(a.x) + " candidates. Only first " + std::to_string(a.x) +
                      " neighbors are searched, results are likely incomplete!");
    }
    std::vector<SeedCand> cands;
    uint32_t hits = 0;
    run_seed(fm_index, seqlen, pSeq[primerId], kmer, distance, indel, a.x, max_locations, cands, hits);
    for (uint32_t fwrvidx = 0; !fatal && fwrvidx < 2; ++fwrvidx) {
      std::set<std::pair<uint32_t, uint32_t>> uphit;
      for (auto const& cd : cands) {
        if (cd.strand != fwrvidx) continue;
        std::string primer = fwrvidx ? forQuery : revQuery;
        std::string genomicseq = cd.genomicseq;
        primer3thal::thal_results o;
        bool ok = primer3thal::thal((const unsigned char*)primer.c_str(), (const unsigned char*)genomicseq.c_str(), &ta, &o);
        if ((!ok) || (o.temp == primer3thal::THAL_ERROR_SCORE)) { msg.push_back("Error: Thermodynamical calculation failed!"); fatal = true; break; }
        if (o.temp > cutTemp) {
          uint32_t chrpos = cd.chrpos, alignpos = cd.alignpos, refIndex = cd.refIndex;
          if (uphit.find(std::make_pair(refIndex, alignpos)) == uphit.end()) {
            uphit.insert(std::make_pair(refIndex, alignpos));
            if (fwrvidx) {
              uint32_t alignshift = alignpos - chrpos;
              chrpos = alignpos;
              genomicseq = genomicseq.substr(alignshift, primer.size());
            } else {
              uint32_t alignshift = alignpos - chrpos;
              chrpos = alignpos - koffset;
              if (alignshift >= koffset) {
                alignshift -= koffset;
                genomicseq = genomicseq.substr(alignshift, primer.size());
              }
            }
            PrimerBind prim;
            prim.refIndex = refIndex; prim.temp = o.temp; prim.perfTemp = matchTemp; prim.primerId = primerId; prim.genome = genomicseq;
            prim.onFor = !fwrvidx;
            prim.pos = chrpos;
            (fwrvidx ? revBind : forBind)[refIndex].push_back(prim);
          }
        }
      }
    }
    if (hits >= max_locations)
      msg.push_back("Warning: More than " + std::to_string(max_locations) + " matches found. Only first " + std::to_string(max_locations) +
                    " matches are reported, results are likely incomplete!");
  }
  if (!fatal) {
    for (uint32_t refIndex = 0; refIndex < nseq; ++refIndex) {
      allp.insert(allp.end(), forBind[refIndex].begin(), forBind[refIndex].end());
      allp.insert(allp.end(), revBind[refIndex].begin(), revBind[refIndex].end());
    }
    std::sort(allp.begin(), allp.end());
    if (!a.prune) {
      for (uint32_t refIndex = 0; refIndex < nseq; ++refIndex) {   // silica.h:591-634
        std::vector<std::pair<uint32_t, uint32_t>> rvByPos;
        for (uint32_t k = 0; k < revBind[refIndex].size(); ++k) rvByPos.push_back(std::make_pair(revBind[refIndex][k].pos, k));
        std::sort(rvByPos.begin(), rvByPos.end());
        std::vector<uint32_t> rvPos(rvByPos.size());
        for (uint32_t k = 0; k < rvByPos.size(); ++k) rvPos[k] = rvByPos[k].first;
        std::vector<uint32_t> cand;
        for (auto fw = forBind[refIndex].begin(); fw != forBind[refIndex].end(); ++fw) {
          auto loIt = std::upper_bound(rvPos.begin(), rvPos.end(), fw->pos);
          std::vector<uint32_t>::iterator hiIt;
          uint64_t hiBound = (uint64_t)fw->pos + (uint64_t)maxProdSize;
          if (hiBound >= ((uint64_t)1 << 32)) hiIt = rvPos.end();
          else hiIt = std::upper_bound(rvPos.begin(), rvPos.end(), (uint32_t)hiBound);
          cand.clear();
          for (auto pit = loIt; pit != hiIt; ++pit) cand.push_back(rvByPos[pit - rvPos.begin()].second);
          std::sort(cand.begin(), cand.end());
          for (auto cit = cand.begin(); cit != cand.end(); ++cit) {
            auto rv = revBind[refIndex].begin() + (*cit);
            if ((rv->pos > fw->pos) && (rv->pos + pSeq[rv->primerId].size() - fw->pos <= maxProdSize)) {
              PcrProduct pp;
              pp.refIndex = refIndex; pp.forPos = fw->pos; pp.forTemp = fw->temp; pp.forId = fw->primerId;
              pp.revPos = rv->pos; pp.revTemp = rv->temp; pp.revId = rv->primerId;
              pp.leng = (rv->pos + pSeq[pp.revId].size()) - fw->pos;
              double pen = (fw->perfTemp - fw->temp) * penDiff;
              if (pen < 0) pen = 0;
              double bpen = (rv->perfTemp - rv->temp) * penDiff;
              if (bpen > 0) pen += bpen;
              pen += std::abs(fw->temp - rv->temp) * penMis;
              pen += pp.leng * penLen;
              pp.penalty = pen;
              if ((cutofPen < 0) || (pen < cutofPen)) pcrColl.push_back(pp);
            }
          }
        }
      }
      std::sort(pcrColl.begin(), pcrColl.end());
    }
  }
  // writeJsonPrimerOut (silica.h:100-187)
  std::ostream& rcfile = std::cout;
  bool errors = false;
  rcfile << "{\"errors\": [";
  for (uint32_t i = 0; i < msg.size(); ++i) {
    std::string msgtype = "warning";
    if (msg[i].compare(0, 5, "Error") == 0) { errors = true; msgtype = "error"; }
    nlohmann::json err;
    err["type"] = msgtype;
    err["title"] = msg[i];
    if (i > 0) rcfile << ',';
    rcfile << err.dump();
  }
  rcfile << "]";
  if (!errors) {
    std::vector<uint64_t> cum(nseq + 1, 0);
    for (uint32_t i = 0; i < nseq; ++i) cum[i + 1] = cum[i] + seqlen[i];
    rcfile << ",\"meta\":";
    nlohmann::json meta;
    meta["version"] = dicey::diceyVersionNumber;
    meta["subcommand"] = "search";
    meta["distance"] = distance;
    meta["genome"] = a.genome;
    meta["outfile"] = a.outfile;
    meta["maxmatches"] = max_locations;
    meta["hamming"] = (!indel);
    rcfile << meta.dump() << ',';
    rcfile << "\"data\":{\"primers\":[";
    for (uint32_t i = 0; i < allp.size(); ++i) {
      if (i > 0) rcfile << ',';
      nlohmann::json j;
      j["Chrom"] = seqname[allp[i].refIndex];
      j["Id"] = i;
      j["Tm"] = allp[i].temp;
      j["Pos"] = allp[i].pos + 1;
      j["End"] = allp[i].pos + pSeq[allp[i].primerId].size();
      if (allp[i].onFor) j["Ori"] = "forward";
      else j["Ori"] = "reverse";
      j["Name"] = pName[allp[i].primerId];
      j["MatchTm"] = allp[i].perfTemp;
      j["Seq"] = pSeq[allp[i].primerId];
      j["Genome"] = allp[i].genome;
      rcfile << j.dump();
    }
    rcfile << "],\"amplicons\":[";
    for (uint32_t i = 0; i < pcrColl.size(); ++i) {
      if (i > 0) rcfile << ',';
      nlohmann::json j;
      j["Chrom"] = seqname[pcrColl[i].refIndex];
      j["Id"] = i;
      j["Length"] = pcrColl[i].leng;
      j["Penalty"] = pcrColl[i].penalty;
      j["ForPos"] = pcrColl[i].forPos + 1;
      j["ForEnd"] = pcrColl[i].forPos + pSeq[pcrColl[i].forId].size();
      j["ForTm"] = pcrColl[i].forTemp;
      j["ForName"] = pName[pcrColl[i].forId];
      j["ForSeq"] = pSeq[pcrColl[i].forId];
      j["RevPos"] = pcrColl[i].revPos + 1;
      j["RevEnd"] = pcrColl[i].revPos + pSeq[pcrColl[i].revId].size();
      j["RevTm"] = pcrColl[i].revTemp;
      j["RevName"] = pName[pcrColl[i].revId];
      j["RevSeq"] = pSeq[pcrColl[i].revId];
      // faidx_fetch_seq(fai, chrom, forPos, revPos + |rev| - 1): the same bases from the index text
      uint64_t lo = cum[pcrColl[i].refIndex] + pcrColl[i].forPos;
      uint64_t hi = cum[pcrColl[i].refIndex] + pcrColl[i].revPos + pSeq[pcrColl[i].revId].size() - 1;
      uint64_t last = cum[pcrColl[i].refIndex] + seqlen[pcrColl[i].refIndex] - 2;   // faidx clips at the end of the record
      if (hi > last) hi = last;
      std::string seqstr;
      if (lo <= hi) { auto sx = extract(fm_index, lo, hi); seqstr.assign(sx.begin(), sx.end()); }
      j["Seq"] = seqstr;
      rcfile << j.dump();
    }
    rcfile << "]}";
  }
  rcfile << '}' << std::endl;
  primer3thal::destroy_thal_structures();
  return fatal ? 1 : 0;
}

// jsonfloat <file of 64-bit hex patterns>: nlohmann::json(double).dump() per line (the number
// format of the Tm / Penalty fields of the search JSON)
static int cmd_jsonfloat(Args const& a) {
  if (a.pos.size() < 1) return usage();
  std::vector<std::string> lines;
  if (!read_lines(a.pos[0], lines)) return 1;
  for (auto const& l : lines) {
    uint64_t u = std::strtoull(l.c_str(), nullptr, 16);
    double d;
    memcpy(&d, &u, 8);
    nlohmann::json j = d;
    std::cout << j.dump() << '\n';
  }
  return 0;
}

static int cmd_dump(Args const& a) {
  if (a.pos.size() < 1) return usage();
  TIndex fm;
  if (!load_index(a.pos[0], fm)) return 1;
  uint64_t n = fm.size();
  std::cout << "n\t" << n << "\nsigma\t" << (uint32_t)fm.sigma << "\nC";
  for (uint32_t i = 0; i <= fm.sigma; ++i) std::cout << '\t' << fm.C[i];
  std::cout << "\ncomp2char";
  for (uint32_t i = 0; i < fm.sigma; ++i) std::cout << '\t' << (uint32_t)(uint8_t)fm.comp2char[i];
  std::cout << "\nBWT";
  for (uint64_t i = 0; i < n; ++i) std::cout << '\t' << (uint32_t)(uint8_t)fm.bwt[i];
  std::cout << "\nSA";
  for (uint64_t i = 0; i < n; ++i) std::cout << '\t' << fm[i];
  std::cout << "\nTEXT";
  auto s = extract(fm, 0, n - 1);
  for (uint64_t i = 0; i < n; ++i) std::cout << '\t' << (uint32_t)(uint8_t)s[i];
  std::cout << "\n";
  return 0;
}

}  // namespace refdrv

int main(int argc, char** argv) {
  using namespace refdrv;
  if (argc < 2) return usage();
  std::string cmd = argv[1];
  Args a;
  parse(argc, argv, a);
  if (cmd == "index") return cmd_index(a);
  if (cmd == "hunt") return cmd_hunt(a);
  if (cmd == "seed") return cmd_seed(a);
  if (cmd == "neighbors") return cmd_neighbors(a);
  if (cmd == "needle") return cmd_needle(a);
  if (cmd == "count") return cmd_count(a, false);
  if (cmd == "locate") return cmd_count(a, true);
  if (cmd == "extract") return cmd_extract(a);
  if (cmd == "padcount") return cmd_padcount(a);
  if (cmd == "dump") return cmd_dump(a);
  if (cmd == "thal") return cmd_thal(a);
  if (cmd == "search") return cmd_search(a);
  if (cmd == "jsonfloat") return cmd_jsonfloat(a);
  return usage();
}
