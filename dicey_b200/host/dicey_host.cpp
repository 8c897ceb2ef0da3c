// dicey_host.cpp -- the C++ host side: dicey's command line and JSON output, with the per-query
// FM-index / neighbourhood / alignment loops replaced by one batched call into the CUDA library
// through the C ABI of include/dicey_b200.h.
//
//   dicey-b200 hunt  [OPTIONS] -g genome.fa.gz <sequence | queries.fasta>     reference src/hunter.h:180-444
//   dicey-b200 index [OPTIONS] genome.fa.gz                                    reference src/index.h:33-141
//
// `hunt` keeps the reference's flags (-g -o -m -x -d -n -f), messages, exit codes and JSON bytes
// (tests/test_host_cli.py compares them with the reference's golden output).  What differs:
// all queries of a FASTA input are searched in ONE dg_hunt_batch call instead of one by one, and
// queries outside the device path's limits (DESIGN.md "Limits") get an error record instead of
// a CPU search -- there is no CPU search path in this program.
// `search` and `padlock` need primer3's thal() Tm model on top of the FM / NW path (SURVEY.md 8f
// rank 1) and are not part of this round; the library entry points they would call
// (dg_hunt_batch with seed_len, dg_count_batch) exist and are tested.
#include <algorithm>
#include <cstdlib>
#include <ctime>
#include <iostream>

#include "../../include/dicey_b200.h"
#include "hostutil.hpp"

using namespace dhost;

static const char* kDiceyVersion = "0.5.1";  // reference src/version.h:8 (meta.version)

namespace {

struct HunterConfig {  // hunter.h:37-50
  bool indel = true, reverse = true, hasOutfile = false;
  uint32_t distance = 1, maxNeighborhood = 10000;
  uint64_t max_locations = 1000;
  std::string sequence, genome, outfile;
  int device = 0;
};

struct DnaHitView {  // one DnaHit (hunter.h:53-66) read out of a dg_result
  int32_t score;
  uint32_t chr, start;
  char strand;
  std::string refalign, queryalign;
};

uint32_t nucleotide_length(const std::string& s) {  // hunter.h:90-97
  uint32_t n = 0;
  for (char c : s) if (c != '-') ++n;
  return n;
}

// writeJsonDnaHitOut (hunter.h:99-160)
std::string hunt_json(const HunterConfig& c, uint32_t distance, const std::string& sequence, const std::string& qname,
                      const std::vector<std::string>& qn, const std::vector<DnaHitView>& ht,
                      const std::vector<std::string>& msg) {
  std::string o = "{\"errors\": [";
  bool errors = false;
  for (size_t i = 0; i < msg.size(); ++i) {
    std::string type = "warning";
    if (msg[i].compare(0, 5, "Error") == 0) { errors = true; type = "error"; }
    JsonObject e;
    e.set_string("type", type);
    e.set_string("title", msg[i]);
    if (i) o += ',';
    o += e.dump();
  }
  o += "]";
  if (!errors) {
    JsonObject meta;
    meta.set_string("version", kDiceyVersion);
    meta.set_string("subcommand", "hunt");
    meta.set_uint("distance", distance);
    meta.set_string("sequence", sequence);
    if (!qname.empty()) meta.set_string("name", qname);
    meta.set_string("genome", c.genome);
    meta.set_string("outfile", c.outfile);
    meta.set_uint("maxmatches", c.max_locations);
    meta.set_bool("hamming", !c.indel);
    meta.set_bool("forwardonly", !c.reverse);
    o += ",\"meta\":" + meta.dump() + ",\"data\":[";
    uint32_t oldchr = 999999, oldstart = 0;
    bool first = true;
    for (const auto& h : ht) {
      if (oldchr != h.chr || oldstart != h.start) {
        if (!first) o += ',';
        first = false;
        JsonObject j;
        j.set_int("distance", std::abs(h.score));
        j.set_string("chr", h.chr < qn.size() ? qn[h.chr] : std::string());
        j.set_uint("start", h.start);
        j.set_uint("end", h.start + nucleotide_length(h.refalign) - 1);
        j.set_string("strand", std::string(1, h.strand));
        j.set_string("refalign", h.refalign);
        j.set_string("queryalign", h.queryalign);
        o += j.dump();
      }
      oldchr = h.chr;
      oldstart = h.start;
    }
    o += ']';
  }
  o += "}\n";
  return o;
}

// jsonDnaHitOut (hunter.h:162-175)
void emit(const HunterConfig& c, const std::string& json) {
  if (c.hasOutfile) {
    if (!gz_append(c.outfile, json)) std::cerr << "Error: cannot write " << c.outfile << std::endl;
  } else {
    std::cout << json << std::flush;
  }
}

void hunt_usage(const char* argv0) {
  std::cout << "Usage: dicey " << argv0 << " [OPTIONS] -g Danio_rerio.fa.gz CATTACTAACATCAGT" << std::endl;
  std::cout << "       dicey " << argv0 << " [OPTIONS] -g Danio_rerio.fa.gz sequences.fasta" << std::endl;
  std::cout << "Generic options:\n"
               "  -? [ --help ]                         show help message\n"
               "  -g [ --genome ] arg                   genome file\n"
               "  -o [ --outfile ] arg                  gzipped output file\n"
               "  -m [ --maxmatches ] arg (=1000)       max. number of matches\n"
               "  -x [ --maxNeighborhood ] arg (=10000) max. neighborhood size\n"
               "  -d [ --distance ] arg (=1)            neighborhood distance\n"
               "  -n [ --hamming ]                      use hamming neighborhood instead of edit \n"
               "                                        distance\n"
               "  -f [ --forward ]                      only forward matches\n"
               "  --device arg (=0)                     CUDA device (dicey-b200 only)\n"
               "\n";
}

int hunter(int argc, char** argv) {
  HunterConfig c;
  Options opt({{"help", '?', false}, {"genome", 'g', true}, {"outfile", 'o', true}, {"maxmatches", 'm', true},
               {"maxNeighborhood", 'x', true}, {"distance", 'd', true}, {"hamming", 'n', false},
               {"forward", 'f', false}, {"input-file", 0, true}, {"device", 0, true}});
  try {
    opt.parse(argc, argv);
    c.max_locations = opt.get_u64("maxmatches", 1000);
    c.maxNeighborhood = (uint32_t)opt.get_u64("maxNeighborhood", 10000);
    c.distance = (uint32_t)opt.get_u64("distance", 1);
    c.device = (int)opt.get_u64("device", getenv("DICEY_B200_DEVICE") ? strtoull(getenv("DICEY_B200_DEVICE"), nullptr, 10) : 0);
  } catch (std::exception& e) {
    std::cerr << "dicey " << argv[0] << ": " << e.what() << std::endl;
    return 1;
  }
  if (opt.has("input-file")) opt.positional.insert(opt.positional.begin(), opt.get("input-file"));
  if (opt.has("help") || opt.positional.empty() || !opt.has("genome")) {
    hunt_usage(argv[0]);
    return -1;
  }
  c.sequence = opt.positional.back();  // program_options keeps the last value of a single-valued positional
  c.genome = opt.get("genome");
  c.outfile = opt.get("outfile");
  c.indel = !opt.has("hamming");
  c.reverse = !opt.has("forward");
  c.hasOutfile = opt.has("outfile");

  std::vector<DnaHitView> none;
  std::vector<std::string> msg, seqname;
  if (c.hasOutfile) {  // truncate (hunter.h:226-230)
    FILE* t = fopen(c.outfile.c_str(), "wb");
    if (t) fclose(t);
  }
  auto fail = [&](const std::string& m) {
    msg.push_back(m);
    emit(c, hunt_json(c, c.distance, c.sequence, "", seqname, none, msg));
    return 1;
  };
  if (!nonempty_regular_file(c.genome)) return fail("Error: Genome does not exist!");
  std::vector<uint64_t> lens;
  if (!read_fai(c.genome, seqname, lens)) {
    std::cerr << "Fail to open genome fai index for " << c.genome << std::endl;
    return fail("Error: Could not retrieve sequence lengths!");
  }
  std::vector<uint32_t> seqlen(lens.size());
  for (size_t i = 0; i < lens.size(); ++i) seqlen[i] = (uint32_t)(lens[i] + 1);  // util.h:201

  // FM-index: <parent>/<stem>.fm9 (hunter.h:248-256), transcoded to the device layout
  std::string index_file = path_join(path_parent(c.genome), path_stem(c.genome)) + ".fm9";
  dg_index* ix = nullptr;
  if (dg_index_open(index_file.c_str(), c.device, &ix) != DG_OK) {
    std::cerr << "dicey-b200: " << dg_last_error() << std::endl;
    return fail("Error: FM-Index cannot be loaded!");
  }
  dg_index_set_records(ix, seqlen.data(), (uint32_t)seqlen.size());

  // queries (hunter.h:262-288)
  std::vector<std::pair<std::string, std::string>> queries;
  if (is_regular_file(c.sequence)) {
    if (!is_fasta(c.sequence)) { dg_index_close(ix); return fail("Error: Input file is not in FASTA format!"); }
    std::ifstream fa(c.sequence.c_str());
    std::string line, fan, faseq;
    while (std::getline(fa, line)) {
      if (line.empty()) continue;
      if (line[0] == '>') {
        if (!fan.empty() && !faseq.empty()) queries.push_back({fan, faseq});
        faseq.clear();
        fan = line.substr(1);
      } else {
        faseq += line;
      }
    }
    if (!fan.empty() && !faseq.empty()) queries.push_back({fan, faseq});
  } else {
    queries.push_back({std::string(), c.sequence});
  }

  // one batched call for every query (hunter.h:289-433 per query)
  std::string cat;
  std::vector<uint64_t> off(1, 0);
  for (const auto& q : queries) { cat += q.second; off.push_back(cat.size()); }
  dg_params par;
  memset(&par, 0, sizeof(par));
  par.distance = c.distance;
  par.max_neighborhood = c.maxNeighborhood;
  par.max_locations = c.max_locations > 0xFFFFFFFFULL ? 0xFFFFFFFFu : (uint32_t)c.max_locations;
  par.indel = c.indel ? 1 : 0;
  par.reverse = c.reverse ? 1 : 0;
  par.seed_len = 0;
  // the device path enumerates distances 0..2; a larger -d is clamped per query to |seq| - 1 by the
  // reference, so only queries that keep d > 2 after clamping are out of reach
  dg_result* res = nullptr;
  bool all_unsupported = false;
  if (par.distance > 2) {
    uint64_t minlen = ~0ULL;
    for (size_t q = 0; q + 1 < off.size(); ++q) minlen = std::min<uint64_t>(minlen, off[q + 1] - off[q]);
    all_unsupported = true;
  }
  int rc = DG_OK;
  if (!all_unsupported) rc = dg_hunt_batch(ix, cat.data(), off.data(), (uint32_t)queries.size(), &par, &res);
  if (rc != DG_OK) {
    std::cerr << "dicey-b200: " << dg_last_error() << std::endl;
    dg_index_close(ix);
    return fail(std::string("Error: GPU search failed (") + dg_last_error() + ")!");
  }
  uint64_t nh = 0, pool_bytes = 0, seq_bytes = 0;
  uint32_t nq = 0;
  const dg_hit* hits = res ? dg_result_hits(res, &nh) : nullptr;
  const uint64_t* qoff = res ? dg_result_query_offsets(res, &nq) : nullptr;
  const uint32_t* status = res ? dg_result_query_status(res) : nullptr;
  const uint32_t* qdist = res ? dg_result_query_distance(res) : nullptr;
  const char* pool = res ? dg_result_pool(res, &pool_bytes) : nullptr;
  const char* norm = res ? dg_result_sequences(res, &seq_bytes) : nullptr;

  for (size_t qi = 0; qi < queries.size(); ++qi) {
    std::vector<std::string> m;
    std::vector<DnaHitView> ht;
    const std::string& raw = queries[qi].second;
    if (raw.size() < 10) {  // hunter.h:299-303
      m.push_back("Error: Input sequence is shorter than 10 nucleotides!");
      emit(c, hunt_json(c, c.distance, raw, queries[qi].first, seqname, ht, m));
      continue;
    }
    uint32_t st = all_unsupported ? (uint32_t)DG_Q_UNSUPPORTED : status[qi];
    if (st & DG_Q_UNSUPPORTED) {
      m.push_back("Error: Query is outside the limits of the GPU search path (length <= 255, distance <= 2)!");
      emit(c, hunt_json(c, c.distance, raw, queries[qi].first, seqname, ht, m));
      continue;
    }
    // replaceNonDna warnings (util.h:208-219), one per replaced character
    for (char ch : raw) {
      char u = (char)toupper((unsigned char)ch);
      if (u != 'A' && u != 'C' && u != 'G' && u != 'T')
        m.push_back("Warning: Non-DNA character in nucleotide sequence detected and replaced by 'N'!");
    }
    if (st & DG_Q_DIST_ADJUSTED) m.push_back("Warning: Distance was adjusted to sequence length!");
    if (st & DG_Q_NBR_CAP) {
      std::string x = std::to_string(c.maxNeighborhood);
      m.push_back("Warning: Neighborhood size exceeds " + x + " candidates. Only first " + x +
                  " neighbors are searched, results are likely incomplete!");
    }
    if (st & DG_Q_HIT_CAP) {
      std::string x = std::to_string(c.max_locations);
      m.push_back("Warning: More than " + x + " matches found. Only first " + x +
                  " matches are reported, results are likely incomplete!");
    }
    std::vector<dg_hit> mine(hits + qoff[qi], hits + qoff[qi + 1]);
    dg_hits_sort(mine.data(), mine.size());  // hunter.h:440
    ht.reserve(mine.size());
    for (const auto& h : mine) {
      DnaHitView v;
      v.score = h.score; v.chr = h.chr; v.start = h.start; v.strand = (char)h.strand;
      v.refalign.assign(pool + h.aln_off, h.aln_len);
      v.queryalign.assign(pool + h.aln_off + h.aln_len, h.aln_len);
      ht.push_back(std::move(v));
    }
    std::string sequence(norm + off[qi], norm + off[qi + 1]);
    emit(c, hunt_json(c, qdist[qi], sequence, queries[qi].first, seqname, ht, m));
  }
  if (res) dg_result_free(res);
  dg_index_close(ix);
  return 0;
}

// ------------------------------------------------------------------------------------------
std::string now_string() {  // boost::posix_time::to_simple_string(second_clock::local_time())
  time_t t = time(nullptr);
  struct tm tmv;
  localtime_r(&t, &tmv);
  static const char* mon[] = {"Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"};
  char buf[64];
  snprintf(buf, sizeof(buf), "%04d-%s-%02d %02d:%02d:%02d", tmv.tm_year + 1900, mon[tmv.tm_mon], tmv.tm_mday, tmv.tm_hour,
           tmv.tm_min, tmv.tm_sec);
  return buf;
}

bool fm9_loadable(const std::string& path) {
  // load_from_checked_file succeeds when <path>_check holds the csa_wt<> type hash and the file parses;
  // opening it on the device is the same test
  uint64_t sz = 0;
  if (!is_regular_file(path, &sz) || !is_regular_file(path + "_check")) return false;
  return true;
}

int index_cmd(int argc, char** argv) {
  Options opt({{"help", '?', false}, {"output", 'o', true}, {"input-file", 0, true}, {"device", 0, true}});
  int device = 0;
  try {
    opt.parse(argc, argv);
    device = (int)opt.get_u64("device", 0);
  } catch (std::exception& e) {
    std::cerr << "dicey " << argv[0] << ": " << e.what() << std::endl;
    return 1;
  }
  if (opt.has("input-file")) opt.positional.insert(opt.positional.begin(), opt.get("input-file"));
  if (opt.has("help") || opt.positional.empty()) {
    std::cout << "Usage: dicey " << argv[0] << " [OPTIONS] genome.fa.gz" << std::endl;
    std::cout << "Generic options:\n"
                 "  -? [ --help ]                     show help message\n"
                 "  -o [ --output ] arg (=genome.fm9) output file\n\n";
    return -1;
  }
  std::string genome = opt.positional.back();
  std::string outfile = opt.has("output") ? opt.get("output") : path_join(path_parent(genome), path_stem(genome) + ".fm9");
  std::cout << '[' << now_string() << "] dicey ";
  for (int i = 0; i < argc; ++i) std::cout << argv[i] << ' ';
  std::cout << std::endl;
  if (!is_regular_file(genome)) {
    std::cerr << "Error: " << genome << " cannot be opened!" << std::endl;
    return -1;
  }
  if (!is_gz(genome)) {
    std::cerr << "Error: Please compress " << genome << " with bgzip." << std::endl;
    return -1;
  }
  if (!fm9_loadable(outfile)) {
    std::cout << '[' << now_string() << "] Prepare FM-Index" << std::endl;
    // index.h:96-114: records upper-cased and joined by '\n', trailing '\n'
    std::string dump, line;
    bool firstSeq = true;
    GzLines in(genome);
    while (in.getline(line)) {
      if (!line.empty() && line[0] == '>') {
        if (!firstSeq) dump += '\n';
        else firstSeq = false;
      } else {
        for (char& ch : line) ch = (char)toupper((unsigned char)ch);
        dump += line;
      }
    }
    dump += '\n';
    std::cout << '[' << now_string() << "] Create FM-Index" << std::endl;
    dg_index* ix = nullptr;
    if (dg_index_build_text((const uint8_t*)dump.data(), dump.size(), device, &ix) != DG_OK ||
        dg_index_write_fm9(ix, outfile.c_str()) != DG_OK) {
      std::cerr << "Error: " << dg_last_error() << std::endl;
      if (ix) dg_index_close(ix);
      return -1;
    }
    dg_index_close(ix);
  }
  std::cout << '[' << now_string() << "] Done." << std::endl;
  return 0;
}

void display_usage() {
  std::cout << "Usage: dicey <command> <arguments>" << std::endl << std::endl;
  std::cout << "Commands:" << std::endl << std::endl;
  std::cout << "    index        index FASTA reference file" << std::endl;
  std::cout << "    hunt         search DNA sequences" << std::endl;
  std::cout << std::endl;
  std::cout << "(search and padlock are not part of dicey-b200 yet: DESIGN.md, scope)" << std::endl << std::endl;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { display_usage(); return 0; }
  std::string cmd = argv[1];
  if (cmd == "version" || cmd == "--version" || cmd == "--version-only" || cmd == "-v") {
    std::cout << "Dicey version: v" << kDiceyVersion << std::endl;
    std::cout << " using " << dg_version() << std::endl;
    return 0;
  }
  if (cmd == "help" || cmd == "--help" || cmd == "-h" || cmd == "-?") { display_usage(); return 0; }
  if (cmd == "index") return index_cmd(argc - 1, argv + 1);
  if (cmd == "hunt") return hunter(argc - 1, argv + 1);
  if (cmd == "search" || cmd == "padlock") {
    std::cerr << "dicey-b200: '" << cmd << "' is not available in this build (needs the thal Tm model; DESIGN.md, scope)" << std::endl;
    return 1;
  }
  std::cerr << "Unrecognized command " << cmd << std::endl;
  return 1;
}
