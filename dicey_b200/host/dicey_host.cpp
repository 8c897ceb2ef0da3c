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
//   dicey-b200 search [OPTIONS] -g genome.fa.gz -i primer3_config/ primers.fasta             reference src/silica.h:208-660
// `padlock` (probe design over GTF regions) is not part of this round; the library entry points it
// would call (dg_count_batch, dg_thal_batch) exist and are tested.
// DICEY_B200_TRACE=1 prints the wall time of every stage on stderr; DICEY_B200_THREADS bounds the
// JSON formatting threads of `hunt`.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <ctime>
#include <iostream>
#include <thread>

#include "../../include/dicey_b200.h"
#include "hostutil.hpp"
#include "jsonnum.hpp"
#include <set>
#include <cmath>

using namespace dhost;

static const char* kDiceyVersion = "0.5.1";  // reference src/version.h:8 (meta.version)

namespace {

struct HunterConfig {  // hunter.h:37-50
  bool indel = true, reverse = true, hasOutfile = false;
  uint32_t distance = 1, maxNeighborhood = 10000;
  uint64_t max_locations = 1000;
  std::string sequence, genome, outfile;
  int device = 0;
  std::vector<int> devices;   // --devices a,b,c: the queries are sharded over these GPUs (index replicated)
};

struct DnaHitView {  // one DnaHit (hunter.h:53-66) read out of a dg_result (the alignment rows stay in its pool)
  int32_t score;
  uint32_t chr, start;
  char strand;
  const char* refalign;
  const char* queryalign;
  uint32_t aln_len;
};

uint32_t nucleotide_length(const char* s, uint32_t len) {  // hunter.h:90-97
  uint32_t n = 0;
  for (uint32_t i = 0; i < len; ++i) if (s[i] != '-') ++n;
  return n;
}

// nlohmann's string escaping appended in place (hostutil.hpp json_escape without the temporary)
void append_escaped(std::string& o, const char* s, size_t len) {
  for (size_t i = 0; i < len; ++i) {
    const unsigned char ch = (unsigned char)s[i];
    if (ch >= 0x20 && ch != '"' && ch != '\\') { o += (char)ch; continue; }
    o += json_escape(std::string(1, (char)ch));
  }
}
void append_uint(std::string& o, uint64_t v) {
  char buf[24];
  int n = 0;
  do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) o += buf[--n];
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
        // nlohmann::json::dump() of the record: keys in std::map order
        o += "{\"chr\":\"";
        if (h.chr < qn.size()) append_escaped(o, qn[h.chr].data(), qn[h.chr].size());
        o += "\",\"distance\":";
        append_uint(o, (uint64_t)std::abs(h.score));
        o += ",\"end\":";
        append_uint(o, (uint32_t)(h.start + nucleotide_length(h.refalign, h.aln_len) - 1));
        o += ",\"queryalign\":\"";
        append_escaped(o, h.queryalign, h.aln_len);
        o += "\",\"refalign\":\"";
        append_escaped(o, h.refalign, h.aln_len);
        o += "\",\"start\":";
        append_uint(o, h.start);
        o += ",\"strand\":\"";
        append_escaped(o, &h.strand, 1);
        o += "\"}";
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
bool g_write_failed = false;   // a failed write of the output file ends the program with a non-zero exit code
void emit(const HunterConfig& c, const std::string& json) {
  if (c.hasOutfile) {
    if (!gz_append(c.outfile, json)) { std::cerr << "Error: cannot write " << c.outfile << std::endl; g_write_failed = true; }
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
               "  --devices arg                         comma-separated CUDA devices: the index is\n"
               "                                        replicated, the queries are sharded (dicey-b200 only)\n"
               "\n";
}

int hunter(int argc, char** argv) {
  HunterConfig c;
  Options opt({{"help", '?', false}, {"genome", 'g', true}, {"outfile", 'o', true}, {"maxmatches", 'm', true},
               {"maxNeighborhood", 'x', true}, {"distance", 'd', true}, {"hamming", 'n', false},
               {"forward", 'f', false}, {"input-file", 0, true}, {"device", 0, true}, {"devices", 0, true}});
  try {
    opt.parse(argc, argv);
    c.max_locations = opt.get_u64("maxmatches", 1000);
    c.maxNeighborhood = (uint32_t)opt.get_u64("maxNeighborhood", 10000);
    c.distance = (uint32_t)opt.get_u64("distance", 1);
    c.device = (int)opt.get_u64("device", getenv("DICEY_B200_DEVICE") ? strtoull(getenv("DICEY_B200_DEVICE"), nullptr, 10) : 0);
    if (opt.has("devices")) {
      std::string list = opt.get("devices"), tok;
      for (size_t i = 0; i <= list.size(); ++i) {
        if (i == list.size() || list[i] == ',') {
          if (!tok.empty()) c.devices.push_back(std::stoi(tok));
          tok.clear();
        } else {
          tok += list[i];
        }
      }
    }
    if (c.devices.empty()) c.devices.push_back(c.device);
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
  const bool trace = getenv("DICEY_B200_TRACE") != nullptr;   // wall time of every stage on stderr
  auto t_last = std::chrono::steady_clock::now();
  auto stage = [&](const char* what, uint64_t items) {
    if (!trace) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[hunt] %-28s %9.1f ms  (%llu)\n", what, std::chrono::duration<double, std::milli>(t - t_last).count(),
            (unsigned long long)items);
    t_last = t;
  };
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

  // FM-index: <parent>/<stem>.fm9 (hunter.h:248-256), transcoded to the device layout -- one replica per
  // GPU of --devices, loaded side by side
  std::string index_file = path_join(path_parent(c.genome), path_stem(c.genome)) + ".fm9";
  const int ndev = (int)c.devices.size();
  struct Shard {
    dg_index* ix = nullptr;
    dg_comm* comm = nullptr;
    dg_result* res = nullptr;
    uint32_t q0 = 0, q1 = 0;
    std::vector<uint64_t> off;     // shard-local query offsets
    int rc = DG_OK;
    std::string err;
    uint64_t gathered = 0;         // hits of all shards, as the all-gather reported them to this shard
    // result views
    const dg_rec* recs = nullptr;
    const dg_hit* hits = nullptr;      // only when the result is not compact
    const char* pool = nullptr;
    const uint64_t* qoff = nullptr;
    const uint32_t* status = nullptr;
    const uint32_t* qdist = nullptr;
    const char* norm = nullptr;
    uint64_t nh = 0;
  };
  std::vector<Shard> shards((size_t)ndev);
  auto close_all = [&]() {
    for (auto& s : shards) {
      if (s.res) dg_result_free(s.res);
      if (s.comm) dg_comm_destroy(s.comm);
      if (s.ix) dg_index_close(s.ix);
      s.res = nullptr; s.comm = nullptr; s.ix = nullptr;
    }
  };
  {
    std::vector<std::thread> th;
    for (int d = 0; d < ndev; ++d)
      th.emplace_back([&, d] {
        Shard& s = shards[(size_t)d];
        s.rc = dg_index_open(index_file.c_str(), c.devices[(size_t)d], &s.ix);
        if (s.rc != DG_OK) { s.err = dg_last_error(); return; }
        dg_index_set_records(s.ix, seqlen.data(), (uint32_t)seqlen.size());
      });
    for (auto& t : th) t.join();
  }
  for (auto& s : shards)
    if (s.rc != DG_OK) {
      std::cerr << "dicey-b200: " << s.err << std::endl;
      close_all();
      return fail("Error: FM-Index cannot be loaded!");
    }
  stage("index load", dg_index_size(shards[0].ix));

  // queries (hunter.h:262-288)
  std::vector<std::pair<std::string, std::string>> queries;
  if (is_regular_file(c.sequence)) {
    if (!is_fasta(c.sequence)) { close_all(); return fail("Error: Input file is not in FASTA format!"); }
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

  stage("queries read", queries.size());
  // one batched call per GPU for its contiguous shard of the queries (hunter.h:289-433 per query)
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
  // (distances 0..2 are enumerated on the device, distance 3 is searched from host-made neighbour lists;
  // a query whose clamped distance stays larger carries DG_Q_UNSUPPORTED)
  const bool all_unsupported = false;
  const uint32_t nq_all = (uint32_t)queries.size();
  uint8_t comm_id[DG_COMM_ID_BYTES];
  bool use_comm = ndev > 1 && !all_unsupported;
  for (int i = 0; i < ndev; ++i)
    for (int j = 0; j < i; ++j)
      if (c.devices[(size_t)i] == c.devices[(size_t)j]) use_comm = false;   // (NCCL wants one rank per GPU; shards on one GPU need no exchange)
  if (use_comm && dg_comm_get_unique_id(comm_id) != DG_OK) {
    std::cerr << "dicey-b200: " << dg_last_error() << " (continuing without the hit all-gather)" << std::endl;
    use_comm = false;
  }
  if (!all_unsupported) {
    std::vector<std::thread> th;
    for (int d = 0; d < ndev; ++d)
      th.emplace_back([&, d] {
        Shard& s = shards[(size_t)d];
        s.q0 = (uint32_t)((uint64_t)nq_all * (uint64_t)d / (uint64_t)ndev);
        s.q1 = (uint32_t)((uint64_t)nq_all * (uint64_t)(d + 1) / (uint64_t)ndev);
        s.off.resize((size_t)(s.q1 - s.q0) + 1);
        for (uint32_t q = s.q0; q <= s.q1; ++q) s.off[q - s.q0] = off[q] - off[s.q0];
        if (use_comm) {
          s.rc = dg_comm_init(ndev, d, comm_id, s.ix, &s.comm);
          if (s.rc != DG_OK) { s.err = dg_last_error(); return; }
        }
        s.rc = dg_hunt_batch(s.ix, cat.data() + off[s.q0], s.off.data(), s.q1 - s.q0, &par, &s.res);
        if (s.rc != DG_OK) { s.err = dg_last_error(); return; }
        uint32_t nq = 0;
        s.recs = dg_result_records(s.res, &s.nh);
        if (!s.recs) {
          uint64_t pool_bytes = 0;
          s.hits = dg_result_hits(s.res, &s.nh);
          s.pool = dg_result_pool(s.res, &pool_bytes);
        }
        s.qoff = dg_result_query_offsets(s.res, &nq);
        s.status = dg_result_query_status(s.res);
        s.qdist = dg_result_query_distance(s.res);
        uint64_t seq_bytes = 0;
        s.norm = dg_result_sequences(s.res, &seq_bytes);
        if (s.comm) {
          // the exchange step: every GPU ends up with the coordinates of every hit of the batch (HBM)
          const dg_wire* table = nullptr;
          uint64_t slot = 0;
          const uint64_t* counts = nullptr;
          s.rc = dg_allgather_hits(s.comm, nullptr, s.q0, &table, &slot, &counts);
          if (s.rc != DG_OK) { s.err = dg_last_error(); return; }
          for (int r = 0; r < ndev; ++r) s.gathered += counts[r];
          if (counts[d] != s.nh) { s.rc = DG_ERR_FORMAT; s.err = "hit all-gather: this rank's count differs from its result"; }
        }
      });
    for (auto& t : th) t.join();
    uint64_t total = 0;
    for (auto& s : shards) total += s.nh;
    for (auto& s : shards) {
      if (s.rc == DG_OK && s.comm && s.gathered != total) { s.rc = DG_ERR_FORMAT; s.err = "hit all-gather: gathered count differs from the sum of the shards"; }
      if (s.rc != DG_OK) {
        std::cerr << "dicey-b200: " << s.err << std::endl;
        std::string m = std::string("Error: GPU search failed (") + s.err + ")!";
        close_all();
        return fail(m);
      }
    }
    stage(ndev > 1 ? "search (dg_hunt_batch per GPU + dg_allgather_hits)" : "search (dg_hunt_batch)", total);
  }
  auto shard_of = [&](size_t qi) -> const Shard& {
    for (size_t d = 0; d + 1 < (size_t)ndev; ++d)
      if (qi < shards[d].q1) return shards[d];
    return shards[(size_t)ndev - 1];
  };

  // one JSON line per query (hunter.h:289-444), formatted by several threads over blocks of queries
  // and written in query order, one write (one gzip member) per block instead of per query: the
  // bytes a reader sees are the same
  auto format_query = [&](size_t qi, std::string& out) {
    std::vector<std::string> m;
    std::vector<DnaHitView> ht;
    const std::string& raw = queries[qi].second;
    if (raw.size() < 10) {  // hunter.h:299-303
      m.push_back("Error: Input sequence is shorter than 10 nucleotides!");
      out += hunt_json(c, c.distance, raw, queries[qi].first, seqname, ht, m);
      return;
    }
    const Shard& sh = shard_of(qi);
    const size_t ql = qi - sh.q0;
    uint32_t st = all_unsupported ? (uint32_t)DG_Q_UNSUPPORTED : sh.status[ql];
    if (st & DG_Q_UNSUPPORTED) {
      m.push_back("Error: Query is outside the limits of the GPU search path (length <= 255; beyond distance 2, sequence length + distance <= 42)!");
      out += hunt_json(c, c.distance, raw, queries[qi].first, seqname, ht, m);
      return;
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
    if (st & DG_Q_NBR_UNVERIFIED) {
      // not a message of the reference: its neighbourhood may have been truncated at -x, and the truncated
      // set of a query this long cannot be reproduced on the device; every neighbour was searched
      m.push_back("Warning: Neighborhood may exceed " + std::to_string(c.maxNeighborhood) +
                  " candidates; the reference's truncation is not reproduced for sequences longer than 40 nucleotides, all neighbors were searched!");
    }
    if (st & DG_Q_HIT_CAP) {
      std::string x = std::to_string(c.max_locations);
      m.push_back("Warning: More than " + x + " matches found. Only first " + x +
                  " matches are reported, results are likely incomplete!");
    }
    const std::string sequence(sh.norm + sh.off[ql], sh.norm + sh.off[ql + 1]);
    if (!sh.recs) {
      // full records (distances beyond three edit operations travel as dg_hit + alignment pool)
      std::vector<dg_hit> full(sh.hits + sh.qoff[ql], sh.hits + sh.qoff[ql + 1]);
      dg_hits_sort(full.data(), full.size());  // hunter.h:440
      ht.reserve(full.size());
      for (const auto& h : full) {
        DnaHitView v;
        v.score = h.score; v.chr = h.chr; v.start = h.start; v.strand = (char)h.strand;
        v.refalign = sh.pool + h.aln_off;
        v.queryalign = sh.pool + h.aln_off + h.aln_len;
        v.aln_len = h.aln_len;
        ht.push_back(v);
      }
      out += hunt_json(c, sh.qdist[ql], sequence, queries[qi].first, seqname, ht, m);
      return;
    }
    std::vector<dg_rec> mine(sh.recs + sh.qoff[ql], sh.recs + sh.qoff[ql + 1]);
    dg_recs_sort(mine.data(), mine.size());  // hunter.h:440
    // the records carry their alignments as edit operations: both rows are rebuilt from the query
    std::string rev(sequence.rbegin(), sequence.rend());
    for (char& ch : rev) ch = ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : ch == 'T' ? 'A' : 'N';
    const size_t stride = sequence.size() + 4;
    std::vector<char> rows(2 * stride * mine.size() + 1);
    ht.reserve(mine.size());
    for (size_t i = 0; i < mine.size(); ++i) {
      const dg_rec& h = mine[i];
      DnaHitView v;
      v.score = h.score; v.chr = h.chr; v.start = h.start; v.strand = (char)h.strand;
      char* ra = rows.data() + 2 * stride * i;
      char* qa = ra + stride;
      const std::string& sq = h.strand == '-' ? rev : sequence;
      const int cols = dg_rec_alignment(&h, sq.data(), (uint32_t)sq.size(), ra, qa);
      v.refalign = ra;
      v.queryalign = qa;
      v.aln_len = cols > 0 ? (uint32_t)cols : 0;
      ht.push_back(v);
    }
    out += hunt_json(c, sh.qdist[ql], sequence, queries[qi].first, seqname, ht, m);
  };
  {
    const size_t nq_fmt = queries.size(), block = 1u << 16;
    unsigned hw = std::thread::hardware_concurrency();
    if (const char* e = getenv("DICEY_B200_THREADS")) hw = (unsigned)std::max(1, atoi(e));
    const size_t nthreads = std::max<size_t>(1, std::min<size_t>(hw ? hw : 1, 32));
    for (size_t b0 = 0; b0 < nq_fmt; b0 += block) {
      const size_t b1 = std::min(nq_fmt, b0 + block), n = b1 - b0;
      const size_t nt = std::max<size_t>(1, std::min(nthreads, n / 64));   // a handful of queries: no threads
      std::vector<std::string> part(nt);
      auto work = [&](size_t t) {
        const size_t lo = b0 + n * t / nt, hi = b0 + n * (t + 1) / nt;
        for (size_t qi = lo; qi < hi; ++qi) format_query(qi, part[t]);
      };
      std::vector<std::thread> th;
      for (size_t t = 1; t < nt; ++t) th.emplace_back(work, t);
      work(0);
      for (auto& x : th) x.join();
      for (size_t t = 0; t < nt; ++t) emit(c, part[t]);   // (one gzip member per part: the bytes a reader sees are the same)
    }
  }
  uint64_t nh_all = 0;
  for (auto& s : shards) nh_all += s.nh;
  stage("sort + JSON + write", nh_all);
  close_all();
  return g_write_failed ? 1 : 0;
}


// ------------------------------------------------------------------------------------------
// `dicey search`: in-silico PCR (reference src/silica.h:208-660).  The FM-index / NW part of every
// primer runs as one dg_hunt_batch call in seed mode, the melting temperatures of every candidate
// site as one dg_thal_batch call (primer3 thal on the GPU, bit-identical); the Tm gate, the
// (refIndex, alignpos) de-duplication, amplicon pairing, penalties and the JSON are host code that
// follows silica.h line by line.
struct SilicaConfig {  // silica.h:38-67
  bool indel = true, pruneprimer = false, hasOutfile = false;
  uint32_t kmer = 15, distance = 1, maxNeighborhood = 10000, maxProdSize = 15000, maxPruneCount = 0;
  uint64_t max_locations = 10000;
  double cutTemp = 45.0, cutofPen = -1.0, penDiff = 0.6, penMis = 0.4, penLen = 0.001;
  double temp = 37.0, mv = 50.0, dv = 1.5, dna_conc = 50.0, dntp = 0.6;
  std::string outfile, infile, genome, primer3Config = "./src/primer3_config/";
  int device = 0;
};
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

std::string revcomp_upper(const std::string& in) {  // util.h:54-114 on an upper-cased sequence
  std::string s(in.rbegin(), in.rend());
  for (char& ch : s) {
    switch (ch) {
      case 'A': ch = 'T'; break; case 'C': ch = 'G'; break; case 'G': ch = 'C'; break; case 'T': ch = 'A'; break;
      case 'U': ch = 'A'; break; case 'R': ch = 'Y'; break; case 'Y': ch = 'R'; break; case 'K': ch = 'M'; break;
      case 'M': ch = 'K'; break; case 'B': ch = 'V'; break; case 'V': ch = 'B'; break; case 'D': ch = 'H'; break;
      case 'H': ch = 'D'; break; case 'S': case 'W': case 'N': break;
      default: ch = 'N';
    }
  }
  return s;
}

// writeJsonPrimerOut (silica.h:100-187)
std::string search_json(const SilicaConfig& c, uint32_t distance, const std::vector<std::string>& qn, const std::vector<PrimerBind>& allp,
                        const std::vector<PcrProduct>& pcrColl, const std::vector<std::string>& pName,
                        const std::vector<std::string>& pSeq, const std::vector<std::string>& msg,
                        const std::vector<std::string>& ampSeq) {
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
    meta.set_string("subcommand", "search");
    meta.set_uint("distance", distance);
    meta.set_string("genome", c.genome);
    meta.set_string("outfile", c.outfile);
    meta.set_uint("maxmatches", c.max_locations);
    meta.set_bool("hamming", !c.indel);
    o += ",\"meta\":" + meta.dump() + ",\"data\":{\"primers\":[";
    // nlohmann::json::dump() of every record, keys in std::map order, appended in place
    auto str = [&](const char* key, const std::string& v) {
      o += key;
      append_escaped(o, v.data(), v.size());
      o += '"';
    };
    auto num = [&](const char* key, uint64_t v) { o += key; append_uint(o, v); };
    auto dbl = [&](const char* key, double v) { o += key; o += json_double(v); };
    size_t bytes = 256;
    for (const auto& b : allp) bytes += 160 + b.genome.size() + pName[b.primerId].size() + pSeq[b.primerId].size();
    for (size_t i = 0; i < pcrColl.size(); ++i) bytes += 320 + ampSeq[i].size() + 2 * 64;
    o.reserve(o.size() + bytes);
    for (size_t i = 0; i < allp.size(); ++i) {
      if (i) o += ',';
      const PrimerBind& b = allp[i];
      str("{\"Chrom\":\"", qn[b.refIndex]);
      num(",\"End\":", (uint64_t)b.pos + pSeq[b.primerId].size());
      str(",\"Genome\":\"", b.genome);
      num(",\"Id\":", i);
      dbl(",\"MatchTm\":", b.perfTemp);
      str(",\"Name\":\"", pName[b.primerId]);
      str(",\"Ori\":\"", b.onFor ? "forward" : "reverse");
      num(",\"Pos\":", (uint32_t)(b.pos + 1));
      str(",\"Seq\":\"", pSeq[b.primerId]);
      dbl(",\"Tm\":", b.temp);
      o += '}';
    }
    o += "],\"amplicons\":[";
    for (size_t i = 0; i < pcrColl.size(); ++i) {
      if (i) o += ',';
      const PcrProduct& p = pcrColl[i];
      str("{\"Chrom\":\"", qn[p.refIndex]);
      num(",\"ForEnd\":", (uint64_t)p.forPos + pSeq[p.forId].size());
      str(",\"ForName\":\"", pName[p.forId]);
      num(",\"ForPos\":", (uint32_t)(p.forPos + 1));
      str(",\"ForSeq\":\"", pSeq[p.forId]);
      dbl(",\"ForTm\":", p.forTemp);
      num(",\"Id\":", i);
      num(",\"Length\":", p.leng);
      dbl(",\"Penalty\":", p.penalty);
      num(",\"RevEnd\":", (uint64_t)p.revPos + pSeq[p.revId].size());
      str(",\"RevName\":\"", pName[p.revId]);
      num(",\"RevPos\":", (uint32_t)(p.revPos + 1));
      str(",\"RevSeq\":\"", pSeq[p.revId]);
      dbl(",\"RevTm\":", p.revTemp);
      str(",\"Seq\":\"", ampSeq[i]);
      o += '}';
    }
    o += "]}";
  }
  o += "}\n";
  return o;
}

void search_usage(const char* argv0) {
  std::cout << "Usage: dicey " << argv0 << " [OPTIONS] -g <ref.fa.gz> sequences.fasta" << std::endl;
  std::cout << "Generic options:\n"
               "  -? [ --help ]                         show help message\n"
               "  -g [ --genome ] arg                   genome file\n"
               "  -i [ --config ] arg (=./src/primer3_config/)\n"
               "                                        primer3 config directory\n"
               "  -o [ --outfile ] arg                  output file\n"
               "\nApproximate Search Options:\n"
               "  -k [ --kmer ] arg (=15)               k-mer size\n"
               "  -m [ --maxmatches ] arg (=10000)      max. number of matches per k-mer\n"
               "  -x [ --maxNeighborhood ] arg (=10000) max. neighborhood size\n"
               "  -d [ --distance ] arg (=1)            neighborhood distance\n"
               "  -q [ --pruneprimer ] arg              prune primer threshold\n"
               "  -n [ --hamming ]                      use hamming neighborhood instead of edit \n"
               "                                        distance\n"
               "\nParameters for Scoring and Penalty Calculation:\n"
               "  -c [ --cutTemp ] arg (=45)            min. primer melting temperature\n"
               "  -l [ --maxProdSize ] arg (=15000)     max. PCR Product size\n"
               "  --cutoffPenalty arg (=-1)             max. penalty for products (-1 = keep all)\n"
               "  --penaltyTmDiff arg (=0.6)            multiplication factor for deviation of \n"
               "                                        primer Tm penalty\n"
               "  --penaltyTmMismatch arg (=0.4)        multiplication factor for Tm pair \n"
               "                                        difference penalty\n"
               "  --penaltyLength arg (=0.001)          multiplication factor for amplicon length \n"
               "                                        penalty\n"
               "\nParameters for Tm Calculation:\n"
               "  --enttemp arg (=37)                   temperature for entropie and entalpie \n"
               "                                        calculation in Celsius\n"
               "  --monovalent arg (=50)                concentration of monovalent ions in mMol\n"
               "  --divalent arg (=1.5)                 concentration of divalent ions in mMol\n"
               "  --dna arg (=50)                       concentration of annealing(!) Oligos in nMol\n"
               "  --dntp arg (=0.6)                     the sum  of all dNTPs in mMol\n"
               "\n";
}

int silica(int argc, char** argv) {
  SilicaConfig c;
  Options opt({{"help", '?', false}, {"genome", 'g', true}, {"config", 'i', true}, {"outfile", 'o', true}, {"kmer", 'k', true},
               {"maxmatches", 'm', true}, {"maxNeighborhood", 'x', true}, {"distance", 'd', true}, {"pruneprimer", 'q', true},
               {"hamming", 'n', false}, {"cutTemp", 'c', true}, {"maxProdSize", 'l', true}, {"cutoffPenalty", 0, true},
               {"penaltyTmDiff", 0, true}, {"penaltyTmMismatch", 0, true}, {"penaltyLength", 0, true}, {"enttemp", 0, true},
               {"monovalent", 0, true}, {"divalent", 0, true}, {"dna", 0, true}, {"dntp", 0, true}, {"input-file", 0, true},
               {"device", 0, true}});
  try {
    opt.parse(argc, argv);
    c.kmer = (uint32_t)opt.get_u64("kmer", 15);
    c.max_locations = opt.get_u64("maxmatches", 10000);
    c.maxNeighborhood = (uint32_t)opt.get_u64("maxNeighborhood", 10000);
    c.distance = (uint32_t)opt.get_u64("distance", 1);
    c.maxPruneCount = (uint32_t)opt.get_u64("pruneprimer", 0);
    c.maxProdSize = (uint32_t)opt.get_u64("maxProdSize", 15000);
    c.cutTemp = opt.get_double("cutTemp", 45.0);
    c.cutofPen = opt.get_double("cutoffPenalty", -1.0);
    c.penDiff = opt.get_double("penaltyTmDiff", 0.6);
    c.penMis = opt.get_double("penaltyTmMismatch", 0.4);
    c.penLen = opt.get_double("penaltyLength", 0.001);
    c.temp = opt.get_double("enttemp", 37.0);
    c.mv = opt.get_double("monovalent", 50.0);
    c.dv = opt.get_double("divalent", 1.5);
    c.dna_conc = opt.get_double("dna", 50.0);
    c.dntp = opt.get_double("dntp", 0.6);
    c.device = (int)opt.get_u64("device", getenv("DICEY_B200_DEVICE") ? strtoull(getenv("DICEY_B200_DEVICE"), nullptr, 10) : 0);
  } catch (std::exception& e) {
    std::cerr << "dicey " << argv[0] << ": " << e.what() << std::endl;
    return 1;
  }
  if (opt.has("input-file")) opt.positional.insert(opt.positional.begin(), opt.get("input-file"));
  if (opt.has("help") || opt.positional.empty() || !opt.has("genome")) {
    search_usage(argv[0]);
    return -1;
  }
  c.infile = opt.positional.back();
  c.genome = opt.get("genome");
  c.outfile = opt.get("outfile");
  if (opt.has("config")) c.primer3Config = opt.get("config");
  c.indel = !opt.has("hamming");
  c.hasOutfile = opt.has("outfile");
  c.pruneprimer = opt.has("pruneprimer");

  std::vector<PrimerBind> allp;
  std::vector<PcrProduct> pcrColl;
  std::vector<std::string> msg, seqname, pName, pSeq, ampSeq;
  uint32_t distance = c.distance;
  // DICEY_B200_TRACE=1: wall time of every stage on stderr
  const bool trace = getenv("DICEY_B200_TRACE") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto stage = [&](const char* what, uint64_t items) {
    if (!trace) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[search] %-28s %9.1f ms  (%llu)\n", what, std::chrono::duration<double, std::milli>(t - t_last).count(),
            (unsigned long long)items);
    t_last = t;
  };
  dg_index* ix = nullptr;
  dg_thal* th = nullptr;
  auto out = [&]() {   // jsonPrimerOut (silica.h:190-205): one gzip stream, truncating
    const std::string json = search_json(c, distance, seqname, allp, pcrColl, pName, pSeq, msg, ampSeq);
    if (c.hasOutfile) {
      FILE* t = fopen(c.outfile.c_str(), "wb");
      if (t) fclose(t);
      if (!gz_append(c.outfile, json)) std::cerr << "Error: cannot write " << c.outfile << std::endl;
    } else {
      std::cout << json << std::flush;
    }
  };
  auto fail = [&](const std::string& m) {
    msg.push_back(m);
    out();
    if (th) dg_thal_close(th);
    if (ix) dg_index_close(ix);
    return 1;
  };
  if (!nonempty_regular_file(c.genome)) return fail("Error: Genome does not exist!");
  // primer3 config directory (silica.h:300-315)
  {
    struct stat stc;
    if (stat(c.primer3Config.c_str(), &stc) != 0 || !S_ISDIR(stc.st_mode)) return fail("Error: Cannot find primer3 config directory!");
    while (c.primer3Config.size() > 1 && c.primer3Config.back() == '/') c.primer3Config.pop_back();
    c.primer3Config += '/';
    struct stat stf;
    if (stat((c.primer3Config + "tetraloop.dh").c_str(), &stf) != 0) return fail("Error: Config directory path appears to be incorrect!");
  }
  std::vector<uint64_t> lens;
  if (!read_fai(c.genome, seqname, lens)) {
    std::cerr << "Fail to open genome fai index for " << c.genome << std::endl;
    return fail("Error: Could not retrieve sequence lengths!");
  }
  const uint32_t nseq = (uint32_t)lens.size();
  std::vector<uint32_t> seqlen(nseq);
  std::vector<uint64_t> cum(nseq + 1, 0);
  for (uint32_t i = 0; i < nseq; ++i) { seqlen[i] = (uint32_t)(lens[i] + 1); cum[i + 1] = cum[i] + seqlen[i]; }
  std::string index_file = path_join(path_parent(c.genome), path_stem(c.genome)) + ".fm9";
  if (dg_index_open(index_file.c_str(), c.device, &ix) != DG_OK) {
    std::cerr << "dicey-b200: " << dg_last_error() << std::endl;
    ix = nullptr;
    return fail("Error: FM-Index cannot be loaded!");
  }
  dg_index_set_records(ix, seqlen.data(), nseq);
  stage("index load", dg_index_size(ix));
  if (dg_thal_open(c.primer3Config.c_str(), c.mv, c.dv, c.dntp, c.dna_conc, c.device, &th) != DG_OK) {
    std::cerr << "dicey-b200: " << dg_last_error() << std::endl;
    th = nullptr;
    return fail("Error: Config directory path appears to be incorrect!");
  }
  // input FASTA (silica.h:349-408)
  {
    struct stat sti;
    if (stat(c.infile.c_str(), &sti) != 0 || S_ISDIR(sti.st_mode)) return fail("Error: Input fasta file is missing!");
  }
  struct Rec { std::string name, seq; };
  std::vector<Rec> recs;   // records that reach the count / length tests, in file order
  {
    std::ifstream fafile(c.infile.c_str());
    std::string fan, tmpfasta, line;
    while (std::getline(fafile, line)) {
      if (line.empty()) continue;
      if (line[0] == '>') {
        if (!fan.empty() && !tmpfasta.empty() && tmpfasta.size() > c.kmer) {
          recs.push_back({fan, tmpfasta});
          tmpfasta.clear();   // (a sequence not longer than k stays and is prepended to the next record, as in the reference)
        }
        fan = line.substr(1);
      } else {
        for (char& ch : line) ch = (char)toupper((unsigned char)ch);
        tmpfasta += line;
      }
    }
    if (!fan.empty() && !tmpfasta.empty() && tmpfasta.size() > c.kmer) recs.push_back({fan, tmpfasta});
  }
  // prune counts of the 3' k-mer and its reverse complement (silica.h:364-367): exact, forward strand
  std::vector<uint64_t> cnt_f(recs.size(), 0), cnt_r(recs.size(), 0);
  if (c.pruneprimer && !recs.empty()) {
    std::string cat_f, cat_r;
    std::vector<uint64_t> off(1, 0);
    for (const auto& r : recs) {
      std::string qr = r.seq.substr(r.seq.size() - c.kmer);
      cat_f += qr;
      cat_r += revcomp_upper(qr);
      off.push_back(cat_f.size());
    }
    dg_params pc;
    memset(&pc, 0, sizeof(pc));
    pc.distance = 0; pc.max_neighborhood = c.maxNeighborhood; pc.max_locations = 1; pc.indel = 1; pc.reverse = 0;
    if (dg_count_batch(ix, cat_f.data(), off.data(), (uint32_t)recs.size(), &pc, cnt_f.data()) != DG_OK ||
        dg_count_batch(ix, cat_r.data(), off.data(), (uint32_t)recs.size(), &pc, cnt_r.data()) != DG_OK)
      return fail(std::string("Error: GPU search failed (") + dg_last_error() + ")!");
  }
  for (size_t i = 0; i < recs.size(); ++i) {
    if (c.pruneprimer && (cnt_f[i] > c.maxPruneCount || cnt_r[i] > c.maxPruneCount)) continue;
    std::string inseq;
    for (char ch : recs[i].seq) {   // replaceNonDna (util.h:208-219)
      if (ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T') inseq += ch;
      else { msg.push_back("Warning: Non-DNA character in nucleotide sequence detected and replaced by 'N'!"); inseq += 'N'; }
    }
    if (inseq.size() < 10 || inseq.size() < c.kmer)
      return fail("Error: Input sequence is shorter than 10 nucleotides or shorter than the selected k-mer length!");
    if (distance >= inseq.size()) {
      distance = (uint32_t)inseq.size() - 1;
      msg.push_back("Warning: Distance was adjusted to sequence length!");
    }
    pName.push_back(recs[i].name);
    pSeq.push_back(inseq);
  }
  const uint32_t np = (uint32_t)pSeq.size();
  stage("primers read / pruned", np);
  // perfect-match temperatures: thal(primer, reverse complement) (silica.h:430-443)
  std::vector<std::string> revQ(np);
  std::vector<double> matchTemp(np, 0.0);
  if (np) {
    std::string c1, c2;
    std::vector<uint64_t> o1(1, 0), o2(1, 0);
    for (uint32_t i = 0; i < np; ++i) {
      revQ[i] = revcomp_upper(pSeq[i]);
      c1 += pSeq[i]; o1.push_back(c1.size());
      c2 += revQ[i]; o2.push_back(c2.size());
    }
    std::vector<uint8_t> ok(np);
    if (dg_thal_batch(th, c1.data(), o1.data(), c2.data(), o2.data(), np, matchTemp.data(), ok.data()) != DG_OK)
      return fail(std::string("Error: GPU search failed (") + dg_last_error() + ")!");
    for (uint32_t i = 0; i < np; ++i)
      if (!ok[i] || matchTemp[i] == -999999.0) return fail("Error: Thermodynamical calculation failed!");
  }
  stage("perfect-match Tm", np);
  // seeds: every candidate site of every primer (silica.h:446-500, 521-532)
  dg_result* res = nullptr;
  if (np) {
    std::string cat;
    std::vector<uint64_t> off(1, 0);
    for (const auto& s2 : pSeq) { cat += s2; off.push_back(cat.size()); }
    dg_params par;
    memset(&par, 0, sizeof(par));
    par.distance = distance;
    par.max_neighborhood = c.maxNeighborhood;
    par.max_locations = c.max_locations > 0xFFFFFFFFULL ? 0xFFFFFFFFu : (uint32_t)c.max_locations;
    par.indel = c.indel ? 1 : 0;
    par.reverse = 1;
    par.seed_len = c.kmer;
    if (dg_hunt_batch(ix, cat.data(), off.data(), np, &par, &res) != DG_OK)
      return fail(std::string("Error: GPU search failed (") + dg_last_error() + ")!");
  }
  uint64_t nh = 0, pool_bytes = 0;
  uint32_t nq = 0;
  const dg_hit* hits = res ? dg_result_hits(res, &nh) : nullptr;
  const uint64_t* qoff = res ? dg_result_query_offsets(res, &nq) : nullptr;
  const uint32_t* status = res ? dg_result_query_status(res) : nullptr;
  const char* pool = res ? dg_result_pool(res, &pool_bytes) : nullptr;
  stage("seeds: FM search + NW", nh);
  // melting temperature of every candidate (silica.h:502-512)
  std::vector<double> tm(nh, 0.0);
  std::vector<uint8_t> tok(nh, 1);
  if (nh) {
    std::string c1, c2;
    std::vector<uint64_t> o1(1, 0), o2(1, 0);
    c1.reserve(nh * 24); c2.reserve(nh * 32);
    for (uint64_t h = 0; h < nh; ++h) {
      const std::string& primer = hits[h].strand == '-' ? pSeq[hits[h].query] : revQ[hits[h].query];
      c1 += primer; o1.push_back(c1.size());
      c2.append(pool + hits[h].aln_off, hits[h].aln_len); o2.push_back(c2.size());
    }
    if (nh > 0xFFFFFFFFULL || dg_thal_batch(th, c1.data(), o1.data(), c2.data(), o2.data(), (uint32_t)nh, tm.data(), tok.data()) != DG_OK) {
      if (res) dg_result_free(res);
      return fail(std::string("Error: GPU search failed (") + dg_last_error() + ")!");
    }
  }
  stage("candidate Tm (thal)", nh);
  std::vector<std::vector<PrimerBind>> forBind(nseq), revBind(nseq);
  for (uint32_t primerId = 0; primerId < np; ++primerId) {
    const uint32_t koffset = (uint32_t)pSeq[primerId].size() - c.kmer;
    if (status[primerId] & DG_Q_NBR_CAP) {
      std::string x = std::to_string(c.maxNeighborhood);
      msg.push_back("Warning: Neighborhood size exceeds " + x + " candidates. Only first " + x +
                    " neighbors are searched, results are likely incomplete!");
    }
    for (int fwrvidx = 0; fwrvidx < 2; ++fwrvidx) {
      std::set<std::pair<uint32_t, uint32_t>> uphit;
      for (uint64_t h = qoff[primerId]; h < qoff[primerId + 1]; ++h) {
        if ((hits[h].strand == '-') != (fwrvidx == 1)) continue;
        if (!tok[h] || tm[h] == -999999.0) {
          if (res) dg_result_free(res);
          return fail("Error: Thermodynamical calculation failed!");
        }
        if (!(tm[h] > c.cutTemp)) continue;
        uint32_t refIndex = hits[h].chr, chrpos = hits[h].start, alignpos = hits[h].alignpos;
        if (!uphit.insert(std::make_pair(refIndex, alignpos)).second) continue;
        const std::string& primer = fwrvidx ? pSeq[primerId] : revQ[primerId];
        std::string genomicseq(pool + hits[h].aln_off, hits[h].aln_len);
        if (fwrvidx) {
          uint32_t alignshift = alignpos - chrpos;
          chrpos = alignpos;
          genomicseq = alignshift <= genomicseq.size() ? genomicseq.substr(alignshift, primer.size()) : std::string();
        } else {
          uint32_t alignshift = alignpos - chrpos;
          chrpos = alignpos - koffset;
          if (alignshift >= koffset) {
            alignshift -= koffset;
            genomicseq = alignshift <= genomicseq.size() ? genomicseq.substr(alignshift, primer.size()) : std::string();
          }
        }
        PrimerBind prim;
        prim.refIndex = refIndex; prim.temp = tm[h]; prim.perfTemp = matchTemp[primerId]; prim.primerId = primerId;
        prim.genome = genomicseq; prim.onFor = !fwrvidx; prim.pos = chrpos;
        (fwrvidx ? revBind : forBind)[refIndex].push_back(prim);
      }
    }
    if (status[primerId] & DG_Q_HIT_CAP) {
      std::string x = std::to_string(c.max_locations);
      msg.push_back("Warning: More than " + x + " matches found. Only first " + x + " matches are reported, results are likely incomplete!");
    }
  }
  if (res) dg_result_free(res);
  for (uint32_t refIndex = 0; refIndex < nseq; ++refIndex) {   // silica.h:580-587
    allp.insert(allp.end(), forBind[refIndex].begin(), forBind[refIndex].end());
    allp.insert(allp.end(), revBind[refIndex].begin(), revBind[refIndex].end());
  }
  std::sort(allp.begin(), allp.end());
  stage("Tm gate, de-duplication", allp.size());
  if (!c.pruneprimer) {
    for (uint32_t refIndex = 0; refIndex < nseq; ++refIndex) {   // silica.h:591-634
      std::vector<std::pair<uint32_t, uint32_t>> rvByPos;
      rvByPos.reserve(revBind[refIndex].size());
      for (uint32_t k = 0; k < revBind[refIndex].size(); ++k) rvByPos.push_back(std::make_pair(revBind[refIndex][k].pos, k));
      std::sort(rvByPos.begin(), rvByPos.end());
      std::vector<uint32_t> rvPos(rvByPos.size());
      for (uint32_t k = 0; k < rvByPos.size(); ++k) rvPos[k] = rvByPos[k].first;
      std::vector<uint32_t> cand;
      for (auto fw = forBind[refIndex].begin(); fw != forBind[refIndex].end(); ++fw) {
        auto loIt = std::upper_bound(rvPos.begin(), rvPos.end(), fw->pos);
        std::vector<uint32_t>::iterator hiIt;
        uint64_t hiBound = (uint64_t)fw->pos + (uint64_t)c.maxProdSize;
        if (hiBound >= ((uint64_t)1 << 32)) hiIt = rvPos.end();
        else hiIt = std::upper_bound(rvPos.begin(), rvPos.end(), (uint32_t)hiBound);
        cand.clear();
        for (auto pit = loIt; pit != hiIt; ++pit) cand.push_back(rvByPos[pit - rvPos.begin()].second);
        std::sort(cand.begin(), cand.end());
        for (auto cit = cand.begin(); cit != cand.end(); ++cit) {
          auto rv = revBind[refIndex].begin() + (*cit);
          if ((rv->pos > fw->pos) && (rv->pos + pSeq[rv->primerId].size() - fw->pos <= c.maxProdSize)) {
            PcrProduct pp;
            pp.refIndex = refIndex; pp.forPos = fw->pos; pp.forTemp = fw->temp; pp.forId = fw->primerId;
            pp.revPos = rv->pos; pp.revTemp = rv->temp; pp.revId = rv->primerId;
            pp.leng = (uint32_t)((rv->pos + pSeq[pp.revId].size()) - fw->pos);
            double pen = (fw->perfTemp - fw->temp) * c.penDiff;
            if (pen < 0) pen = 0;
            double bpen = (rv->perfTemp - rv->temp) * c.penDiff;
            if (bpen > 0) pen += bpen;
            pen += std::abs(fw->temp - rv->temp) * c.penMis;
            pen += pp.leng * c.penLen;
            pp.penalty = pen;
            if ((c.cutofPen < 0) || (pen < c.cutofPen)) pcrColl.push_back(pp);
          }
        }
      }
    }
    std::sort(pcrColl.begin(), pcrColl.end());
    // amplicon sequences: faidx_fetch_seq(chrom, forPos, revPos + |rev| - 1), clipped to the record
    if (!pcrColl.empty()) {
      std::vector<uint64_t> pos(pcrColl.size()), len(pcrColl.size());
      uint64_t total = 0;
      for (size_t i = 0; i < pcrColl.size(); ++i) {
        const PcrProduct& p = pcrColl[i];
        uint64_t lo = p.forPos, hi = (uint64_t)p.revPos + pSeq[p.revId].size();   // [lo, hi) inside the record
        if (hi > lens[p.refIndex]) hi = lens[p.refIndex];
        if (lo > hi) lo = hi;
        pos[i] = cum[p.refIndex] + lo;
        len[i] = hi - lo;
        total += len[i];
      }
      std::string buf(total, '\0');
      if (dg_index_fetch_text(ix, pos.data(), len.data(), (uint32_t)pcrColl.size(), &buf[0]) != DG_OK)
        return fail(std::string("Error: GPU search failed (") + dg_last_error() + ")!");
      ampSeq.resize(pcrColl.size());
      uint64_t at = 0;
      for (size_t i = 0; i < pcrColl.size(); ++i) { ampSeq[i] = buf.substr(at, len[i]); at += len[i]; }
    }
  }
  stage("amplicons", pcrColl.size());
  out();
  stage("JSON", allp.size());
  dg_thal_close(th);
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
  return dg_fm9_check(path.c_str()) == DG_OK;   // a truncated, corrupt or foreign-type file is rebuilt (index.h:94-95)
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
  std::cout << "    search       in-silico PCR" << std::endl;
  std::cout << std::endl;
  std::cout << "(padlock is not part of dicey-b200 yet: DESIGN.md, scope)" << std::endl << std::endl;
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
  if (cmd == "search") return silica(argc - 1, argv + 1);
  if (cmd == "padlock") {
    std::cerr << "dicey-b200: 'padlock' is not part of this build yet (DESIGN.md, scope)" << std::endl;
    return 1;
  }
  std::cerr << "Unrecognized command " << cmd << std::endl;
  return 1;
}
