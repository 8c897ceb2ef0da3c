// oracle/dicey_oracle -- TEST INFRASTRUCTURE, not product code.
//
// A plain C++ restatement (no SDSL, no Boost, no CUDA) of the reference's algorithm for the
// FM-index primer-matching path, used as the portable checker beside oracle/_ref/dicey_ref (the
// reference's own code).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may run it.
//
// PARITY PIN: tests/test_oracle.py compares every output of this program with the golden files
// in tests/golden/, which were produced by oracle/_ref/dicey_ref, i.e. by the reference's own
// SDSL / neighbors.h / needle.h compiled from /root/reference (tests/golden/make_golden.py).
// The reference ships no tests or golden vectors of its own for this path.
//
// What each function follows (all paths relative to the reference tree):
//   Fm / load            src/xxsds/include/sdsl/csa_wt.hpp:381-388, wt_pc.hpp:627-636,
//                        rank_support_v.hpp:131-135, select_support_mcl.hpp:470-499,
//                        wt_helper.hpp:329-337, csa_alphabet_strategy.hpp:246-252,
//                        int_vector.hpp:1824-1838, io.hpp:917-936 (the _check sidecar)
//   rank1                rank_support_v.hpp:104-115
//   wt_rank              wt_pc.hpp:325-347
//   inverse_select       wt_pc.hpp:359-374
//   backward_search      suffix_array_algorithm.hpp:151-179 (one step), :207-226 (pattern)
//   sa / locate          csa_wt.hpp:340-354, suffix_array_algorithm.hpp:521-535
//   isa / extract        suffix_array_helper.hpp:416-430, suffix_array_algorithm.hpp:588-609
//   neighbors            src/neighbors.h:29-92
//   needle               src/needle.h:59-138, src/align.h:52-80,176-203
//   run_hunt             src/hunter.h:289-444 (+ :53-97)
//   run_seed             src/silica.h:449-573 (FM / NW part; the thal gate is left to the caller)
//   padcount             src/padlock.h:381-427
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <chrono>
#include <vector>

namespace orc {

// ------------------------------------------------------------------------------ .fm9
struct Node {
  uint64_t bv_pos, bv_pos_rank;
  uint16_t parent, child[2];
};
struct Fm {
  uint64_t n = 0, sigma_wt = 0, bv_bits = 0;
  std::vector<uint64_t> bv, bb;
  std::vector<Node> nodes;
  uint16_t c_to_leaf[256];
  uint64_t path[256];
  std::vector<uint64_t> sa_w, isa_w;
  uint8_t sa_width = 0, isa_width = 0;
  uint8_t char2comp[256];
  std::vector<uint8_t> comp2char;
  std::vector<uint64_t> C;
  uint16_t sigma = 0;
  uint64_t size() const { return n; }
};

static bool rd(std::istream& in, void* p, size_t n) { return n == 0 || (bool)in.read((char*)p, (std::streamsize)n); }
static bool rd_iv(std::istream& in, std::vector<uint64_t>* w, uint64_t& bits, uint8_t& width) {
  uint64_t h;
  if (!rd(in, &h, 8)) return false;
  bits = h & ((1ULL << 56) - 1);
  width = (uint8_t)(h >> 56);
  uint64_t nw = (bits + 63) >> 6;
  if (w) { w->resize(nw); return rd(in, w->data(), nw * 8); }
  in.seekg((std::streamoff)(nw * 8), std::ios::cur);
  return (bool)in;
}
static bool skip_select(std::istream& in) {
  uint64_t cnt, bits;
  uint8_t w;
  if (!rd(in, &cnt, 8)) return false;
  if (!cnt) return true;
  if (!rd_iv(in, nullptr, bits, w) || !rd_iv(in, nullptr, bits, w)) return false;
  for (uint64_t i = 0; i < ((cnt + 4095) >> 12); ++i)
    if (!rd_iv(in, nullptr, bits, w)) return false;
  return true;
}
static bool load(const std::string& path, Fm& f) {
  {  // load_from_checked_file: the sidecar holds std::hash of the demangled type name
    std::ifstream c((path + "_check").c_str(), std::ios::binary);
    uint64_t h = 0;
    if (!c || !rd(c, &h, 8)) return false;
    const char* name =
        "csa_wt<wt_pc<huff_shape, bit_vector, rank_support_v<1, 1>, select_support_mcl<1, 1>, "
        "select_support_mcl<0, 1>, byte_tree<false> >, 32u, 64u, sa_order_sa_sampling<0>, "
        "isa_sampling<0>, byte_alphabet>";
    if (h != std::hash<std::string>()(name)) return false;
  }
  std::ifstream in(path.c_str(), std::ios::binary);
  if (!in) return false;
  uint64_t bits;
  uint8_t w;
  if (!rd(in, &f.n, 8) || !rd(in, &f.sigma_wt, 8)) return false;
  if (!rd_iv(in, &f.bv, f.bv_bits, w)) return false;
  f.bv.push_back(0);
  if (!rd_iv(in, &f.bb, bits, w)) return false;
  if (!skip_select(in) || !skip_select(in)) return false;
  uint64_t nn;
  if (!rd(in, &nn, 8)) return false;
  f.nodes.resize(nn);
  for (auto& nd : f.nodes) {
    uint8_t rec[22];
    if (!rd(in, rec, 22)) return false;
    memcpy(&nd.bv_pos, rec, 8); memcpy(&nd.bv_pos_rank, rec + 8, 8); memcpy(&nd.parent, rec + 16, 2);
    memcpy(&nd.child[0], rec + 18, 2); memcpy(&nd.child[1], rec + 20, 2);
  }
  if (!rd(in, f.c_to_leaf, sizeof(f.c_to_leaf)) || !rd(in, f.path, sizeof(f.path))) return false;
  if (!rd_iv(in, &f.sa_w, bits, f.sa_width) || !rd_iv(in, &f.isa_w, bits, f.isa_width)) return false;
  f.sa_w.push_back(0); f.isa_w.push_back(0);
  std::vector<uint64_t> t;
  if (!rd_iv(in, &t, bits, w)) return false;
  memcpy(f.char2comp, t.data(), 256);
  if (!rd_iv(in, &t, bits, w)) return false;
  f.comp2char.assign((uint8_t*)t.data(), (uint8_t*)t.data() + bits / 8);
  if (!rd_iv(in, &f.C, bits, w)) return false;
  return rd(in, &f.sigma, 2);
}
static uint64_t get_int(const std::vector<uint64_t>& w, uint64_t i, uint8_t width) {
  uint64_t bit = i * width, k = bit >> 6, o = bit & 63;
  uint64_t v = w[k] >> o;
  if (o + width > 64) v |= w[k + 1] << (64 - o);
  return width == 64 ? v : (v & ((1ULL << width) - 1));
}

// ------------------------------------------------------------------------------ FM queries
struct Counters { uint64_t ranks = 0, lf = 0; };
static Counters g_cnt;

static uint64_t rank1(const Fm& f, uint64_t idx) {  // rank_support_v.hpp:104-115
  const uint64_t* p = f.bb.data() + ((idx >> 8) & 0xFFFFFFFFFFFFFFFEULL);
  uint64_t r = p[0] + ((p[1] >> (63 - 9 * ((idx & 0x1FF) >> 6))) & 0x1FF);
  if (idx & 0x3F) r += (uint64_t)__builtin_popcountll(f.bv[idx >> 6] & ((1ULL << (idx & 0x3F)) - 1));
  return r;
}
static uint64_t wt_rank(const Fm& f, uint64_t i, uint8_t c) {  // wt_pc.hpp:325-347
  if (f.c_to_leaf[c] == 0xFFFF) return 0;
  if (f.sigma_wt == 1) return i;
  uint64_t p = f.path[c];
  uint32_t len = (uint32_t)(p >> 56);
  uint64_t result = i;
  uint32_t v = 0;
  for (uint32_t l = 0; l < len && result; ++l, p >>= 1) {
    uint64_t k = rank1(f, f.nodes[v].bv_pos + result) - f.nodes[v].bv_pos_rank;
    if (p & 1) result = k; else result -= k;
    v = f.nodes[v].child[p & 1];
  }
  ++g_cnt.ranks;
  return result;
}
static std::pair<uint64_t, uint8_t> inverse_select(const Fm& f, uint64_t i) {  // wt_pc.hpp:359-374
  uint32_t v = 0;
  while (f.nodes[v].child[0] != 0xFFFF) {
    uint64_t at = f.nodes[v].bv_pos + i;
    uint64_t k = rank1(f, at) - f.nodes[v].bv_pos_rank;
    if ((f.bv[at >> 6] >> (at & 63)) & 1) { i = k; v = f.nodes[v].child[1]; }
    else { i -= k; v = f.nodes[v].child[0]; }
  }
  return std::make_pair(i, (uint8_t)f.nodes[v].bv_pos_rank);
}
static uint64_t lf(const Fm& f, uint64_t i) {  // suffix_array_helper.hpp:280-292
  auto rc = inverse_select(f, i);
  ++g_cnt.lf;
  return f.C[f.char2comp[rc.second]] + rc.first;
}
// suffix_array_algorithm.hpp:151-179; closed interval [l, r]
static uint64_t backward_step(const Fm& f, uint64_t l, uint64_t r, uint8_t c, uint64_t& lo, uint64_t& ro) {
  uint8_t cc = f.char2comp[c];
  if (cc == 0 && c > 0) { lo = 1; ro = 0; return 0; }
  uint64_t cb = f.C[cc];
  if (l == 0 && r + 1 == f.size()) { lo = cb; ro = f.C[cc + 1] - 1; }
  else { lo = cb + wt_rank(f, l, c); ro = cb + wt_rank(f, r + 1, c) - 1; }
  return ro + 1 - lo;
}
// suffix_array_algorithm.hpp:207-226; *steps = iterations executed
static uint64_t backward_search(const Fm& f, const std::string& s, uint64_t& l, uint64_t& r, uint32_t* steps = nullptr) {
  l = 0; r = f.size() - 1;
  uint32_t e = 0;
  size_t i = s.size();
  while (i > 0 && r + 1 - l > 0) {
    --i;
    backward_step(f, l, r, (uint8_t)s[i], l, r);
    ++e;
  }
  if (steps) *steps = e;
  return r + 1 - l;
}
static uint64_t count(const Fm& f, const std::string& s) {  // :447-454
  if (s.size() > f.size()) return 0;
  uint64_t l, r;
  return backward_search(f, s, l, r);
}
static uint64_t sa_at(const Fm& f, uint64_t i) {  // csa_wt.hpp:340-354
  uint64_t off = 0;
  while (i % 32 != 0) { i = lf(f, i); ++off; }
  uint64_t v = get_int(f.sa_w, i / 32, f.sa_width);
  return v + off < f.size() ? v + off : v + off - f.size();
}
static std::vector<uint64_t> locate(const Fm& f, const std::string& s) {  // :521-535
  uint64_t l, r;
  uint64_t occ = backward_search(f, s, l, r);
  std::vector<uint64_t> out(occ);
  for (uint64_t i = 0; i < occ; ++i) out[i] = sa_at(f, l + i);
  return out;
}
static uint64_t isa_at(const Fm& f, uint64_t i) {  // suffix_array_helper.hpp:416-430
  // sample_qeq (csa_sampling_strategy.hpp:702-706): the next ISA sample after floor(i / 64), cyclically
  uint64_t nsamp = (f.size() - 1) / 64 + 1;
  uint64_t ci = (i / 64 + 1) % nsamp;
  uint64_t row = get_int(f.isa_w, ci, f.isa_width), pos = ci * 64;
  uint64_t steps = pos < i ? pos + f.size() - i : pos - i;
  while (steps--) row = lf(f, row);
  return row;
}
static uint8_t first_row_symbol(const Fm& f, uint64_t row) {  // suffix_array_helper.hpp:27-50
  uint32_t c = 0;
  while (c + 1 < f.sigma && f.C[c + 1] <= row) ++c;
  return f.comp2char[c];
}
static std::string extract(const Fm& f, uint64_t lo, uint64_t hi) {  // suffix_array_algorithm.hpp:588-609
  uint64_t steps = hi - lo + 1;
  std::string out(steps, '\0');
  uint64_t order = isa_at(f, hi);
  out[--steps] = (char)first_row_symbol(f, order);
  while (steps != 0) {
    auto rc = inverse_select(f, order);
    order = f.C[f.char2comp[rc.second]] + rc.first;
    out[--steps] = (char)rc.second;
    ++g_cnt.lf;
  }
  return out;
}

// ------------------------------------------------------------------------------ neighbors.h
static void nb_insert(std::set<std::string>& st, const std::string& s, bool indel) {  // :29-45
  if (!indel) { st.insert(s); return; }
  bool ins = true;
  for (auto it = st.begin(); it != st.end();) {
    if (it->find(s) != std::string::npos) st.erase(it++);
    else {
      if (s.find(*it) != std::string::npos) ins = false;
      ++it;
    }
  }
  if (ins) st.insert(s);
}
static void nb_rec(std::string& q, int inputdist, int dist, bool indel, int pos, uint32_t maxsize,
                   std::set<std::string>& st) {  // :47-83
  static const char alphabet[4] = {'A', 'C', 'G', 'T'};
  if (st.size() >= maxsize) return;
  if (pos < (int)q.size()) {
    if (dist > 0 && indel) {
      std::string nx = q.substr(0, pos) + q.substr(pos + 1);
      nb_rec(nx, inputdist, dist - 1, indel, pos, maxsize, st);
    }
    nb_rec(q, inputdist, dist, indel, pos + 1, maxsize, st);
    if (dist > 0) {
      char orig = q[pos];
      for (char a : alphabet)
        if (a != orig) {
          q[pos] = a;
          nb_rec(q, inputdist, dist - 1, indel, pos + 1, maxsize, st);
        }
      q[pos] = orig;
      if (indel)
        for (char a : alphabet) {
          std::string nx = q.substr(0, pos) + std::string(1, a) + q.substr(pos);
          nb_rec(nx, inputdist, dist - 1, indel, pos + 1, maxsize, st);
        }
    }
  } else if (dist < inputdist) {
    nb_insert(st, q, indel);
  }
}
static void neighbors(const std::string& query, int dist, bool indel, uint32_t maxsize, std::set<std::string>& st) {  // :86-92
  std::string q(query);
  nb_insert(st, q, indel);
  nb_rec(q, dist, dist, indel, 0, maxsize, st);
}

// ------------------------------------------------------------------------------ needle.h
// AlignConfig<false,true>, DnaScore(0,-1,-1,-1); returns the score, fills the two alignment rows
static int needle(const std::string& a1, const std::string& a2, std::string& r0, std::string& r1) {
  size_t m = a1.size(), n = a2.size(), mf = n + 1;
  std::vector<int> s(n + 1, 0);
  std::vector<bool> bit3((m + 1) * (n + 1), false), bit4((m + 1) * (n + 1), false);
  int prevsub = 0;
  auto vgap = [&](size_t col) { return (col == 0 || col == n) ? 0 : -1; };  // align.h:59-65
  for (size_t row = 0; row <= m; ++row)
    for (size_t col = 0; col <= n; ++col) {
      if (row == 0 && col == 0) { s[0] = 0; prevsub = 0; }
      else if (row == 0) { s[col] = -(int)col; bit3[col] = true; }
      else if (col == 0) { s[0] = 0; prevsub = 0; bit4[row * mf] = true; }
      else {
        int pp = prevsub;
        prevsub = s[col];
        int diag = pp + (a1[row - 1] == a2[col - 1] ? 0 : -1);
        int up = prevsub + vgap(col);
        int left = s[col - 1] - 1;
        s[col] = std::max(std::max(diag, up), left);
        if (s[col] == left) bit3[row * mf + col] = true;
        else if (s[col] == up) bit4[row * mf + col] = true;
      }
    }
  std::string trace;
  size_t row = m, col = n;
  while (row > 0 || col > 0) {
    if (bit3[row * mf + col]) { --col; trace.push_back('h'); }
    else if (bit4[row * mf + col]) { --row; trace.push_back('v'); }
    else { --row; --col; trace.push_back('s'); }
  }
  r0.clear(); r1.clear();
  size_t i = 0, j = 0;
  for (auto it = trace.rbegin(); it != trace.rend(); ++it) {  // align.h:176-203
    if (*it == 's') { r0 += a1[i++]; r1 += a2[j++]; }
    else if (*it == 'h') { r0 += '-'; r1 += a2[j++]; }
    else { r0 += a1[i++]; r1 += '-'; }
  }
  return s[n];
}
static uint32_t trail_gap(const std::string& r1) {  // hunter.h:69-77
  uint32_t last = (uint32_t)r1.size() - 1;
  for (uint32_t j = 0; j < r1.size(); ++j) if (r1[j] != '-') last = j;
  return (uint32_t)r1.size() - last - 1;
}

// ------------------------------------------------------------------------------ util.h
static char complement(char n) {  // util.h:54-93
  switch (n) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'U': return 'A';
    case 'R': return 'Y'; case 'Y': return 'R'; case 'S': return 'S'; case 'W': return 'W'; case 'K': return 'M';
    case 'M': return 'K'; case 'B': return 'V'; case 'V': return 'B'; case 'D': return 'H'; case 'H': return 'D';
  }
  return 'N';
}
static std::string revcomp(std::string s) {
  for (auto& c : s) c = complement((char)std::toupper((unsigned char)c));
  std::reverse(s.begin(), s.end());
  return s;
}

// ------------------------------------------------------------------------------ hunter.h
struct Hit {
  int32_t score; uint32_t chr, start; char strand; std::string ra, qa;
  bool operator<(const Hit& b) const {  // hunter.h:63-65
    return (score > b.score) || (score == b.score && chr < b.chr) || (score == b.score && chr == b.chr && start < b.start);
  }
};
struct Params { bool indel = true, reverse = true; uint32_t distance = 1, maxnbr = 10000; size_t maxloc = 1000; };
struct QRes { std::string seq; uint32_t distance = 0; std::vector<std::string> msg; std::vector<Hit> push, sorted; };

static void locate_chr(const std::vector<uint32_t>& seqlen, int64_t pos, uint32_t& ref, uint32_t& chrpos) {  // :358-362
  int64_t cum = 0;
  ref = 0;
  for (; (ref + 1 < seqlen.size()) && (pos >= cum + seqlen[ref]); ++ref) cum += seqlen[ref];
  chrpos = (uint32_t)(pos - cum);
}

static void run_hunt(const Fm& fm, const std::vector<uint32_t>& seqlen, const std::string& qseq, const Params& p, QRes& out) {
  out.seq = qseq;
  out.distance = p.distance;
  if (out.seq.size() < 10) { out.msg.push_back("Error: Input sequence is shorter than 10 nucleotides!"); return; }
  for (auto& c : out.seq) c = (char)std::toupper((unsigned char)c);
  {
    std::string t;
    for (char c : out.seq) {
      if (c == 'A' || c == 'C' || c == 'G' || c == 'T') t += c;
      else { out.msg.push_back("Warning: Non-DNA character in nucleotide sequence detected and replaced by 'N'!"); t += 'N'; }
    }
    out.seq = t;
  }
  std::string rev = revcomp(out.seq);
  if (out.distance >= out.seq.size()) { out.distance = (uint32_t)out.seq.size() - 1; out.msg.push_back("Warning: Distance was adjusted to sequence length!"); }
  size_t ctx = p.indel ? out.distance : 0;
  std::vector<std::set<std::string>> fwrv(2);
  neighbors(out.seq, (int)out.distance, p.indel, p.maxnbr, fwrv[0]);
  if (p.reverse) neighbors(rev, (int)out.distance, p.indel, p.maxnbr, fwrv[1]);
  if (fwrv[0].size() >= p.maxnbr || fwrv[1].size() >= p.maxnbr)
    out.msg.push_back("Warning: Neighborhood size exceeds " + std::to_string(p.maxnbr) + " candidates. Only first " +
                      std::to_string(p.maxnbr) + " neighbors are searched, results are likely incomplete!");
  uint32_t hits = 0;
  for (uint32_t fr = 0; fr < 2; ++fr)
    for (auto it = fwrv[fr].begin(); it != fwrv[fr].end() && hits < p.maxloc; ++it) {
      const std::string& query = *it;
      size_t m = query.size();
      size_t occs = count(fm, query);
      if (!occs) continue;
      auto loc = locate(fm, query);
      std::sort(loc.begin(), loc.end());
      for (size_t i = 0; i < std::min(occs, p.maxloc) && hits < p.maxloc; ++i) {
        uint32_t ref, chrpos;
        locate_chr(seqlen, (int64_t)loc[i], ref, chrpos);
        size_t pre_x = ctx, post_x = ctx;
        if (pre_x > loc[i]) pre_x = loc[i];
        if (loc[i] + m + post_x > fm.size()) post_x = fm.size() - loc[i] - m;
        std::string s = extract(fm, loc[i] - pre_x, loc[i] + m + post_x - 1);
        std::string pre = s.substr(0, pre_x);
        s = s.substr(pre_x);
        if (pre.find_last_of('\n') != std::string::npos) pre = pre.substr(pre.find_last_of('\n') + 1);
        std::string post = s.substr(m);
        post = post.substr(0, post.find_first_of('\n'));
        std::string g = pre + s.substr(0, m) + post;
        if (pre.size() < chrpos) chrpos -= (uint32_t)pre.size();
        const std::string& qq = fr ? rev : out.seq;
        char strand = fr ? '-' : '+';
        if (p.indel) {
          std::string r0, r1, ra, qa;
          int score = needle(g, qq, r0, r1);
          bool lead = true;
          for (uint32_t j = 0; j < r1.size() - trail_gap(r1); ++j) {  // hunter.h:391-401
            if (r1[j] != '-') lead = false;
            if (!lead) { ra += r0[j]; qa += r1[j]; } else ++chrpos;
          }
          out.push.push_back(Hit{score, ref, chrpos + 1, strand, ra, qa});
        } else {
          int score = 0;
          for (size_t k = 0; k < g.size() && k < qq.size(); ++k) if (g[k] != qq[k]) --score;  // hunter.h:79-88
          out.push.push_back(Hit{score, ref, chrpos + 1, strand, g, qq});
        }
        ++hits;
      }
    }
  if (hits >= p.maxloc)
    out.msg.push_back("Warning: More than " + std::to_string(p.maxloc) + " matches found. Only first " +
                      std::to_string(p.maxloc) + " matches are reported, results are likely incomplete!");
  out.sorted = out.push;
  std::sort(out.sorted.begin(), out.sorted.end());  // hunter.h:440
}

// ------------------------------------------------------------------------------ I/O
static bool read_lines(const std::string& path, std::vector<std::string>& lines) {
  std::ifstream f(path.c_str());
  if (!f) return false;
  std::string l;
  while (std::getline(f, l)) { if (!l.empty() && l.back() == '\r') l.pop_back(); if (!l.empty()) lines.push_back(l); }
  return true;
}
static void split_query(const std::string& line, std::string& name, std::string& seq) {
  size_t t = line.find('\t');
  if (t == std::string::npos) { name.clear(); seq = line; } else { name = line.substr(0, t); seq = line.substr(t + 1); }
}
struct Args {
  std::vector<std::string> pos;
  uint32_t d = 1, x = 10000, k = 15;
  size_t m = 1000; bool mset = false, hamming = false, forward = false;
  std::string records;
  unsigned threads = 1;
};

}  // namespace orc

int main(int argc, char** argv) {
  using namespace orc;
  if (argc < 2) { std::cerr << "dicey_oracle hunt|seed|neighbors|needle|count|locate|extract|padcount ... (same arguments as oracle/_ref/dicey_ref)\n"; return 2; }
  std::string cmd = argv[1];
  Args a;
  for (int i = 2; i < argc; ++i) {
    std::string s = argv[i];
    auto val = [&]() -> const char* { if (i + 1 >= argc) exit(2); return argv[++i]; };
    if (s == "-d") a.d = (uint32_t)atoi(val());
    else if (s == "-x") a.x = (uint32_t)atoi(val());
    else if (s == "-k") a.k = (uint32_t)atoi(val());
    else if (s == "-m") { a.m = (size_t)atoll(val()); a.mset = true; }
    else if (s == "-n") a.hamming = true;
    else if (s == "-f") a.forward = true;
    else if (s == "--records") a.records = val();
    else if (s == "--threads") a.threads = (unsigned)atoi(val());
    else if (s == "--counters") { /* accepted for command-line compatibility with dicey_ref; the port counts no work */ }
    else a.pos.push_back(s);
  }
  if (cmd == "neighbors") {
    std::vector<std::string> qs;
    if (!read_lines(a.pos[0], qs)) qs.push_back(a.pos[0]);
    for (auto& q : qs) {
      std::set<std::string> st;
      neighbors(q, (int)a.d, !a.hamming, a.x, st);
      std::cout << "Q\t" << q << '\t' << st.size() << '\n';
      for (auto& s : st) std::cout << s << '\n';
    }
    return 0;
  }
  if (cmd == "needle") {
    std::vector<std::string> lines;
    read_lines(a.pos[0], lines);
    for (auto& l : lines) {
      size_t t = l.find('\t');
      std::string r0, r1;
      int sc = needle(l.substr(0, t), l.substr(t + 1), r0, r1);
      std::cout << sc << '\t' << r0 << '\t' << r1 << '\n';
    }
    return 0;
  }
  Fm fm;
  if (a.pos.empty() || !load(a.pos[0], fm)) { std::cerr << "Error: FM-Index cannot be loaded!\n"; return 1; }
  if (cmd == "count" || cmd == "locate") {
    std::vector<std::string> pats;
    read_lines(a.pos[1], pats);
    for (auto& s : pats) {
      uint64_t l, r;
      uint64_t occ = backward_search(fm, s, l, r);
      if (cmd == "count") std::cout << l << '\t' << r << '\t' << occ << '\n';
      else {
        auto loc = locate(fm, s);
        std::sort(loc.begin(), loc.end());
        std::cout << occ;
        for (auto v : loc) std::cout << '\t' << v;
        std::cout << '\n';
      }
    }
    return 0;
  }
  if (cmd == "extract") {
    std::string s = extract(fm, strtoull(a.pos[1].c_str(), 0, 10), strtoull(a.pos[2].c_str(), 0, 10));
    std::cout.write(s.data(), (std::streamsize)s.size());
    return 0;
  }
  if (cmd == "padcount") {  // padlock.h:381-427
    std::vector<std::string> arms;
    read_lines(a.pos[1], arms);
    for (auto& arm : arms) {
      std::string rarm = revcomp(arm);
      uint64_t exact = count(fm, arm) + count(fm, rarm);
      std::set<std::string> fw, rv;
      neighbors(arm, (int)a.d, !a.hamming, 10000, fw);
      neighbors(rarm, (int)a.d, !a.hamming, 10000, rv);
      uint64_t total = 0;
      for (auto& s : fw) total += count(fm, s);
      for (auto& s : rv) total += count(fm, s);
      std::cout << arm << '\t' << exact << '\t' << total << '\n';
    }
    return 0;
  }
  // hunt / seed need the record table: "name<TAB>length" lines, seqlen = length + 1 (util.h:201)
  std::vector<std::string> rl;
  std::vector<uint32_t> seqlen;
  if (a.pos.size() < 3 || !read_lines(a.pos[1], rl)) return 2;
  for (auto& l : rl) { std::istringstream is(l); std::string nm; uint64_t len; is >> nm >> len; seqlen.push_back((uint32_t)len + 1); }
  std::vector<std::string> ql;
  read_lines(a.pos[2], ql);
  if (cmd == "hunt") {
    Params p;
    p.indel = !a.hamming; p.reverse = !a.forward; p.distance = a.d; p.maxnbr = a.x; p.maxloc = a.mset ? a.m : 1000;
    // queries are independent: P threads take them in an interleaved fashion (bench.py's CPU baseline
    // when oracle/_ref/dicey_ref is not available); the timing line mirrors dicey_ref's
    std::vector<QRes> res(ql.size());
    unsigned P = a.threads ? a.threads : 1;
    auto t0 = std::chrono::steady_clock::now();
    {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < P; ++t)
        th.emplace_back([&, t] {
          for (size_t i = t; i < ql.size(); i += P) {
            std::string name, seq;
            split_query(ql[i], name, seq);
            run_hunt(fm, seqlen, seq, p, res[i]);
          }
        });
      for (auto& t : th) t.join();
    }
    double loop_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    uint64_t nhits = 0;
    for (auto& r : res) nhits += r.push.size();
    if (!a.records.empty()) {
      std::ofstream os(a.records.c_str());
      for (size_t i = 0; i < ql.size(); ++i) {
        const QRes& r = res[i];
        os << "Q\t" << i << '\t' << r.seq << '\t' << r.distance << '\t' << r.msg.size() << '\t' << r.push.size() << '\n';
        for (auto& m : r.msg) os << "M\t" << m << '\n';
        for (auto& h : r.push) os << "P\t" << h.score << '\t' << h.chr << '\t' << h.start << '\t' << h.strand << '\t' << h.ra << '\t' << h.qa << '\n';
        for (auto& h : r.sorted) os << "S\t" << h.score << '\t' << h.chr << '\t' << h.start << '\t' << h.strand << '\t' << h.ra << '\t' << h.qa << '\n';
      }
    }
    std::cout << "{\"queries\": " << ql.size() << ", \"threads\": " << P << ", \"loop_s\": " << loop_s << ", \"queries_per_s\": "
              << (loop_s > 0 ? ql.size() / loop_s : 0.0) << ", \"hits\": " << nhits << ", \"n\": " << fm.size() << "}" << std::endl;
    return 0;
  }
  if (cmd == "seed") {  // silica.h:449-573 without the thal gate
    size_t maxloc = a.mset ? a.m : 10000;
    bool indel = !a.hamming;
    for (size_t qi = 0; qi < ql.size(); ++qi) {
      std::string name, seq;
      split_query(ql[qi], name, seq);
      for (auto& c : seq) c = (char)std::toupper((unsigned char)c);
      if (seq.size() <= a.k) { std::cout << "Q\t" << qi << "\tskipped\n"; continue; }
      uint32_t koffset = (uint32_t)seq.size() - a.k;
      std::string sequence = seq.substr(seq.size() - a.k), rev = revcomp(sequence);
      std::vector<std::set<std::string>> fwrv(2);
      neighbors(sequence, (int)a.d, indel, a.x, fwrv[0]);
      neighbors(rev, (int)a.d, indel, a.x, fwrv[1]);
      size_t ctx = indel ? a.d : 0;
      uint32_t hits = 0;
      std::ostringstream body;
      for (uint32_t fr = 0; fr < 2; ++fr)
        for (auto it = fwrv[fr].begin(); it != fwrv[fr].end() && hits < maxloc; ++it) {
          const std::string& query = *it;
          size_t m = query.size(), occs = count(fm, query);
          if (!occs) continue;
          auto loc = locate(fm, query);
          std::sort(loc.begin(), loc.end());
          for (size_t i = 0; i < std::min(occs, maxloc) && hits < maxloc; ++i) {
            uint32_t ref, chrpos;
            locate_chr(seqlen, (int64_t)loc[i], ref, chrpos);
            size_t pre_x = ctx, post_x = ctx;
            if (fr) post_x += koffset; else pre_x += koffset;
            if (pre_x > loc[i]) pre_x = loc[i];
            if (loc[i] + m + post_x > fm.size()) post_x = fm.size() - loc[i] - m;
            std::string s = extract(fm, loc[i] - pre_x, loc[i] + m + post_x - 1);
            std::string pre = s.substr(0, pre_x);
            s = s.substr(pre_x);
            if (pre.find_last_of('\n') != std::string::npos) pre = pre.substr(pre.find_last_of('\n') + 1);
            std::string post = s.substr(m);
            post = post.substr(0, post.find_first_of('\n'));
            std::string g = pre + s.substr(0, m) + post;
            if (pre.size() <= chrpos) chrpos -= (uint32_t)pre.size();  // silica.h:501
            std::string r0, r1;
            needle(g, fr ? rev : sequence, r0, r1);
            uint32_t alignpos = chrpos;
            bool lead = true;
            for (uint32_t j = 0; j < r1.size() - trail_gap(r1); ++j) { if (r1[j] != '-') lead = false; if (lead) ++alignpos; }
            body << "C\t" << fr << '\t' << ref << '\t' << chrpos << '\t' << alignpos << '\t' << g << '\t' << query << '\n';
            ++hits;
          }
        }
      std::cout << "Q\t" << qi << '\t' << seq << '\t' << hits << '\n' << body.str();
    }
    return 0;
  }
  return 2;
}
