// tests/hostsim -- TEST INFRASTRUCTURE.  Compiles the per-item arithmetic of
// dicey_b200/csrc/dg_core.cuh with g++ so that the script enumeration, the antichain rule and
// the NW traceback can be checked against the reference on a box without a GPU.  Nothing here
// is linked into the product library.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <fstream>
#include <set>
#include <string>
#include <vector>
#include <pthread.h>
#include <thread>
#include "../../dicey_b200/csrc/dg_core.cuh"
#include "../../dicey_b200/csrc/nbr_trunc.hpp"
#include "../../dicey_b200/csrc/fm9.hpp"
#include "../../dicey_b200/csrc/fm9_select.hpp"
#include "../../dicey_b200/csrc/dg_thal.cuh"
#include "../../dicey_b200/csrc/thal_params.hpp"
#include "../../dicey_b200/host/jsonnum.hpp"

using namespace dg;

static void neighbors(const std::string& q, int d, bool indel, std::set<std::string>& out) {
  int m = (int)q.size();
  const uint8_t* base = (const uint8_t*)q.data();
  int E = slots_per_pos(indel) * m;
  std::set<std::string> all;
  std::vector<uint8_t> buf(m + 8);
  auto add = [&](const Script& sc) {
    int L = script_ltr(base, m, sc, buf.data());
    // cross-check the right-to-left emitter
    std::string rt;
    script_rtl(base, m, sc, [&](uint8_t x) { rt.push_back((char)x); return true; });
    std::string lt((char*)buf.data(), L);
    if (std::string(rt.rbegin(), rt.rend()) != lt || L != script_len(m, sc)) { fprintf(stderr, "rtl/ltr mismatch\n"); exit(3); }
    // cross-check the packed fast path (ACGT-only, short strings)
    bool clean = true;
    for (char ch : q) if (base_code((uint8_t)ch) == 4) clean = false;
    if (clean && m + sc.nev <= kMaxPacked) {
      uint64_t code = 0;
      for (int i = 0; i < m; ++i) code = (code << 2) | (uint64_t)base_code(base[i]);
      for (int e = 0; e < sc.nev; ++e) code = apply_event_packed(code, m - 1 - sc.pos[e], sc.k[e]);
      std::string pk;
      for (int t = L - 1; t >= 0; --t) pk.push_back((char)code_base((int)((code >> (2 * t)) & 3)));
      if (pk != lt || (L < 32 && (code >> (2 * L)) != 0)) { fprintf(stderr, "packed mismatch %s vs %s\n", pk.c_str(), lt.c_str()); exit(3); }
    }
    all.insert(lt);
  };
  Script sc; sc.nev = 0; sc.pos[0] = sc.pos[1] = sc.k[0] = sc.k[1] = 0;
  add(sc);
  if (d >= 1)
    for (int e1 = 0; e1 < E; ++e1) {
      int p1, k1;
      if (!decode_event(base, m, indel, e1, p1, k1)) continue;
      sc.nev = 1; sc.pos[0] = p1; sc.k[0] = k1;
      add(sc);
      if (d >= 2)
        for (int e2 = second_event_start(p1, k1, indel); e2 < E; ++e2) {
          int p2, k2;
          if (!decode_event(base, m, indel, e2, p2, k2)) continue;
          if (!pair_ok(p1, k1, p2)) continue;
          sc.nev = 2; sc.pos[1] = p2; sc.k[1] = k2;
          add(sc);
        }
    }
  if (!indel) { out = all; return; }
  std::vector<uint8_t> s0(m + 8), s1(m + 8);
  for (auto const& t : all) {
    bool keep = is_minimal(base, m, d, (const uint8_t*)t.data(), (int)t.size(), s0.data(), s1.data());
    if (m + d <= 31) {  // the register-resident form used by k_cand_keys must agree on every string
      bool keep2 = is_minimal_small(pack4(base, m), m, d, pack4((const uint8_t*)t.data(), (int)t.size()), (int)t.size());
      if (keep2 != keep) { fprintf(stderr, "is_minimal_small mismatch on %s / %s\n", q.c_str(), t.c_str()); exit(3); }
      bool keep4 = is_minimal_band(pack4(base, m), m, d, pack4((const uint8_t*)t.data(), (int)t.size()), (int)t.size());
      if (keep4 != keep) { fprintf(stderr, "is_minimal_band mismatch on %s / %s (d %d): %d vs %d\n", q.c_str(), t.c_str(), d, (int)keep4, (int)keep); exit(3); }
      if (d == 1 && m >= 2) {
        bool keep3 = is_minimal_d1(pack4(base, m), m, pack4((const uint8_t*)t.data(), (int)t.size()), (int)t.size());
        if (keep3 != keep) { fprintf(stderr, "is_minimal_d1 mismatch on %s / %s\n", q.c_str(), t.c_str()); exit(3); }
      }
    }
    if (keep) out.insert(t);
  }
}

// Cooperating lanes as threads: the Warp concept of dg_thal.cuh with its collectives built on a
// barrier, so that the group logic the GPU kernels rely on (ballot compaction, one lane group per
// cell of a row, arg-min within a group) runs -- with real concurrency -- in the CPU suite.
// Eight lanes for the shared-memory form (thal8), a full warp of 32 for the wide form (thalw32).
template <int N>
struct LaneSharedN {
  pthread_barrier_t bar;
  unsigned pred[N];
  double g[N], S[N], H[N];
  uint32_t o[N];
};
template <int N>
struct ThalThreadLanesN {
  static constexpr int n = N;
  static constexpr unsigned kAll = N >= 32 ? 0xffffffffu : ((1u << (N & 31)) - 1u);
  int lane = 0;
  LaneSharedN<N>* sh = nullptr;
  void sync() const { pthread_barrier_wait(&sh->bar); }
  unsigned ballot(bool p) const {
    sh->pred[lane] = p ? 1u : 0u;
    sync();
    unsigned m = 0;
    for (int i = 0; i < n; ++i) m |= sh->pred[i] << i;
    sync();
    return m;
  }
  unsigned lanemask_lt() const { return (1u << lane) - 1u; }
  bool all(bool p) const { return ballot(p) == kAll; }
  bool any(bool p) const { return ballot(p) != 0u; }
  unsigned group_mask(int first_lane, int lanes, bool member) const {
    if (!member) return 1u << lane;
    return lanes >= n ? kAll : (((1u << lanes) - 1u) << first_lane);
  }
  void argmin(unsigned mask, double& g, uint32_t& o, double& S, double& H) const {
    sh->g[lane] = g; sh->o[lane] = o; sh->S[lane] = S; sh->H[lane] = H;
    sync();
    int best = -1;
    for (int i = 0; i < n; ++i) {
      if (!((mask >> i) & 1u)) continue;
      if (best < 0 || sh->g[i] < sh->g[best] || (sh->g[i] == sh->g[best] && sh->o[i] < sh->o[best])) best = i;
    }
    const double bg = sh->g[best], bS = sh->S[best], bH = sh->H[best];
    const uint32_t bo = sh->o[best];
    sync();
    g = bg; o = bo; S = bS; H = bH;
  }
};
typedef LaneSharedN<8> LaneShared;
typedef ThalThreadLanesN<8> ThalThreadLanes;

static void print_tm(int rc, double tm) {
  uint64_t u;
  memcpy(&u, &tm, 8);
  char buf[64];
  snprintf(buf, sizeof(buf), "%.17g", tm);
  std::cout << rc << '\t' << buf << '\t' << std::hex << u << std::dec << '\n';
}

// thal_end1_tm_lanes (the shared-memory form of k_thal_warp) on N concurrent lanes
template <int N>
static int run_thal_lanes(const ThalParams& tp, const char* pairs) {
  std::ifstream f(pairs);
  std::string line;
  LaneSharedN<N> sh;
  pthread_barrier_init(&sh.bar, nullptr, N);
  while (std::getline(f, line)) {
    size_t t = line.find('\t');
    if (t == std::string::npos) continue;
    std::string o1 = line.substr(0, t), o2 = line.substr(t + 1);
    const size_t cells = o1.size() * o2.size();
    std::vector<uint8_t> n1(o1.size() + 2), n2(o2.size() + 2), codes(256);
    std::vector<double> tab(2 * cells + 2);
    std::vector<uint16_t> plist(cells + 1), rstart(o1.size() + 2);
    double tms[N];
    int rcs[N];
    std::vector<std::thread> lanes;
    for (int l = 0; l < N; ++l)
      lanes.emplace_back([&, l] {
        ThalThreadLanesN<N> wp;
        wp.lane = l;
        wp.sh = &sh;
        rcs[l] = thal_end1_tm_lanes(wp, &tp, (const uint8_t*)o1.data(), (int)o1.size(), (const uint8_t*)o2.data(), (int)o2.size(),
                                    n1.data(), n2.data(), tab.data(), plist.data(), rstart.data(), codes.data(), &tms[l]);
      });
    for (auto& th : lanes) th.join();
    for (int l = 1; l < N; ++l)
      if (rcs[l] != rcs[0] || memcmp(&tms[l], &tms[0], 8) != 0) { fprintf(stderr, "lanes disagree\n"); return 3; }
    if (rcs[0] == 2) { fprintf(stderr, "sequential form requested\n"); return 3; }
    print_tm(rcs[0], tms[0]);
  }
  pthread_barrier_destroy(&sh.bar);
  return 0;
}

// thal_end1_tm_wide on N concurrent lanes (N = 1: the plain one-lane concept)
template <int N>
static int run_thal_wide(const ThalParams& tp, const char* pairs) {
  std::ifstream f(pairs);
  std::string line;
  LaneSharedN<(N > 1 ? N : 2)> sh;
  if (N > 1) pthread_barrier_init(&sh.bar, nullptr, N);
  while (std::getline(f, line)) {
    size_t t = line.find('\t');
    if (t == std::string::npos) continue;
    std::string o1 = line.substr(0, t), o2 = line.substr(t + 1);
    const bool fits = thal_lengths_ok((int)o1.size(), (int)o2.size());
    const size_t cells = fits ? o1.size() * o2.size() : 1;
    std::vector<uint8_t> n1(o1.size() + 2), n2(o2.size() + 2), a1(o1.size() + 2), ra(o1.size() + 2), b(o2.size() + 2), rb(o2.size() + 2);
    std::vector<double> tab(2 * cells + 2);
    double tms[N > 1 ? N : 1];
    int rcs[N > 1 ? N : 1];
    if (N == 1) {
      ThalOneLane wp;
      rcs[0] = thal_end1_tm_wide(wp, &tp, (const uint8_t*)o1.data(), (int)o1.size(), (const uint8_t*)o2.data(), (int)o2.size(), n1.data(),
                                 n2.data(), tab.data(), a1.data(), ra.data(), b.data(), rb.data(), &tms[0]);
    } else {
      std::vector<std::thread> lanes;
      for (int l = 0; l < N; ++l)
        lanes.emplace_back([&, l] {
          ThalThreadLanesN<(N > 1 ? N : 2)> wp;
          wp.lane = l;
          wp.sh = &sh;
          rcs[l] = thal_end1_tm_wide(wp, &tp, (const uint8_t*)o1.data(), (int)o1.size(), (const uint8_t*)o2.data(), (int)o2.size(),
                                     n1.data(), n2.data(), tab.data(), a1.data(), ra.data(), b.data(), rb.data(), &tms[l]);
        });
      for (auto& th : lanes) th.join();
      for (int l = 1; l < N; ++l)
        if (rcs[l] != rcs[0] || memcmp(&tms[l], &tms[0], 8) != 0) { fprintf(stderr, "lanes disagree\n"); return 3; }
    }
    if (rcs[0] == 2) { fprintf(stderr, "sequential form requested\n"); return 3; }
    print_tm(rcs[0], tms[0]);
  }
  if (N > 1) pthread_barrier_destroy(&sh.bar);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::string cmd = argv[1];
  if (cmd == "neighbors" && argc >= 5) {
    std::ifstream f(argv[2]);
    int d = atoi(argv[3]);
    bool indel = atoi(argv[4]) != 0;
    std::string q;
    while (std::getline(f, q)) {
      if (q.empty()) continue;
      std::set<std::string> st;
      neighbors(q, d, indel, st);
      std::cout << "Q\t" << q << '\t' << st.size() << '\n';
      for (auto const& s : st) std::cout << s << '\n';
    }
    return 0;
  }
  if (cmd == "pairslots") {
    // pair_slot (the index arithmetic of k_probe_pairs / k_resolve): every slot of every length maps to the
    // (first position, first kind, second position) the row-major layout defines, for both kind counts
    long checked = 0;
    for (int m = 1; m <= 31; ++m)
      for (int S : {3, 8}) {
        uint32_t u = 0;
        for (int p1 = 0; p1 < m; ++p1)
          for (int k1 = 0; k1 < S; ++k1)
            for (int p2 = p1; p2 < m; ++p2, ++u) {
              int a, b, c;
              if (S == 3) pair_slot<3>(u, m, a, b, c); else pair_slot<8>(u, m, a, b, c);
              if (a != p1 || b != k1 || c != p2) { fprintf(stderr, "pair_slot(%u, m=%d, S=%d) = (%d,%d,%d), want (%d,%d,%d)\n", u, m, S, a, b, c, p1, k1, p2); return 3; }
              ++checked;
            }
        if (u != (uint32_t)(S * m * (m + 1) / 2)) { fprintf(stderr, "slot count\n"); return 3; }
      }
    printf("pair_slot checked on %ld slots\n", checked);
    return 0;
  }
  if (cmd == "replay" && argc >= 6) {
    // the reference's generation order replayed (nbr_trunc.hpp): sets as `dicey_ref neighbors -x` prints them
    std::ifstream f(argv[2]);
    int d = atoi(argv[3]);
    bool indel = atoi(argv[4]) != 0;
    uint32_t x = (uint32_t)strtoul(argv[5], nullptr, 10);
    std::string q;
    NeighborReplay nr;
    NeighborReplayPacked np;
    while (std::getline(f, q)) {
      if (q.empty()) continue;
      const bool capped = nr.run(q, d, indel, x);
      std::vector<std::string> v = nr.strings();
      std::sort(v.begin(), v.end());
      if (NeighborReplayPacked::fits(q, d)) {   // the packed form must give the same set and the same verdict
        const bool capped2 = np.run(q, d, indel, x);
        std::vector<std::string> v2 = np.strings();
        std::sort(v2.begin(), v2.end());
        if (capped2 != capped || v2 != v || np.peak() != nr.peak()) {
          fprintf(stderr, "packed replay differs on %s (d %d indel %d x %u): %zu vs %zu strings\n", q.c_str(), d, (int)indel, x, v2.size(), v.size());
          return 3;
        }
      }
      std::cout << "Q\t" << q << '\t' << v.size() << '\n';
      for (auto const& s : v) std::cout << s << '\n';
    }
    return 0;
  }
  if (cmd == "nbrbound" && argc >= 4) {
    // nbr_upper_bound_part (the certificate that a neighbourhood stays under the cap) must never fall
    // below the number of distinct strings the scripts spell: random and low-complexity queries,
    // with and without N, d = 1 and 2.  argv[2] = number of queries, argv[3] = seed
    int nq = atoi(argv[2]);
    uint64_t x = strtoull(argv[3], nullptr, 10) * 2654435761ULL + 12345;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    double worst = 0, sum20 = 0; int n20 = 0; uint32_t max20 = 0;
    for (int i = 0; i < nq; ++i) {
      int m = 10 + (int)(rnd() % 13);
      if (i % 3 == 0) m = 20;
      int sigma = 1 + (int)(rnd() % 4);
      if (i % 2) sigma = 4;
      std::string q;
      for (int j = 0; j < m; ++j) q.push_back("ACGT"[rnd() % sigma]);
      if (i % 7 == 0) q[rnd() % m] = 'N';
      if (i % 11 == 0) for (int j = 1; j < m; ++j) if (rnd() % 2) q[j] = q[j - 1];   // long runs
      for (int d = 1; d <= 2; ++d) {
        // distinct strings, by enumeration
        std::set<std::string> all;
        const uint8_t* base = (const uint8_t*)q.data();
        std::vector<uint8_t> buf(m + 8);
        Script sc; sc.nev = 0; sc.pos[0] = sc.pos[1] = sc.k[0] = sc.k[1] = 0;
        all.insert(q);
        const int E = 9 * m;
        for (int e1 = 0; e1 < E; ++e1) {
          int p1, k1;
          if (!decode_event(base, m, true, e1, p1, k1)) continue;
          sc.nev = 1; sc.pos[0] = p1; sc.k[0] = k1;
          all.insert(std::string((char*)buf.data(), script_ltr(base, m, sc, buf.data())));
          if (d >= 2)
            for (int e2 = second_event_start(p1, k1, true); e2 < E; ++e2) {
              int p2, k2;
              if (!decode_event(base, m, true, e2, p2, k2) || !pair_ok(p1, k1, p2)) continue;
              sc.nev = 2; sc.pos[1] = p2; sc.k[1] = k2;
              all.insert(std::string((char*)buf.data(), script_ltr(base, m, sc, buf.data())));
            }
          sc.nev = 0;
        }
        if (i % 8 == 0 && m + d <= 31) {   // the one-pass antichain test against the per-substring DP, on every string
          std::vector<uint8_t> s0(m + 8), s1(m + 8);
          const Packed4 q4 = pack4(base, m);
          for (auto const& t : all) {
            const bool a = is_minimal(base, m, d, (const uint8_t*)t.data(), (int)t.size(), s0.data(), s1.data());
            const bool b2 = is_minimal_band(q4, m, d, pack4((const uint8_t*)t.data(), (int)t.size()), (int)t.size());
            if (a != b2) { fprintf(stderr, "is_minimal_band mismatch on %s / %s (d %d)\n", q.c_str(), t.c_str(), d); return 3; }
          }
        }
        auto bq = [&](int j) { return base_code(base[j]); };
        uint32_t U = 1 + nbr_upper_bound_part(bq, m, d, 0, E, 1);
        // the same split over 32 "lanes" must add up
        uint32_t U2 = 1;
        for (int lane = 0; lane < 32; ++lane) U2 += nbr_upper_bound_part(bq, m, d, lane, E, 32);
        if (U != U2) { fprintf(stderr, "lane split mismatch\n"); return 3; }
        const uint32_t U3 = nbr_upper_bound_closed(bq, m, d);   // what the kernel evaluates
        if (U3 != U) { fprintf(stderr, "closed form %u != enumeration %u on %s d=%d\n", U3, U, q.c_str(), d); return 3; }
        if (U < all.size()) { fprintf(stderr, "UNSOUND: %s d=%d bound %u < distinct %zu\n", q.c_str(), d, U, all.size()); return 3; }
        double r = (double)U / (double)all.size();
        if (r > worst) worst = r;
        if (m == 20 && d == 2 && q.find('N') == std::string::npos) { sum20 += U; ++n20; if (U > max20) max20 = U; }
      }
    }
    printf("bound >= distinct on %d queries; worst bound/distinct %.3f; 20-mers at d=2: mean bound %.0f, max %u\n", nq, worst,
           n20 ? sum20 / n20 : 0.0, max20);
    return 0;
  }
  if (cmd == "needle" && argc >= 3) {
    std::ifstream f(argv[2]);
    std::string line;
    int banded_checked = 0, ops_checked = 0;
    while (std::getline(f, line)) {
      size_t t = line.find('\t');
      if (t == std::string::npos) continue;
      std::string g = line.substr(0, t), s = line.substr(t + 1);
      int mg = (int)g.size(), n = (int)s.size();
      std::vector<uint8_t> trace(((mg + 1) * (n + 1) + 3) / 4 + 1), ops(mg + n + 1), ra(mg + n + 1), qa(mg + n + 1);
      std::vector<int> srow(n + 1);
      int lead = 0, score = 0;
      TraceBytes tb{trace.data(), n + 1};
      int kept = needle_align((const uint8_t*)g.data(), mg, (const uint8_t*)s.data(), n, tb, srow.data(),
                              ops.data(), ra.data(), qa.data(), &lead, &score);
      if (n <= 31) {  // the thread-local trace layout of k_verify must agree with the byte layout
        std::vector<uint64_t> rows(mg + 1, 0);
        std::vector<uint8_t> ra2(mg + n + 1), qa2(mg + n + 1);
        int lead2 = 0, score2 = 0;
        TraceRows64 tr{rows.data()};
        int kept2 = needle_align((const uint8_t*)g.data(), mg, (const uint8_t*)s.data(), n, tr, srow.data(), ops.data(),
                                 ra2.data(), qa2.data(), &lead2, &score2);
        if (kept2 != kept || lead2 != lead || score2 != score || memcmp(ra.data(), ra2.data(), kept) || memcmp(qa.data(), qa2.data(), kept)) {
          fprintf(stderr, "TraceRows64 / TraceBytes mismatch\n");
          return 3;
        }
      }
      // the banded register form used by k_verify in hunt mode: exact whenever score >= -dmax
      for (int dmax = -score; dmax <= -score + 1; ++dmax) {
        int hi = mg - n + dmax, W = hi + dmax + 1;
        if (n > 31 || hi < 0 || W > kBandMax || dmax < 0) continue;
        std::vector<uint32_t> tr(mg + 1, 0);
        std::vector<uint8_t> ra3(mg + n + 1), qa3(mg + n + 1);
        int score3 = needle_banded_fill((const uint8_t*)g.data(), mg, (const uint8_t*)s.data(), n, dmax, tr.data());
        int nops = 0, lead3 = 0, trail3 = 0;
        needle_banded_shape(tr.data(), mg, n, dmax, &nops, &lead3, &trail3);
        int kept3 = nops - lead3 - trail3;
        needle_banded_emit(tr.data(), (const uint8_t*)g.data(), mg, (const uint8_t*)s.data(), n, dmax, nops, lead3, trail3,
                           ra3.data(), qa3.data());
        ++banded_checked;
        {  // the compact record form: ops from the trace, rows rebuilt from the query alone
          int nop = 0;
          uint64_t bits = needle_banded_ops(tr.data(), (const uint8_t*)g.data(), mg, (const uint8_t*)s.data(), n, dmax, nops, lead3,
                                            trail3, &nop);
          if (nop != -score) { fprintf(stderr, "ops count %d != -score %d on %s %s\n", nop, -score, g.c_str(), s.c_str()); return 3; }
          if (nop <= kRecOps) {
            std::vector<uint8_t> ra4(n + kRecOps + 1), qa4(n + kRecOps + 1);
            int cols = rec_expand_rows(bits, nop, (const uint8_t*)s.data(), n, ra4.data(), qa4.data());
            if (cols != kept || memcmp(ra.data(), ra4.data(), kept) || memcmp(qa.data(), qa4.data(), kept)) {
              fprintf(stderr, "compact ops do not rebuild the alignment of %s %s\n", g.c_str(), s.c_str());
              return 3;
            }
            ++ops_checked;
          }
        }
        if (score3 != score || kept3 != kept || lead3 != lead || memcmp(ra.data(), ra3.data(), kept) || memcmp(qa.data(), qa3.data(), kept)) {
          fprintf(stderr, "banded / full needle mismatch on %s %s (dmax %d)\n", g.c_str(), s.c_str(), dmax);
          return 3;
        }
      }
      {  // the same from materialised rows (full-matrix paths of k_verify)
        int nop = 0;
        uint64_t bits = rows_to_ops(ra.data(), qa.data(), kept, &nop);
        if (nop <= kRecOps) {
          std::vector<uint8_t> ra4(n + kRecOps + 1), qa4(n + kRecOps + 1);
          int cols = rec_expand_rows(bits, nop, (const uint8_t*)s.data(), n, ra4.data(), qa4.data());
          if (cols != kept || memcmp(ra.data(), ra4.data(), kept) || memcmp(qa.data(), qa4.data(), kept)) {
            fprintf(stderr, "rows_to_ops does not round-trip on %s %s\n", g.c_str(), s.c_str());
            return 3;
          }
          ++ops_checked;
        }
      }
      std::cout << score << '\t' << lead << '\t' << std::string((char*)ra.data(), kept) << '\t'
                << std::string((char*)qa.data(), kept) << '\n';
    }
    fprintf(stderr, "banded needle checked on %d (pair, dmax) cases, compact ops on %d\n", banded_checked, ops_checked);
    return banded_checked >= 100 && ops_checked >= 200 ? 0 : 4;
  }
  if ((cmd == "thal8" || cmd == "thal32") && argc >= 4) {
    // the lane-cooperative arrangement on eight / thirty-two concurrent lanes (threads)
    ThalParams tp;
    std::string err;
    if (!thal_params_from_dump(argv[2], tp, err)) { fprintf(stderr, "%s\n", err.c_str()); return 2; }
    return cmd == "thal8" ? run_thal_lanes<8>(tp, argv[3]) : run_thal_lanes<32>(tp, argv[3]);
  }
  if ((cmd == "thalw" || cmd == "thalw8" || cmd == "thalw32") && argc >= 4) {
    // thal_end1_tm_wide (one side up to THAL_MAX_SEQ; what k_thal_wide runs) on 1 / 8 / 32 lanes
    ThalParams tp;
    std::string err;
    if (!thal_params_from_dump(argv[2], tp, err)) { fprintf(stderr, "%s\n", err.c_str()); return 2; }
    return cmd == "thalw" ? run_thal_wide<1>(tp, argv[3]) : cmd == "thalw8" ? run_thal_wide<8>(tp, argv[3]) : run_thal_wide<32>(tp, argv[3]);
  }
  if (cmd == "thalany" && argc >= 4) {
    // thal_end1_tm_any: the sequential form for every pair of lengths the reference accepts
    ThalParams tp;
    std::string err;
    if (!thal_params_from_dump(argv[2], tp, err)) { fprintf(stderr, "%s\n", err.c_str()); return 2; }
    std::ifstream f(argv[3]);
    std::string line;
    while (std::getline(f, line)) {
      size_t t = line.find('\t');
      if (t == std::string::npos) continue;
      std::string o1 = line.substr(0, t), o2 = line.substr(t + 1);
      const size_t cells = thal_lengths_ok((int)o1.size(), (int)o2.size()) ? o1.size() * o2.size() : 1;
      std::vector<uint8_t> n1(o1.size() + 2), n2(o2.size() + 2);
      std::vector<double> ds(cells + 1), dh(cells + 1);
      double tm = 0;
      bool ok = thal_end1_tm_any(&tp, (const uint8_t*)o1.data(), (int)o1.size(), (const uint8_t*)o2.data(), (int)o2.size(), n1.data(),
                                 n2.data(), ds.data(), dh.data(), &tm);
      print_tm(ok ? 1 : 0, tm);
    }
    return 0;
  }
  if (cmd == "thal2" && argc >= 4) {
    // the lane-cooperative arrangement of dg_thal.cuh (what the GPU kernel runs) with one lane
    ThalParams tp;
    std::string err;
    if (!thal_params_from_dump(argv[2], tp, err)) { fprintf(stderr, "%s\n", err.c_str()); return 2; }
    std::ifstream f(argv[3]);
    std::string line;
    while (std::getline(f, line)) {
      size_t t = line.find('\t');
      if (t == std::string::npos) continue;
      std::string o1 = line.substr(0, t), o2 = line.substr(t + 1);
      const size_t cells = o1.size() * o2.size();
      std::vector<uint8_t> n1(o1.size() + 2), n2(o2.size() + 2), codes(256);
      std::vector<double> tab(2 * cells + 2);
      std::vector<uint16_t> plist(cells + 1), rstart(o1.size() + 2);
      double tm = 0;
      ThalOneLane wp;
      int rc = thal_end1_tm_lanes(wp, &tp, (const uint8_t*)o1.data(), (int)o1.size(), (const uint8_t*)o2.data(), (int)o2.size(), n1.data(),
                                  n2.data(), tab.data(), plist.data(), rstart.data(), codes.data(), &tm);
      if (rc == 2) { fprintf(stderr, "sequential form requested\n"); return 3; }
      uint64_t u;
      memcpy(&u, &tm, 8);
      char buf[64];
      snprintf(buf, sizeof(buf), "%.17g", tm);
      std::cout << rc << '\t' << buf << '\t' << std::hex << u << std::dec << '\n';
    }
    return 0;
  }
  if (cmd == "thal" && argc >= 4) {
    // dg_thal.cuh on the host: <params.tsv> <pairs.tsv> -> the lines `dicey_ref thal` prints
    ThalParams tp;
    std::string err;
    if (!thal_params_from_dump(argv[2], tp, err)) { fprintf(stderr, "%s\n", err.c_str()); return 2; }
    std::ifstream f(argv[3]);
    std::string line;
    while (std::getline(f, line)) {
      size_t t = line.find('\t');
      if (t == std::string::npos) continue;
      std::string o1 = line.substr(0, t), o2 = line.substr(t + 1);
      std::vector<uint8_t> n1(o1.size() + 2), n2(o2.size() + 2);
      std::vector<double> ds(o1.size() * o2.size() + 1), dh(o1.size() * o2.size() + 1);
      double tm = 0;
      bool ok = thal_end1_tm(&tp, (const uint8_t*)o1.data(), (int)o1.size(), (const uint8_t*)o2.data(), (int)o2.size(), n1.data(),
                             n2.data(), ds.data(), dh.data(), &tm);
      uint64_t u;
      memcpy(&u, &tm, 8);
      char buf[64];
      snprintf(buf, sizeof(buf), "%.17g", tm);
      std::cout << (ok ? 1 : 0) << '\t' << buf << '\t' << std::hex << u << std::dec << '\n';
    }
    return 0;
  }
  if (cmd == "thalcfg" && argc >= 4) {
    // the primer3_config loader must reproduce the tables the reference holds: <config dir> <params.tsv>
    ThalParams a, b;
    std::string err;
    if (!thal_params_from_config(argv[2], 50.0, 1.5, 0.6, 50.0, a, err) || !thal_params_from_dump(argv[3], b, err)) {
      fprintf(stderr, "%s\n", err.c_str());
      return 2;
    }
    if (memcmp(&a, &b, sizeof(a)) != 0) {
      const double* x = (const double*)&a; const double* y = (const double*)&b;
      for (size_t i = 0; i < sizeof(a) / 8; ++i) if (memcmp(x + i, y + i, 8)) { fprintf(stderr, "first difference at double %zu: %.17g vs %.17g\n", i, x[i], y[i]); break; }
      return 3;
    }
    std::cout << "tables identical (" << sizeof(a) / 8 << " doubles)\n";
    return 0;
  }
  if (cmd == "jsonfloat" && argc >= 3) {
    // jsonnum.hpp: one 64-bit pattern per line -> what nlohmann::json(double).dump() prints
    std::ifstream f(argv[2]);
    std::string line;
    while (std::getline(f, line)) {
      uint64_t u = strtoull(line.c_str(), nullptr, 16);
      double d;
      memcpy(&d, &u, 8);
      std::cout << dhost::json_double(d) << '\n';
    }
    return 0;
  }
  if (cmd == "select" && argc >= 3) {
    // the select_support_mcl sections rebuilt from m_bv must equal the bytes SDSL wrote into the file
    FILE* f = fopen(argv[2], "rb");
    if (!f) return 2;
    Fm9Reader rd(f);
    uint64_t n, sigma, bits, rbits;
    uint8_t w;
    std::vector<uint64_t> bv;
    if (!rd.u64(n) || !rd.u64(sigma) || !rd.int_vector(&bv, bits, w) || !rd.int_vector(nullptr, rbits, w)) return 2;
    for (int one = 1; one >= 0; --one) {
      long a = ftell(f);
      if (!fm9_skip_select(rd)) return 2;
      long b = ftell(f);
      std::vector<uint8_t> want((size_t)(b - a));
      fseek(f, a, SEEK_SET);
      if (fread(want.data(), 1, want.size(), f) != want.size()) return 2;
      SelectMclWriter sel(bv.data(), bits, one == 1);
      if (sel.bytes() != want) {
        fprintf(stderr, "select_support_mcl<%d> differs: %zu bytes built, %zu in the file\n", one, sel.bytes().size(), want.size());
        return 3;
      }
      std::cout << "select_support_mcl<" << one << "> " << want.size() << " bytes identical\n";
    }
    fclose(f);
    return 0;
  }
  return 2;
}
