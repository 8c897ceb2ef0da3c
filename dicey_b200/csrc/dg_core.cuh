// dg_core.cuh -- per-item arithmetic of the hot path, shared by every kernel.
//
// Everything here is a plain inline function over raw pointers so that the same source is
// (a) inlined into the sm_100a kernels of dg_search.cu / dg_build.cu and (b) compiled by g++
// into tests/hostsim (unit tests of the arithmetic against the reference on a box without a
// GPU).  The product library contains the device instantiation only.
//
// What of the reference each part replaces:
//   OccBlock / rank_acgt / lf_step      SDSL wt_pc::rank + rank_support_v::rank (wt_pc.hpp:325-347,
//                                       rank_support_v.hpp:104-115), inverse_select (wt_pc.hpp:359-374)
//   backward_step                       backward_search (suffix_array_algorithm.hpp:151-179)
//   sa_value                            csa_wt::operator[] (csa_wt.hpp:340-354)
//   Script / script_rtl / script_ltr    the DFS of neighbors.h:47-83, unrolled into edit scripts
//   restricted_distance / is_minimal    the antichain rule of _insert (neighbors.h:29-45)
//   needle_align                        needle() (needle.h:59-138) for AlignConfig<false,true>,
//                                       DnaScore(0,-1,-1,-1), plus the gap stripping of
//                                       hunter.h:391-401 / silica.h:527-532
//   locate_record                       hunter.h:358-362 (text position -> refIndex, chrpos)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DG_HD __host__ __device__ __forceinline__
#else
#define DG_HD inline
struct uint2 { uint32_t x, y; };  // host-side stand-in for the CUDA vector type (tests/hostsim)
#endif

namespace dg {

constexpr int kMaxQuery = 255;     // longest query the device path accepts
constexpr int kMaxDist = 2;        // largest distance the device path enumerates
constexpr int kMaxListDist = 32;   // largest distance searched at all: beyond kMaxDist the neighbourhood comes as a list from
                                   // the host replay (nbr_trunc.hpp), strings of up to 42 characters
constexpr int kMaxCompactDist = 3; // a compact record (dg_rec) carries at most three edit operations: larger distances
                                   // return full records (dg_hit + alignment pool)
constexpr int kFlagShift = 12;     // one exception-flag bit per 4096 BWT rows
constexpr int kSaSample = 32;      // csa_wt<> t_dens

// ------------------------------------------------------------------------------------------
// Occurrence blocks: 64 BWT symbols per 32-byte block (one DRAM sector).  cnt[c] = number of
// symbol c (A,C,G,T = 0..3) in BWT[0, 64*b); lo/hi = bit planes of the 2-bit codes.  BWT
// symbols outside ACGT are stored as code 0 and listed in the exception arrays.
struct OccBlock {
  uint32_t cnt[4];
  uint64_t lo, hi;
};

struct IndexView {
  const OccBlock* occ;
  uint64_t n;                 // text length incl. sentinel; rows are 0..n-1
  const uint32_t* excflag;    // bit (row >> kFlagShift): region holds a non-ACGT BWT symbol
  const uint32_t* exc_pos;    // rows of non-ACGT BWT symbols, ascending
  const uint8_t* exc_sym;     // their byte values
  uint32_t n_exc;
  const uint32_t* rare_pos;   // the same rows grouped by symbol (ascending inside a group)
  const uint32_t* rare_off;   // 257 entries: group of byte s is [rare_off[s], rare_off[s+1])
  const uint32_t* Cb;         // 256 entries: number of text symbols smaller than byte s
  const uint8_t* present;     // 256 entries: byte s occurs in the text
  uint32_t C4[4];             // Cb['A'], Cb['C'], Cb['G'], Cb['T']
  const uint2* kmer;          // 4^K half-open SA intervals [x, y) of the ACGT K-mers
  uint32_t K;
  const uint32_t* sa_samples; // SA[32 k]
  const uint32_t* sa_full;    // SA[row] for every row, or null (then sa_samples + LF walks)
  const uint8_t* text;        // the n text bytes (sentinel included)
  const uint64_t* cum;        // nseq + 1 cumulative seqlen (util.h:201 lengths)
  uint32_t nseq;
  const uint32_t* present_kb; // 4^KB bits: the ACGT KB-mer with this packed code occurs in the text
  uint32_t KB;                // 0 = no presence bitmap
  // the same for KB + 1 (strings at least that long are tested there: 4 x fewer false survivors)
  // and KB - 1 (strings one base too short for the primary bitmap); null = not built
  const uint32_t* present_hi;
  const uint32_t* present_lo;
  const uint32_t* present_kb_l;   // the same sets addressed by presence_bit_left() (null: not built)
  const uint32_t* present_hi_l;
};

DG_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}

// A,C,G,T -> 0..3; anything else -> 4.
DG_HD int base_code(uint8_t b) {
  if (b == 'A') return 0;
  if (b == 'C') return 1;
  if (b == 'G') return 2;
  if (b == 'T') return 3;
  return 4;
}
DG_HD uint8_t code_base(int c) { return (uint8_t)("ACGT"[c & 3]); }

DG_HD uint32_t lower_bound_u32(const uint32_t* a, uint32_t lo, uint32_t hi, uint32_t key) {
  while (lo < hi) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
DG_HD uint32_t upper_bound_u64(const uint64_t* a, uint32_t lo, uint32_t hi, uint64_t key) {
  while (lo < hi) {  // first index with a[idx] > key
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (a[mid] <= key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

DG_HD OccBlock load_block(const OccBlock* p) {
#if defined(__CUDA_ARCH__)
  // one 32-byte sector, two 128-bit read-only loads
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  OccBlock o;
  o.cnt[0] = a.x; o.cnt[1] = a.y; o.cnt[2] = a.z; o.cnt[3] = a.w;
  o.lo = ((uint64_t)b.y << 32) | b.x;
  o.hi = ((uint64_t)b.w << 32) | b.z;
  return o;
#else
  return *p;
#endif
}

DG_HD bool region_flag(const IndexView& ix, uint64_t row) {
  uint64_t g = row >> kFlagShift;
  return (ix.excflag[g >> 5] >> (g & 31)) & 1u;
}

// occurrences of code c (0..3) in BWT[0, i), 0 <= i <= n, given the loaded block of i.
DG_HD uint32_t rank_in_block(const IndexView& ix, const OccBlock& b, uint64_t i, int c) {
  uint32_t o = (uint32_t)(i & 63);
  uint64_t mask = o ? (~0ULL >> (64 - o)) : 0ULL;
  uint64_t m = ((c & 1) ? b.lo : ~b.lo) & ((c & 2) ? b.hi : ~b.hi) & mask;
  uint32_t r = b.cnt[c] + (uint32_t)popc64(m);
  if (c == 0 && o && region_flag(ix, i - 1)) {
    // non-ACGT symbols are stored as code 0: take those inside [64*blk, i) back out
    uint32_t base = (uint32_t)(i & ~63ULL);
    uint32_t a = lower_bound_u32(ix.exc_pos, 0, ix.n_exc, base);
    uint32_t e = lower_bound_u32(ix.exc_pos, a, ix.n_exc, (uint32_t)i);
    r -= (e - a);
  }
  return r;
}
DG_HD uint32_t rank_acgt(const IndexView& ix, uint64_t i, int c) {
  OccBlock b = load_block(ix.occ + (i >> 6));
  return rank_in_block(ix, b, i, c);
}
// occurrences of an arbitrary byte s in BWT[0, i)
DG_HD uint32_t rank_byte(const IndexView& ix, uint64_t i, uint8_t s) {
  int c = base_code(s);
  if (c < 4) return rank_acgt(ix, i, c);
  uint32_t a = ix.rare_off[s], e = ix.rare_off[s + 1];
  return lower_bound_u32(ix.rare_pos, a, e, (uint32_t)(i > 0xFFFFFFFFULL ? 0xFFFFFFFFULL : i)) - a;
}

// One backward-search step on the half-open interval [l, r): suffix_array_algorithm.hpp:151-179.
// (The reference special-cases the full range; rank(0) = 0 and rank(n) = count make the general
// formula give the same interval.)  A byte that does not occur in the text empties the interval.
DG_HD void backward_step(const IndexView& ix, uint32_t& l, uint32_t& r, uint8_t s) {
  int c = base_code(s);
  if (c < 4) {
    OccBlock bl = load_block(ix.occ + (l >> 6));
    uint32_t nl = ix.C4[c] + rank_in_block(ix, bl, l, c);
    uint32_t nr;
    if ((r >> 6) == (l >> 6)) nr = ix.C4[c] + rank_in_block(ix, bl, r, c);
    else nr = ix.C4[c] + rank_acgt(ix, r, c);
    l = nl; r = nr;
  } else {
    if (!ix.present[s]) { r = l; return; }
    uint32_t nl = ix.Cb[s] + rank_byte(ix, l, s);
    uint32_t nr = ix.Cb[s] + rank_byte(ix, r, s);
    l = nl; r = nr;
  }
}

// BWT[row] and LF(row): traverse_csa_wt_traits::access -> wt_pc::inverse_select
// (suffix_array_helper.hpp:280-292, wt_pc.hpp:359-374).
DG_HD uint32_t lf_step(const IndexView& ix, uint32_t row, uint8_t* sym) {
  OccBlock b = load_block(ix.occ + (row >> 6));
  uint32_t o = row & 63;
  int c = (int)(((b.hi >> o) & 1) << 1 | ((b.lo >> o) & 1));
  if (c == 0 && region_flag(ix, row)) {
    uint32_t k = lower_bound_u32(ix.exc_pos, 0, ix.n_exc, row);
    if (k < ix.n_exc && ix.exc_pos[k] == row) {
      uint8_t s = ix.exc_sym[k];
      if (sym) *sym = s;
      uint32_t a = ix.rare_off[s], e = ix.rare_off[s + 1];
      return ix.Cb[s] + (lower_bound_u32(ix.rare_pos, a, e, row) - a);
    }
  }
  if (sym) *sym = code_base(c);
  return ix.C4[c] + rank_in_block(ix, b, row, c);
}

// SA[row]: csa_wt::operator[] (csa_wt.hpp:340-354) with sa_order_sa_sampling, t_dens = 32.
DG_HD uint32_t sa_value(const IndexView& ix, uint32_t row) {
  if (ix.sa_full) return ix.sa_full[row];
  uint32_t off = 0;
  while (row & (kSaSample - 1)) { row = lf_step(ix, row, nullptr); ++off; }
  uint64_t v = (uint64_t)ix.sa_samples[row / kSaSample] + off;
  if (v >= ix.n) v -= ix.n;
  return (uint32_t)v;
}

// ------------------------------------------------------------------------------------------
// Edit scripts.  The reference DFS (neighbors.h:47-83) walks the query left to right and at
// each position may delete the base, keep it, substitute one of the other alphabet letters, or
// insert one of the 4 letters before it; a string is recorded when at least one edit was spent.
// Every recorded string is therefore "query + a list of <= d events sorted by position, where
// several insertions (ordered) may precede one substitution/deletion at the same position".
// An event is a slot number e = pos * slots + k; edit mode has 9 slots per position
// (k 0..3 substitute ACGT[k], 4 delete, 5..8 insert ACGT[k-5] before pos), Hamming mode 4.
struct Script {
  int nev;      // 0..2
  int pos[2];
  int k[2];
};

DG_HD int slots_per_pos(bool indel) { return indel ? 9 : 4; }

// Decodes slot e and checks it against the base string (a substitution must change the base).
DG_HD bool decode_event(const uint8_t* base, int m, bool indel, int e, int& pos, int& k) {
  int s = indel ? 9 : 4;
  pos = e / s;
  k = e - pos * s;
  if (pos >= m) return false;
  if (k < 4 && code_base(k) == base[pos]) return false;
  return true;
}
// Is (e1, e2) an admissible ordered pair?  pos1 < pos2, or equal positions with e1 an insertion.
DG_HD bool pair_ok(int pos1, int k1, int pos2) { return pos1 < pos2 || (pos1 == pos2 && k1 >= 5); }
// First slot a second event may take after first event (pos1, k1).
DG_HD int second_event_start(int pos1, int k1, bool indel) {
  int s = indel ? 9 : 4;
  return (k1 >= 5 ? pos1 : pos1 + 1) * s;
}

// Calls f(byte) for every character of the edited string from the LAST to the first; f returns
// false to stop early.  Returns false if stopped.
template <typename F>
DG_HD bool script_rtl(const uint8_t* base, int m, const Script& sc, F&& f) {
  int ev = sc.nev - 1;
  for (int p = m - 1; p >= 0; --p) {
    uint8_t x = base[p];
    bool emit = true;
    if (ev >= 0 && sc.pos[ev] == p && sc.k[ev] <= 4) {
      if (sc.k[ev] == 4) emit = false; else x = code_base(sc.k[ev]);
      --ev;
    }
    if (emit && !f(x)) return false;
    while (ev >= 0 && sc.pos[ev] == p) {
      if (!f(code_base(sc.k[ev] - 5))) return false;
      --ev;
    }
  }
  return true;
}
// Left-to-right materialisation; returns the length (<= m + nev).
DG_HD int script_ltr(const uint8_t* base, int m, const Script& sc, uint8_t* out) {
  int ev = 0, L = 0;
  for (int p = 0; p < m; ++p) {
    while (ev < sc.nev && sc.pos[ev] == p && sc.k[ev] >= 5) { out[L++] = code_base(sc.k[ev] - 5); ++ev; }
    uint8_t x = base[p];
    bool emit = true;
    if (ev < sc.nev && sc.pos[ev] == p) {
      if (sc.k[ev] == 4) emit = false; else x = code_base(sc.k[ev]);
      ++ev;
    }
    if (emit) out[L++] = x;
  }
  return L;
}
DG_HD int script_len(int m, const Script& sc) {
  int L = m;
  for (int i = 0; i < sc.nev; ++i) { if (sc.k[i] == 4) --L; else if (sc.k[i] >= 5) ++L; }
  return L;
}

// Packed script code carried by a candidate: strand | nev | e1 | e2.
DG_HD uint32_t pack_script(int strand, int nev, int e1, int e2) {
  return (uint32_t)strand | ((uint32_t)nev << 1) | ((uint32_t)e1 << 3) | ((uint32_t)e2 << 15);
}
DG_HD void unpack_script(uint32_t code, bool indel, int& strand, Script& sc) {
  strand = code & 1;
  sc.nev = (code >> 1) & 3;
  int s = indel ? 9 : 4;
  int e1 = (code >> 3) & 0xFFF, e2 = (code >> 15) & 0xFFF;
  sc.pos[0] = e1 / s; sc.k[0] = e1 - sc.pos[0] * s;
  sc.pos[1] = e2 / s; sc.k[1] = e2 - sc.pos[1] * s;
}

// ------------------------------------------------------------------------------------------
// Membership in the generated set G_d(q) of neighbors.h: minimum number of DFS edits turning q
// into u, where an insertion is only possible while a query base is still pending (the DFS
// never inserts after the last base, neighbors.h:49,71-77) and only ACGT can be written.
// Returns min(cost, cap + 1).  row/prev are caller scratch of >= ulen + 1 entries.
DG_HD int restricted_distance(const uint8_t* q, int m, const uint8_t* u, int ulen, int cap,
                              uint8_t* prev, uint8_t* row) {
  const int INF = 100;
  int dl = m - ulen;
  if (dl > cap || -dl > cap) return cap + 1;  // every insertion / deletion changes the length by one
  // only the diagonal band |i - j| <= cap can hold values <= cap; cells just outside it are INF
  {
    int jhi = cap < ulen ? cap : ulen;
    prev[0] = 0;
    for (int j = 1; j <= jhi; ++j)  // D[0][j]: j insertions before q[0] (ACGT letters only)
      prev[j] = (uint8_t)((m > 0 && prev[j - 1] < INF && base_code(u[j - 1]) < 4) ? prev[j - 1] + 1 : INF);
    if (jhi + 1 <= ulen) prev[jhi + 1] = (uint8_t)INF;
  }
  for (int i = 1; i <= m; ++i) {
    int jlo = i - cap > 0 ? i - cap : 0;
    int jhi = i + cap < ulen ? i + cap : ulen;
    int best = INF;
    for (int j = jlo; j <= jhi; ++j) {
      int v;
      if (j == 0) {
        v = i;  // i deletions
      } else {
        int uc = base_code(u[j - 1]);
        int sub = prev[j - 1];
        if (q[i - 1] != u[j - 1]) sub = (uc < 4) ? sub + 1 : INF;
        int del = prev[j] + 1;
        int ins = (j > jlo && i < m && uc < 4) ? row[j - 1] + 1 : INF;
        v = sub < del ? sub : del;
        if (ins < v) v = ins;
      }
      if (v > INF) v = INF;
      row[j] = (uint8_t)v;
      if (v < best) best = v;
    }
    if (jhi + 1 <= ulen) row[jhi + 1] = (uint8_t)INF;
    if (best > cap) return cap + 1;
    uint8_t* t = prev; prev = row; row = t;
  }
  int r = prev[ulen];
  return r > cap ? cap + 1 : r;
}

// _insert (neighbors.h:29-45) keeps exactly the substring-minimal strings of everything the DFS
// generates (plus the query).  t (length L) is kept iff no proper substring of it is generated.
// Generated strings have length >= m - d, so only those substrings are tried.
DG_HD bool is_minimal(const uint8_t* q, int m, int d, const uint8_t* t, int L, uint8_t* s0, uint8_t* s1) {
  int minlen = m - d;
  if (minlen < 1) minlen = 1;
  for (int len = minlen; len < L; ++len)
    for (int a = 0; a + len <= L; ++a)
      if (restricted_distance(q, m, t + a, len, d, s0, s1) <= d) return false;
  return true;
}

// Register-resident form of restricted_distance / is_minimal for m + d <= 31: both strings as 4-bit
// symbol classes (A C G T = 1..4, N = 5), 16 per 64-bit word, the DP as a band of 2 cap + 1 cells
// (k = j - i + cap) held in scalars.  Same recurrence, same result (tests/hostsim cross-checks the
// two on every generated string).
struct Packed4 {
  uint64_t w0, w1;
  DG_HD uint32_t at(int i) const { return (uint32_t)(((i < 16 ? w0 : w1) >> (4 * (i & 15))) & 15u); }
};
DG_HD Packed4 pack4(const uint8_t* s, int n) {
  Packed4 p;
  p.w0 = p.w1 = 0;
  for (int i = 0; i < n; ++i) {
    uint8_t b = s[i];
    uint64_t c = b == 'A' ? 1u : b == 'C' ? 2u : b == 'G' ? 3u : b == 'T' ? 4u : 5u;
    if (i < 16) p.w0 |= c << (4 * i); else p.w1 |= c << (4 * (i - 16));
  }
  return p;
}
constexpr int kSmallBand = 2 * kMaxDist + 1;

// min(cost, cap + 1) of turning q (length m) into t[a, a + ulen) under the DFS rules
DG_HD int restricted_distance_small(const Packed4& q, int m, const Packed4& t, int a, int ulen, int cap) {
  const int INF = 100;
  const int dl = m - ulen;
  if (dl > cap || -dl > cap) return cap + 1;
  const int W = 2 * cap + 1;
  int v[kSmallBand];
  // row 0: D[0][j] = j insertions before q[0] (ACGT letters only); k = j + cap
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < kSmallBand; ++k) {
    int j = k - cap, val = INF;
    if (k < W && j >= 0 && j <= ulen) {
      val = j;
      for (int x = 0; x < j; ++x) if (t.at(a + x) > 4u || m == 0) val = INF;
    }
    v[k] = val;
  }
  for (int i = 1; i <= m; ++i) {
    const uint32_t qc = q.at(i - 1);
    int best = INF;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < kSmallBand; ++k) {
      if (k < W) {
        const int j = i + k - cap;
        int val = INF;
        if (j == 0) {
          val = i;
        } else if (j > 0 && j <= ulen) {
          const uint32_t uc = t.at(a + j - 1);
          int sub = v[k];                                        // D[i-1][j-1]
          if (qc != uc) sub = uc <= 4u ? sub + 1 : INF;
          const int del = (k + 1 < W ? v[k + 1] : INF) + 1;      // D[i-1][j] + 1
          const int ins = (k > 0 && i < m && uc <= 4u) ? v[k - 1] + 1 : INF;   // D[i][j-1] + 1 (v[k-1] already holds row i)
          val = sub < del ? sub : del;
          if (ins < val) val = ins;
          if (val > INF) val = INF;
        }
        v[k] = val;
        if (val < best) best = val;
      }
    }
    if (best > cap) return cap + 1;
  }
  const int kf = ulen - m + cap;   // j = ulen at i = m
  int r = INF;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < kSmallBand; ++k) if (k == kf) r = v[k];
  return r > cap ? cap + 1 : r;
}
DG_HD bool is_minimal_small(const Packed4& q, int m, int d, const Packed4& t, int L) {
  int minlen = m - d;
  if (minlen < 1) minlen = 1;
  for (int len = minlen; len < L; ++len)
    for (int a = 0; a + len <= L; ++a)
      if (restricted_distance_small(q, m, t, a, len, d) <= d) return false;
  return true;
}

// The same decision with ONE pass over the query instead of one DP per substring: a semi-global form
// of the recurrence above in which the substring may start anywhere in t (row 0 costs nothing at
// any column), so that cell (m, j) holds the smallest distance over all substrings ending at j.
// t is non-minimal iff some substring ending before its last character is within d (v0: any start),
// or one ending at the last character that does not start at the first (v1: row 0 is open from
// column 1 on only).  Only cells with -d <= j - i <= 3 d can stay within d (a substring of length
// >= m - d starts at most 2 d characters into t), a band of 4 d + 1 cells.
constexpr int kWideBand = 4 * kMaxDist + 1;
DG_HD bool is_minimal_band(const Packed4& q, int m, int d, const Packed4& t, int L) {
  const int INF = 100;
  if (L < m - d + 1) return true;          // no proper substring is long enough
  const int W = 4 * d + 1;                 // k = j - i + d
  int v0[kWideBand], v1[kWideBand];
  // row 0: the empty prefix of q matches the empty substring at any start: cost 0 (insertions before
  // q[0] are then never needed: starting later is cheaper)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < kWideBand; ++k) {
    const int j = k - d;
    const bool in = k < W && j >= 0 && j <= L;
    v0[k] = in ? 0 : INF;
    v1[k] = (in && j >= 1) ? 0 : INF;
  }
  for (int i = 1; i <= m; ++i) {
    const uint32_t qc = q.at(i - 1);
    int best = INF;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < kWideBand; ++k) {
      if (k < W) {
        const int j = i + k - d;
        int a0 = INF, a1 = INF;
        if (j == 0) {
          a0 = i;                          // i deletions, substring starting (and ending) at 0
        } else if (j > 0 && j <= L) {
          const uint32_t uc = t.at(j - 1);
          const bool letter = uc <= 4u;
          const int miss = (qc != uc) ? (letter ? 1 : INF) : 0;
          const int up0 = (k + 1 < W ? v0[k + 1] : INF) + 1, up1 = (k + 1 < W ? v1[k + 1] : INF) + 1;   // D[i-1][j] + 1
          const bool can_ins = k > 0 && i < m && letter;
          a0 = v0[k] + miss;
          if (up0 < a0) a0 = up0;
          if (can_ins && v0[k - 1] + 1 < a0) a0 = v0[k - 1] + 1;   // D[i][j-1] + 1 (already row i)
          a1 = v1[k] + miss;
          if (up1 < a1) a1 = up1;
          if (can_ins && v1[k - 1] + 1 < a1) a1 = v1[k - 1] + 1;
          if (a0 > INF) a0 = INF;
          if (a1 > INF) a1 = INF;
        }
        v0[k] = a0;
        v1[k] = a1;
        if (a0 < best) best = a0;
      }
    }
    if (best > d) return true;             // (v1 >= v0 everywhere)
  }
  // row m: substrings ending at j = m + k - d
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < kWideBand; ++k) {
    if (k < W) {
      const int j = m + k - d;
      if (j >= 1 && j < L && v0[k] <= d) return false;
      if (j == L && v1[k] <= d) return false;
    }
  }
  return true;
}

// Distance 1 needs no DP at all: a string of G_1(q) of length m - 1 / m / m + 1 is q with one
// deletion / substitution / insertion, so "some proper substring of t (length >= m - 1) is
// generated" reduces to two predicates on packed strings, each a handful of shifts and XORs:
//   one_deletion(q, u)  |u| = m - 1 and u is q with one base removed (any position)
//   hamming_le1(q, u)   |u| = m and u differs from q in at most one place, the new letter in ACGT
DG_HD Packed4 shr4(const Packed4& p, int a) {   // drop the first a symbols
  Packed4 r;
  if (a >= 16) { r.w0 = a >= 32 ? 0 : p.w1 >> (4 * (a - 16)); r.w1 = 0; }
  else if (a == 0) { r = p; }
  else { r.w0 = (p.w0 >> (4 * a)) | (p.w1 << (64 - 4 * a)); r.w1 = p.w1 >> (4 * a); }
  return r;
}
DG_HD Packed4 keep4(const Packed4& p, int len) {   // first len symbols, the rest cleared
  Packed4 r;
  r.w0 = len >= 16 ? p.w0 : (len <= 0 ? 0 : p.w0 & ((1ULL << (4 * len)) - 1ULL));
  r.w1 = len >= 32 ? p.w1 : (len <= 16 ? 0 : p.w1 & ((1ULL << (4 * (len - 16))) - 1ULL));
  return r;
}
DG_HD int ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)x) - 1;
#else
  return __builtin_ctzll(x);
#endif
}
// index of the first symbol where x and y differ among the first len, or len
DG_HD int first_diff4(const Packed4& x, const Packed4& y, int len) {
  const Packed4 a = keep4(x, len), b = keep4(y, len);
  const uint64_t d0 = a.w0 ^ b.w0, d1 = a.w1 ^ b.w1;
  if (d0) return ctz64(d0) >> 2;
  if (d1) return 16 + (ctz64(d1) >> 2);
  return len;
}
DG_HD bool one_deletion4(const Packed4& q, int m, const Packed4& u) {   // |u| = m - 1
  const int f = first_diff4(q, u, m - 1);
  if (f >= m - 1) return true;                       // u = q without its last base
  return first_diff4(shr4(q, f + 1), shr4(u, f), m - 1 - f) >= m - 1 - f;
}
DG_HD bool hamming_le1_4(const Packed4& q, int m, const Packed4& u) {   // |u| = m
  const int f = first_diff4(q, u, m);
  if (f >= m) return true;                           // u = q
  if (u.at(f) > 4u) return false;                    // only ACGT can be written
  return first_diff4(shr4(q, f + 1), shr4(u, f + 1), m - 1 - f) >= m - 1 - f;
}
DG_HD bool is_minimal_d1(const Packed4& q, int m, const Packed4& t, int L) {
  if (L < m) return true;                            // nothing of length >= m - 1 fits inside
  if (L == m) return !(one_deletion4(q, m, t) || one_deletion4(q, m, shr4(t, 1)));
  // L == m + 1
  if (hamming_le1_4(q, m, t) || hamming_le1_4(q, m, shr4(t, 1))) return false;
  return !(one_deletion4(q, m, t) || one_deletion4(q, m, shr4(t, 1)) || one_deletion4(q, m, shr4(t, 2)));
}

// ------------------------------------------------------------------------------------------
// An upper bound on the number of DISTINCT strings neighbors() generates for q at edit distance
// d <= 2 -- hence on the size its set can ever have, hence a certificate that the cap of
// neighbors.h:50 is never reached when the bound stays below max_neighborhood.  (Queries that miss
// the certificate are replayed exactly on the host, nbr_trunc.hpp.)
// The scripts (one or two events, sorted as in Script) are counted except those a rule below maps
// to a strictly smaller script -- fewer events, else smaller (pos1, kind1, pos2, kind2) with
// substitution < deletion < insertion -- that spells the same string.  The smallest script of a
// string is never discounted, so the count never falls below the number of distinct strings.
//   L(p, k): the event could move one base to the left:  deletion inside a run (q[p] == q[p-1]),
//            insertion of the letter that precedes it (c == q[p-1]).
//   A  L(event 1)                                   -> event 1 one base to the left
//   B  L(event 2), base pos2 - 1 untouched by event 1 -> event 2 one base to the left
//   C  (delete p, insert before p + 1)              -> one substitution at p, or nothing
//   D  (insert before p, delete p)                  -> one substitution at p, or nothing
//   E  (insert c before p, substitute p), p + 1 < m -> (substitute p by c, insert before p + 1), or one insertion
//   F  (delete p, substitute p + 1 by x)            -> (substitute p by x, delete p + 1), or one deletion
// bq(i) returns the base code of q[i] (0..3, 4 = not ACGT).
template <typename BaseAt>
DG_HD bool nbr_event_shiftable(BaseAt bq, int p, int k) {
  if (p == 0 || k < 4) return false;
  if (k == 4) return bq(p) == bq(p - 1);
  return (k - 5) == bq(p - 1);
}
template <typename BaseAt>
DG_HD bool nbr_pair_discounted(BaseAt bq, int m, int p1, int k1, int p2, int k2) {
  if (nbr_event_shiftable(bq, p1, k1)) return true;                                   // A
  if ((p2 > p1 + 1 || (p2 == p1 + 1 && k1 >= 5)) && nbr_event_shiftable(bq, p2, k2)) return true;   // B
  if (k1 == 4 && k2 >= 5 && p2 == p1 + 1) return true;                                // C
  if (k1 >= 5 && k2 == 4 && p2 == p1) return true;                                    // D
  if (k1 >= 5 && k2 < 4 && p2 == p1 && p1 + 1 < m) return true;                       // E
  if (k1 == 4 && k2 < 4 && p2 == p1 + 1) return true;                                 // F
  return false;
}
// events e = pos * 9 + k in [e_lo, e_hi) are taken as FIRST events (so that the lanes of a warp can
// share one query); the unedited string and the single events are counted by the caller that owns
// e_lo == 0.
template <typename BaseAt>
DG_HD uint32_t nbr_upper_bound_part(BaseAt bq, int m, int d, int e_lo, int e_hi, int e_step) {
  uint32_t n = 0;
  const int E = 9 * m;
  for (int e1 = e_lo; e1 < e_hi; e1 += e_step) {
    const int p1 = e1 / 9, k1 = e1 - 9 * p1;
    if (k1 < 4 && k1 == bq(p1)) continue;                        // not an edit
    if (!nbr_event_shiftable(bq, p1, k1)) ++n;                   // the single event
    if (d < 2 || nbr_event_shiftable(bq, p1, k1)) continue;      // (rule A discounts every pair of it)
    for (int e2 = second_event_start(p1, k1, true); e2 < E; ++e2) {
      const int p2 = e2 / 9, k2 = e2 - 9 * p2;
      if (k2 < 4 && k2 == bq(p2)) continue;
      if (!nbr_pair_discounted(bq, m, p1, k1, p2, k2)) ++n;
    }
  }
  return n;
}

// The same count in closed form, O(m): per position p the valid events are nsub(p) substitutions
// (3, or 4 at a non-ACGT base), one deletion and four insertions; of these the deletion is shiftable
// inside a run and one insertion is when the preceding base is a letter.  With c(p) the events that
// are not shiftable (ci / cdel: the insertions / deletion among them) and w(p) all of them:
//   second event two or more bases to the right: rules A and B only         -> c(p1) * sum c(p2)
//   second event at p1 + 1: after an insertion rule B still holds (c), after a substitution nothing
//   applies (w), after a deletion rules C and F leave the deletion alone (1)
//   second event at p1 (the first is an insertion): D drops the deletion, E the substitutions unless
//   p1 is the last base, the four insertions stay.
template <typename BaseAt>
DG_HD uint32_t nbr_upper_bound_closed(BaseAt bq, int m, int d) {
  if (d <= 0) return 1;
  uint64_t total = 1, suffix = 0;      // suffix = sum of c(p2) over p2 >= p + 2 while walking p downwards
  uint32_t c_next = 0, w_next = 0;     // c(p + 1), w(p + 1)
  for (int p = m - 1; p >= 0; --p) {
    const int b = bq(p), prev = p > 0 ? bq(p - 1) : -1;
    const uint32_t nsub = b < 4 ? 3u : 4u;
    const uint32_t cdel = (p > 0 && b == prev) ? 0u : 1u;
    const uint32_t ci = (p > 0 && prev < 4) ? 3u : 4u;
    const uint32_t c = nsub + cdel + ci, w = nsub + 5u;
    total += c;
    if (d >= 2) {
      const bool last = p + 1 >= m;
      total += (uint64_t)c * suffix;                                   // p2 >= p + 2
      total += (uint64_t)ci * c_next + (uint64_t)nsub * w_next + (last ? 0u : cdel);   // p2 == p + 1
      total += (uint64_t)ci * ((last ? nsub : 0u) + 4u);               // p2 == p
    }
    suffix += c_next;
    c_next = c;
    w_next = w;
  }
  return total > 0xFFFFFFFFULL ? 0xFFFFFFFFu : (uint32_t)total;
}

// ------------------------------------------------------------------------------------------
// needle() for std::string x std::string, AlignConfig<false,true>, DnaScore(0,-1,-1,-1)
// (needle.h:59-138, align.h:52-80): rows = genomic g (length mg), columns = query s (length n).
// tr: trace storage (TraceBytes: >= ((mg+1)*(n+1)+3)/4 bytes; TraceRows64: mg+1 words, n <= 31);
// srow: >= n+1 ints; ops: >= mg+n bytes.  Outputs the alignment with leading / trailing
// query-gap columns stripped (hunter.h:391-401), the number of stripped leading columns, and
// the score.  Returns the number of kept columns.
// Trace storage: 2 bits per DP cell (1 = bit3 "left", 2 = bit4 "up", 0 = diagonal).
struct TraceBytes {   // any size: (mg+1)*(n+1) cells packed 4 per byte in caller memory
  uint8_t* t;
  int mf;
  DG_HD void set(int row, int col, int v) {
    int cell = row * mf + col, sh = (cell & 3) * 2;
    t[cell >> 2] = (uint8_t)((t[cell >> 2] & ~(3 << sh)) | (v << sh));
  }
  DG_HD int get(int row, int col) const {
    int cell = row * mf + col;
    return (t[cell >> 2] >> ((cell & 3) * 2)) & 3;
  }
};
struct TraceRows64 {  // n <= 31: one 64-bit word per DP row (thread-local array)
  uint64_t* w;
  DG_HD void set(int row, int col, int v) {
    w[row] = (w[row] & ~(3ULL << (2 * col))) | ((uint64_t)v << (2 * col));
  }
  DG_HD int get(int row, int col) const { return (int)((w[row] >> (2 * col)) & 3); }
};

template <typename Tr>
DG_HD int needle_align(const uint8_t* g, int mg, const uint8_t* s, int n, Tr tr, int* srow,
                       uint8_t* ops, uint8_t* refalign, uint8_t* queryalign, int* lead_out, int* score_out) {
  auto setbits = [&](int row, int col, int v) { tr.set(row, col, v); };
  auto getbits = [&](int row, int col) { return tr.get(row, col); };
  int prevsub = 0;
  for (int row = 0; row <= mg; ++row) {
    for (int col = 0; col <= n; ++col) {
      if (row == 0 && col == 0) {
        srow[0] = 0; prevsub = 0; setbits(row, col, 0);
      } else if (row == 0) {
        srow[col] = -col;            // _horizontalGap(AlignConfig<false,*>) = col * ge
        setbits(row, col, 1);
      } else if (col == 0) {
        srow[0] = 0;                 // _verticalGap(AlignConfig<*,true>, 0, n, ..) = 0
        prevsub = 0;
        setbits(row, col, 2);
      } else {
        int prevprevsub = prevsub;
        prevsub = srow[col];
        int vg = (col == n) ? 0 : -1;  // _verticalGap: free at col == 0 or col == n
        int diag = prevprevsub + (g[row - 1] == s[col - 1] ? 0 : -1);
        int up = prevsub + vg;
        int left = srow[col - 1] - 1;
        int v = diag > up ? diag : up;
        if (left > v) v = left;
        srow[col] = v;
        setbits(row, col, v == left ? 1 : (v == up ? 2 : 0));
      }
    }
  }
  *score_out = srow[n];
  // traceback (needle.h:114-131): bit3 -> 'h', bit4 -> 'v', else 's'
  int row = mg, col = n, nops = 0;
  while (row > 0 || col > 0) {
    int b = getbits(row, col);
    if (b == 1) { --col; ops[nops++] = 'h'; }
    else if (b == 2) { --row; ops[nops++] = 'v'; }
    else { --row; --col; ops[nops++] = 's'; }
  }
  // _createAlignment (align.h:176-203) read left to right = ops back to front.
  // trailing columns whose query row is '-' are dropped (_trailGap, hunter.h:69-77) ...
  int last_aligned = nops - 1;   // column index of the last column with a query character
  {
    int j = 0, found = -1;
    for (int t = nops - 1; t >= 0; --t, ++j) if (ops[t] != 'v') found = j;
    if (found >= 0) last_aligned = found;
  }
  int ncols = last_aligned + 1;
  // ... and leading ones advance chrpos instead of being copied (hunter.h:393-400).
  int r = 0, c = 0, kept = 0, lead = 0;
  bool leadGap = true;
  for (int j = 0, t = nops - 1; j < ncols; ++j, --t) {
    uint8_t a0, a1;
    if (ops[t] == 's') { a0 = g[r++]; a1 = s[c++]; }
    else if (ops[t] == 'h') { a0 = '-'; a1 = s[c++]; }
    else { a0 = g[r++]; a1 = '-'; }
    if (a1 != '-') leadGap = false;
    if (!leadGap) { refalign[kept] = a0; queryalign[kept] = a1; ++kept; }
    else ++lead;
  }
  *lead_out = lead;
  return kept;
}


// ------------------------------------------------------------------------------------------
// Banded form of needle_align for the hunt loop.  In `hunt` the genomic context always contains
// the matched neighbour string, so the optimal score is >= -dmax (dmax = the query's distance).
// Every cell on an optimal full path then satisfies  -dmax <= row - col <= (mg - n) + dmax:
//   * 0 < col < n, score >= -dmax  =>  col - (row - j) <= dmax for some free start j >= 0;
//   * the rest of the path aligns s[col, n) with at most mg - row genomic symbols (trailing ones
//     are free), so (n - col) - (mg - row) <= dmax.
// Cells outside the band are treated as -inf.  Banded values never exceed the true ones and are
// equal on every optimal full path (its cells' optimal predecessors are again such cells), so a
// candidate that loses or ties in the full matrix (needle.h:87-101: horizontal, then vertical, then
// diagonal) loses or ties identically here: score, traceback and alignment are those of needle().
// Band width W = mg - n + 2 dmax + 1 <= kBandMax; trace = 2 bits x W per row in one 32-bit word.
constexpr int kBandMax = 11;   // distance 2: mg - n <= 3 d, W <= 5 d + 1
constexpr int kBandNeg = -1000;

// 4-bit symbol classes for the register-resident query window: equal bytes <=> equal classes for
// everything a normalised query can hold (A C G T N); any other text byte gets a class of its own.
DG_HD uint32_t sym_class(uint8_t b) {
  return b == 'A' ? 1u : b == 'C' ? 2u : b == 'G' ? 3u : b == 'T' ? 4u : b == 'N' ? 5u : 6u;
}

// DP + trace.  tr: mg + 1 words.  Returns the score (srow[n] of needle()).  BAND is the compile-time
// capacity of the register band (W <= BAND): 6 covers distance 1, kBandMax distance 2.
template <int BAND>
DG_HD int needle_banded_fill_t(const uint8_t* g, int mg, const uint8_t* s, int n, int dmax, uint32_t* tr) {
  const int hi = mg - n + dmax;            // largest row - col in the band
  const int W = hi + dmax + 1;
  // query as 4-bit classes, 16 per word (n <= 31)
  uint64_t sq0 = 0, sq1 = 0;
  for (int i = 0; i < n; ++i) {
    uint64_t c = sym_class(s[i]);
    if (i < 16) sq0 |= c << (4 * i); else sq1 |= c << (4 * (i - 16));
  }
  auto qclass = [&](int i) -> uint32_t {  // class of s[i], 0 outside the query
    if (i < 0 || i >= n) return 0u;
    return (uint32_t)(((i < 16 ? sq0 : sq1) >> (4 * (i & 15))) & 15u);
  };
  int v[BAND];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < BAND; ++k) v[k] = kBandNeg;
  // win holds the classes of s[col - 1] for the W columns of the current row: col = row - hi + k
  uint64_t win = 0;
  for (int k = 0; k < W; ++k) win |= (uint64_t)qclass(0 - hi + k - 1) << (4 * k);
  int score = 0;
  // Row 0 (no genomic symbol yet): -col with a horizontal trace.  The rows after it are written
  // without data-dependent branches: every band slot computes the recurrence from its three
  // neighbours (slots outside the matrix hold kBandNeg and stay there) and the special cells
  // (column 0, the free last column) are selected afterwards.
  {
    uint32_t bits = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < BAND; ++k) {
      if (k < W) {
        const int col = k - hi;
        const bool inr = (unsigned)col <= (unsigned)n;
        v[k] = inr ? -col : kBandNeg;
        bits |= (uint32_t)((inr && col) ? 1 : 0) << (2 * k);
        if (col == n) score = -col;
      }
    }
    tr[0] = bits;
    win = (win >> 4) | ((uint64_t)qclass(-hi + W - 1) << (4 * (W - 1)));
  }
  for (int row = 1; row <= mg; ++row) {
    const uint32_t gc = sym_class(g[row - 1]);
    uint32_t bits = 0;
    const int c0 = row - hi;   // column of k = 0
    int left_new = kBandNeg;   // value of the slot to the left in this row
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < BAND; ++k) {
      if (k < W) {
        const int col = c0 + k;
        const bool inr = (unsigned)col <= (unsigned)n;
        const uint32_t qc = (uint32_t)((win >> (4 * k)) & 15u);
        const int diag = v[k] + (gc == qc ? 0 : -1);
        const int up = (k + 1 < W ? v[k + 1] : kBandNeg) + (col == n ? 0 : -1);
        const int left = left_new - 1;
        int val = diag > up ? diag : up;
        if (left > val) val = left;
        int t = val == left ? 1 : (val == up ? 2 : 0);
        if (col == 0) { val = 0; t = 2; }
        if (!inr) { val = kBandNeg; t = 0; }
        if (col == n) score = val;
        v[k] = val;
        left_new = val;
        bits |= (uint32_t)t << (2 * k);
      }
    }
    tr[row] = bits;
    // next row: every column index grows by one
    win = (win >> 4) | ((uint64_t)qclass(c0 + W - 1) << (4 * (W - 1)));
  }
  return score;
}

DG_HD int needle_banded_fill(const uint8_t* g, int mg, const uint8_t* s, int n, int dmax, uint32_t* tr) {
  const int W = mg - n + 2 * dmax + 1;
  if (W <= 6) return needle_banded_fill_t<6>(g, mg, s, n, dmax, tr);
  return needle_banded_fill_t<kBandMax>(g, mg, s, n, dmax, tr);
}

// Traceback, first pass: number of alignment columns, leading and trailing columns whose query
// row is a gap (hunter.h:69-77, 393-400).
DG_HD void needle_banded_shape(const uint32_t* tr, int mg, int n, int dmax, int* nops_out, int* lead_out, int* trail_out) {
  const int hi = mg - n + dmax;
  int row = mg, col = n, nops = 0, trail = 0, run_v = 0;
  bool seen_query = false;
  while (row > 0 || col > 0) {
    const int t = (int)((tr[row] >> (2 * (col - (row - hi)))) & 3u);
    if (t == 1) { --col; seen_query = true; run_v = 0; }
    else if (t == 2) { --row; if (!seen_query) ++trail; ++run_v; }
    else { --row; --col; seen_query = true; run_v = 0; }
    ++nops;
  }
  *nops_out = nops;
  *trail_out = trail;
  *lead_out = seen_query ? run_v : 0;   // the vertical moves made last are the leftmost columns
}

// Traceback, second pass: writes the kept columns (left to right) of both alignment rows.
DG_HD void needle_banded_emit(const uint32_t* tr, const uint8_t* g, int mg, const uint8_t* s, int n, int dmax, int nops,
                              int lead, int trail, uint8_t* refalign, uint8_t* queryalign) {
  const int hi = mg - n + dmax;
  int row = mg, col = n;
  for (int step = 0; step < nops; ++step) {
    const int t = (int)((tr[row] >> (2 * (col - (row - hi)))) & 3u);
    uint8_t a0, a1;
    if (t == 1) { --col; a0 = '-'; a1 = s[col]; }
    else if (t == 2) { --row; a0 = g[row]; a1 = '-'; }
    else { --row; --col; a0 = g[row]; a1 = s[col]; }
    const int j = nops - 1 - step;           // column index from the left
    if (j >= lead && j < nops - trail) { refalign[j - lead] = a0; queryalign[j - lead] = a1; }
  }
}

// ------------------------------------------------------------------------------------------
// Compact alignments (dg_rec of include/dicey_b200.h).  The kept columns of a `hunt` alignment hold
// every query base exactly once, and every column that is not "genomic base == query base" costs
// one unit of the score (leading / trailing query-gap columns are free and stripped, hunter.h:391-401),
// so with score >= -d there are at most d such columns.  A record therefore carries up to kRecOps
// edit operations instead of two strings; both rows are rebuilt on the host from the query:
//   op = column (10 bits) | type << 10 (1 mismatch, 2 gap in the reference row, 3 gap in the query
//        row) | genomic byte << 12; ops ascending by column from bit 0, 20 bits each;
//   bits 60-63: (start - 1) - (text position - record start) + 8, which gives text_pos back.
constexpr int kRecOps = 3;
constexpr int kRecOpBits = 20;
DG_HD uint64_t rec_op(int col, int type, uint8_t ref) {
  return (uint64_t)(col & 1023) | ((uint64_t)type << 10) | ((uint64_t)ref << 12);
}

// Second traceback pass of the banded form: the operations of the kept columns (replaces
// needle_banded_emit when the record is compact).  Returns the packed ops; *nop_out may exceed
// kRecOps (then the packing is meaningless and the caller must fall back).
DG_HD uint64_t needle_banded_ops(const uint32_t* tr, const uint8_t* g, int mg, const uint8_t* s, int n, int dmax, int nops,
                                 int lead, int trail, int* nop_out) {
  const int hi = mg - n + dmax;
  int row = mg, col = n, nop = 0;
  uint64_t bits = 0;
  for (int step = 0; step < nops; ++step) {
    const int t = (int)((tr[row] >> (2 * (col - (row - hi)))) & 3u);
    int type = 0;
    uint8_t ref = 0;
    if (t == 1) { --col; type = 2; }
    else if (t == 2) { --row; type = 3; ref = g[row]; }
    else { --row; --col; if (g[row] != s[col]) { type = 1; ref = g[row]; } }
    const int j = nops - 1 - step;           // column index from the left
    if (type && j >= lead && j < nops - trail) {
      bits = (bits << kRecOpBits) | rec_op(j - lead, type, ref);   // found right to left: the leftmost ends in the low bits
      ++nop;
    }
  }
  *nop_out = nop;
  return bits & 0x0FFFFFFFFFFFFFFFULL;
}

// The same from materialised rows (the full-matrix paths).
DG_HD uint64_t rows_to_ops(const uint8_t* ra, const uint8_t* qa, int kept, int* nop_out) {
  uint64_t bits = 0;
  int nop = 0;
  for (int j = kept - 1; j >= 0; --j) {
    if (ra[j] == qa[j]) continue;
    const int type = ra[j] == '-' ? 2 : (qa[j] == '-' ? 3 : 1);
    bits = (bits << kRecOpBits) | rec_op(j, type, type == 2 ? 0 : ra[j]);
    ++nop;
  }
  *nop_out = nop;
  return bits & 0x0FFFFFFFFFFFFFFFULL;
}

// Both alignment rows from the query (the strand's search string, m bases) and the ops; returns the
// number of columns (m + query-gap columns).  ra / qa need m + kRecOps bytes.
DG_HD int rec_expand_rows(uint64_t ops, int nop, const uint8_t* s, int m, uint8_t* ra, uint8_t* qa) {
  int c = 0, col = 0, k = 0;
  while (c < m || k < nop) {
    const uint64_t op = ops >> (kRecOpBits * k);
    if (k < nop && (int)(op & 1023) == col) {
      const int type = (int)((op >> 10) & 3);
      const uint8_t ref = (uint8_t)((op >> 12) & 255);
      if (type == 1) { ra[col] = ref; qa[col] = s[c++]; }
      else if (type == 2) { ra[col] = '-'; qa[col] = s[c++]; }
      else { ra[col] = ref; qa[col] = '-'; }
      ++k;
    } else {
      if (c >= m) break;   // malformed ops: never loop forever
      ra[col] = qa[col] = s[c++];
    }
    ++col;
  }
  return col;
}

// ------------------------------------------------------------------------------------------
// The pair enumeration of k_probe_pairs / k_resolve (dg_search.cu): for one string of m bases the
// S * m (m + 1) / 2 slots are laid out by first position p1 (S * (m - p1) slots each), then first kind,
// then second position p2 = p1 .. m - 1.
DG_HD float dg_sqrtf(float x) {
#if defined(__CUDA_ARCH__)
  return sqrtf(x);
#else
  return __builtin_sqrtf(x);
#endif
}
// slot of the pair enumeration -> (first position, first kind, second position)
template <int S>
DG_HD void pair_slot(uint32_t u, int m, int& p1, int& k1i, int& p2) {
  // rows of S * (m - p1) slots; T(p1) = p1 m - p1 (p1 - 1) / 2 rows-of-S precede first position p1
  const uint32_t v = u / (uint32_t)S;
  const float b2 = (float)(2 * m + 1);
  int g = (int)((b2 - dg_sqrtf(b2 * b2 - 8.0f * (float)v)) * 0.5f);
  if (g < 0) g = 0;
  if (g > m - 1) g = m - 1;
  auto T = [&](int x) { return (uint32_t)(x * m - (x * (x - 1)) / 2); };
  while (g > 0 && T(g) > v) --g;
  while (g + 1 < m && T(g + 1) <= v) ++g;
  p1 = g;
  const uint32_t r = u - (uint32_t)S * T(g);
  const uint32_t npos = (uint32_t)(m - g);
  k1i = (int)(r / npos);
  p2 = g + (int)(r - (uint32_t)k1i * npos);
}

// ------------------------------------------------------------------------------------------
// Packed fast path: an ACGT-only string of length <= 31 as 2-bit codes, LAST base in the low
// bits (so the low 2K bits are the K-mer table index and base t from the right is bits 2t..2t+1).
// apply_event_packed applies one canonical event (k < 4 substitute code k, 4 delete, 5.. insert
// code k-5 before the base) at right-based index j of the base it refers to.  Two events are
// applied left one first: an edit never moves anything to its right.
DG_HD uint64_t apply_event_packed(uint64_t code, int j, int k) {
  int sh = 2 * j;
  if (k < 4) return (code & ~(3ULL << sh)) | ((uint64_t)k << sh);
  uint64_t low_excl = code & ((1ULL << sh) - 1);          // bases right of j
  if (k == 4) return (((code >> sh) >> 2) << sh) | low_excl;
  uint64_t low_incl = code & ((1ULL << (sh + 2)) - 1);     // base j and everything right of it
  return (((code >> sh) >> 2) << (sh + 4)) | ((uint64_t)(k - 5) << (sh + 2)) | low_incl;
}
constexpr int kMaxPacked = 31;  // longest edited string the packed path handles

// Bit address of a KB-mer window (packed, last base in the low bits) in the presence bitmap.
// Random probes are bounded by DRAM row activations, not bytes, so addresses are arranged for
// siblings to land together: the region is selected by the LAST KB - 7 bases and the bit inside it
// by the 7 bases before them (2 KB: 16 lines of one DRAM row; the 4 outermost bases stay within
// one 32-byte sector).  Every neighbour string whose edits lie left of its last KB - 7 bases
// (right-anchored, so indels there shift nothing) probes the same region as its siblings.
// Measured on the headline workload (search stage, ms): 5 bases 3.20, 6: 3.08, 7: 2.92, 8: 3.02,
// 10: 3.18, 12: 3.50.
#ifndef DG_BP
#define DG_BP 7
#endif
constexpr int kPresenceBitBases = DG_BP;   // bases that select the bit inside a region
DG_HD int presence_bit_bases(int KB) { return KB - 2 < kPresenceBitBases ? KB - 2 : kPresenceBitBases; }
DG_HD uint64_t presence_bit(uint64_t window, int KB) {
  const int bp = presence_bit_bases(KB), lo = 2 * (KB - bp);
  return ((window & ((1ULL << lo) - 1ULL)) << (2 * bp)) | (window >> lo);
}

// The mirror-image addressing: the region is selected by the FIRST KB - 7 bases of the window and
// the bit by the 7 bases after them (with the first base in the high bits that is the packed code
// itself).  A string whose edits all lie right of its first KB - 7 bases probes, with its first
// KB bases, the same region as its siblings; between the two layouts only edits in the middle of
// the string (5 of the 19 window positions of a 20-mer) still open a DRAM row of their own.
DG_HD uint64_t presence_bit_left(uint64_t window, int KB) { (void)KB; return window; }

// hunter.h:358-362 / silica.h:475-479: text position -> (refIndex, chrpos).
DG_HD void locate_record(const uint64_t* cum, uint32_t nseq, uint64_t pos, uint32_t& refIndex, uint32_t& chrpos) {
  if (nseq == 0) { refIndex = 0; chrpos = (uint32_t)pos; return; }
  // largest r with cum[r] <= pos, clamped to the last record
  uint32_t r = upper_bound_u64(cum, 0, nseq + 1, pos);  // first cum[idx] > pos
  r = r ? r - 1 : 0;
  if (r > nseq - 1) r = nseq - 1;
  refIndex = r;
  chrpos = (uint32_t)((int64_t)pos - (int64_t)cum[r]);
}

}  // namespace dg
