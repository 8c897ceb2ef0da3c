// dg_search.cu -- the batched hot path: the per-query loop of `dicey hunt` (reference
// src/hunter.h:289-433) and the FM / NW part of `dicey search` (src/silica.h:449-573) for a whole
// batch of queries at once.
//
//   k_prepare   upper-case, non-ACGT -> 'N', reverse complement (hunter.h:306-309,
//               util.h:54-114,208-219), per-query distance clamp (hunter.h:312-315)
//   k_search    neighbors() x sdsl::count (neighbors.h:47-92, suffix_array_algorithm.hpp:447-454):
//               every edit script of every query on both strands is one lane; the last K bases
//               come from the K-mer interval table, the rest are backward-search steps on the
//               32-byte occurrence blocks; scripts whose interval survives become candidates
//   k_minimal   the antichain rule of _insert (neighbors.h:29-45) applied to the survivors
//   sort/unique std::set<std::string> iteration order (hunter.h:350) and de-duplication
//   k_take      the hit budget of hunter.h:350,357 (hits < max_locations)
//   k_locate    sdsl::locate (suffix_array_algorithm.hpp:521-535) + std::sort (hunter.h:356)
//   k_verify    hunter.h:358-432: record lookup, context with '\n' trimming, needle() /
//               needleScore(), gap stripping, DnaHit
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cub/cub.cuh>
#include <map>

#include "dg_common.cuh"

namespace dg {

struct Cand {
  uint32_t q, l, r, code;
};

struct BatchDev {
  const uint8_t* fwd;
  const uint8_t* rc;
  const uint64_t* off;
  uint32_t nq;
  uint32_t* status;
  uint32_t* dist;
  uint32_t seed_len, distance, max_loc, max_nbr;
  uint8_t indel, reverse;
};

DG_HD void query_geom(const BatchDev& b, uint32_t q, int strand, const uint8_t*& base, int& m, int& koff) {
  uint64_t o = b.off[q];
  int L = (int)(b.off[q + 1] - o);
  m = b.seed_len ? (int)b.seed_len : L;
  koff = L - m;
  base = strand == 0 ? b.fwd + o + koff : b.rc + o;
}

// Per-length unit tables: a "unit" is a group of <= 32 consecutive scripts run by one warp.
struct UnitTabs {
  const uint32_t* tab;      // packed (row + 1) << 12 | first   (row 0 = "no first event")
  const uint32_t* tab_off;  // 256 entries
  const uint32_t* tab_cnt;  // 256 entries: units per strand for query length m
  const uint32_t* script_ub; // 256 entries: scripts per strand if every slot were valid (saturating)
};

namespace {

template <typename T>
struct ABuf {  // stream-ordered allocation (cudaMallocAsync pool; reuse across batches is cheap)
  T* p = nullptr;
  size_t count = 0;
  cudaStream_t st = nullptr;
  ABuf() = default;
  ABuf(const ABuf&) = delete;
  ABuf& operator=(const ABuf&) = delete;
  ~ABuf() { release(); }
  void alloc(size_t n, cudaStream_t s) {
    release();
    st = s;
    count = n;
    DG_CUDA(cudaMallocAsync((void**)&p, (n ? n : 1) * sizeof(T), s));
  }
  void release() {
    if (p) cudaFreeAsync(p, st);
    p = nullptr;
    count = 0;
  }
};

inline unsigned grid_for(uint64_t items, unsigned block) { return (unsigned)((items + block - 1) / block); }

__device__ __forceinline__ uint8_t norm_base(uint8_t ch) {
  if (ch >= 'a' && ch <= 'z') ch -= 32;
  return (ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T') ? ch : (uint8_t)'N';
}
__device__ __forceinline__ uint8_t comp_base(uint8_t ch) {
  return ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : ch == 'T' ? 'A' : 'N';
}

__global__ void k_prepare(const uint8_t* __restrict__ raw, BatchDev b, uint8_t* __restrict__ fwd, uint8_t* __restrict__ rc,
                          UnitTabs ut, uint64_t* __restrict__ units) {
  uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= b.nq) return;
  uint64_t o = b.off[q];
  int L = (int)(b.off[q + 1] - o);
  for (int i = 0; i < L; ++i) {
    uint8_t ch = norm_base(raw[o + i]);
    fwd[o + i] = ch;
    rc[o + L - 1 - i] = comp_base(ch);
  }
  uint32_t st = 0;
  int m = b.seed_len ? (int)b.seed_len : L;
  uint32_t d = b.distance;
  if (b.seed_len) {
    if (L <= (int)b.seed_len) st |= DG_Q_SKIPPED;               // silica.h:363,388
  } else {
    if (L < 10) st |= DG_Q_TOO_SHORT;                          // hunter.h:299-303
  }
  if (L > kMaxQuery) st |= DG_Q_UNSUPPORTED;
  if (!(st & (DG_Q_SKIPPED | DG_Q_TOO_SHORT)) && d >= (uint32_t)L) {  // hunter.h:312-315, silica.h:376-379
    d = (uint32_t)L - 1;
    st |= DG_Q_DIST_ADJUSTED;
  }
  if (d > (uint32_t)kMaxDist || (b.seed_len && d >= b.seed_len)) st |= DG_Q_UNSUPPORTED;
  bool run = !(st & (DG_Q_SKIPPED | DG_Q_TOO_SHORT | DG_Q_UNSUPPORTED));
  if (run && m <= kMaxQuery) {
    // neighbors.h:50 stops the DFS once the set holds max_neighborhood strings.  The set never
    // holds more strings than scripts were generated, so fewer scripts than the cap certifies an
    // untruncated neighbourhood.  Hamming sets hold exactly one string per script.
    if (b.indel) {
      if (ut.script_ub[m] >= b.max_nbr) st |= DG_Q_NBR_UNVERIFIED;
    } else {
      uint64_t w = 0, w2 = 0;  // per-position substitution choices: 3, or 4 at an 'N'
      const uint8_t* s0 = fwd + o + (L - m);
      for (int i = 0; i < m; ++i) { uint64_t c = base_code(s0[i]) < 4 ? 3 : 4; w += c; w2 += c * c; }
      uint64_t size = 1 + (d >= 1 ? w : 0) + (d >= 2 ? (w * w - w2) / 2 : 0);
      if (size >= b.max_nbr) st |= DG_Q_NBR_CAP;
    }
  }
  b.status[q] = st;
  b.dist[q] = d;
  units[q] = run ? (uint64_t)ut.tab_cnt[m] * (b.reverse ? 2 : 1) : 0;
}

// ------------------------------------------------------------------------------------------
// k_search: one lane = one edit script.  Persistent warps stride over the units of the batch.
struct SearchOut {
  Cand* cands;
  uint32_t cap;
  unsigned int* n_cand;       // atomic cursor
  unsigned int* overflow;
  unsigned long long* n_scripts;
};

__global__ void __launch_bounds__(256) k_search(IndexView ix, BatchDev b, UnitTabs ut, const uint64_t* __restrict__ unit_off,
                                                uint64_t uniform_units, SearchOut out) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t total = unit_off[b.nq];
  const bool indel = b.indel != 0;
  const int nstrand = b.reverse ? 2 : 1;
  unsigned long long my_scripts = 0;
  for (uint64_t unit = warp; unit < total; unit += nwarps) {
    // unit -> (query, strand, local unit)
    uint32_t q;
    uint64_t local;
    if (uniform_units) {
      q = (uint32_t)(unit / uniform_units);
      local = unit - (uint64_t)q * uniform_units;
    } else {
      uint32_t lo = 0, hi = b.nq;  // last q with unit_off[q] <= unit
      while (hi - lo > 1) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (unit_off[mid] <= unit) lo = mid; else hi = mid;
      }
      q = lo;
      local = unit - unit_off[q];
    }
    const uint8_t* base;
    int m, koff;
    query_geom(b, q, 0, base, m, koff);
    uint32_t per_strand = ut.tab_cnt[m];
    int strand = (int)(local / per_strand);
    uint32_t u = (uint32_t)(local - (uint64_t)strand * per_strand);
    if (strand) query_geom(b, q, 1, base, m, koff);
    (void)nstrand;
    uint32_t packed = ut.tab[ut.tab_off[m] + u];
    int row = (int)(packed >> 12);          // 0 = no first event
    int idx = (int)(packed & 0xFFF) + (int)lane;
    const int dq = (int)b.dist[q];
    const int E = slots_per_pos(indel) * m;
    Script sc;
    sc.nev = 0; sc.pos[0] = sc.pos[1] = 0; sc.k[0] = sc.k[1] = 0;
    bool valid = true;
    int e1 = 0, e2 = 0;
    if (row == 0) {
      // index 0 = the unedited string, index s >= 1 = single event s-1
      if (idx > 0) {
        e1 = idx - 1;
        valid = dq >= 1 && e1 < E && decode_event(base, m, indel, e1, sc.pos[0], sc.k[0]);
        sc.nev = 1;
      }
    } else {
      e1 = row - 1;
      e2 = idx;
      valid = dq >= 2 && e2 < E && decode_event(base, m, indel, e1, sc.pos[0], sc.k[0]) &&
              decode_event(base, m, indel, e2, sc.pos[1], sc.k[1]) && pair_ok(sc.pos[0], sc.k[0], sc.pos[1]);
      sc.nev = 2;
    }
    if (!valid) continue;
    ++my_scripts;
    const int L = script_len(m, sc);
    if (L <= 0) continue;
    const int K = (int)ix.K;
    bool collecting = L >= K;
    int cnt = 0;
    uint32_t kc = 0;
    uint32_t l = 0, r = (uint32_t)ix.n;
    bool alive = script_rtl(base, m, sc, [&](uint8_t x) -> bool {
      if (collecting) {
        int c = base_code(x);
        if (c < 4) {
          kc |= (uint32_t)c << (2 * cnt);
          if (++cnt == K) {
            uint2 iv = __ldg(&ix.kmer[kc]);
            l = iv.x; r = iv.y;
            collecting = false;
            return l < r;
          }
          return true;
        }
        // a non-ACGT letter inside the last K: replay what was collected, then step normally
        collecting = false;
        for (int t = 0; t < cnt; ++t) {
          backward_step(ix, l, r, code_base((int)((kc >> (2 * t)) & 3)));
          if (l >= r) return false;
        }
      }
      backward_step(ix, l, r, x);
      return l < r;
    });
    if (alive && l < r) {
      unsigned int slot = atomicAdd(out.n_cand, 1u);
      if (slot < out.cap) {
        Cand c;
        c.q = q; c.l = l; c.r = r; c.code = pack_script(strand, sc.nev, e1, e2);
        out.cands[slot] = c;
      } else {
        atomicExch(out.overflow, 1u);
      }
    }
  }
  // one statistics update per warp
  for (int o = 16; o; o >>= 1) my_scripts += __shfl_down_sync(0xFFFFFFFFu, my_scripts, o);
  if (lane == 0 && my_scripts) atomicAdd(out.n_scripts, my_scripts);
}

// ------------------------------------------------------------------------------------------
// left-to-right generator of an edited string (no materialisation)
struct LtrIter {
  const uint8_t* base;
  int m;
  Script sc;
  int p, ev;
  __device__ __forceinline__ void init(const uint8_t* b, int mm, const Script& s) { base = b; m = mm; sc = s; p = 0; ev = 0; }
  __device__ __forceinline__ int next() {
    while (p < m) {
      if (ev < sc.nev && sc.pos[ev] == p && sc.k[ev] >= 5) { int c = code_base(sc.k[ev] - 5); ++ev; return c; }
      uint8_t x = base[p];
      bool emit = true;
      if (ev < sc.nev && sc.pos[ev] == p) {
        if (sc.k[ev] == 4) emit = false; else x = code_base(sc.k[ev]);
        ++ev;
      }
      ++p;
      if (emit) return x;
    }
    return -1;
  }
};

// three-way comparison of the strings of two candidates of the same query: std::string order
__device__ int cand_string_cmp(const BatchDev& b, const Cand& x, const Cand& y) {
  bool indel = b.indel != 0;
  int sx, sy;
  Script scx, scy;
  unpack_script(x.code, indel, sx, scx);
  unpack_script(y.code, indel, sy, scy);
  const uint8_t *bx, *by;
  int mx, my, kx, ky;
  query_geom(b, x.q, sx, bx, mx, kx);
  query_geom(b, y.q, sy, by, my, ky);
  LtrIter ix, iy;
  ix.init(bx, mx, scx);
  iy.init(by, my, scy);
  for (;;) {
    int cx = ix.next(), cy = iy.next();
    if (cx != cy) return cx < cy ? -1 : 1;  // -1 (end) sorts first: shorter prefix first
    if (cx < 0) return 0;
  }
}

struct CandLess {
  BatchDev b;
  __device__ bool operator()(const Cand& x, const Cand& y) const {
    if (x.q != y.q) return x.q < y.q;
    int sx = x.code & 1, sy = y.code & 1;
    if (sx != sy) return sx < sy;
    int c = cand_string_cmp(b, x, y);
    if (c) return c < 0;
    return x.code < y.code;
  }
};

__global__ void k_minimal(BatchDev b, const Cand* __restrict__ cands, uint32_t n, uint8_t* __restrict__ keep) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Cand c = cands[i];
  int strand;
  Script sc;
  unpack_script(c.code, true, strand, sc);
  const uint8_t* base;
  int m, koff;
  query_geom(b, c.q, strand, base, m, koff);
  uint8_t t[kMaxQuery + 8], s0[kMaxQuery + 8], s1[kMaxQuery + 8];
  int L = script_ltr(base, m, sc, t);
  keep[i] = is_minimal(base, m, (int)b.dist[c.q], t, L, s0, s1) ? 1 : 0;
}

__global__ void k_unique(BatchDev b, const Cand* __restrict__ cands, uint32_t n, uint8_t* __restrict__ keep) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool first = true;
  if (i > 0) {
    Cand a = cands[i - 1], c = cands[i];
    if (a.q == c.q && ((a.code ^ c.code) & 1) == 0 && cand_string_cmp(b, a, c) == 0) first = false;
  }
  keep[i] = first ? 1 : 0;
}

// hit budget: take[i] = min(occ, max_loc) per candidate
__global__ void k_take_in(const Cand* __restrict__ cands, uint32_t n, uint32_t max_loc, uint64_t* __restrict__ take,
                          uint32_t* __restrict__ qkey) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Cand c = cands[i];
  uint32_t occ = c.r - c.l;
  take[i] = occ < max_loc ? occ : max_loc;
  qkey[i] = c.q;
}
// before[i] = sum of take over earlier candidates of the same query -> hits taken / rows located
__global__ void k_take_out(const Cand* __restrict__ cands, uint32_t n, uint32_t max_loc, const uint64_t* __restrict__ before,
                           const uint64_t* __restrict__ take, uint64_t* __restrict__ ntake, uint64_t* __restrict__ nloc,
                           unsigned long long* __restrict__ qhits, uint32_t* __restrict__ status) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Cand c = cands[i];
  uint64_t bf = before[i], tk = take[i];
  uint64_t nt = bf < max_loc ? (tk < max_loc - bf ? tk : max_loc - bf) : 0;
  ntake[i] = nt;
  nloc[i] = nt ? (uint64_t)(c.r - c.l) : 0;
  if (nt) atomicAdd(&qhits[c.q], (unsigned long long)nt);
  if (bf + tk >= max_loc) atomicOr(&status[c.q], (uint32_t)DG_Q_HIT_CAP);  // hunter.h:434-437
}

__global__ void k_locate(IndexView ix, const Cand* __restrict__ cands, uint32_t n, const uint64_t* __restrict__ loc_off,
                         uint64_t total, uint64_t* __restrict__ keys) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  // candidate i with loc_off[i] <= t < loc_off[i+1]
  uint32_t lo = 0, hi = n;
  while (hi - lo > 1) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (loc_off[mid] <= t) lo = mid; else hi = mid;
  }
  // skip candidates with zero rows that share the offset
  Cand c = cands[lo];
  uint32_t row = c.l + (uint32_t)(t - loc_off[lo]);
  uint32_t pos = sa_value(ix, row);
  keys[t] = ((uint64_t)lo << 32) | pos;
}

struct VerifyArgs {
  const Cand* cands;
  uint32_t ncand;
  const uint64_t* hit_off;   // per candidate (exclusive scan of ntake), ncand + 1 entries
  const uint64_t* loc_off;   // per candidate, ncand + 1 entries
  const uint64_t* keys;      // sorted (candidate << 32 | position)
  uint64_t nhits;
  dg_hit* hits;
  uint8_t* pool;
  uint32_t pool_stride;      // bytes per hit in the pool
  uint8_t* scratch;          // per-thread NW scratch
  uint32_t scratch_stride;
  uint32_t trace_bytes, srow_ints;
  uint64_t first_hit;        // chunk start
  uint64_t chunk;            // hits in this launch
};

__global__ void k_verify(IndexView ix, BatchDev b, VerifyArgs a) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.chunk) return;
  uint64_t h = a.first_hit + t;
  uint32_t lo = 0, hi = a.ncand;
  while (hi - lo > 1) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (a.hit_off[mid] <= h) lo = mid; else hi = mid;
  }
  Cand c = a.cands[lo];
  uint64_t j = h - a.hit_off[lo];
  uint64_t pos = a.keys[a.loc_off[lo] + j] & 0xFFFFFFFFULL;
  const bool indel = b.indel != 0;
  int strand;
  Script sc;
  unpack_script(c.code, indel, strand, sc);
  const uint8_t* base;
  int mq, koff;
  query_geom(b, c.q, strand, base, mq, koff);
  const int m = script_len(mq, sc);             // neighbour length
  const int d = (int)b.dist[c.q];
  // hunter.h:358-362
  uint32_t refIndex, chrpos;
  locate_record(ix.cum, ix.nseq, pos, refIndex, chrpos);
  // context (hunter.h:318-323,363-378; silica.h:480-497)
  uint64_t pre_extract = indel ? d : 0, post_extract = indel ? d : 0;
  if (b.seed_len) { if (strand) post_extract += koff; else pre_extract += koff; }
  if (pre_extract > pos) pre_extract = pos;
  if (pos + m + post_extract > ix.n) post_extract = ix.n - pos - m;
  const uint8_t* T = ix.text;
  uint64_t pre = 0;
  while (pre < pre_extract && T[pos - 1 - pre] != '\n') ++pre;   // keep what follows the last '\n'
  uint64_t post = 0;
  while (post < post_extract && T[pos + m + post] != '\n') ++post;
  const uint8_t* g = T + pos - pre;
  const int mg = (int)(pre + m + post);
  if (b.seed_len ? (pre <= chrpos) : (pre < chrpos)) chrpos -= (uint32_t)pre;  // silica.h:501 / hunter.h:382
  dg_hit out;
  memset(&out, 0, sizeof(out));
  out.query = c.q;
  out.chr = refIndex;
  out.text_pos = pos;
  out.strand = strand ? '-' : '+';
  out.aln_off = h * a.pool_stride;
  uint8_t* slot = a.pool + out.aln_off;
  if (b.seed_len) {
    // search: genomic context for the Tm gate + alignpos (silica.h:522-532)
    uint8_t* scr = a.scratch + t * (uint64_t)a.scratch_stride;
    int* srow = (int*)scr;
    uint8_t* trace = scr + a.srow_ints * 4;
    uint8_t* ops = trace + a.trace_bytes;
    uint8_t* ra = ops + (mg + mq + 4);
    uint8_t* qa = ra + (mg + mq + 4);
    int lead = 0, score = 0;
    needle_align(g, mg, base, mq, trace, srow, ops, ra, qa, &lead, &score);
    for (int i = 0; i < mg; ++i) slot[i] = g[i];
    out.aln_len = (uint32_t)mg;
    out.score = score;
    out.start = chrpos;
    out.alignpos = chrpos + (uint32_t)lead;
  } else if (indel) {
    uint8_t* scr = a.scratch + t * (uint64_t)a.scratch_stride;
    int* srow = (int*)scr;
    uint8_t* trace = scr + a.srow_ints * 4;
    uint8_t* ops = trace + a.trace_bytes;
    uint8_t* ra = ops + (mg + mq + 4);
    uint8_t* qa = ra + (mg + mq + 4);
    int lead = 0, score = 0;
    int kept = needle_align(g, mg, base, mq, trace, srow, ops, ra, qa, &lead, &score);
    for (int i = 0; i < kept; ++i) { slot[i] = ra[i]; slot[kept + i] = qa[i]; }
    out.aln_len = (uint32_t)kept;
    out.score = score;
    out.start = chrpos + (uint32_t)lead + 1;  // hunter.h:399,402
    out.alignpos = out.start;
  } else {
    // needleScore (hunter.h:79-88): mismatches over min(|genomic|, |query|)
    int score = 0;
    int lim = mg < mq ? mg : mq;
    for (int i = 0; i < lim; ++i) if (g[i] != base[i]) --score;
    for (int i = 0; i < mg; ++i) slot[i] = g[i];
    // Hamming records carry refalign = genomicseq and queryalign = the query; both have length
    // mq here (no context in Hamming mode, neighbour length = query length)
    for (int i = 0; i < mq; ++i) slot[mg + i] = base[i];
    out.aln_len = (uint32_t)mg;
    out.score = score;
    out.start = chrpos + 1;
    out.alignpos = out.start;
  }
  a.hits[h] = out;
}

__global__ void k_count(const Cand* __restrict__ cands, uint32_t n, unsigned long long* __restrict__ counts) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Cand c = cands[i];
  atomicAdd(&counts[c.q], (unsigned long long)(c.r - c.l));
}

// exact backward search of literal patterns (one thread per pattern)
__global__ void k_backward_search(IndexView ix, const uint8_t* __restrict__ seqs, const uint64_t* __restrict__ off, uint32_t nq,
                                  uint64_t* __restrict__ lout, uint64_t* __restrict__ rout) {
  uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  uint64_t o = off[q];
  int L = (int)(off[q + 1] - o);
  uint32_t l = 0, r = (uint32_t)ix.n;
  for (int i = L - 1; i >= 0 && l < r; --i) backward_step(ix, l, r, seqs[o + i]);
  if (l < r) { lout[q] = l; rout[q] = (uint64_t)r - 1; }
  else { lout[q] = l; rout[q] = (uint64_t)l - 1; }  // SDSL reports an empty interval as r = l - 1 (r + 1 - l == 0)
}

// unit table of one query length (host)
void build_unit_table(int m, int d, bool indel, std::vector<uint32_t>& tab) {
  int E = slots_per_pos(indel) * m;
  tab.clear();
  int n0 = 1 + (d >= 1 ? E : 0);  // row 0: the unedited string + the single events
  for (int s = 0; s < n0; s += 32) tab.push_back((0u << 12) | (uint32_t)s);
  if (d >= 2) {
    int sl = slots_per_pos(indel);
    for (int e1 = 0; e1 < E; ++e1) {
      int p1 = e1 / sl, k1 = e1 - p1 * sl;
      int start = second_event_start(p1, k1, indel);
      for (int s = start; s < E; s += 32) tab.push_back(((uint32_t)(e1 + 1) << 12) | (uint32_t)s);
    }
  }
}

}  // namespace

}  // namespace dg

using namespace dg;

// ============================================================================================
struct dg_result {
  std::vector<dg_hit> hits;
  std::vector<uint64_t> qoff;
  std::vector<uint32_t> status, dist;
  std::vector<char> pool;
  std::vector<char> seqs;
};

struct dg_batch {
  dg_index* ix = nullptr;
  dg_params par;
  uint32_t nq = 0;
  uint64_t nbytes = 0;
  bool counts_only = false;
  // inputs
  ABuf<uint8_t> raw, fwd, rc;
  ABuf<uint64_t> off, units, unit_off;
  ABuf<uint32_t> status, dist, tab, tab_off, tab_cnt, script_ub;
  uint64_t uniform_units = 0;
  int max_len = 0, min_len = 0;
  // outputs of run()
  ABuf<Cand> cands;
  uint32_t ncand = 0;
  ABuf<dg_hit> hits;
  ABuf<uint8_t> pool;
  ABuf<unsigned long long> qhits;
  ABuf<uint64_t> qoff;
  ABuf<unsigned long long> counts;
  uint64_t nhits = 0;
  uint32_t pool_stride = 0;
  bool ran = false;
};

static BatchDev batch_dev(const dg_batch* b) {
  BatchDev d;
  d.fwd = b->fwd.p; d.rc = b->rc.p; d.off = b->off.p; d.nq = b->nq; d.status = b->status.p; d.dist = b->dist.p;
  d.seed_len = b->par.seed_len; d.distance = b->par.distance; d.max_loc = b->par.max_locations;
  d.max_nbr = b->par.max_neighborhood ? b->par.max_neighborhood : 10000;
  d.indel = b->par.indel; d.reverse = b->par.reverse;
  return d;
}

static int stage_impl(dg_index* ix, const char* seqs, const uint64_t* offsets, uint32_t nq, const dg_params* par,
                      dg_batch** out) {
  if (!ix || !offsets || !par || !out || (nq && !seqs)) { set_error("null argument"); return DG_ERR_ARG; }
  if (par->distance > (uint32_t)kMaxDist) {
    set_error("distance > 2 is outside the device path (DESIGN.md, Limits)");
    return DG_ERR_UNSUPPORTED;
  }
  if (par->seed_len > (uint32_t)kMaxQuery || (par->seed_len && par->distance >= par->seed_len)) {
    set_error("bad seed length");
    return DG_ERR_ARG;
  }
  if (offsets[0] != 0) { set_error("offsets[0] must be 0"); return DG_ERR_ARG; }
  dg_batch* b = new dg_batch();
  try {
    DG_CUDA(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    b->ix = ix;
    b->par = *par;
    if (b->par.max_locations == 0) b->par.max_locations = 1;
    b->nq = nq;
    b->nbytes = offsets[nq];
    // distinct search-string lengths -> unit tables
    bool have[256];
    memset(have, 0, sizeof(have));
    int minL = 1 << 30, maxL = 0;
    for (uint32_t q = 0; q < nq; ++q) {
      if (offsets[q + 1] < offsets[q]) { set_error("offsets must be non-decreasing"); delete b; return DG_ERR_ARG; }
      uint64_t L = offsets[q + 1] - offsets[q];
      int Li = L > 100000 ? 100000 : (int)L;
      minL = std::min(minL, Li);
      maxL = std::max(maxL, Li);
      if (!par->seed_len && L <= (uint64_t)kMaxQuery) have[L] = true;
    }
    if (par->seed_len) have[par->seed_len] = true;
    if (nq == 0) { minL = maxL = 0; }
    b->min_len = minL;
    b->max_len = maxL;
    std::vector<uint32_t> tab, tab_off(256, 0), tab_cnt(256, 0), sub(256, 0), one;
    for (int m = 1; m < 256; ++m) {
      if (!have[m]) continue;
      int d = std::min<int>((int)par->distance, m - 1);
      build_unit_table(m, d, par->indel != 0, one);
      tab_off[m] = (uint32_t)tab.size();
      tab_cnt[m] = (uint32_t)one.size();
      {
        int sl = slots_per_pos(par->indel != 0), E = sl * m;
        uint64_t ub = 1 + (d >= 1 ? (uint64_t)E : 0);
        if (d >= 2)
          for (int e1 = 0; e1 < E; ++e1) ub += (uint64_t)(E - second_event_start(e1 / sl, e1 % sl, par->indel != 0));
        sub[m] = ub > 0xFFFFFFFFULL ? 0xFFFFFFFFu : (uint32_t)ub;
      }
      tab.insert(tab.end(), one.begin(), one.end());
    }
    b->tab.alloc(tab.size(), st);
    b->tab_off.alloc(256, st);
    b->tab_cnt.alloc(256, st);
    b->script_ub.alloc(256, st);
    DG_CUDA(cudaMemcpyAsync(b->script_ub.p, sub.data(), 256 * 4, cudaMemcpyHostToDevice, st));
    DG_CUDA(cudaMemcpyAsync(b->tab.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, st));
    DG_CUDA(cudaMemcpyAsync(b->tab_off.p, tab_off.data(), 256 * 4, cudaMemcpyHostToDevice, st));
    DG_CUDA(cudaMemcpyAsync(b->tab_cnt.p, tab_cnt.data(), 256 * 4, cudaMemcpyHostToDevice, st));
    // uniform batches map unit -> query by a division instead of a binary search
    bool uniform = nq > 0 && minL == maxL && maxL <= kMaxQuery &&
                   (par->seed_len ? (maxL > (int)par->seed_len) : (maxL >= 10)) && par->distance < (uint32_t)maxL;
    if (uniform) {
      int m = par->seed_len ? (int)par->seed_len : maxL;
      b->uniform_units = (uint64_t)tab_cnt[m] * (par->reverse ? 2 : 1);
    }
    b->raw.alloc(b->nbytes + 1, st);
    b->fwd.alloc(b->nbytes + 1, st);
    b->rc.alloc(b->nbytes + 1, st);
    b->off.alloc((size_t)nq + 1, st);
    b->units.alloc((size_t)nq + 1, st);
    b->unit_off.alloc((size_t)nq + 1, st);
    b->status.alloc(nq, st);
    b->dist.alloc(nq, st);
    DG_CUDA(cudaMemcpyAsync(b->raw.p, seqs, b->nbytes, cudaMemcpyHostToDevice, st));
    DG_CUDA(cudaMemcpyAsync(b->off.p, offsets, ((size_t)nq + 1) * 8, cudaMemcpyHostToDevice, st));
    // the copies above read caller memory: finish them before returning ownership
    DG_CUDA(cudaStreamSynchronize(st));
    *out = b;
    return DG_OK;
  } catch (CudaFail& e) {
    delete b;
    return e.code;
  }
}

static void prof_mark(dg_index* ix, int i) {
  if (!ix->prof.enabled) return;
  if (!ix->prof.created) {
    for (auto& e : ix->prof.ev) cudaEventCreate(&e);
    ix->prof.created = true;
  }
  cudaEventRecord(ix->prof.ev[i], ix->stream);
}

static int run_impl(dg_batch* b) {
  dg_index* ix = b->ix;
  try {
    DG_CUDA(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    const unsigned B = 256;
    const uint32_t nq = b->nq;
    uint64_t launches = 0;
    b->nhits = 0;
    b->ncand = 0;
    BatchDev bd = batch_dev(b);
    UnitTabs ut{b->tab.p, b->tab_off.p, b->tab_cnt.p, b->script_ub.p};
    IndexView v = ix->view();
    ABuf<uint8_t> tmp;
    size_t tmp_cap = 0;
    auto ensure_tmp = [&](size_t bytes) -> void* {
      if (bytes > tmp_cap) { tmp.alloc(bytes + (bytes >> 3) + 256, st); tmp_cap = tmp.count; }
      return tmp.p;
    };
    prof_mark(ix, 0);
    // ---- prepare
    DG_CUDA(cudaMemsetAsync(b->units.p, 0, ((size_t)nq + 1) * 8, st));
    if (nq) { k_prepare<<<grid_for(nq, B), B, 0, st>>>(b->raw.p, bd, b->fwd.p, b->rc.p, ut, b->units.p); ++launches; }
    {
      size_t tb = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb, b->units.p, b->unit_off.p, (int)(nq + 1), st);
      cub::DeviceScan::ExclusiveSum(ensure_tmp(tb), tb, b->units.p, b->unit_off.p, (int)(nq + 1), st);
      launches += 2;
    }
    prof_mark(ix, 1);
    // ---- search
    uint64_t cap64 = 32ULL * nq + (1ULL << 20);
    if (b->par.distance >= 2) cap64 = 256ULL * nq + (1ULL << 20);
    if (const char* e = getenv("DG_CAND_CAP")) cap64 = strtoull(e, nullptr, 10);
    if (cap64 > (1ULL << 30)) cap64 = 1ULL << 30;
    b->cands.alloc(cap64, st);
    ABuf<unsigned int> ctr;     // [0] n_cand, [1] overflow
    ABuf<unsigned long long> nscripts;
    ctr.alloc(2, st);
    nscripts.alloc(1, st);
    DG_CUDA(cudaMemsetAsync(ctr.p, 0, 8, st));
    DG_CUDA(cudaMemsetAsync(nscripts.p, 0, 8, st));
    SearchOut so{b->cands.p, (uint32_t)cap64, ctr.p, ctr.p + 1, nscripts.p};
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ix->device);
    if (nq) { k_search<<<nsm * 8, 256, 0, st>>>(v, bd, ut, b->unit_off.p, b->uniform_units, so); ++launches; }
    prof_mark(ix, 2);
    unsigned int hc[2] = {0, 0};
    unsigned long long h_scripts = 0;
    DG_CUDA(cudaMemcpyAsync(hc, ctr.p, 8, cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaMemcpyAsync(&h_scripts, nscripts.p, 8, cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaStreamSynchronize(st));
    DG_CUDA(cudaGetLastError());
    if (hc[1]) {
      set_error("candidate buffer overflow (" + std::to_string(hc[0]) + " neighbour strings matched; raise DG_CAND_CAP)");
      return DG_ERR_OVERFLOW;
    }
    uint32_t n = hc[0];
    uint64_t n_candidates = n;
    ABuf<Cand> c2;
    ABuf<uint8_t> keep;
    ABuf<uint32_t> nsel;
    nsel.alloc(1, st);
    Cand* cur = b->cands.p;
    // ---- antichain rule (edit mode), lexicographic order, de-duplication
    if (n && b->par.indel) {
      keep.alloc(n, st);
      c2.alloc(n, st);
      k_minimal<<<grid_for(n, 128), 128, 0, st>>>(bd, cur, n, keep.p);
      size_t tb = 0;
      cub::DeviceSelect::Flagged(nullptr, tb, cur, keep.p, c2.p, nsel.p, (int)n, st);
      cub::DeviceSelect::Flagged(ensure_tmp(tb), tb, cur, keep.p, c2.p, nsel.p, (int)n, st);
      launches += 3;
      uint32_t n2 = 0;
      DG_CUDA(cudaMemcpyAsync(&n2, nsel.p, 4, cudaMemcpyDeviceToHost, st));
      DG_CUDA(cudaStreamSynchronize(st));
      // result now in c2; keep b->cands as the other buffer
      cur = c2.p;
      n = n2;
    }
    if (n) {
      size_t tb = 0;
      CandLess less{bd};
      cub::DeviceMergeSort::SortKeys(nullptr, tb, cur, (int)n, less, st);
      cub::DeviceMergeSort::SortKeys(ensure_tmp(tb), tb, cur, (int)n, less, st);
      launches += 3;
      if (!keep.p || keep.count < n) keep.alloc(n, st);
      Cand* other = (cur == c2.p) ? b->cands.p : nullptr;
      ABuf<Cand> c3;
      if (!other) { c3.alloc(n, st); other = c3.p; }
      k_unique<<<grid_for(n, B), B, 0, st>>>(bd, cur, n, keep.p);
      cub::DeviceSelect::Flagged(nullptr, tb, cur, keep.p, other, nsel.p, (int)n, st);
      cub::DeviceSelect::Flagged(ensure_tmp(tb), tb, cur, keep.p, other, nsel.p, (int)n, st);
      launches += 3;
      uint32_t n3 = 0;
      DG_CUDA(cudaMemcpyAsync(&n3, nsel.p, 4, cudaMemcpyDeviceToHost, st));
      DG_CUDA(cudaStreamSynchronize(st));
      if (other == c3.p) {
        // final list must outlive this scope: move it into b->cands
        DG_CUDA(cudaMemcpyAsync(b->cands.p, c3.p, (size_t)n3 * sizeof(Cand), cudaMemcpyDeviceToDevice, st));
        cur = b->cands.p;
      } else {
        cur = other;  // == b->cands.p
      }
      n = n3;
    }
    b->ncand = n;
    prof_mark(ix, 3);
    // ---- count-only mode (padlock.h:381-427, silica.h:365-394)
    if (b->counts_only) {
      b->counts.alloc(nq ? nq : 1, st);
      DG_CUDA(cudaMemsetAsync(b->counts.p, 0, (size_t)(nq ? nq : 1) * 8, st));
      if (n) { k_count<<<grid_for(n, B), B, 0, st>>>(cur, n, b->counts.p); ++launches; }
      prof_mark(ix, 4);
      prof_mark(ix, 5);
      b->ran = true;
      ix->prof.launches = launches;
      ix->prof.last.scripts = h_scripts;
      ix->prof.last.candidates = n_candidates;
      return DG_OK;
    }
    // ---- hit budget
    b->qhits.alloc((size_t)nq + 1, st);
    b->qoff.alloc((size_t)nq + 1, st);
    DG_CUDA(cudaMemsetAsync(b->qhits.p, 0, ((size_t)nq + 1) * 8, st));
    ABuf<uint64_t> take, before, ntake, nloc, hit_off, loc_off, keys, keys2;
    ABuf<uint32_t> qkey;
    uint64_t nhits = 0, nlocate = 0;
    if (n) {
      take.alloc(n, st); before.alloc(n, st); ntake.alloc((size_t)n + 1, st); nloc.alloc((size_t)n + 1, st);
      hit_off.alloc((size_t)n + 1, st); loc_off.alloc((size_t)n + 1, st); qkey.alloc(n, st);
      DG_CUDA(cudaMemsetAsync(ntake.p + n, 0, 8, st));
      DG_CUDA(cudaMemsetAsync(nloc.p + n, 0, 8, st));
      k_take_in<<<grid_for(n, B), B, 0, st>>>(cur, n, b->par.max_locations, take.p, qkey.p);
      size_t tb = 0;
      cub::DeviceScan::ExclusiveSumByKey(nullptr, tb, qkey.p, take.p, before.p, (int)n, cub::Equality(), st);
      cub::DeviceScan::ExclusiveSumByKey(ensure_tmp(tb), tb, qkey.p, take.p, before.p, (int)n, cub::Equality(), st);
      k_take_out<<<grid_for(n, B), B, 0, st>>>(cur, n, b->par.max_locations, before.p, take.p, ntake.p, nloc.p, b->qhits.p,
                                              b->status.p);
      cub::DeviceScan::ExclusiveSum(nullptr, tb, ntake.p, hit_off.p, (int)(n + 1), st);
      cub::DeviceScan::ExclusiveSum(ensure_tmp(tb), tb, ntake.p, hit_off.p, (int)(n + 1), st);
      cub::DeviceScan::ExclusiveSum(nullptr, tb, nloc.p, loc_off.p, (int)(n + 1), st);
      cub::DeviceScan::ExclusiveSum(ensure_tmp(tb), tb, nloc.p, loc_off.p, (int)(n + 1), st);
      launches += 8;
      DG_CUDA(cudaMemcpyAsync(&nhits, hit_off.p + n, 8, cudaMemcpyDeviceToHost, st));
      DG_CUDA(cudaMemcpyAsync(&nlocate, loc_off.p + n, 8, cudaMemcpyDeviceToHost, st));
    }
    {
      size_t tb = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb, (uint64_t*)b->qhits.p, b->qoff.p, (int)(nq + 1), st);
      cub::DeviceScan::ExclusiveSum(ensure_tmp(tb), tb, (uint64_t*)b->qhits.p, b->qoff.p, (int)(nq + 1), st);
      launches += 2;
    }
    DG_CUDA(cudaStreamSynchronize(st));
    uint64_t loc_cap = 1ULL << 28;
    if (const char* e = getenv("DG_LOCATE_CAP")) loc_cap = strtoull(e, nullptr, 10);
    if (nlocate > loc_cap || nlocate >= (1ULL << 31)) {
      set_error("too many occurrences to locate in one batch (" + std::to_string(nlocate) + "; raise DG_LOCATE_CAP or split the batch)");
      return DG_ERR_OVERFLOW;
    }
    // ---- locate + per-candidate ascending order
    const uint64_t* sorted_keys = nullptr;
    if (nlocate) {
      keys.alloc(nlocate, st);
      keys2.alloc(nlocate, st);
      k_locate<<<grid_for(nlocate, 128), 128, 0, st>>>(v, cur, n, loc_off.p, nlocate, keys.p);
      int cbits = 1;
      while ((1ULL << cbits) < (uint64_t)n + 1 && cbits < 32) ++cbits;
      size_t tb = 0;
      cub::DeviceRadixSort::SortKeys(nullptr, tb, keys.p, keys2.p, (int)nlocate, 0, 32 + cbits, st);
      cub::DeviceRadixSort::SortKeys(ensure_tmp(tb), tb, keys.p, keys2.p, (int)nlocate, 0, 32 + cbits, st);
      launches += 4;
      sorted_keys = keys2.p;
    }
    prof_mark(ix, 4);
    // ---- verify
    b->nhits = nhits;
    int maxq = b->par.seed_len ? (int)b->par.seed_len : std::min(b->max_len, kMaxQuery);
    int maxg = std::min(b->max_len, kMaxQuery) + 2 * (int)b->par.distance;
    if (!b->par.indel && !b->par.seed_len) maxg = maxq;
    uint32_t aln_max = (uint32_t)(maxg + maxq);
    b->pool_stride = b->par.seed_len ? (uint32_t)maxg : (b->par.indel ? 2 * aln_max : (uint32_t)(maxg + maxq));
    b->hits.alloc(nhits ? nhits : 1, st);
    b->pool.alloc(nhits ? nhits * b->pool_stride : 1, st);
    if (nhits) {
      VerifyArgs a;
      a.cands = cur; a.ncand = n; a.hit_off = hit_off.p; a.loc_off = loc_off.p; a.keys = sorted_keys; a.nhits = nhits;
      a.hits = b->hits.p; a.pool = b->pool.p; a.pool_stride = b->pool_stride;
      a.srow_ints = (uint32_t)(maxq + 2);
      a.trace_bytes = (uint32_t)(((maxg + 1) * (maxq + 1) + 3) / 4 + 4);
      a.scratch_stride = a.srow_ints * 4 + a.trace_bytes + 3 * (aln_max + 8);
      a.scratch_stride = (a.scratch_stride + 15) & ~15u;
      bool need_scratch = b->par.indel || b->par.seed_len;
      uint64_t chunk = nhits;
      if (need_scratch) {
        uint64_t budget = 1ULL << 30;
        chunk = std::max<uint64_t>(1, std::min<uint64_t>(nhits, budget / a.scratch_stride));
      }
      ABuf<uint8_t> scratch;
      scratch.alloc(need_scratch ? chunk * a.scratch_stride : 1, st);
      a.scratch = scratch.p;
      for (uint64_t first = 0; first < nhits; first += chunk) {
        a.first_hit = first;
        a.chunk = std::min<uint64_t>(chunk, nhits - first);
        k_verify<<<grid_for(a.chunk, 128), 128, 0, st>>>(v, bd, a);
        ++launches;
      }
    }
    prof_mark(ix, 5);
    DG_CUDA(cudaGetLastError());
    b->ran = true;
    ix->prof.launches = launches;
    ix->prof.last.scripts = h_scripts;
    ix->prof.last.candidates = n_candidates;
    ix->prof.last.located = nlocate;
    ix->prof.last.hits = nhits;
    return DG_OK;
  } catch (CudaFail& e) {
    return e.code;
  }
}

static void prof_collect(dg_index* ix) {
  if (!ix->prof.enabled || !ix->prof.created) return;
  cudaEventSynchronize(ix->prof.ev[5]);
  float ms[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&ms[i], ix->prof.ev[i], ix->prof.ev[i + 1]);
  ix->prof.last.ms_prepare = ms[0];
  ix->prof.last.ms_search = ms[1];
  ix->prof.last.ms_filter = ms[2];
  ix->prof.last.ms_locate = ms[3];
  ix->prof.last.ms_verify = ms[4];
  cudaEventElapsedTime(&ix->prof.last.ms_total, ix->prof.ev[0], ix->prof.ev[5]);
  ix->prof.last.launches = ix->prof.launches;
}

static int fetch_impl(dg_batch* b, dg_result** out) {
  if (!b->ran) { set_error("dg_batch_fetch before dg_batch_run"); return DG_ERR_ARG; }
  dg_index* ix = b->ix;
  try {
    DG_CUDA(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    dg_result* r = new dg_result();
    uint32_t nq = b->nq;
    r->hits.resize(b->nhits);
    r->qoff.resize((size_t)nq + 1);
    r->status.resize(nq);
    r->dist.resize(nq);
    r->pool.resize(b->nhits * b->pool_stride);
    r->seqs.resize(b->nbytes);
    if (b->nhits) {
      DG_CUDA(cudaMemcpyAsync(r->hits.data(), b->hits.p, b->nhits * sizeof(dg_hit), cudaMemcpyDeviceToHost, st));
      DG_CUDA(cudaMemcpyAsync(r->pool.data(), b->pool.p, b->nhits * b->pool_stride, cudaMemcpyDeviceToHost, st));
    }
    DG_CUDA(cudaMemcpyAsync(r->qoff.data(), b->qoff.p, ((size_t)nq + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (nq) {
      DG_CUDA(cudaMemcpyAsync(r->status.data(), b->status.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
      DG_CUDA(cudaMemcpyAsync(r->dist.data(), b->dist.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    }
    if (b->nbytes) DG_CUDA(cudaMemcpyAsync(r->seqs.data(), b->fwd.p, b->nbytes, cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaStreamSynchronize(st));
    prof_collect(ix);
    *out = r;
    return DG_OK;
  } catch (CudaFail& e) {
    return e.code;
  }
}

// ============================================================================================
extern "C" {

int dg_batch_stage(dg_index* idx, const char* seqs, const uint64_t* offsets, uint32_t nq, const dg_params* params,
                   dg_batch** out) {
  return stage_impl(idx, seqs, offsets, nq, params, out);
}
int dg_batch_run(dg_batch* b) {
  if (!b) { set_error("null batch"); return DG_ERR_ARG; }
  return run_impl(b);
}
int dg_batch_fetch(dg_batch* b, dg_result** out) {
  if (!b || !out) { set_error("null argument"); return DG_ERR_ARG; }
  return fetch_impl(b, out);
}
int dg_batch_summary(dg_batch* b, uint64_t* n_hits, uint64_t* n_candidates) {
  if (!b || !b->ran) { set_error("batch has not run"); return DG_ERR_ARG; }
  cudaStreamSynchronize(b->ix->stream);
  prof_collect(b->ix);
  if (n_hits) *n_hits = b->nhits;
  if (n_candidates) *n_candidates = b->ncand;
  return DG_OK;
}
void dg_batch_free(dg_batch* b) {
  if (!b) return;
  cudaSetDevice(b->ix->device);
  delete b;
}

int dg_hunt_batch(dg_index* idx, const char* seqs, const uint64_t* offsets, uint32_t nq, const dg_params* params,
                  dg_result** out) {
  dg_batch* b = nullptr;
  int rc = stage_impl(idx, seqs, offsets, nq, params, &b);
  if (rc) return rc;
  rc = run_impl(b);
  if (!rc) rc = fetch_impl(b, out);
  dg_batch_free(b);
  return rc;
}

int dg_count_batch(dg_index* idx, const char* seqs, const uint64_t* offsets, uint32_t nq, const dg_params* params,
                   uint64_t* counts) {
  if (!counts) { set_error("null argument"); return DG_ERR_ARG; }
  dg_batch* b = nullptr;
  int rc = stage_impl(idx, seqs, offsets, nq, params, &b);
  if (rc) return rc;
  b->counts_only = true;
  rc = run_impl(b);
  if (!rc && nq) {
    if (cudaMemcpyAsync(counts, b->counts.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, idx->stream) != cudaSuccess ||
        cudaStreamSynchronize(idx->stream) != cudaSuccess) {
      set_error("copy of counts failed");
      rc = DG_ERR_CUDA;
    }
  }
  dg_batch_free(b);
  return rc;
}

int dg_backward_search_batch(dg_index* idx, const char* seqs, const uint64_t* offsets, uint32_t nq, uint64_t* l, uint64_t* r) {
  if (!idx || !offsets || !l || !r || (nq && !seqs)) { set_error("null argument"); return DG_ERR_ARG; }
  try {
    DG_CUDA(cudaSetDevice(idx->device));
    cudaStream_t st = idx->stream;
    ABuf<uint8_t> d_s;
    ABuf<uint64_t> d_off, d_l, d_r;
    uint64_t nb = offsets[nq];
    d_s.alloc(nb + 1, st); d_off.alloc((size_t)nq + 1, st); d_l.alloc(nq, st); d_r.alloc(nq, st);
    DG_CUDA(cudaMemcpyAsync(d_s.p, seqs, nb, cudaMemcpyHostToDevice, st));
    DG_CUDA(cudaMemcpyAsync(d_off.p, offsets, ((size_t)nq + 1) * 8, cudaMemcpyHostToDevice, st));
    if (nq) k_backward_search<<<grid_for(nq, 128), 128, 0, st>>>(idx->view(), d_s.p, d_off.p, nq, d_l.p, d_r.p);
    DG_CUDA(cudaMemcpyAsync(l, d_l.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaMemcpyAsync(r, d_r.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaStreamSynchronize(st));
    DG_CUDA(cudaGetLastError());
    return DG_OK;
  } catch (CudaFail& e) {
    return e.code;
  }
}

const dg_hit* dg_result_hits(const dg_result* r, uint64_t* n) {
  if (n) *n = r ? r->hits.size() : 0;
  return r && !r->hits.empty() ? r->hits.data() : nullptr;
}
const uint64_t* dg_result_query_offsets(const dg_result* r, uint32_t* nq) {
  if (nq) *nq = r ? (uint32_t)(r->qoff.size() - 1) : 0;
  return r ? r->qoff.data() : nullptr;
}
const uint32_t* dg_result_query_status(const dg_result* r) { return r ? r->status.data() : nullptr; }
const uint32_t* dg_result_query_distance(const dg_result* r) { return r ? r->dist.data() : nullptr; }
const char* dg_result_pool(const dg_result* r, uint64_t* bytes) {
  if (bytes) *bytes = r ? r->pool.size() : 0;
  return r ? r->pool.data() : nullptr;
}
const char* dg_result_sequences(const dg_result* r, uint64_t* bytes) {
  if (bytes) *bytes = r ? r->seqs.size() : 0;
  return r ? r->seqs.data() : nullptr;
}
void dg_result_free(dg_result* r) { delete r; }

// wire format of the hit all-gather: u64 nq, u64 nhits, u64 pool bytes, u64 seq bytes, then
// qoff[nq+1], status[nq], dist[nq], hits[nhits], pool, seqs
int dg_result_pack(const dg_result* r, void* buf, uint64_t* bytes) {
  if (!r || !bytes) { set_error("null argument"); return DG_ERR_ARG; }
  uint64_t nq = r->qoff.size() - 1, nh = r->hits.size(), np = r->pool.size(), ns = r->seqs.size();
  uint64_t need = 32 + (nq + 1) * 8 + nq * 8 + nh * sizeof(dg_hit) + np + ns;
  if (!buf) { *bytes = need; return DG_OK; }
  if (*bytes < need) { set_error("buffer too small"); return DG_ERR_ARG; }
  uint8_t* p = (uint8_t*)buf;
  uint64_t hdr[4] = {nq, nh, np, ns};
  memcpy(p, hdr, 32); p += 32;
  memcpy(p, r->qoff.data(), (nq + 1) * 8); p += (nq + 1) * 8;
  memcpy(p, r->status.data(), nq * 4); p += nq * 4;
  memcpy(p, r->dist.data(), nq * 4); p += nq * 4;
  memcpy(p, r->hits.data(), nh * sizeof(dg_hit)); p += nh * sizeof(dg_hit);
  memcpy(p, r->pool.data(), np); p += np;
  memcpy(p, r->seqs.data(), ns);
  *bytes = need;
  return DG_OK;
}
int dg_result_unpack(const void* buf, uint64_t bytes, dg_result** out) {
  if (!buf || !out || bytes < 32) { set_error("bad buffer"); return DG_ERR_ARG; }
  const uint8_t* p = (const uint8_t*)buf;
  uint64_t hdr[4];
  memcpy(hdr, p, 32); p += 32;
  uint64_t nq = hdr[0], nh = hdr[1], np = hdr[2], ns = hdr[3];
  uint64_t need = 32 + (nq + 1) * 8 + nq * 8 + nh * sizeof(dg_hit) + np + ns;
  if (bytes < need) { set_error("truncated buffer"); return DG_ERR_ARG; }
  dg_result* r = new dg_result();
  r->qoff.resize(nq + 1); r->status.resize(nq); r->dist.resize(nq); r->hits.resize(nh); r->pool.resize(np); r->seqs.resize(ns);
  memcpy(r->qoff.data(), p, (nq + 1) * 8); p += (nq + 1) * 8;
  memcpy(r->status.data(), p, nq * 4); p += nq * 4;
  memcpy(r->dist.data(), p, nq * 4); p += nq * 4;
  memcpy(r->hits.data(), p, nh * sizeof(dg_hit)); p += nh * sizeof(dg_hit);
  memcpy(r->pool.data(), p, np); p += np;
  memcpy(r->seqs.data(), p, ns);
  *out = r;
  return DG_OK;
}

}  // extern "C"
