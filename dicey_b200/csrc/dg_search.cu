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
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <cub/cub.cuh>
#include <map>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <new>

#include "dg_common.cuh"
#include "nbr_trunc.hpp"

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
  uint64_t* qcode;          // 2 per query: packed forward / reverse-complement search string
  uint8_t* qflag;           // bit0 = ACGT only ("clean"), bit1 = packed code valid
  uint32_t* irregular;      // bit 0: a query departs from the uniform batch shape; bit 1: a Hamming neighbourhood
                            // reaches the cap; bit 2: an edit neighbourhood is not certified below the cap
  uint32_t uniform_len;     // common raw length of the batch (0 = mixed lengths)
  uint32_t key_sortable;    // every string of the batch fits the 126-bit sort key (listed neighbourhoods need that)
  // neighbourhoods the cap -x truncated (neighbors.h:50), replayed on the host (nbr_trunc.hpp): for the
  // (query, strand) pairs listed in trunc_qs (ascending (q << 1) | strand) the set IS the sorted key list
  // trunc_keys[trunc_off[i] .. trunc_off[i + 1]) instead of the substring-minimal strings
  const uint32_t* trunc_qs;
  const uint32_t* trunc_off;
  const ulonglong2* trunc_keys;
  uint32_t n_trunc;
};

// A candidate whose string came from a host-made list (distance 3: nbr_trunc.hpp replays the whole
// neighbourhood) instead of an edit script: code = strand | 3 << 1 | length << 3 | rank in the
// (query, strand) set << 11 (the set is sorted as std::set<std::string> iterates it).
constexpr uint32_t kListedNev = 3;
DG_HD bool cand_is_listed(uint32_t code) { return ((code >> 1) & 3u) == kListedNev; }
DG_HD uint32_t pack_listed(int strand, int len, uint32_t rank) { return (uint32_t)strand | (kListedNev << 1) | ((uint32_t)len << 3) | (rank << 11); }
struct ListedStrings {      // uploaded by resolve_special
  const uint8_t* chars;     // the strings back to back
  const uint32_t* off;      // n + 1 offsets
  const uint32_t* q;        // query of each string
  const uint32_t* code;     // pack_listed(...) of each string
  uint32_t n;
};

DG_HD void query_geom(const BatchDev& b, uint32_t q, int strand, const uint8_t*& base, int& m, int& koff) {
  uint64_t o = b.off[q];
  int L = (int)(b.off[q + 1] - o);
  m = b.seed_len ? (int)b.seed_len : L;
  koff = L - m;
  base = strand == 0 ? b.fwd + o + koff : b.rc + o;
}

// Script enumeration.  A "clean" (ACGT-only) query enumerates 3 substitutions per position (the
// three other letters), a query holding 'N' enumerates 4 (neighbors.h:61-69 substitutes every
// alphabet letter different from the base); edit mode adds 1 deletion + 4 insertions per position.
DG_HD int enum_slots(bool indel, bool clean) { return indel ? (clean ? 8 : 9) : (clean ? 3 : 4); }
DG_HD bool enum_is_ins(bool indel, bool clean, int kk) { return indel && kk >= (clean ? 4 : 5); }
// enumeration slot kk at a position whose base has 2-bit code bc (4 = not ACGT) -> canonical k
// (0..3 substitute that letter, 4 delete, 5..8 insert); false if the slot is not a real edit
DG_HD bool enum_to_canonical(bool clean, int kk, int bc, int& k) {
  if (clean) {
    if (kk < 3) { k = (bc + 1 + kk) & 3; return true; }
    k = kk + 1;  // 3 -> delete (4), 4..7 -> insert (5..8)
    return true;
  }
  k = kk;
  return !(kk < 4 && kk == bc);
}

// Per-length unit tables: a "unit" is a group of <= 32 consecutive scripts run by one warp.
// Row 0 holds the single-event scripts (preceded by the unedited string unless edit mode with
// d >= 1, where the query itself can never be substring-minimal); row e1 + 1 holds the pairs
// whose first event is enumeration slot e1.
struct UnitTabs {
  const uint32_t* tab;       // packed (row << 12) | first index
  const uint32_t* tab_off;   // [2][256]: variant 0 = clean, 1 = with 'N'
  const uint32_t* tab_cnt;   // [2][256]: units per strand for search-string length m
  const uint32_t* script_ub; // 256 entries: scripts per strand (9- / 4-slot count, saturating)
};

namespace {

// Per-stream cache of device blocks.  One batch makes ~40 short-lived allocations; with three
// pipeline workers the CUDA API calls behind them serialise on the context lock and were the
// largest part of the per-chunk fixed cost.  A freed block goes back to the cache of ITS stream
// and is only ever handed out again for work on the same stream, so stream order keeps reuse
// safe without events; misses fall through to cudaMallocAsync.
struct StreamPool {
  std::mutex mu;                               // (a stream's pool may be shared by host threads: DG_WORKERS, callers of one index)
  std::multimap<size_t, void*> free_blocks;   // capacity -> block
  size_t cached = 0;
  static size_t round_up(size_t bytes) {       // (8 + k) * 2^e classes: at most 12.5 % slack
    if (bytes < 4096) return 4096;
    int e = 63 - __builtin_clzll((unsigned long long)bytes);
    size_t step = (size_t)1 << (e - 3);
    return (bytes + step - 1) & ~(step - 1);
  }
  void* get(size_t cap) {
    std::lock_guard<std::mutex> g(mu);
    auto it = free_blocks.find(cap);
    if (it == free_blocks.end()) return nullptr;
    void* p = it->second;
    free_blocks.erase(it);
    cached -= cap;
    return p;
  }
  bool put(void* p, size_t cap) {
    std::lock_guard<std::mutex> g(mu);
    if (cached + cap > (24ULL << 30)) return false;
    free_blocks.emplace(cap, p);
    cached += cap;
    return true;
  }
};
std::mutex g_pools_mu;
std::map<cudaStream_t, StreamPool*> g_pools;
StreamPool* pool_of(cudaStream_t st) {
  std::lock_guard<std::mutex> g(g_pools_mu);
  auto it = g_pools.find(st);
  if (it != g_pools.end()) return it->second;
  StreamPool* p = new StreamPool();
  g_pools[st] = p;
  return p;
}

// A side stream per batch stream: the slow path of one tile (k_resolve: latency-bound) runs there while
// the probes of the next tile (k_probe_*: DRAM-bound) run on the batch stream.
struct AuxStream {
  cudaStream_t s = nullptr;
  cudaEvent_t ev[8] = {};
  unsigned next = 0;
  cudaEvent_t event() { return ev[next++ & 7u]; }
};
std::mutex g_aux_mu;
std::map<cudaStream_t, AuxStream*> g_aux;
AuxStream* aux_of(cudaStream_t st) {
  std::lock_guard<std::mutex> g(g_aux_mu);
  auto it = g_aux.find(st);
  if (it != g_aux.end()) return it->second;
  AuxStream* a = new AuxStream();
  int least = 0, greatest = 0;
  cudaDeviceGetStreamPriorityRange(&least, &greatest);
  if (cudaStreamCreateWithPriority(&a->s, cudaStreamNonBlocking, greatest) != cudaSuccess) { delete a; return nullptr; }
  for (auto& e : a->ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  g_aux[st] = a;
  return a;
}

template <typename T>
struct ABuf {  // stream-ordered allocation through the stream's block cache
  T* p = nullptr;
  size_t count = 0, cap = 0;
  cudaStream_t st = nullptr;
  ABuf() = default;
  ABuf(const ABuf&) = delete;
  ABuf& operator=(const ABuf&) = delete;
  ~ABuf() { release(); }
  void alloc(size_t n, cudaStream_t s) {
    release();
    st = s;
    count = n;
    cap = StreamPool::round_up((n ? n : 1) * sizeof(T));
    p = (T*)pool_of(s)->get(cap);
    if (!p) DG_CUDA(cudaMallocAsync((void**)&p, cap, s));
  }
  void release() {
    if (p && !pool_of(st)->put(p, cap)) cudaFreeAsync(p, st);
    p = nullptr;
    count = cap = 0;
  }
  void swap(ABuf& o) { std::swap(p, o.p); std::swap(count, o.count); std::swap(cap, o.cap); std::swap(st, o.st); }
};

inline unsigned grid_for(uint64_t items, unsigned block) { return (unsigned)((items + block - 1) / block); }

// Waiting for a stream.  cudaStreamSynchronize spins (lowest latency: one process per box); with one
// process per GPU and several pipeline threads each, spinning waiters outnumber the cores, so
// DG_SYNC=block makes a waiter sleep on a blocking event instead.
inline cudaError_t sync_stream(cudaStream_t st) {
  static const bool block = getenv("DG_SYNC") && std::string(getenv("DG_SYNC")) == "block";
  if (!block) return cudaStreamSynchronize(st);
  thread_local cudaEvent_t ev = nullptr;
  thread_local int ev_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!ev || ev_dev != dev) {
    cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    ev_dev = dev;
  }
  cudaError_t e = cudaEventRecord(ev, st);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(ev);
}

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
  bool changed = false;   // normalisation altered the sequence (lower case, non-ACGT): irregular bit 3
  if (((o | (uint64_t)L) & 3) == 0) {
    // word path (offsets and length multiples of 4, e.g. batches of 20-mers): 4 bases per access
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(raw + o);
    uint32_t* fw = reinterpret_cast<uint32_t*>(fwd + o);
    uint32_t* cw = reinterpret_cast<uint32_t*>(rc + o);
    const int nw = L >> 2;
    for (int j = 0; j < nw; ++j) {
      const uint32_t w = rw[j];
      const uint8_t c0 = norm_base((uint8_t)w), c1 = norm_base((uint8_t)(w >> 8)), c2 = norm_base((uint8_t)(w >> 16)),
                    c3 = norm_base((uint8_t)(w >> 24));
      fw[j] = (uint32_t)c0 | ((uint32_t)c1 << 8) | ((uint32_t)c2 << 16) | ((uint32_t)c3 << 24);
      changed |= fw[j] != w;
      cw[nw - 1 - j] = (uint32_t)comp_base(c3) | ((uint32_t)comp_base(c2) << 8) | ((uint32_t)comp_base(c1) << 16) |
                       ((uint32_t)comp_base(c0) << 24);
    }
  } else {
    for (int i = 0; i < L; ++i) {
      uint8_t ch = norm_base(raw[o + i]);
      changed |= ch != raw[o + i];
      fwd[o + i] = ch;
      rc[o + L - 1 - i] = comp_base(ch);
    }
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
  if (d > (uint32_t)kMaxListDist || (b.seed_len && d >= b.seed_len) || (d > (uint32_t)kMaxDist && m + (int)d > 42)) st |= DG_Q_UNSUPPORTED;
  bool run = !(st & (DG_Q_SKIPPED | DG_Q_TOO_SHORT | DG_Q_UNSUPPORTED));
  if (run && b.max_loc == 0) st |= DG_Q_HIT_CAP;   // -m 0: no hit is taken and hunter.h:434 (0 >= 0) warns for every query
  // beyond the enumerated distances the host replays the neighbourhood and uploads it as a list
  // (resolve_special): the query takes no part in the script enumeration below
  if (run && d > (uint32_t)kMaxDist && !b.key_sortable) { st |= DG_Q_UNSUPPORTED; run = false; }
  const bool listed = run && d > (uint32_t)kMaxDist;
  if (listed) { st |= DG_Q_NBR_UNVERIFIED; atomicOr(b.irregular, 16u); run = false; }
  bool clean = true;
  uint64_t cf = 0, cr = 0;
  if (run) {
    // search strings: forward = last m bases, reverse = first m bases of the reverse complement
    const uint8_t* s0 = fwd + o + (L - m);
    const uint8_t* s1 = rc + o;
    uint64_t w = 0, w2 = 0;  // per-position substitution choices: 3, or 4 at an 'N'
    for (int i = 0; i < m; ++i) {
      int c0 = base_code(s0[i]), c1 = base_code(s1[i]);
      if (c0 == 4) clean = false;
      uint64_t c = c0 < 4 ? 3 : 4;
      w += c; w2 += c * c;
      if (m <= 32) { cf = (cf << 2) | (uint64_t)(c0 & 3); cr = (cr << 2) | (uint64_t)(c1 & 3); }
    }
    // neighbors.h:50 stops the DFS once the set holds max_neighborhood strings.  The set never
    // holds more strings than scripts were generated, so fewer scripts than the cap certifies an
    // untruncated neighbourhood.  Hamming sets hold exactly one string per script.
    // Edit mode: k_nbr_bound certifies most of the flagged queries afterwards; what stays flagged, and
    // every Hamming query whose (exactly known) set size reaches the cap, is replayed on the host.
    if (b.indel) {
      if (ut.script_ub[m] >= b.max_nbr) st |= DG_Q_NBR_UNVERIFIED;
    } else {
      uint64_t size = 1 + (d >= 1 ? w : 0) + (d >= 2 ? (w * w - w2) / 2 : 0);
      if (size >= b.max_nbr) { st |= DG_Q_NBR_CAP; atomicOr(b.irregular, 2u); }
    }
  }
  bool packed = run && clean && (m + (int)d <= kMaxPacked);
  b.qcode[2 * (uint64_t)q] = cf;
  b.qcode[2 * (uint64_t)q + 1] = cr;
  b.qflag[q] = (uint8_t)((clean ? 1 : 0) | (packed ? 2 : 0));
  b.status[q] = st;
  b.dist[q] = d;
  int variant = clean ? 0 : 1;
  // packed queries are searched by k_search_packed; k_search takes the rest
  units[q] = (run && !packed) ? (uint64_t)ut.tab_cnt[variant * 256 + m] * (b.reverse ? 2 : 1) : 0;
  if (listed) run = true;   // (searched, through the list)
  if (!run || listed || !clean || (uint32_t)L != b.uniform_len || d != b.distance) atomicOr(b.irregular, 1u);
  if (changed) atomicOr(b.irregular, 8u);
}

// One thread per query still flagged DG_Q_NBR_UNVERIFIED: the bound of dg_core.cuh (nbr_upper_bound_closed)
// on the number of distinct strings neighbors() can generate, for both strands; below the cap the
// flag goes (the reference cannot have truncated), otherwise the host replays the query exactly.
__global__ void k_nbr_bound(BatchDev b) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= b.nq) return;
  const uint32_t st = b.status[q];
  if (!(st & DG_Q_NBR_UNVERIFIED)) return;
  const int d = (int)b.dist[q];
  bool ok = d <= 2;
  for (int strand = 0; ok && strand < (b.reverse ? 2 : 1); ++strand) {
    const uint8_t* base;
    int m, koff;
    query_geom(b, q, strand, base, m, koff);
    auto bq = [&](int j) { return base_code(base[j]); };
    if (nbr_upper_bound_closed(bq, m, d) >= b.max_nbr) ok = false;
  }
  if (ok) b.status[q] = st & ~(uint32_t)DG_Q_NBR_UNVERIFIED;
  else atomicOr(b.irregular, 4u);
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

// One backward-search step for an ACGT code.  Only the count of code c and the two bit planes of
// the 32-byte block are loaded (no dynamically indexed register array -> no local memory).
__device__ __forceinline__ uint32_t rank_planes(const IndexView& ix, uint64_t lo, uint64_t hi, uint32_t i, int c) {
  const uint32_t o = i & 63;
  const uint64_t mask = o ? (~0ULL >> (64 - o)) : 0ULL;
  const uint64_t m = ((c & 1) ? lo : ~lo) & ((c & 2) ? hi : ~hi) & mask;
  uint32_t r = (uint32_t)__popcll(m);
  if (c == 0 && o && region_flag(ix, (uint64_t)i - 1)) {
    // non-ACGT symbols are stored as code 0: take those inside [64*blk, i) back out
    const uint32_t base = i & ~63u;
    const uint32_t a = lower_bound_u32(ix.exc_pos, 0, ix.n_exc, base);
    const uint32_t e = lower_bound_u32(ix.exc_pos, a, ix.n_exc, i);
    r -= (e - a);
  }
  return r;
}
__device__ __forceinline__ void step_acgt(const IndexView& ix, uint32_t& l, uint32_t& r, int c) {
  const OccBlock* pl = ix.occ + (l >> 6);
  const uint32_t cl = __ldg(&pl->cnt[c]);
  const uint4 wl = __ldg(reinterpret_cast<const uint4*>(pl) + 1);
  const uint64_t llo = ((uint64_t)wl.y << 32) | wl.x, lhi = ((uint64_t)wl.w << 32) | wl.z;
  const uint32_t c4 = ix.C4[c];
  uint32_t nr;
  if ((r >> 6) == (l >> 6)) {
    nr = c4 + cl + rank_planes(ix, llo, lhi, r, c);
  } else {
    const OccBlock* pr = ix.occ + (r >> 6);
    const uint32_t cr = __ldg(&pr->cnt[c]);
    const uint4 wr = __ldg(reinterpret_cast<const uint4*>(pr) + 1);
    nr = c4 + cr + rank_planes(ix, ((uint64_t)wr.y << 32) | wr.x, ((uint64_t)wr.w << 32) | wr.z, r, c);
  }
  l = c4 + cl + rank_planes(ix, llo, lhi, l, c);
  r = nr;
}

__global__ void __launch_bounds__(256) k_search(IndexView ix, BatchDev b, UnitTabs ut, const uint64_t* __restrict__ unit_off,
                                                uint64_t uniform_units, SearchOut out) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t total = unit_off[b.nq];
  const bool indel = b.indel != 0;
  const bool uniform = uniform_units != 0 && *b.irregular == 0;
  const int K = (int)ix.K;
  const uint32_t kmask = (K >= 16) ? 0xFFFFFFFFu : ((1u << (2 * K)) - 1u);
  unsigned long long my_scripts = 0;
  for (uint64_t unit = warp; unit < total; unit += nwarps) {
    // unit -> (query, strand, local unit)
    uint32_t q;
    uint64_t local;
    if (uniform) {
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
    const uint8_t flags = b.qflag[q];
    const bool clean = flags & 1, packed = flags & 2;
    const uint8_t* base;
    int m, koff;
    query_geom(b, q, 0, base, m, koff);
    const int variant = clean ? 0 : 1;
    uint32_t per_strand = ut.tab_cnt[variant * 256 + m];
    int strand = (int)(local / per_strand);
    uint32_t u = (uint32_t)(local - (uint64_t)strand * per_strand);
    if (strand) query_geom(b, q, 1, base, m, koff);
    uint32_t pk = ut.tab[ut.tab_off[variant * 256 + m] + u];
    const int row = (int)(pk >> 12);          // 0 = singles (+ the unedited string)
    int idx = (int)(pk & 0xFFF) + (int)lane;
    const int dq = (int)b.dist[q];
    const int S = enum_slots(indel, clean);
    const int E = S * m;
    const bool with_base = !(indel && dq >= 1);
    uint64_t code = packed ? b.qcode[2 * (uint64_t)q + strand] : 0;
    auto bcode = [&](int pos) -> int {
      return packed ? (int)((code >> (2 * (m - 1 - pos))) & 3) : base_code(base[pos]);
    };
    Script sc;
    sc.nev = 0; sc.pos[0] = sc.pos[1] = 0; sc.k[0] = sc.k[1] = 0;
    bool valid = true;
    if (row == 0) {
      int e = with_base ? idx - 1 : idx;
      if (e >= 0) {
        valid = dq >= 1 && e < E;
        if (valid) {
          sc.pos[0] = e / S;
          valid = enum_to_canonical(clean, e - sc.pos[0] * S, bcode(sc.pos[0]), sc.k[0]);
        }
        sc.nev = 1;
      }
    } else {
      int e1 = row - 1, e2 = idx;
      valid = dq >= 2 && e2 < E;
      if (valid) {
        sc.pos[0] = e1 / S;
        sc.pos[1] = e2 / S;
        valid = enum_to_canonical(clean, e1 - sc.pos[0] * S, bcode(sc.pos[0]), sc.k[0]) &&
                enum_to_canonical(clean, e2 - sc.pos[1] * S, bcode(sc.pos[1]), sc.k[1]) &&
                pair_ok(sc.pos[0], sc.k[0], sc.pos[1]);
      }
      sc.nev = 2;
    }
    if (!valid) continue;
    ++my_scripts;
    const int L = script_len(m, sc);
    if (L <= 0) continue;
    uint32_t l = 0, r = (uint32_t)ix.n;
    bool alive;
    if (packed) {
      // the edited string as one 64-bit code: table lookup on its low 2K bits, then one
      // backward step per remaining base
      for (int e = 0; e < sc.nev; ++e) code = apply_event_packed(code, m - 1 - sc.pos[e], sc.k[e]);
      int t = 0;
      if (L >= K) {
        uint2 iv = __ldg(&ix.kmer[(uint32_t)code & kmask]);
        l = iv.x; r = iv.y; t = K;
      }
      while (l < r && t < L) {
        step_acgt(ix, l, r, (int)((code >> (2 * t)) & 3));
        ++t;
      }
      alive = l < r;
    } else {
      bool collecting = L >= K;
      int cnt = 0;
      uint32_t kc = 0;
      alive = script_rtl(base, m, sc, [&](uint8_t x) -> bool {
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
      }) && l < r;
    }
    if (alive) {
      unsigned int slot = atomicAdd(out.n_cand, 1u);
      if (slot < out.cap) {
        const int cs = indel ? 9 : 4;  // canonical slot numbering carried by the candidate
        Cand c;
        c.q = q; c.l = l; c.r = r;
        c.code = pack_script(strand, sc.nev, sc.pos[0] * cs + sc.k[0], sc.pos[1] * cs + sc.k[1]);
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
// k_search_packed: the fast path for ACGT-only queries whose edited strings fit one 64-bit code
// (qflag bit 1).  One warp owns one (query, strand) and walks its edit scripts 32 at a time:
//   1. every lane builds its edited string as a packed code (a few shifts) and probes the KB-mer
//      presence bitmap with the last KB bases: one DRAM access that empties ~95 % of the lanes
//      on a 3 Gb text (the string cannot occur if its suffix does not);
//   2. the survivors are compacted with a warp ballot into a per-warp shared-memory queue, and
//      each time 32 of them are queued they run the slow path together with full lanes:
//      K-mer table lookup on the last K bases + one backward-search step per remaining base.
// The strings enumerated are exactly those of k_search for a clean query (neighbors.h:47-83).
#ifndef DG_PACKED_MIN_BLOCKS
#define DG_PACKED_MIN_BLOCKS 4
#endif
// What the kernel reads of the batch and of the index (the whole BatchDev / IndexView as parameters
// costs registers the enumeration loop needs).
struct PackedArgs {
  const uint64_t* qcode;
  const uint8_t* qflag;
  const uint32_t* dist;
  const uint64_t* off;
  uint32_t nq, seed_len;
  uint32_t reverse;
  // presence windows by string length class: 0 = KB - 1, 1 = KB, 2 = KB + 1 and longer (right- and
  // left-anchored bitmap, window bases; a null right bitmap = no filter for that class)
  const uint32_t* win_r[3];
  const uint32_t* win_l[3];
  int win_k[3];
  int KB;
};

// The edited string of one event at right-based index j of `code` (an ACGT-only string as 2-bit
// codes, last base in the low bits): enumeration kind kk = 0..2 substitute the base by the three
// other letters, 3 delete it, 4..7 insert A C G T before it.  T / low / bb are the pieces of `code`
// around j, shared by the eight kinds of one position.
struct EditSite {
  uint64_t T;      // the bases left of j, moved down onto j
  uint64_t low;    // the bases right of j
  uint32_t bb;     // the base at j
  int sh;          // 2 j
};
__device__ __forceinline__ EditSite edit_site(uint64_t code, int j) {
  EditSite s;
  s.sh = 2 * j;
  s.low = code & ((1ULL << s.sh) - 1ULL);
  s.bb = (uint32_t)(code >> s.sh) & 3u;
  s.T = ((code >> s.sh) >> 2) << s.sh;
  return s;
}
template <bool INDEL>
__device__ __forceinline__ uint64_t edit_apply(const EditSite& s, int kk, int& dL, int& kcanon) {
  if (!INDEL || kk < 3) {
    const uint32_t c = (s.bb + 1u + (uint32_t)kk) & 3u;
    dL = 0; kcanon = (int)c;
    return (s.T << 2) | ((uint64_t)c << s.sh) | s.low;
  }
  if (kk == 3) { dL = -1; kcanon = 4; return s.T | s.low; }
  const uint32_t c = (uint32_t)(kk - 4);
  dL = 1; kcanon = 5 + (int)c;
  return (s.T << 4) | ((uint64_t)((c << 2) | s.bb) << s.sh) | s.low;
}

// Address of the presence-bitmap word for the longest window the string covers (KB + 1, KB or KB - 1
// bases): its last bases or -- when all its edits lie right of its first kb - 7 bases -- its first
// bases, so that siblings probe the same 2 KB region (right- and left-anchored bitmaps).  Branch-free:
// the caller issues the loads of several probes back to back and tests the bits afterwards.
// state: 0 = test the bit, 1 = passes without a filter (no bitmap for that length), 2 = cannot occur.
struct ProbeAddr {
  const uint32_t* word;
  uint32_t bit;     // bit inside the word
  uint32_t state;
};
// the window of one string length, picked once for all the strings of that length
struct WinSel {
  const uint32_t* bm;
  const uint32_t* bml;
  int kb;
  int cls;
};
__device__ __forceinline__ WinSel win_select(const PackedArgs& a, int L) {
  WinSel w;
  w.cls = L - a.KB + 1;
  const int c = w.cls > 2 ? 2 : (w.cls < 0 ? 0 : w.cls);
  w.bm = c == 2 ? a.win_r[2] : (c == 1 ? a.win_r[1] : a.win_r[0]);
  w.bml = c == 2 ? a.win_l[2] : (c == 1 ? a.win_l[1] : a.win_l[0]);
  w.kb = c == 2 ? a.win_k[2] : (c == 1 ? a.win_k[1] : a.win_k[0]);
  return w;
}
template <bool OTHER = false>   // OTHER: the window at the opposite end (the second opinion of the slow path)
__device__ __forceinline__ ProbeAddr presence_addr(const PackedArgs& a, const WinSel& ws, uint64_t code, int L, int p_left) {
  const int cls = ws.cls;
  const uint32_t* bm = ws.bm;
  const uint32_t* bml = ws.bml;
  const int kb = ws.kb;
  const uint64_t wmask = (1ULL << (2 * kb)) - 1ULL;
  const int sr = 2 * (L - kb);
  const uint64_t bit_l = presence_bit_left((code >> (sr > 0 ? sr : 0)) & wmask, kb);
  const uint64_t bit_r = presence_bit(code & wmask, kb);
  const bool first = bml != nullptr && p_left >= kb - presence_bit_bases(kb);
  const bool left = OTHER ? !first : first;
  const uint64_t bit = left ? bit_l : bit_r;
  const uint32_t* base = left ? bml : bm;
  ProbeAddr r;
  r.state = L <= 0 ? 2u : ((a.KB == 0 || cls < 0 || bm == nullptr || (OTHER && (bml == nullptr || L <= kb))) ? 1u : 0u);
  r.word = r.state ? reinterpret_cast<const uint32_t*>(a.qcode) : base + (bit >> 5);   // (a harmless, always valid address when no test is needed)
  r.bit = (uint32_t)bit & 31u;
  return r;
}
__device__ __forceinline__ uint32_t ld_probe(const uint32_t* p) {
  uint32_t word;   // one random 4-byte read: ask L2 to fill 64 bytes instead of the whole 128-byte line
#ifndef DG_PROBE_LD
#define DG_PROBE_LD 64
#endif
#if DG_PROBE_LD == 64
  asm("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(word) : "l"(p));
#elif DG_PROBE_LD == 128
  asm("ld.global.nc.L2::128B.u32 %0, [%1];" : "=r"(word) : "l"(p));
#else
  asm("ld.global.nc.u32 %0, [%1];" : "=r"(word) : "l"(p));
#endif
  return word;
}
// the probes of all kinds of one position: addresses first, then the loads (in flight together), then the bits
template <bool INDEL>
__device__ __forceinline__ uint32_t probe_site(const PackedArgs& a, const EditSite& site, int L1, int p_left, bool have) {
  constexpr int S = INDEL ? 8 : 3;
  ProbeAddr pa[S];
  const WinSel w0 = win_select(a, L1), wm = INDEL ? win_select(a, L1 - 1) : w0, wp = INDEL ? win_select(a, L1 + 1) : w0;
#pragma unroll
  for (int kk = 0; kk < S; ++kk) {
    int dL, kc;
    const uint64_t code = edit_apply<INDEL>(site, kk, dL, kc);
    pa[kk] = presence_addr(a, (!INDEL || kk < 3) ? w0 : (kk == 3 ? wm : wp), code, L1 + dL, p_left);
  }
  uint32_t w[S];
#pragma unroll
  for (int kk = 0; kk < S; ++kk) w[kk] = ld_probe(pa[kk].word);
  uint32_t mask = 0;
#pragma unroll
  for (int kk = 0; kk < S; ++kk) {
    const uint32_t bit = pa[kk].state == 0 ? ((w[kk] >> pa[kk].bit) & 1u) : (pa[kk].state == 1 ? 1u : 0u);
    mask |= bit << kk;
  }
  return have ? mask : 0u;
}
__device__ __forceinline__ bool presence_probe(const PackedArgs& a, uint64_t code, int L, int p_left) {
  const ProbeAddr pa = presence_addr(a, win_select(a, L), code, L, p_left);
  if (pa.state) return pa.state == 1;
  return (ld_probe(pa.word) >> pa.bit) & 1u;
}

// Slow path of 32 queued strings (one per lane; `have` marks the real ones): K-mer table lookup on
// the last K bases + one backward-search step per remaining base; survivors become candidates.
// Not inlined: it is reached from several places of the kernel and rarely (~4 % of the strings).
// A string longer than its window first takes a second opinion: the window at its other end (the
// bitmap with the opposite anchoring) must hold it as well.  A false survivor of the first probe --
// its 19 last bases occur somewhere, say -- walks the whole backward search before the last base
// fails it (~10 dependent random reads); about three in four of them are stopped here by one read.
__device__ __forceinline__ bool second_opinion(const PackedArgs& a, uint64_t code, uint2 meta, int cs) {
  const int L = (int)(meta.y >> 27);
  const int nev = (int)((meta.y >> 1) & 3u);
  const int p_left = nev ? (int)((meta.y >> 3) & 0xFFFu) / cs : 0;
  const ProbeAddr pa = presence_addr<true>(a, win_select(a, L), code, L, p_left);
  return pa.state != 0 || ((ld_probe(pa.word) >> pa.bit) & 1u);
}
__device__ __noinline__ void resolve_chain(const IndexView& ix, const SearchOut& out, bool have, uint64_t code, uint2 meta) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  const int K = (int)ix.K;
  const uint32_t kmask = (K >= 16) ? 0xFFFFFFFFu : ((1u << (2 * K)) - 1u);
  bool alive = false;
  uint32_t l = 0, r = (uint32_t)ix.n;
  if (have) {
    const int L = (int)(meta.y >> 27);
    int t = 0;
    if (L >= K) {
      uint2 iv = __ldg(&ix.kmer[(uint32_t)code & kmask]);
      l = iv.x; r = iv.y; t = K;
    }
    while (l < r && t < L) {
      step_acgt(ix, l, r, (int)((code >> (2 * t)) & 3));
      ++t;
    }
    alive = l < r;
  }
  unsigned am = __ballot_sync(FULL, alive);
  if (am) {
    unsigned int first = 0;
    if (lane == (uint32_t)(__ffs(am) - 1)) first = atomicAdd(out.n_cand, (unsigned int)__popc(am));
    first = __shfl_sync(FULL, first, __ffs(am) - 1);
    if (alive) {
      unsigned int slot = first + (unsigned int)__popc(am & lt);
      if (slot < out.cap) {
        Cand c;
        c.q = meta.x; c.l = l; c.r = r; c.code = meta.y & 0x07FFFFFFu;
        out.cands[slot] = c;
      } else {
        atomicExch(out.overflow, 1u);
      }
    }
  }
}

#ifdef DG_RESOLVE_ASYNC
// EXPERIMENT (north_star: "occ blocks staged through TMA into shared memory").
// The backward search of one string is a chain of dependent steps: the blocks of step t + 1 are only
// known once step t is done, so there is nothing to prefetch WITHIN a chain.  What asynchronous
// staging can buy is a second, independent chain per lane: two batches of 32 queued strings advance
// in lockstep, every step issues the l- and r-block of both batches as cp.async (LDGSTS) copies into
// a per-warp shared-memory slab and only then waits -- twice the requests in flight per warp, the
// rank arithmetic on shared memory.  Measured against the plain version in DESIGN.md section 4.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
struct Chain {
  uint64_t code;
  uint2 meta;
  uint32_t l, r;
  int t, L;
  bool alive;
};
__device__ __noinline__ void resolve_chain2(const IndexView& ix, const SearchOut& out, uint4 (*slab)[4], bool haveA, uint64_t codeA, uint2 metaA,
                                            bool haveB, uint64_t codeB, uint2 metaB) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  const int K = (int)ix.K;
  const uint32_t kmask = (K >= 16) ? 0xFFFFFFFFu : ((1u << (2 * K)) - 1u);
  Chain ch[2];
  ch[0].code = codeA; ch[0].meta = metaA; ch[1].code = codeB; ch[1].meta = metaB;
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    const bool have = b ? haveB : haveA;
    ch[b].L = (int)(ch[b].meta.y >> 27);
    ch[b].l = 0; ch[b].r = (uint32_t)ix.n; ch[b].t = 0;
    if (have && ch[b].L >= K) {
      const uint2 iv = __ldg(&ix.kmer[(uint32_t)ch[b].code & kmask]);
      ch[b].l = iv.x; ch[b].r = iv.y; ch[b].t = K;
    }
    ch[b].alive = have && ch[b].l < ch[b].r;
  }
  uint4* mine = &slab[lane][0];   // [0..1] batch A: l-block planes, r-block planes; [2..3] batch B (counts come by plain loads)
  for (;;) {
    const bool goA = ch[0].alive && ch[0].t < ch[0].L, goB = ch[1].alive && ch[1].t < ch[1].L;
    if (!__any_sync(FULL, goA || goB)) break;
    uint32_t cl[2] = {0, 0}, cr[2] = {0, 0};
    int cc[2] = {0, 0};
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const bool go = b ? goB : goA;
      if (go) {
        cc[b] = (int)((ch[b].code >> (2 * ch[b].t)) & 3);
        const OccBlock* pl = ix.occ + (ch[b].l >> 6);
        const OccBlock* pr = ix.occ + (ch[b].r >> 6);
        cp_async16(&mine[2 * b], reinterpret_cast<const uint4*>(pl) + 1);
        cp_async16(&mine[2 * b + 1], reinterpret_cast<const uint4*>(pr) + 1);
        cl[b] = __ldg(&pl->cnt[cc[b]]);
        cr[b] = __ldg(&pr->cnt[cc[b]]);
      }
    }
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const bool go = b ? goB : goA;
      if (go) {
        const uint4 wl = mine[2 * b], wr = mine[2 * b + 1];
        const uint32_t c4 = ix.C4[cc[b]];
        const uint32_t nl = c4 + cl[b] + rank_planes(ix, ((uint64_t)wl.y << 32) | wl.x, ((uint64_t)wl.w << 32) | wl.z, ch[b].l, cc[b]);
        const uint32_t nr = c4 + cr[b] + rank_planes(ix, ((uint64_t)wr.y << 32) | wr.x, ((uint64_t)wr.w << 32) | wr.z, ch[b].r, cc[b]);
        ch[b].l = nl; ch[b].r = nr; ++ch[b].t;
        ch[b].alive = nl < nr;
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    const bool alive = ch[b].alive;
    const unsigned am = __ballot_sync(FULL, alive);
    if (am) {
      unsigned int first = 0;
      if (lane == (uint32_t)(__ffs(am) - 1)) first = atomicAdd(out.n_cand, (unsigned int)__popc(am));
      first = __shfl_sync(FULL, first, __ffs(am) - 1);
      if (alive) {
        const unsigned int slot = first + (unsigned int)__popc(am & lt);
        if (slot < out.cap) {
          Cand c;
          c.q = ch[b].meta.x; c.l = ch[b].l; c.r = ch[b].r; c.code = ch[b].meta.y & 0x07FFFFFFu;
          out.cands[slot] = c;
        } else {
          atomicExch(out.overflow, 1u);
        }
      }
    }
  }
}
#endif

__device__ __forceinline__ void resolve_queued(const IndexView& ix, const PackedArgs& a, const SearchOut& out, bool have, uint64_t code, uint2 meta,
                                               int cs) {
  if (have && !second_opinion(a, code, meta, cs)) have = false;
  resolve_chain(ix, out, have, code, meta);
}

// k_search_packed.  One warp owns a short run of (query, strand) pairs.  The lanes are POSITIONS of
// the string and the edit kinds are walked in a loop that is uniform across the warp, so building an
// edited string is a handful of shifts without divergence, and the (up to eight) probes of one
// position are independent loads in flight together:
//   single events   the positions of up to eight pairs of equal length are laid side by side
//                   (8 pairs x 20 bases = 5 full passes of 32 lanes);
//   pairs of events (distance 2) for one first position p1 the lanes are (first kind, second
//                   position) combinations; the first event is applied per lane, the second kinds
//                   are walked uniformly.
// The strings enumerated are exactly those of k_search for a clean query (neighbors.h:47-83).
template <bool INDEL>
__global__ void __launch_bounds__(256, DG_PACKED_MIN_BLOCKS) k_search_packed(const __grid_constant__ IndexView ix, const __grid_constant__ PackedArgs a,
                                                                             const __grid_constant__ SearchOut out, uint32_t pairs_per_warp,
                                                                             const uint32_t* __restrict__ skip_if_regular) {
  if (skip_if_regular && !(*skip_if_regular & 1u)) return;   // the regular batch went through k_probe_* / k_resolve
  constexpr int S = INDEL ? 8 : 3;    // enumeration kinds per position of a clean query
  constexpr int CS = INDEL ? 9 : 4;   // canonical slot numbering carried by the candidate
  constexpr unsigned FULL = 0xFFFFFFFFu;
  __shared__ uint64_t q_code[8][64];
  __shared__ uint2 q_meta[8][64];
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t npairs = a.reverse ? 2ULL * a.nq : (uint64_t)a.nq;
  const uint64_t pair_lo = warp * pairs_per_warp;
  const uint64_t pair_hi = pair_lo + pairs_per_warp < npairs ? pair_lo + pairs_per_warp : npairs;
  uint32_t queued = 0;
  unsigned long long my_scripts = 0;

  // survivors of the bitmap -> the per-warp queue; 32 queued strings run the slow path together
  auto push = [&](bool pass, uint64_t code, uint32_t q, uint32_t scode, int L) {
    const unsigned pm = __ballot_sync(FULL, pass);
    if (!pm) return;
    if (pass) {
      const uint32_t slot = queued + (uint32_t)__popc(pm & lt);
      q_code[wib][slot] = code;
      q_meta[wib][slot] = make_uint2(q, scode | ((uint32_t)L << 27));
    }
    queued += (uint32_t)__popc(pm);
    __syncwarp();
    if (queued >= 32) {
      queued -= 32;
      const uint64_t c = q_code[wib][queued + lane];
      const uint2 mt = q_meta[wib][queued + lane];
      __syncwarp();
      resolve_queued(ix, a, out, true, c, mt, CS);
    }
  };
  auto pair_shape = [&](uint64_t pair, uint32_t& q, int& strand, int& m, int& dq) -> bool {
    q = a.reverse ? (uint32_t)(pair >> 1) : (uint32_t)pair;
    strand = a.reverse ? (int)(pair & 1) : 0;
    if (pair >= pair_hi || !(a.qflag[q] & 2)) return false;
    m = a.seed_len ? (int)a.seed_len : (int)(a.off[q + 1] - a.off[q]);
    dq = (int)a.dist[q];
    return true;
  };

  for (uint64_t g0 = pair_lo; g0 < pair_hi;) {
    // a group: consecutive packed pairs of one length and distance (at most 8)
    uint32_t q0;
    int s0, m = 0, dq = 0;
    if (!pair_shape(g0, q0, s0, m, dq)) { ++g0; continue; }
    int g = 1;
    {
      uint32_t ql;
      int sl, ml = 0, dl = 0;
      const bool same = lane < 8 && pair_shape(g0 + lane, ql, sl, ml, dl) && ml == m && dl == dq;
      const unsigned sm = __ballot_sync(FULL, same) & 0xFFu;
      g = __ffs(~sm) - 1;               // leading run of matching pairs (lane 0 always matches)
      if (g < 1) g = 1;
      if (g > 8) g = 8;
    }
    const bool with_base = !(INDEL && dq >= 1);
    // ---- the unedited strings
    if (with_base) {
      const bool have = lane < (uint32_t)g;
      uint32_t q = 0;
      uint64_t code = 0;
      uint32_t scode = 0;
      bool pass = false;
      if (have) {
        const uint64_t pair = g0 + lane;
        q = a.reverse ? (uint32_t)(pair >> 1) : (uint32_t)pair;
        const int strand = a.reverse ? (int)(pair & 1) : 0;
        code = a.qcode[2 * (uint64_t)q + strand];
        scode = pack_script(strand, 0, 0, 0);
        ++my_scripts;
        pass = presence_probe(a, code, m, 0);
      }
      push(pass, code, q, scode, m);
    }
    // ---- single events: lanes = (pair of the group, position)
    if (dq >= 1) {
      const int nslots = g * m;
      for (int s = 0; s < nslots; s += 32) {
        const int slot = s + (int)lane;
        const bool have = slot < nslots;
        const int pi = have ? slot / m : 0;
        const int p = slot - pi * m;
        const uint64_t pair = g0 + (uint64_t)pi;
        const uint32_t q = a.reverse ? (uint32_t)(pair >> 1) : (uint32_t)pair;
        const int strand = a.reverse ? (int)(pair & 1) : 0;
        const uint64_t code0 = a.qcode[2 * (uint64_t)q + strand];
        const EditSite site = edit_site(code0, m - 1 - p);
        const uint32_t passmask = probe_site<INDEL>(a, site, m, p, have);
        if (have) my_scripts += S;
        uint32_t anyk = __reduce_or_sync(FULL, passmask);   // kinds that passed on some lane
#pragma unroll 1
        while (anyk) {
          const int kk = __ffs(anyk) - 1;
          anyk &= anyk - 1;
          int dL, kc;
          const uint64_t code = edit_apply<INDEL>(site, kk, dL, kc);
          push((passmask >> kk) & 1u, code, q, pack_script(strand, 1, p * CS + kc, 0), m + dL);
        }
      }
    }
    // ---- pairs of events: per pair and first position, lanes = (first kind, second position)
    if (dq >= 2) {
      for (int pi = 0; pi < g; ++pi) {
        const uint64_t pair = g0 + (uint64_t)pi;
        const uint32_t q = a.reverse ? (uint32_t)(pair >> 1) : (uint32_t)pair;
        const int strand = a.reverse ? (int)(pair & 1) : 0;
        const uint64_t code0 = a.qcode[2 * (uint64_t)q + strand];
        for (int p1 = 0; p1 < m; ++p1) {
          // second positions p1 .. m-1 (p1 itself only after an insertion)
          const int npos = m - p1;
          const int nslots = S * npos;
          const EditSite site1 = edit_site(code0, m - 1 - p1);
          for (int s = 0; s < nslots; s += 32) {
            const int slot = s + (int)lane;
            const int k1i = slot / npos;
            const int p2 = p1 + (slot - k1i * npos);
            int dL1 = 0, k1c = 0;
            uint64_t code1 = code0;
            bool have = slot < nslots;
            if (have) {
              code1 = edit_apply<INDEL>(site1, k1i, dL1, k1c);
              if (p2 == p1 && k1c < 5) have = false;       // pair_ok: equal positions only after an insertion
            }
            const EditSite site2 = edit_site(code1, m - 1 - p2);
            const int L1 = m + dL1;
            const uint32_t passmask = probe_site<INDEL>(a, site2, L1, p1, have);
            if (have) my_scripts += S;
            uint32_t anyk = __reduce_or_sync(FULL, passmask);   // kinds that passed on some lane
#pragma unroll 1
            while (anyk) {
              const int kk = __ffs(anyk) - 1;
              anyk &= anyk - 1;
              int dL, kc;
              const uint64_t code = edit_apply<INDEL>(site2, kk, dL, kc);
              push((passmask >> kk) & 1u, code, q, pack_script(strand, 2, p1 * CS + k1c, p2 * CS + kc), L1 + dL);
            }
          }
        }
      }
    }
    g0 += (uint64_t)g;
  }
  if (queued) {
    const bool have = lane < queued;
    const uint64_t c = have ? q_code[wib][lane] : 0;
    const uint2 mt = have ? q_meta[wib][lane] : make_uint2(0, 0);
    resolve_queued(ix, a, out, have, c, mt, CS);
  }
  for (int o = 16; o; o >>= 1) my_scripts += __shfl_down_sync(FULL, my_scripts, o);
  if (lane == 0 && my_scripts) atomicAdd(out.n_scripts, my_scripts);
}

// ------------------------------------------------------------------------------------------
// The regular batch (every query ACGT-only, one length m, one distance >= 1 -- the shape of the headline
// workload and of BASELINE configs 2 and 4) takes the same enumeration in TWO kernels:
//   k_probe_*   nothing but the presence probes: one thread per (string, position), eight independent
//               reads in flight per thread, few registers, high occupancy -- a gather kernel that runs
//               at the rate the memory system serves random sectors; one result byte per thread
//               (bit kk = edit kind kk survives);
//   k_resolve   walks the result bytes (~4 % of the bits are set), rebuilds those strings and runs the
//               queue + slow path of k_search_packed on them.
// Both leave at once when k_prepare found the batch irregular (*irregular & 1): k_search_packed, which
// handles any mixture, then does the work instead (and leaves at once when the batch IS regular).
struct ProbeShape {
  int m, dq;
  uint32_t n2;           // pair slots per string: S * m (m + 1) / 2 (first position, first kind, second position >= first)
  uint64_t npairs;       // (query, strand) pairs
  const uint32_t* irregular;
  unsigned long long scripts_per_pair;   // strings enumerated per pair (statistics)
  unsigned long long* n_scripts;
};
template <bool INDEL>
__global__ void __launch_bounds__(256, 6) k_probe_singles(const __grid_constant__ PackedArgs a, const __grid_constant__ ProbeShape sh,
                                                          uint64_t slot0, uint64_t nslots, uint8_t* __restrict__ masks) {
  if (*sh.irregular & 1u) return;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int m = sh.m;
  if (tid >= nslots) return;
  const uint64_t slot = slot0 + tid;
  if (slot == 0) atomicAdd(sh.n_scripts, sh.scripts_per_pair * sh.npairs);
  const uint64_t pair = slot / (uint32_t)m;
  const int p = (int)(slot - pair * (uint64_t)m);
  const uint64_t code0 = a.qcode[pair + (a.reverse ? 0 : pair)];   // qcode holds two codes per query
  const EditSite site = edit_site(code0, m - 1 - p);
  uint32_t mask = probe_site<INDEL>(a, site, m, p, true);
  if (!INDEL && p == 0) {   // Hamming sets hold the unedited string as well: bit 7 of position 0
    if (presence_probe(a, code0, m, 0)) mask |= 0x80u;
  }
  masks[tid] = (uint8_t)mask;
}

template <bool INDEL>
__global__ void __launch_bounds__(256, 5) k_probe_pairs(const __grid_constant__ PackedArgs a, const __grid_constant__ ProbeShape sh,
                                                        uint64_t pair0, uint64_t npairs_tile, uint8_t* __restrict__ masks) {
  constexpr int S = INDEL ? 8 : 3;
  if (*sh.irregular & 1u) return;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npairs_tile * (uint64_t)sh.n2) return;
  const uint64_t pl = t / sh.n2;
  const uint32_t u = (uint32_t)(t - pl * (uint64_t)sh.n2);
  const uint64_t pair = pair0 + pl;
  const int m = sh.m;
  int p1, k1i, p2;
  pair_slot<S>(u, m, p1, k1i, p2);
  const uint64_t code0 = a.qcode[pair + (a.reverse ? 0 : pair)];
  int dL1, k1c;
  const uint64_t code1 = edit_apply<INDEL>(edit_site(code0, m - 1 - p1), k1i, dL1, k1c);
  const bool have = !(p2 == p1 && k1c < 5);       // pair_ok: equal positions only after an insertion
  const EditSite site2 = edit_site(code1, m - 1 - p2);
  masks[t] = (uint8_t)probe_site<INDEL>(a, site2, m + dL1, p1, have);
}

// masks of k_probe_singles (pairs == false: npairs * m bytes) or of one tile of k_probe_pairs -> candidates
template <bool INDEL>
__global__ void __launch_bounds__(256, 4) k_resolve(const __grid_constant__ IndexView ix, const __grid_constant__ PackedArgs a,
                                                    const __grid_constant__ SearchOut out, const __grid_constant__ ProbeShape sh,
                                                    const uint8_t* __restrict__ masks, uint64_t nbytes, uint64_t pair0, int pairs,
                                                    uint64_t slot0) {
  constexpr int S = INDEL ? 8 : 3;
  constexpr int CS = INDEL ? 9 : 4;
  constexpr unsigned FULL = 0xFFFFFFFFu;
  if (*sh.irregular & 1u) return;
  // two queues per warp: survivors of the first probe wait for the second opinion (A); what that leaves
  // (about three in ten) waits for the backward search (B), so that the long dependent chain of the
  // search always runs with 32 live lanes
#ifdef DG_RESOLVE_ASYNC
  constexpr uint32_t kChainBatch = 64;                 // two batches of 32 per chain call
  __shared__ uint4 s_slab[8][32][4];
#else
  constexpr uint32_t kChainBatch = 32;
#endif
  __shared__ uint64_t q_code[8][32], r_code[8][kChainBatch + 32];   // (q_code: 64 four-byte tokens per warp)
  __shared__ uint2 r_meta[8][kChainBatch + 32];
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const int m = sh.m;
  uint32_t queued = 0, rqueued = 0;
  auto stage_b = [&](bool have, uint64_t c, uint2 mt) {   // second opinion of up to 32 strings -> queue B -> chain
    const bool keep = have && second_opinion(a, c, mt, CS);
    const unsigned km = __ballot_sync(FULL, keep);
    if (keep) {
      const uint32_t sl = rqueued + (uint32_t)__popc(km & lt);
      r_code[wib][sl] = c;
      r_meta[wib][sl] = mt;
    }
    rqueued += (uint32_t)__popc(km);
    __syncwarp();
    if (rqueued >= kChainBatch) {
      rqueued -= kChainBatch;
      const uint64_t c2 = r_code[wib][rqueued + lane];
      const uint2 m2 = r_meta[wib][rqueued + lane];
#ifdef DG_RESOLVE_ASYNC
      const uint64_t c3 = r_code[wib][rqueued + 32 + lane];
      const uint2 m3 = r_meta[wib][rqueued + 32 + lane];
      __syncwarp();
      resolve_chain2(ix, out, s_slab[wib], true, c2, m2, true, c3, m3);
#else
      __syncwarp();
      resolve_chain(ix, out, true, c2, m2);
#endif
    }
  };
  // queue A holds TOKENS (result byte index * 8 + kind): the strings are rebuilt 32 at a time, one per lane
  uint32_t* q_tok = reinterpret_cast<uint32_t*>(&q_code[wib][0]);   // 64 tokens
  auto stage_a = [&](bool have, uint32_t tok) {
    uint64_t code = 0;
    uint32_t q = 0, scode = 0;
    int L = 0;
    if (have) {
      const uint64_t slot = tok >> 3;
      const int kk = (int)(tok & 7u);
      if (!pairs) {
        const uint64_t gslot = slot0 + slot;
        const uint64_t pair = gslot / (uint32_t)m;
        const int p = (int)(gslot - pair * (uint64_t)m);
        q = a.reverse ? (uint32_t)(pair >> 1) : (uint32_t)pair;
        const int strand = a.reverse ? (int)(pair & 1) : 0;
        const uint64_t code0 = a.qcode[2 * (uint64_t)q + strand];
        if (kk == 7 && !INDEL) {
          code = code0; L = m; scode = pack_script(strand, 0, 0, 0);
        } else {
          int dL, kc;
          code = edit_apply<INDEL>(edit_site(code0, m - 1 - p), kk, dL, kc);
          L = m + dL;
          scode = pack_script(strand, 1, p * CS + kc, 0);
        }
      } else {
        const uint64_t pl = slot / sh.n2;
        const uint32_t u = (uint32_t)(slot - pl * (uint64_t)sh.n2);
        const uint64_t pair = pair0 + pl;
        q = a.reverse ? (uint32_t)(pair >> 1) : (uint32_t)pair;
        const int strand = a.reverse ? (int)(pair & 1) : 0;
        const uint64_t code0 = a.qcode[2 * (uint64_t)q + strand];
        int p1, k1i, p2, dL1, k1c, dL, kc;
        pair_slot<S>(u, m, p1, k1i, p2);
        const uint64_t code1 = edit_apply<INDEL>(edit_site(code0, m - 1 - p1), k1i, dL1, k1c);
        code = edit_apply<INDEL>(edit_site(code1, m - 1 - p2), kk, dL, kc);
        L = m + dL1 + dL;
        scode = pack_script(strand, 2, p1 * CS + k1c, p2 * CS + kc);
      }
    }
    stage_b(have, code, make_uint2(q, scode | ((uint32_t)L << 27)));
  };
  const uint64_t nwords = (nbytes + 3) >> 2;
  const uint32_t* mw = reinterpret_cast<const uint32_t*>(masks);
  for (uint64_t w0 = warp * 32; w0 < nwords; w0 += nwarps * 32) {
    const uint64_t wi = w0 + lane;
    uint32_t bits = wi < nwords ? __ldg(mw + wi) : 0u;
    if (wi * 4 + 4 > nbytes && wi < nwords) bits &= (1u << (8 * (uint32_t)(nbytes - wi * 4))) - 1u;   // the tail word
    while (__any_sync(FULL, bits != 0)) {   // one set bit per lane and round
      const bool have = bits != 0;
      const unsigned pm = __ballot_sync(FULL, have);
      if (have) {
        const int bpos = __ffs(bits) - 1;
        bits &= bits - 1;
        q_tok[queued + (uint32_t)__popc(pm & lt)] = (uint32_t)(wi * 32 + (uint64_t)bpos);   // (byte index * 8 + kind)
      }
      queued += (uint32_t)__popc(pm);
      __syncwarp();
      if (queued >= 32) {
        queued -= 32;
        const uint32_t tok = q_tok[queued + lane];
        __syncwarp();
        stage_a(true, tok);
      }
    }
  }
  if (queued) {
    const bool have = lane < queued;
    const uint32_t tok = have ? q_tok[lane] : 0u;
    __syncwarp();
    stage_a(have, tok);
  }
#ifdef DG_RESOLVE_ASYNC
  if (rqueued) {
    const bool haveA = lane < rqueued, haveB = lane + 32 < rqueued;
    const uint64_t cA = haveA ? r_code[wib][lane] : 0, cB = haveB ? r_code[wib][lane + 32] : 0;
    const uint2 mA = haveA ? r_meta[wib][lane] : make_uint2(0, 0), mB = haveB ? r_meta[wib][lane + 32] : make_uint2(0, 0);
    resolve_chain2(ix, out, s_slab[wib], haveA, cA, mA, haveB, cB, mB);
  }
#else
  if (rqueued) {
    const bool have = lane < rqueued;
    const uint64_t c = have ? r_code[wib][lane] : 0;
    const uint2 mt = have ? r_meta[wib][lane] : make_uint2(0, 0);
    resolve_chain(ix, out, have, c, mt);
  }
#endif
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

// ------------------------------------------------------------------------------------------
// Radix-sortable image of a candidate: (query, strand) then the neighbour string as 3-bit codes
// (end = 0 < A < C < G < N < T, the byte order of std::string's operator<), 21 characters per
// word.  Strings of up to 42 characters are ordered and de-duplicated by the key alone; batches
// with longer strings take the comparison sort below instead.
struct CandKey {
  uint64_t hi, lo;
  uint32_t qs;    // (query << 1) | strand, or the sentinel (dropped by the antichain rule)
  uint32_t idx;   // position in the unsorted candidate array
};
struct CandKeyDecomposer {
  __host__ __device__ ::cuda::std::tuple<uint32_t&, uint64_t&, uint64_t&> operator()(CandKey& k) const {
    return {k.qs, k.hi, k.lo};
  }
};
constexpr int kKeyChars = 42;

// 2-bit packed string (last base in the low bits, L <= 32 bases) -> 4-bit classes A C G T = 1..4,
// first base in the low nibble (Packed4): reverse the base order, spread every 2-bit field into a
// nibble, add one to the L valid nibbles.
__device__ __forceinline__ uint64_t spread2to4(uint64_t v) {   // 16 bases in the low 32 bits
  v = (v | (v << 16)) & 0x0000FFFF0000FFFFULL;
  v = (v | (v << 8)) & 0x00FF00FF00FF00FFULL;
  v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0FULL;
  v = (v | (v << 2)) & 0x3333333333333333ULL;
  return v;
}
__device__ __forceinline__ Packed4 packed2_to_packed4(uint64_t code, int L) {
  Packed4 p;
  p.w0 = p.w1 = 0;
  if (L <= 0) return p;
  uint64_t r = __brevll(code);
  r = ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);   // base j (from the left) ...
  r >>= 2 * (32 - L);                                                             // ... now at bits 2j
  const uint64_t ones = 0x1111111111111111ULL;
  const uint64_t m0 = L >= 16 ? ~0ULL : ((1ULL << (4 * L)) - 1ULL);
  const uint64_t m1 = L <= 16 ? 0ULL : (L >= 32 ? ~0ULL : ((1ULL << (4 * (L - 16))) - 1ULL));
  p.w0 = spread2to4(r & 0xFFFFFFFFULL) + (ones & m0);
  p.w1 = (spread2to4(r >> 32) + (ones & m1)) & m1;
  p.w0 &= m0;
  return p;
}

// group_cnt / group_rank (may be null): the first half of a counting sort by (query, strand) -- every
// candidate takes the next place of its group; the dropped ones (sentinel) count per warp
__device__ __forceinline__ void key_count(uint32_t qs, uint32_t sentinel, uint32_t i, uint32_t* group_cnt, uint32_t* group_rank) {
  if (!group_cnt) return;
  if (qs != sentinel) { group_rank[i] = atomicAdd(&group_cnt[qs], 1u); return; }
  const unsigned peers = __activemask();   // (all dropped candidates of the warp that reach this point together)
  const uint32_t lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(&group_cnt[sentinel], (uint32_t)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  group_rank[i] = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
}
__global__ void k_cand_keys(BatchDev b, const Cand* __restrict__ cands, uint32_t n, uint32_t sentinel, CandKey* __restrict__ keys,
                            int slow_keys, uint32_t* __restrict__ group_cnt, uint32_t* __restrict__ group_rank) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Cand c = cands[i];
  const bool indel = b.indel != 0;
  int strand;
  Script sc;
  unpack_script(c.code, indel, strand, sc);
  const uint8_t* base;
  int m, koff;
  query_geom(b, c.q, strand, base, m, koff);
  const int d = (int)b.dist[c.q];
  CandKey k;
  k.hi = k.lo = 0;
  bool keep;
  if (cand_is_listed(c.code)) {
    // a string of a host-made list: the list is the set (already an antichain, already in set order)
    k.hi = c.code >> 11;
    k.qs = (c.q << 1) | (c.code & 1u);
    k.idx = i;
    keys[i] = k;
    key_count(k.qs, sentinel, i, group_cnt, group_rank);
    return;
  }
  if ((b.qflag[c.q] & 2) && !slow_keys) {
    // packed path (ACGT-only query, m + d <= 31): the string k_search_packed searched, rebuilt the
    // way it built it, then widened by shifts and masks instead of a walk over the bases
    const uint64_t code0 = b.qcode[2 * (uint64_t)c.q + strand];
    uint64_t code = code0;
    int L = m;
    for (int ev = 0; ev < sc.nev; ++ev) {   // the left event first: an edit never moves anything to its right
      const int kk = sc.k[ev];
      code = apply_event_packed(code, m - 1 - sc.pos[ev], kk);
      L += kk == 4 ? -1 : (kk >= 5 ? 1 : 0);
    }
    const uint64_t al = L > 0 ? code << (64 - 2 * L) : 0;   // base j at bits 62 - 2j
#pragma unroll
    for (int j = 0; j < kMaxPacked; ++j) {
      uint64_t v = (al >> (62 - 2 * j)) & 3ULL;
      v = v + 1ULL + (v == 3ULL ? 1ULL : 0ULL);               // key order: A < C < G < N < T
      if (j >= L) v = 0;
      if (j < 21) k.hi |= v << (60 - 3 * j); else k.lo |= v << (60 - 3 * (j - 21));
    }
    if (!indel) {
      keep = true;
    } else {
      const Packed4 q4 = packed2_to_packed4(code0, m), t4 = packed2_to_packed4(code, L);
      keep = (d == 1 && m >= 2) ? is_minimal_d1(q4, m, t4, L) : is_minimal_band(q4, m, d, t4, L);
    }
  } else
  if (m + d <= 31) {
    // register path (a query holding N): the edited string is generated into 4-bit classes and key codes
    const Packed4 q4 = pack4(base, m);
    Packed4 t4;
    t4.w0 = t4.w1 = 0;
    int L = 0;
    auto emit = [&](uint32_t cls) {   // cls: A C G T N = 1..5
      if (L < 16) t4.w0 |= (uint64_t)cls << (4 * L); else t4.w1 |= (uint64_t)cls << (4 * (L - 16));
      const uint64_t v = cls == 4u ? 5u : (cls == 5u ? 4u : cls);   // key order: A < C < G < N < T
      if (L < 21) k.hi |= v << (60 - 3 * L); else k.lo |= v << (60 - 3 * (L - 21));
      ++L;
    };
    int ev = 0;
    for (int p = 0; p < m; ++p) {
      while (ev < sc.nev && sc.pos[ev] == p && sc.k[ev] >= 5) { emit((uint32_t)(sc.k[ev] - 5) + 1u); ++ev; }
      uint32_t cls = q4.at(p);
      bool out = true;
      if (ev < sc.nev && sc.pos[ev] == p) {
        if (sc.k[ev] == 4) out = false; else cls = (uint32_t)sc.k[ev] + 1u;
        ++ev;
      }
      if (out) emit(cls);
    }
    keep = !indel || (d == 1 && m >= 2 ? is_minimal_d1(q4, m, t4, L) : is_minimal_band(q4, m, d, t4, L));
  } else {
    uint8_t t[kMaxQuery + 8], s0[kMaxQuery + 8], s1[kMaxQuery + 8];
    int L = script_ltr(base, m, sc, t);
    keep = !indel || is_minimal(base, m, d, t, L, s0, s1);
    int lim = L < kKeyChars ? L : kKeyChars;
    for (int j = 0; j < lim; ++j) {
      uint8_t ch = t[j];
      uint64_t v = ch == 'A' ? 1 : ch == 'C' ? 2 : ch == 'G' ? 3 : ch == 'T' ? 5 : 4;
      if (j < 21) k.hi |= v << (60 - 3 * j); else k.lo |= v << (60 - 3 * (j - 21));
    }
  }
  if (b.n_trunc) {
    // a neighbourhood the cap truncated: the set is the replayed list, not the minimal strings
    const uint32_t qs = (c.q << 1) | (uint32_t)strand;
    const uint32_t t = lower_bound_u32(b.trunc_qs, 0, b.n_trunc, qs);
    if (t < b.n_trunc && b.trunc_qs[t] == qs) {
      uint32_t lo = b.trunc_off[t], hi = b.trunc_off[t + 1];
      while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        const ulonglong2 e = b.trunc_keys[mid];
        if (e.x < k.hi || (e.x == k.hi && e.y < k.lo)) lo = mid + 1; else hi = mid;
      }
      keep = false;
      if (lo < b.trunc_off[t + 1]) {
        const ulonglong2 e = b.trunc_keys[lo];
        keep = e.x == k.hi && e.y == k.lo;
      }
    }
  }
  k.qs = keep ? ((c.q << 1) | (uint32_t)strand) : sentinel;
  k.idx = i;
  keys[i] = k;
  key_count(k.qs, sentinel, i, group_cnt, group_rank);
}
// second half of the counting sort: groups laid out by the scanned counts (order inside a group: k_group_order)
__global__ void k_scatter_keys(const CandKey* __restrict__ keys, uint32_t n, const uint32_t* __restrict__ group_off,
                               const uint32_t* __restrict__ group_rank, CandKey* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const CandKey k = keys[i];
  out[group_off[k.qs] + group_rank[i]] = k;
}
__global__ void k_unique_keys(const CandKey* __restrict__ keys, uint32_t n, uint32_t sentinel, uint8_t* __restrict__ keep) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  CandKey c = keys[i];
  bool first = c.qs != sentinel;
  if (first && i > 0) {
    CandKey a = keys[i - 1];
    if (a.qs == c.qs && a.hi == c.hi && a.lo == c.lo) first = false;
  }
  keep[i] = first ? 1 : 0;
}
// After a radix sort on (query, strand) alone: orders every small group by its string key with a
// rank count (groups hold ~1-40 candidates); a group larger than kGroupMax raises `big` and the
// host falls back to the full-key radix sort.
constexpr int kGroupMax = 256;
__global__ void k_group_order(const CandKey* __restrict__ in, uint32_t n, uint32_t sentinel, CandKey* __restrict__ out,
                              unsigned int* __restrict__ big) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const CandKey me = in[i];
  if (me.qs == sentinel) { out[i] = me; return; }
  uint32_t s = i, e = i + 1;
  while (s > 0 && i - s < (uint32_t)kGroupMax && in[s - 1].qs == me.qs) --s;
  while (e < n && e - i < (uint32_t)kGroupMax && in[e].qs == me.qs) ++e;
  if (i - s >= (uint32_t)kGroupMax || e - i >= (uint32_t)kGroupMax) { atomicExch(big, 1u); out[i] = me; return; }
  uint32_t rank = 0;
  for (uint32_t j = s; j < e; ++j) {
    const CandKey o = in[j];
    if (o.hi < me.hi || (o.hi == me.hi && (o.lo < me.lo || (o.lo == me.lo && o.idx < me.idx)))) ++rank;
  }
  out[s + rank] = me;
}
__global__ void k_gather_cands(const Cand* __restrict__ cands, const CandKey* __restrict__ keys, uint32_t n, Cand* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = cands[keys[i].idx];
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
                           unsigned long long* __restrict__ qhits, uint32_t* __restrict__ status,
                           unsigned long long* __restrict__ max_rows) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Cand c = cands[i];
  uint64_t bf = before[i], tk = take[i];
  uint64_t nt = bf < max_loc ? (tk < max_loc - bf ? tk : max_loc - bf) : 0;
  ntake[i] = nt;
  nloc[i] = nt ? (uint64_t)(c.r - c.l) : 0;
  if (nt && c.r - c.l > 1) atomicMax(max_rows, (unsigned long long)(c.r - c.l));
  if (nt) atomicAdd(&qhits[c.q], (unsigned long long)nt);
  if (bf + tk >= max_loc) atomicOr(&status[c.q], (uint32_t)DG_Q_HIT_CAP);  // hunter.h:434-437
}

// Largest i in [0, n) with off[i] <= x (off ascending, off[0] <= x).  Segments are mostly one
// element long, so i is close to x: gallop out from that guess, then bisect the bracket --
// two or three dependent loads instead of log2(n).
__device__ __forceinline__ uint32_t find_segment(const uint64_t* __restrict__ off, uint32_t n, uint64_t x) {
  uint32_t g = x < (uint64_t)n ? (uint32_t)x : n - 1;
  uint32_t lo, hi;   // invariant: off[lo] <= x, (hi == n or off[hi] > x)
  if (off[g] <= x) {
    lo = g;
    uint32_t step = 1;
    hi = n;
    while (lo + step < n) {
      if (off[lo + step] > x) { hi = lo + step; break; }
      lo += step;
      step <<= 1;
    }
  } else {
    hi = g;
    uint32_t step = 1;
    lo = 0;
    while (hi > step) {
      if (off[hi - step] <= x) { lo = hi - step; break; }
      hi -= step;
      step <<= 1;
    }
  }
  while (hi - lo > 1) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (off[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void k_locate(IndexView ix, const Cand* __restrict__ cands, uint32_t n, const uint64_t* __restrict__ loc_off,
                         uint64_t first_row, uint64_t nrows, uint64_t* __restrict__ keys) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nrows) return;
  const uint64_t t = first_row + k;   // row number over the whole batch; keys holds this slice
  // candidate with loc_off[i] <= t < loc_off[i+1] (the last one of a run sharing the offset:
  // candidates without rows are skipped)
  const uint32_t lo = find_segment(loc_off, n, t);
  Cand c = cands[lo];
  uint32_t row = c.l + (uint32_t)(t - loc_off[lo]);
  uint32_t pos = sa_value(ix, row);
  keys[k] = ((uint64_t)lo << 32) | pos;
}

// Ascending text positions inside each candidate's segment of located rows (hunter.h:356) by a
// rank count; used when no candidate holds more than kGroupMax rows (else: radix sort).
__global__ void k_locate_order(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ loc_off, uint64_t first_row,
                               uint64_t nrows, uint64_t* __restrict__ out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nrows) return;
  const uint64_t me = keys[t];
  const uint32_t c = (uint32_t)(me >> 32);
  const uint64_t s = loc_off[c] - first_row, e = loc_off[c + 1] - first_row;
  uint64_t rank = 0;
  if (e - s > 1)
    for (uint64_t j = s; j < e; ++j) {
      const uint64_t o = keys[j];
      if (o < me || (o == me && j < t)) ++rank;
    }
  out[s + rank] = me;
}

struct VerifyArgs {
  const Cand* cands;
  uint32_t ncand;
  const uint64_t* hit_off;   // per candidate (exclusive scan of ntake), ncand + 1 entries
  const uint64_t* loc_off;   // per candidate, ncand + 1 entries
  const uint64_t* keys;      // sorted (candidate << 32 | position) of the rows [key_base, ...)
  uint64_t key_base;
  uint64_t nhits;
  dg_hit* hits;
  dg_rec* recs;              // compact form (hunt): replaces hits + pool
  int4* wire;                // 16-byte wire record per hit (dg_wire): what the multi-GPU all-gather moves
  uint8_t* pool;             // alignment pool: strings back to back, claimed block by block
  unsigned long long* pool_cursor;
  uint8_t* scratch;          // NW scratch for alignments too large for the thread-local paths
  uint32_t scratch_stride;
  uint32_t trace_bytes, srow_ints;
  uint64_t first_hit;        // chunk start
  uint64_t chunk;            // hits in this launch
  PeerOut peer;              // peer mode of a bound communicator: the wire record also goes to every rank's table
};

// one wire record into this rank's place in the table of every rank (NVLink stores; the fence orders
// them before anything this thread -- and, at kernel end, this stream -- does next)
__device__ __forceinline__ void peer_store(const PeerOut& po, uint64_t h, int4 w) {
  if (!po.nranks) return;
  if (h < po.cap) {
    w.x = (int)((uint32_t)w.x + po.query_base);
    for (uint32_t r = 0; r < po.nranks; ++r) po.tab[r][h] = w;
  }
  __threadfence_system();
}

constexpr int kLocalQ = 31;   // thread-local NW: query columns
constexpr int kLocalG = 40;   // thread-local NW: genomic rows
constexpr int kVerifyBlock = 128;

// One thread per hit (hunter.h:358-432 / silica.h:475-573).  Three alignment paths:
//   banded   hunt, edit mode, query <= 31, band <= kBandMax: register-resident banded DP, one trace
//            word per row, alignment written straight to the pool (needle_banded_* in dg_core.cuh)
//   local    anything else that fits 31 x 40: the full matrix in thread-local memory
//   scratch  larger alignments: the full matrix in a global scratch slab
// The pool has no per-hit stride: each block sums the bytes its hits need (one block scan) and
// claims that many with one atomicAdd, so only the bytes that travel to the host are written.
template <bool COMPACT>
__global__ void __launch_bounds__(kVerifyBlock) k_verify(IndexView ix, BatchDev b, VerifyArgs a) {
  __shared__ unsigned long long s_warp[kVerifyBlock / 32];
  __shared__ unsigned long long s_base;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < a.chunk;
  const bool indel = b.indel != 0;
  dg_hit out;
  memset(&out, 0, sizeof(out));
  uint32_t tr[kLocalG + 1];
  uint64_t rows[kLocalG + 1];
  const uint8_t* g = nullptr;
  const uint8_t* base = nullptr;
  int mq = 0, mg = 0, d = 0, mode = 0;  // 1 banded, 2 local, 3 scratch, 4 Hamming
  int nops = 0, lead = 0, trail = 0, kept = 0, score = 0;
  uint32_t chrpos = 0, chrpos0 = 0, bytes = 0;
  uint64_t h = 0;
  if (valid) {
    h = a.first_hit + t;
    const uint32_t lo = find_segment(a.hit_off, a.ncand, h);
    Cand c = a.cands[lo];
    uint64_t j = h - a.hit_off[lo];
    uint64_t pos = a.keys[a.loc_off[lo] - a.key_base + j] & 0xFFFFFFFFULL;
    int strand, koff;
    Script sc;
    unpack_script(c.code, indel, strand, sc);
    query_geom(b, c.q, strand, base, mq, koff);
    const int m = cand_is_listed(c.code) ? (int)((c.code >> 3) & 0xFFu) : script_len(mq, sc);   // neighbour length
    d = (int)b.dist[c.q];
    // hunter.h:358-362
    uint32_t refIndex;
    locate_record(ix.cum, ix.nseq, pos, refIndex, chrpos);
    chrpos0 = chrpos;
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
    g = T + pos - pre;
    mg = (int)(pre + m + post);
    if (b.seed_len ? (pre <= chrpos) : (pre < chrpos)) chrpos -= (uint32_t)pre;  // silica.h:501 / hunter.h:382
    out.query = c.q;
    out.chr = refIndex;
    out.text_pos = pos;
    out.strand = strand ? '-' : '+';
    if (b.seed_len || indel) {
      const int band_hi = mg - mq + d;
      if (!b.seed_len && mq <= kLocalQ && mg <= kLocalG && band_hi >= 0 && band_hi + d + 1 <= kBandMax) {
        mode = 1;
        score = needle_banded_fill(g, mg, base, mq, d, tr);
        needle_banded_shape(tr, mg, mq, d, &nops, &lead, &trail);
        kept = nops - lead - trail;
      } else if (mq <= kLocalQ && mg <= kLocalG) {
        mode = 2;
      } else {
        mode = 3;
      }
      bytes = b.seed_len ? (uint32_t)mg : 0u;  // modes 2 / 3 (hunt) learn their size below
    } else {
      mode = 4;
      bytes = (uint32_t)(mg + mq);
    }
  }
  // modes 2 and 3 run the full-matrix DP before the pool offset is known: keep the alignment in
  // thread-local / scratch memory and copy it afterwards
  uint8_t ra[kLocalQ + kLocalG + 1], qa[kLocalQ + kLocalG + 1];
  uint8_t *ra_p = ra, *qa_p = qa;
  if (mode == 2) {
    int srow[kLocalQ + 1];
    uint8_t ops[kLocalQ + kLocalG + 1];
    for (int i = 0; i <= mg; ++i) rows[i] = 0;
    kept = needle_align(g, mg, base, mq, TraceRows64{rows}, srow, ops, ra, qa, &lead, &score);
    if (!b.seed_len) bytes = 2u * (uint32_t)kept;
  } else if (mode == 3) {
    uint8_t* scr = a.scratch + t * (uint64_t)a.scratch_stride;
    int* srow = (int*)scr;
    uint8_t* trace = scr + a.srow_ints * 4;
    uint8_t* ops = trace + a.trace_bytes;
    ra_p = ops + (mg + mq + 4);
    qa_p = ra_p + (mg + mq + 4);
    kept = needle_align(g, mg, base, mq, TraceBytes{trace, mq + 1}, srow, ops, ra_p, qa_p, &lead, &score);
    if (!b.seed_len) bytes = 2u * (uint32_t)kept;
  } else if (mode == 1) {
    bytes = 2u * (uint32_t)kept;
  }
  if (COMPACT) {
    // hunt: the record carries the alignment as at most kRecOps edit operations (dg_core.cuh)
    if (!valid) return;
    uint64_t ops = 0;
    int nop = 0;
    uint32_t start;
    if (mode == 4) {
      int lim = mg < mq ? mg : mq;
      for (int i = lim - 1; i >= 0; --i)
        if (g[i] != base[i]) { ops = (ops << kRecOpBits) | rec_op(i, 1, g[i]); ++nop; }
      ops &= 0x0FFFFFFFFFFFFFFFULL;
      score = -nop;
      start = chrpos + 1;
    } else {
      if (mode == 1) ops = needle_banded_ops(tr, g, mg, base, mq, d, nops, lead, trail, &nop);
      else ops = rows_to_ops(ra_p, qa_p, kept, &nop);
      start = chrpos + (uint32_t)lead + 1;   // hunter.h:399,402
    }
    const int delta = (int)(start - 1) - (int)chrpos0;   // in [-d, 2 d]
    dg_rec rec;
    rec.query = out.query;
    rec.chr = out.chr;
    rec.start = start;
    rec.score = (int16_t)score;
    rec.strand = out.strand;
    rec.nops = (uint8_t)(nop > kRecOps ? 255 : nop);
    rec.ops = ops | ((uint64_t)((delta + 8) & 15) << 60);
    a.recs[h] = rec;
    const int4 w = make_int4((int)rec.query, (int)rec.chr, (int)rec.start, (int)(((uint32_t)score & 0xFFFFu) | ((uint32_t)rec.strand << 16)));
    a.wire[h] = w;
    peer_store(a.peer, h, w);
    return;
  }
  // block-wide exclusive scan of `bytes`, one atomicAdd per block
  unsigned long long incl = bytes;
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= (uint32_t)o) incl += up;
  }
  if (lane == 31) s_warp[wib] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tot = 0;
    for (int w = 0; w < kVerifyBlock / 32; ++w) { unsigned long long x = s_warp[w]; s_warp[w] = tot; tot += x; }
    s_base = tot ? atomicAdd(a.pool_cursor, tot) : 0ULL;
  }
  __syncthreads();
  if (!valid) return;
  out.aln_off = s_base + s_warp[wib] + (incl - bytes);
  uint8_t* slot = a.pool + out.aln_off;
  if (mode == 4) {
    // needleScore (hunter.h:79-88): mismatches over min(|genomic|, |query|); the record carries
    // refalign = genomicseq and queryalign = the query
    int sc2 = 0;
    int lim = mg < mq ? mg : mq;
    for (int i = 0; i < lim; ++i) if (g[i] != base[i]) --sc2;
    for (int i = 0; i < mg; ++i) slot[i] = g[i];
    for (int i = 0; i < mq; ++i) slot[mg + i] = base[i];
    out.aln_len = (uint32_t)mg;
    out.score = sc2;
    out.start = chrpos + 1;
    out.alignpos = out.start;
  } else {
    out.score = score;
    if (b.seed_len) {
      // search: genomic context for the Tm gate + alignpos (silica.h:522-532)
      for (int i = 0; i < mg; ++i) slot[i] = g[i];
      out.aln_len = (uint32_t)mg;
      out.start = chrpos;
      out.alignpos = chrpos + (uint32_t)lead;
    } else {
      if (mode == 1) needle_banded_emit(tr, g, mg, base, mq, d, nops, lead, trail, slot, slot + kept);
      else for (int i = 0; i < kept; ++i) { slot[i] = ra_p[i]; slot[kept + i] = qa_p[i]; }
      out.aln_len = (uint32_t)kept;
      out.start = chrpos + (uint32_t)lead + 1;  // hunter.h:399,402
      out.alignpos = out.start;
    }
  }
  a.hits[h] = out;
  const int4 w = make_int4((int)out.query, (int)out.chr, (int)out.start, (int)(((uint32_t)out.score & 0xFFFFu) | ((uint32_t)out.strand << 16)));
  a.wire[h] = w;
  peer_store(a.peer, h, w);
}

__global__ void k_fill_offsets(uint64_t* __restrict__ off, uint64_t n, uint64_t len) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) off[i] = i * len;
}

__global__ void k_count(const Cand* __restrict__ cands, uint32_t n, unsigned long long* __restrict__ counts) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Cand c = cands[i];
  atomicAdd(&counts[c.q], (unsigned long long)(c.r - c.l));
}

// chunked dg_hunt_batch: chunk-local query ids, pool offsets and hit offsets -> batch-global
// ... and the 16-byte wire record of every hit (dg_index_wire_records)
__global__ void k_rebase(dg_hit* __restrict__ hits, uint64_t nhits, uint64_t* __restrict__ qoff, uint32_t nq1,
                         uint32_t q0, uint64_t pool_base, uint64_t hit_base, int4* __restrict__ wire) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nhits) {
    dg_hit h = hits[i];
    h.query += q0;
    h.aln_off += pool_base;
    hits[i].query = h.query;
    hits[i].aln_off = h.aln_off;
    if (wire) wire[i] = make_int4((int)h.query, (int)h.chr, (int)h.start, (int)(((uint32_t)h.score & 0xFFFFu) | ((uint32_t)h.strand << 16)));
  }
  if (i < nq1) qoff[i] += hit_base;
}

// the compact form of the same: query ids -> batch-global, wire record = the record's first 16 bytes
__global__ void k_rebase_recs(dg_rec* __restrict__ recs, uint64_t nhits, uint32_t q0, int4* __restrict__ wire, PeerOut peer,
                              uint64_t hit_base) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nhits) return;
  const uint2* p = reinterpret_cast<const uint2*>(recs + i);   // (24-byte records: 8-byte aligned)
  const uint2 a = p[0], c = p[1];   // query, chr | start, score | strand << 16 | nops << 24
  const uint32_t q = a.x + q0;
  if (q0) recs[i].query = q;
  const int4 w = make_int4((int)q, (int)a.y, (int)c.x, (int)(c.y & 0x00FFFFFFu));
  if (wire) wire[i] = w;
  peer_store(peer, hit_base + i, w);
}
__global__ void k_pack_qmeta(const uint32_t* __restrict__ status, const uint32_t* __restrict__ dist, uint32_t nq,
                             uint16_t* __restrict__ out, const uint64_t* __restrict__ qcode, uint64_t* __restrict__ qpack) {
  uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  out[q] = (uint16_t)((status[q] & 0xFFu) | ((dist[q] & 0xFFu) << 8));
  if (qpack) qpack[q] = qcode[2 * (uint64_t)q];   // the forward code (meaningful when the batch is regular)
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

// listed neighbourhoods (distance 3): literal backward search of every string of the lists
__global__ void k_search_listed(IndexView ix, ListedStrings ls, SearchOut out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ls.n) return;
  const uint32_t o = ls.off[i];
  const int L = (int)(ls.off[i + 1] - o);
  uint32_t l = 0, r = (uint32_t)ix.n;
  for (int j = L - 1; j >= 0 && l < r; --j) backward_step(ix, l, r, ls.chars[o + j]);
  if (l < r && L > 0) {
    const unsigned int slot = atomicAdd(out.n_cand, 1u);
    if (slot < out.cap) {
      Cand c;
      c.q = ls.q[i]; c.l = l; c.r = r; c.code = ls.code[i];
      out.cands[slot] = c;
    } else {
      atomicExch(out.overflow, 1u);
    }
  }
}

// unit table of one search-string length (host)
void build_unit_table(int m, int d, bool indel, bool clean, std::vector<uint32_t>& tab) {
  int S = enum_slots(indel, clean), E = S * m;
  bool with_base = !(indel && d >= 1);
  tab.clear();
  int n0 = (with_base ? 1 : 0) + (d >= 1 ? E : 0);  // row 0: (the unedited string +) the single events
  for (int s = 0; s < n0; s += 32) tab.push_back((0u << 12) | (uint32_t)s);
  if (d >= 2) {
    for (int e1 = 0; e1 < E; ++e1) {
      int p1 = e1 / S, kk = e1 - p1 * S;
      int start = (enum_is_ins(indel, clean, kk) ? p1 : p1 + 1) * S;
      for (int s = start; s < E; s += 32) tab.push_back(((uint32_t)(e1 + 1) << 12) | (uint32_t)s);
    }
  }
}

}  // namespace

// called by dg_index_close before the streams are destroyed
void release_stream_pools(const cudaStream_t* streams, int n) {
  for (int i = 0; i < n; ++i) {
    if (!streams[i]) continue;
    StreamPool* p = nullptr;
    {
      std::lock_guard<std::mutex> g(g_pools_mu);
      auto it = g_pools.find(streams[i]);
      if (it != g_pools.end()) { p = it->second; g_pools.erase(it); }
    }
    {
      AuxStream* a = nullptr;
      {
        std::lock_guard<std::mutex> g(g_aux_mu);
        auto it = g_aux.find(streams[i]);
        if (it != g_aux.end()) { a = it->second; g_aux.erase(it); }
      }
      if (a) {
        cudaStreamSynchronize(a->s);
        for (auto& e : a->ev) if (e) cudaEventDestroy(e);
        cudaStreamDestroy(a->s);
        delete a;
      }
    }
    if (!p) continue;
    cudaStreamSynchronize(streams[i]);
    for (auto& kv : p->free_blocks) cudaFree(kv.second);
    delete p;
  }
}

}  // namespace dg

using namespace dg;

struct dg_batch;
namespace dg {
int batch_wire_view(dg_batch* b, const int4** wire, uint64_t* n, cudaStream_t* st, dg_index** ix);
}

// ============================================================================================
// Host-side result storage.  Results fetched from the device land in page-locked memory (so the
// D2H copies run at full PCIe rate); freed blocks are kept in a small process-wide cache because
// cudaHostAlloc costs milliseconds.
namespace {
struct HostBlock { void* p; size_t cap; bool pinned; };
struct PinCache {
  std::mutex mu;
  std::vector<HostBlock> free_blocks;
  size_t cached = 0;
  bool get(size_t need, HostBlock& out) {
    std::lock_guard<std::mutex> g(mu);
    size_t best = free_blocks.size();
    for (size_t i = 0; i < free_blocks.size(); ++i) {
      const HostBlock& f = free_blocks[i];
      if (f.cap < need || f.cap > 4 * need + (1u << 20)) continue;
      if (best == free_blocks.size() || (f.pinned && !free_blocks[best].pinned) ||
          (f.pinned == free_blocks[best].pinned && f.cap < free_blocks[best].cap)) best = i;
    }
    if (best == free_blocks.size()) return false;
    out = free_blocks[best];
    cached -= out.cap;
    free_blocks.erase(free_blocks.begin() + best);
    return true;
  }
  void put(const HostBlock& b) {
    std::lock_guard<std::mutex> g(mu);
    if (cached + b.cap > (3ULL << 30) || free_blocks.size() >= 64) {
      if (b.pinned) cudaFreeHost(b.p); else free(b.p);
      return;
    }
    free_blocks.push_back(b);
    cached += b.cap;
  }
};
PinCache g_pin;

// Result storage on the host.  Page-locked when the driver grants it; when it does not (locked-
// memory limits of the container) the block is ordinary memory, still recycled through the cache
// so that a steady stream of batches neither re-pins nor re-faults its result buffers.
struct HostBuf {
  void* p = nullptr;
  size_t bytes = 0, cap = 0;
  bool pinned = false;
  HostBuf() = default;
  HostBuf(const HostBuf&) = delete;
  HostBuf& operator=(const HostBuf&) = delete;
  ~HostBuf() { release(); }
  void alloc(size_t n, bool pin) {
    release();
    bytes = n;
    if (!n) return;
    HostBlock b;
    if (g_pin.get(n, b)) { p = b.p; cap = b.cap; pinned = b.pinned; return; }
    cap = n + (n >> 3) + 4096;
    if (pin) {
      if (cudaHostAlloc(&p, cap, cudaHostAllocDefault) == cudaSuccess) { pinned = true; return; }
      p = nullptr;
      cudaGetLastError();
      static bool warned = false;
      if (!warned && getenv("DG_TRACE")) { warned = true; fprintf(stderr, "[dicey_b200] cudaHostAlloc(%zu) failed: results land in pageable memory\n", cap); }
    }
    p = malloc(cap);
    pinned = false;
    if (!p) throw std::bad_alloc();
  }
  void release() {
    if (p) g_pin.put(HostBlock{p, cap, pinned});
    p = nullptr; bytes = cap = 0; pinned = false;
  }
  // capacity >= need, keeping the first `used` bytes (the caller has drained copies into p)
  void reserve(size_t need, size_t used, bool pin) {
    if (need <= cap) return;
    HostBuf nb;
    nb.alloc(need + (need >> 2), pin);
    if (used) memcpy(nb.p, p, used);
    std::swap(p, nb.p); std::swap(cap, nb.cap); std::swap(pinned, nb.pinned);
    nb.bytes = 0;
  }
};
}  // namespace

// A result is either FULL (dg_hit records + alignment pool, per-query offsets / status / distance:
// `search` results and unpacked ones) or COMPACT (`hunt`: dg_rec records + one 16-bit word per query
// came from the device; everything the older accessors hand out is derived on the host on first use).
struct dg_result {
  HostBuf hits, qoff, status, dist, pool, seqs;
  HostBuf recs, qmeta;            // compact form: dg_rec[nhits], uint16 status | distance << 8 per query
  bool compact = false;
  uint64_t nhits = 0;
  uint32_t nq = 0;
  uint64_t transfer_bytes = 0;    // device -> host bytes of this result
  uint64_t uniform_len = 0;       // > 0: every query has this length; else seq_off holds nq + 1 offsets
  std::vector<uint64_t> seq_off;
  std::vector<uint64_t> cum;      // record starts of the index (text_pos of an expanded dg_hit)
  // sequences of chunks whose queries were all ACGT-only and of one length came from the device as
  // 2-bit codes (8 bytes per query instead of its characters); decoded into `seqs` on first use
  struct PackedSeqs { uint32_t q0, q1, len; };
  std::vector<PackedSeqs> packed_seqs;
  HostBuf qpack;                  // uint64 per query (only the entries of packed chunks are meaningful)
  mutable std::mutex mu;
  mutable bool have_qinfo = false, have_hits = false, have_seqs = true;
  uint64_t qstart(uint32_t q) const { return uniform_len ? (uint64_t)q * uniform_len : seq_off[q]; }
  uint32_t qlen(uint32_t q) const { return uniform_len ? (uint32_t)uniform_len : (uint32_t)(seq_off[q + 1] - seq_off[q]); }
};

namespace {
// the characters of the chunks that travelled as 2-bit codes
void ensure_seqs(const dg_result* cr) {
  dg_result* r = const_cast<dg_result*>(cr);
  if (r->have_seqs) return;
  std::lock_guard<std::mutex> g(r->mu);
  if (r->have_seqs) return;
  const uint64_t* codes = (const uint64_t*)r->qpack.p;
  uint8_t* out = (uint8_t*)r->seqs.p;
  for (const auto& ch : r->packed_seqs) {
    const uint32_t L = ch.len;
    auto work = [&](uint32_t lo, uint32_t hi) {
      for (uint32_t q = lo; q < hi; ++q) {
        const uint64_t code = codes[q];
        uint8_t* s = out + r->qstart(q);
        for (uint32_t j = 0; j < L; ++j) s[j] = (uint8_t)"ACGT"[(code >> (2 * (L - 1 - j))) & 3];
      }
    };
    const uint32_t n = ch.q1 - ch.q0;
    const unsigned nt = (unsigned)std::max<uint32_t>(1, std::min<uint32_t>(8, n / 32768));
    if (nt <= 1) {
      work(ch.q0, ch.q1);
    } else {
      std::vector<std::thread> ts;
      for (unsigned t = 0; t < nt; ++t) ts.emplace_back(work, ch.q0 + (uint32_t)((uint64_t)n * t / nt), ch.q0 + (uint32_t)((uint64_t)n * (t + 1) / nt));
      for (auto& t : ts) t.join();
    }
  }
  r->have_seqs = true;
}

// the strand's search string of query q: the normalised query, or its reverse complement
inline void strand_query(const dg_result* r, uint32_t q, bool minus, std::vector<uint8_t>& out) {
  ensure_seqs(r);
  const uint8_t* s = (const uint8_t*)r->seqs.p + r->qstart(q);
  const uint32_t m = r->qlen(q);
  out.resize(m);
  if (!minus) { memcpy(out.data(), s, m); return; }
  for (uint32_t i = 0; i < m; ++i) {
    const uint8_t ch = s[m - 1 - i];
    out[i] = ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : ch == 'T' ? 'A' : 'N';
  }
}

// per-query offsets / status / distance of a compact result (records are in query order)
void ensure_qinfo(const dg_result* cr) {
  dg_result* r = const_cast<dg_result*>(cr);
  if (!r->compact) return;
  std::lock_guard<std::mutex> g(r->mu);
  if (r->have_qinfo) return;
  const uint32_t nq = r->nq;
  r->qoff.alloc(((size_t)nq + 1) * 8, false);
  r->status.alloc((size_t)nq * 4, false);
  r->dist.alloc((size_t)nq * 4, false);
  uint64_t* qoff = (uint64_t*)r->qoff.p;
  const dg_rec* recs = (const dg_rec*)r->recs.p;
  uint64_t i = 0;
  for (uint32_t q = 0; q < nq; ++q) {
    qoff[q] = i;
    while (i < r->nhits && recs[i].query == q) ++i;
  }
  qoff[nq] = r->nhits;
  const uint16_t* qm = (const uint16_t*)r->qmeta.p;
  for (uint32_t q = 0; q < nq; ++q) {
    ((uint32_t*)r->status.p)[q] = qm[q] & 0xFFu;
    ((uint32_t*)r->dist.p)[q] = qm[q] >> 8;
  }
  r->have_qinfo = true;
}

// dg_hit + pool of a compact result
void ensure_hits(const dg_result* cr) {
  dg_result* r = const_cast<dg_result*>(cr);
  if (!r->compact) return;
  ensure_seqs(r);
  std::lock_guard<std::mutex> g(r->mu);
  if (r->have_hits) return;
  const uint64_t n = r->nhits;
  const dg_rec* recs = (const dg_rec*)r->recs.p;
  std::vector<uint64_t> off((size_t)n + 1, 0);
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t cols = r->qlen(recs[i].query);
    for (int k = 0; k < recs[i].nops && k < kRecOps; ++k) cols += ((recs[i].ops >> (kRecOpBits * k + 10)) & 3) == 3;
    off[i + 1] = off[i] + 2ull * cols;
  }
  r->hits.alloc(n * sizeof(dg_hit), false);
  r->pool.alloc(off[n], false);
  dg_hit* hits = (dg_hit*)r->hits.p;
  uint8_t* pool = (uint8_t*)r->pool.p;
  const unsigned nthreads = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(8, n / 65536));
  auto work = [&](uint64_t lo, uint64_t hi) {
    std::vector<uint8_t> sq, ra, qa;
    for (uint64_t i = lo; i < hi; ++i) {
      const dg_rec& c = recs[i];
      strand_query(r, c.query, c.strand == '-', sq);
      const int m = (int)sq.size();
      ra.resize((size_t)m + kRecOps + 1);
      qa.resize((size_t)m + kRecOps + 1);
      const int cols = rec_expand_rows(c.ops, c.nops <= kRecOps ? c.nops : 0, sq.data(), m, ra.data(), qa.data());
      dg_hit h;
      memset(&h, 0, sizeof(h));
      h.query = c.query; h.score = c.score; h.chr = c.chr; h.start = c.start; h.strand = c.strand;
      h.alignpos = c.start;
      h.aln_off = off[i];
      h.aln_len = (uint32_t)cols;
      const int delta = (int)((c.ops >> 60) & 15) - 8;
      const uint64_t rec0 = c.chr < r->cum.size() ? r->cum[c.chr] : 0;
      h.text_pos = rec0 + (uint64_t)((int64_t)c.start - 1 - delta);
      hits[i] = h;
      memcpy(pool + off[i], ra.data(), (size_t)cols);
      memcpy(pool + off[i] + cols, qa.data(), (size_t)cols);
    }
  };
  if (nthreads <= 1) {
    work(0, n);
  } else {
    std::vector<std::thread> ts;
    for (unsigned t = 0; t < nthreads; ++t) ts.emplace_back(work, n * t / nthreads, n * (t + 1) / nthreads);
    for (auto& t : ts) t.join();
  }
  r->have_hits = true;
}

// query offsets of the caller -> the result (equal-length batches keep one number)
void keep_offsets(dg_result* r, const uint64_t* offsets, uint32_t nq) {
  r->uniform_len = 0;
  r->seq_off.clear();
  if (!nq) return;
  const uint64_t L0 = offsets[1];
  uint64_t bad = offsets[0];
  for (uint32_t q = 0; q <= nq; ++q) bad |= offsets[q] ^ ((uint64_t)q * L0);
  if (!bad && L0) { r->uniform_len = L0; return; }
  r->seq_off.assign(offsets, offsets + nq + 1);
}
}  // namespace

struct dg_batch {
  dg_index* ix = nullptr;
  cudaStream_t st = nullptr;   // every kernel, allocation and copy of this batch
  dg_params par;
  uint32_t nq = 0;
  uint64_t nbytes = 0;
  bool counts_only = false;
  // inputs
  ABuf<uint8_t> raw, fwd, rc;
  ABuf<uint64_t> off, units, unit_off;
  ABuf<uint32_t> status, dist, irregular;
  std::shared_ptr<TabEntry> tabs;   // unit tables (cached in the index)
  bool maybe_capped = false;        // some query length could reach the cap -x: certify / replay after k_prepare
  bool peer_from_verify = false;    // dg_batch_run of a staged batch: k_verify feeds a bound communicator's peer tables
  ABuf<uint32_t> trunc_qs, trunc_off;
  ABuf<ulonglong2> trunc_keys;
  uint32_t n_trunc = 0;
  ABuf<uint8_t> ls_chars;           // listed neighbourhoods (distance 3)
  ABuf<uint32_t> ls_off, ls_q, ls_code;
  uint32_t n_listed = 0;
  ABuf<uint64_t> qcode;
  ABuf<uint8_t> qflag;
  uint32_t uniform_len = 0;
  uint64_t uniform_units = 0;
  int max_len = 0, min_len = 0;
  uint64_t uniform_host_len = 0;        // the caller's offsets, for the result: one length, or a copy
  std::vector<uint64_t> host_off;
  uint32_t h_irregular = 0;             // host copy of *irregular after the search stage (bit 3: normalisation changed a sequence)
  // outputs of run()
  ABuf<Cand> cands;
  uint32_t ncand = 0;
  ABuf<dg_hit> hits;
  ABuf<dg_rec> recs;           // compact form (hunt): instead of hits + pool
  ABuf<uint16_t> qmeta;        // compact form: status | distance << 8 per query
  ABuf<uint64_t> qpack;        // compact form: the forward 2-bit code per query (the sequence of a regular batch)
  bool compact = false;
  ABuf<int4> wire;
  ABuf<uint8_t> pool;
  ABuf<unsigned long long> qhits;
  ABuf<uint64_t> qoff;
  ABuf<unsigned long long> counts;
  uint64_t nhits = 0;
  uint64_t pool_bytes = 0;   // compacted alignment pool
  uint32_t pool_stride = 0;
  bool ran = false;
};

int dg::batch_wire_view(dg_batch* b, const int4** wire, uint64_t* n, cudaStream_t* st, dg_index** ix) {
  if (!b || !b->ran || b->counts_only) { set_error("batch has not run"); return DG_ERR_ARG; }
  *wire = b->nhits ? b->wire.p : nullptr;
  *n = b->nhits;
  *st = b->st;
  *ix = b->ix;
  return DG_OK;
}

static BatchDev batch_dev(const dg_batch* b) {
  BatchDev d;
  d.fwd = b->fwd.p; d.rc = b->rc.p; d.off = b->off.p; d.nq = b->nq; d.status = b->status.p; d.dist = b->dist.p;
  d.seed_len = b->par.seed_len; d.distance = b->par.distance; d.max_loc = b->par.max_locations;
  d.max_nbr = b->par.max_neighborhood ? b->par.max_neighborhood : 10000;
  d.qcode = b->qcode.p; d.qflag = b->qflag.p; d.irregular = b->irregular.p; d.uniform_len = b->uniform_len;
  d.indel = b->par.indel; d.reverse = b->par.reverse;
  d.trunc_qs = b->trunc_qs.p; d.trunc_off = b->trunc_off.p; d.trunc_keys = b->trunc_keys.p; d.n_trunc = b->n_trunc;
  {
    const int maxq = b->par.seed_len ? (int)b->par.seed_len : std::min(b->max_len, kMaxQuery);
    d.key_sortable = (maxq + (int)std::min<uint32_t>(b->par.distance, (uint32_t)std::max(maxq - 1, 0)) <= 42 && !getenv("DG_MERGE_SORT")) ? 1u : 0u;
  }
  return d;
}

// d_seqs / ready: the sequence bytes are already in device memory (uploaded by the chunk pipeline's
// uploader; `ready` is recorded after that copy) -- the host pointer is then only scanned.
static int stage_impl(dg_index* ix, const char* seqs, const uint64_t* offsets, uint32_t nq, const dg_params* par,
                      dg_batch** out, cudaStream_t on_stream = nullptr, const uint8_t* d_seqs = nullptr,
                      cudaEvent_t ready = nullptr, bool keep_host_offsets = true, uint64_t uniform_L = 0) {
  // uniform_L > 0: the caller has checked that every query has this length; offsets is not read
  uint64_t two_off[2] = {0, uniform_L};
  const uint64_t total_bytes = uniform_L ? (uint64_t)nq * uniform_L : (offsets ? offsets[nq] : 0);
  if (uniform_L) offsets = two_off;
  if (!ix || !offsets || !par || !out || (nq && !seqs)) { set_error("null argument"); return DG_ERR_ARG; }
  // (distances beyond kMaxListDist are not refused here: the reference clamps -d per query to |seq| - 1,
  // and what stays too large is flagged DG_Q_UNSUPPORTED per query by k_prepare)
  if (par->seed_len > (uint32_t)kMaxQuery || (par->seed_len && par->distance >= par->seed_len)) {
    set_error("bad seed length");
    return DG_ERR_ARG;
  }
  if (offsets[0] != 0) { set_error("offsets[0] must be 0"); return DG_ERR_ARG; }
  dg_batch* b = new dg_batch();
  static const bool trace = getenv("DG_TRACE") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double ts0 = now();
  try {
    DG_CUDA(cudaSetDevice(ix->device));
    cudaStream_t st = on_stream ? on_stream : ix->stream;
    b->ix = ix;
    b->st = st;
    b->par = *par;
    b->nq = nq;
    b->nbytes = total_bytes;
    // distinct search-string lengths -> unit tables
    bool have[256];
    memset(have, 0, sizeof(have));
    int minL = 1 << 30, maxL = 0;
    // equal-length batches (the common case) are recognised by one branch-free pass
    bool equal_len = nq > 0;
    if (nq) {
      const uint64_t L0 = offsets[1];
      uint64_t bad = 0;
      if (!uniform_L) for (uint32_t q = 0; q <= nq; ++q) bad |= offsets[q] ^ ((uint64_t)q * L0);
      equal_len = bad == 0;
      if (equal_len) {
        minL = maxL = L0 > 100000 ? 100000 : (int)L0;
        if (!par->seed_len && L0 <= (uint64_t)kMaxQuery) have[L0] = true;
      }
    }
    for (uint32_t q = 0; q < nq && !equal_len; ++q) {
      if (offsets[q + 1] < offsets[q]) { set_error("offsets must be non-decreasing"); delete b; return DG_ERR_ARG; }
      uint64_t L = offsets[q + 1] - offsets[q];
      int Li = L > 100000 ? 100000 : (int)L;
      minL = std::min(minL, Li);
      maxL = std::max(maxL, Li);
      if (!par->seed_len && L <= (uint64_t)kMaxQuery) have[L] = true;
    }
    if (par->seed_len) have[par->seed_len] = true;
    if (nq == 0) { minL = maxL = 0; }
    if (equal_len && offsets[1] > 0) b->uniform_host_len = offsets[1];
    else if (keep_host_offsets) b->host_off.assign(offsets, offsets + nq + 1);
    b->min_len = minL;
    b->max_len = maxL;
    const bool indel = par->indel != 0;
    {
      std::string key((const char*)have, sizeof(have));
      key.push_back((char)par->distance);
      key.push_back((char)indel);
      std::lock_guard<std::mutex> g(ix->tab_mu);
      auto it = ix->tab_cache.find(key);
      if (it != ix->tab_cache.end()) {
        b->tabs = it->second;
      } else {
        std::vector<uint32_t> tab, tab_off(512, 0), tab_cnt(512, 0), sub(256, 0), one;
        for (int m = 1; m < 256; ++m) {
          if (!have[m]) continue;
          int d = std::min<int>((int)par->distance, m - 1);
          for (int variant = 0; variant < 2; ++variant) {
            build_unit_table(m, d, indel, variant == 0, one);
            tab_off[variant * 256 + m] = (uint32_t)tab.size();
            tab_cnt[variant * 256 + m] = (uint32_t)one.size();
            tab.insert(tab.end(), one.begin(), one.end());
          }
          {
            int sl = slots_per_pos(indel), E = sl * m;
            uint64_t ub = 1 + (d >= 1 ? (uint64_t)E : 0);
            if (d >= 2)
              for (int e1 = 0; e1 < E; ++e1) ub += (uint64_t)(E - second_event_start(e1 / sl, e1 % sl, indel));
            sub[m] = ub > 0xFFFFFFFFULL ? 0xFFFFFFFFu : (uint32_t)ub;
          }
        }
        std::vector<uint32_t> blob;
        blob.reserve(1280 + tab.size());
        blob.insert(blob.end(), sub.begin(), sub.end());
        blob.insert(blob.end(), tab_off.begin(), tab_off.end());
        blob.insert(blob.end(), tab_cnt.begin(), tab_cnt.end());
        blob.insert(blob.end(), tab.begin(), tab.end());
        auto e = std::make_shared<TabEntry>();
        e->blob.alloc(blob.size() + 1);
        DG_CUDA(cudaMemcpy(e->blob.p, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice));
        e->tab_cnt = tab_cnt;
        e->script_ub = sub;
        if (ix->tab_cache.size() >= 64) ix->tab_cache.clear();   // (batches in flight keep their entry alive)
        ix->tab_cache[key] = e;
        b->tabs = e;
      }
    }
    const std::vector<uint32_t>& tab_cnt = b->tabs->tab_cnt;
    {
      // can any query of this batch reach the cap?  (edit mode: by script count; Hamming: by the size of
      // an all-N query of that length)
      const uint64_t cap = par->max_neighborhood ? par->max_neighborhood : 10000;
      for (int m = 1; m < 256 && !b->maybe_capped; ++m) {
        if (!have[m]) continue;
        const uint64_t d = std::min<uint64_t>(par->distance, (uint64_t)m - 1), w = 4ull * m;
        const uint64_t hsize = 1 + (d >= 1 ? w : 0) + (d >= 2 ? (w * w - 16ull * m) / 2 : 0);
        if (indel ? b->tabs->script_ub[m] >= cap : hsize >= cap) b->maybe_capped = true;
      }
      if (par->distance > (uint32_t)kMaxDist) b->maybe_capped = true;   // listed neighbourhoods are resolved by the same host pass
    }
    const double ts1 = now();
    b->irregular.alloc(1, st);
    // uniform batches (one length, all ACGT, nothing skipped) map unit -> query by a division
    // instead of a binary search; k_prepare clears the shortcut if any query is irregular
    bool uniform = nq > 0 && minL == maxL && maxL <= kMaxQuery &&
                   (par->seed_len ? (maxL > (int)par->seed_len) : (maxL >= 10)) && par->distance < (uint32_t)maxL;
    if (uniform) {
      int m = par->seed_len ? (int)par->seed_len : maxL;
      b->uniform_units = (uint64_t)tab_cnt[m] * (par->reverse ? 2 : 1);
      b->uniform_len = (uint32_t)maxL;
    }
    b->qcode.alloc(2 * (size_t)nq + 2, st);
    b->qflag.alloc((size_t)nq + 1, st);
    b->raw.alloc(b->nbytes + 1, st);
    b->fwd.alloc(b->nbytes + 1, st);
    b->rc.alloc(b->nbytes + 1, st);
    b->off.alloc((size_t)nq + 1, st);
    b->units.alloc((size_t)nq + 1, st);
    b->unit_off.alloc((size_t)nq + 1, st);
    b->status.alloc(nq, st);
    b->dist.alloc(nq, st);
    const double ts2 = now();
    if (d_seqs) {
      if (ready) DG_CUDA(cudaStreamWaitEvent(st, ready, 0));
      if (b->nbytes) DG_CUDA(cudaMemcpyAsync(b->raw.p, d_seqs, b->nbytes, cudaMemcpyDeviceToDevice, st));
    } else if (b->nbytes) {
      DG_CUDA(cudaMemcpyAsync(b->raw.p, seqs, b->nbytes, cudaMemcpyHostToDevice, st));
    }
    if (equal_len) {   // equal-length batches do not send their offsets: the device writes q * L
      k_fill_offsets<<<grid_for((uint64_t)nq + 1, 256), 256, 0, st>>>(b->off.p, (uint64_t)nq + 1, offsets[1]);
      DG_CUDA(cudaGetLastError());
    } else {
      DG_CUDA(cudaMemcpyAsync(b->off.p, offsets, ((size_t)nq + 1) * 8, cudaMemcpyHostToDevice, st));
    }
    // the copies above read caller memory: finish them before returning ownership
    if (!d_seqs || !equal_len) DG_CUDA(sync_stream(st));
    if (trace) fprintf(stderr, "[dg_batch_stage] scan+tables %.3f ms, allocs %.3f ms, copies %.3f ms\n", ts1 - ts0, ts2 - ts1, now() - ts2);
    *out = b;
    return DG_OK;
  } catch (CudaFail& e) {
    delete b;
    return e.code;
  }
}

// The queries k_prepare / k_nbr_bound left flagged (their neighbourhood may reach the cap -x) are
// replayed in the reference's own generation order on the host (nbr_trunc.hpp).  A strand whose set
// stays below the cap needs nothing: the device's substring-minimal set is the reference's.  A strand
// that reaches it gets its truncated set as a sorted key list (BatchDev::trunc_*), and the query the
// warning of hunter.h:342-345 (DG_Q_NBR_CAP).  Rare by construction; costs one stream synchronisation
// when the batch holds a query length that could reach the cap at all.
static int resolve_truncation(dg_batch* b, cudaStream_t st) {
  try {
    uint32_t irr = 0;
    DG_CUDA(cudaMemcpyAsync(&irr, b->irregular.p, 4, cudaMemcpyDeviceToHost, st));
    DG_CUDA(sync_stream(st));
    if (!(irr & 22u)) return DG_OK;
    const uint32_t nq = b->nq;
    const bool indel = b->par.indel != 0;
    const uint32_t cap = b->par.max_neighborhood ? b->par.max_neighborhood : 10000;
    std::vector<uint32_t> status(nq), dist(nq);
    std::vector<uint64_t> off((size_t)nq + 1);
    std::vector<uint8_t> fwd((size_t)b->nbytes + 1);
    DG_CUDA(cudaMemcpyAsync(status.data(), b->status.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaMemcpyAsync(dist.data(), b->dist.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaMemcpyAsync(off.data(), b->off.p, ((size_t)nq + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (b->nbytes) DG_CUDA(cudaMemcpyAsync(fwd.data(), b->fwd.p, b->nbytes, cudaMemcpyDeviceToHost, st));
    DG_CUDA(sync_stream(st));
    const uint32_t want = (uint32_t)DG_Q_NBR_UNVERIFIED | (indel ? 0u : (uint32_t)DG_Q_NBR_CAP);   // (listed queries carry UNVERIFIED in both modes)
    std::vector<uint32_t> flagged;
    for (uint32_t q = 0; q < nq; ++q)
      if ((status[q] & want) && !(status[q] & (DG_Q_TOO_SHORT | DG_Q_SKIPPED | DG_Q_UNSUPPORTED))) flagged.push_back(q);
    if (flagged.empty()) return DG_OK;
    const int nstrand = b->par.reverse ? 2 : 1;
    struct Out { bool capped[2] = {false, false}; bool unkeyed = false; std::vector<ulonglong2> keys[2]; std::vector<std::string> all[2]; };
    std::vector<Out> outs(flagged.size());
    auto work = [&](size_t lo, size_t hi) {
      NeighborReplay nr_any;
      NeighborReplayPacked nr_packed;
      for (size_t i = lo; i < hi; ++i) {
        const uint32_t q = flagged[i];
        const uint8_t* s = fwd.data() + off[q];
        const int L = (int)(off[q + 1] - off[q]);
        const int m = b->par.seed_len ? (int)b->par.seed_len : L;
        for (int strand = 0; strand < nstrand; ++strand) {
          std::string str;
          if (strand == 0) {
            str.assign((const char*)s + (L - m), (size_t)m);
          } else {
            for (int j = 0; j < m; ++j) {
              const uint8_t ch = s[L - 1 - j];
              str.push_back(ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : ch == 'T' ? 'A' : 'N');
            }
          }
          // (ACGT-only strings that fit one word take the packed form of the replay, ~10 times faster)
          const bool use_packed = NeighborReplayPacked::fits(str, (int)dist[q]);
          const bool capped = use_packed ? nr_packed.run(str, (int)dist[q], indel, cap) : nr_any.run(str, (int)dist[q], indel, cap);
          auto replayed = [&]() { return use_packed ? nr_packed.strings() : nr_any.strings(); };
          outs[i].capped[strand] = capped;
          if ((int)dist[q] > kMaxDist) {   // beyond the enumerated distances: the whole set goes to the device as a list
            outs[i].all[strand] = replayed();
            std::sort(outs[i].all[strand].begin(), outs[i].all[strand].end());
            continue;
          }
          if (!capped) continue;
          for (const std::string& t : replayed()) {
            if ((int)t.size() > kKeyChars) { outs[i].unkeyed = true; break; }
            ulonglong2 k = make_ulonglong2(0, 0);
            for (int j = 0; j < (int)t.size(); ++j) {
              const char ch = t[j];
              const unsigned long long v = ch == 'A' ? 1 : ch == 'C' ? 2 : ch == 'G' ? 3 : ch == 'T' ? 5 : 4;
              if (j < 21) k.x |= v << (60 - 3 * j); else k.y |= v << (60 - 3 * (j - 21));
            }
            outs[i].keys[strand].push_back(k);
          }
          std::sort(outs[i].keys[strand].begin(), outs[i].keys[strand].end(),
                    [](const ulonglong2& a, const ulonglong2& c) { return a.x < c.x || (a.x == c.x && a.y < c.y); });
        }
      }
    };
    const size_t nthreads = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(flagged.size(), 64), std::thread::hardware_concurrency()));
    if (nthreads <= 1) {
      work(0, flagged.size());
    } else {
      std::vector<std::thread> ts;
      for (size_t t = 0; t < nthreads; ++t) ts.emplace_back(work, flagged.size() * t / nthreads, flagged.size() * (t + 1) / nthreads);
      for (auto& t : ts) t.join();
    }
    std::vector<uint32_t> qs, loff(1, 0);
    std::vector<ulonglong2> keys;
    std::vector<uint8_t> lchars;
    std::vector<uint32_t> lo(1, 0), lq, lcode;
    for (size_t i = 0; i < flagged.size(); ++i) {
      const uint32_t q = flagged[i];
      Out& o = outs[i];
      if ((int)dist[q] > kMaxDist) {
        status[q] &= ~(uint32_t)(DG_Q_NBR_UNVERIFIED | DG_Q_NBR_CAP);
        if (o.capped[0] || o.capped[1]) status[q] |= DG_Q_NBR_CAP;
        for (int strand = 0; strand < nstrand; ++strand)
          for (size_t r = 0; r < o.all[strand].size(); ++r) {
            const std::string& t = o.all[strand][r];
            lchars.insert(lchars.end(), t.begin(), t.end());
            lo.push_back((uint32_t)lchars.size());
            lq.push_back(q);
            lcode.push_back(pack_listed(strand, (int)t.size(), (uint32_t)r));
          }
        continue;
      }
      if (o.unkeyed) {   // strings longer than the sort key: cannot be listed; the flag stays
        status[q] |= DG_Q_NBR_UNVERIFIED;
        continue;
      }
      status[q] &= ~(uint32_t)(DG_Q_NBR_UNVERIFIED | DG_Q_NBR_CAP);
      if (o.capped[0] || o.capped[1]) status[q] |= DG_Q_NBR_CAP;
      for (int strand = 0; strand < nstrand; ++strand) {
        if (!o.capped[strand]) continue;
        qs.push_back((q << 1) | (uint32_t)strand);
        keys.insert(keys.end(), o.keys[strand].begin(), o.keys[strand].end());
        loff.push_back((uint32_t)keys.size());
      }
    }
    DG_CUDA(cudaMemcpyAsync(b->status.p, status.data(), (size_t)nq * 4, cudaMemcpyHostToDevice, st));
    b->n_trunc = (uint32_t)qs.size();
    if (b->n_trunc) {
      b->trunc_qs.alloc(qs.size(), st);
      b->trunc_off.alloc(loff.size(), st);
      b->trunc_keys.alloc(keys.size() + 1, st);
      DG_CUDA(cudaMemcpyAsync(b->trunc_qs.p, qs.data(), qs.size() * 4, cudaMemcpyHostToDevice, st));
      DG_CUDA(cudaMemcpyAsync(b->trunc_off.p, loff.data(), loff.size() * 4, cudaMemcpyHostToDevice, st));
      if (!keys.empty()) DG_CUDA(cudaMemcpyAsync(b->trunc_keys.p, keys.data(), keys.size() * sizeof(ulonglong2), cudaMemcpyHostToDevice, st));
    }
    b->n_listed = (uint32_t)lq.size();
    if (b->n_listed) {
      b->ls_chars.alloc(lchars.size() + 1, st);
      b->ls_off.alloc(lo.size(), st);
      b->ls_q.alloc(lq.size(), st);
      b->ls_code.alloc(lcode.size(), st);
      DG_CUDA(cudaMemcpyAsync(b->ls_chars.p, lchars.data(), lchars.size(), cudaMemcpyHostToDevice, st));
      DG_CUDA(cudaMemcpyAsync(b->ls_off.p, lo.data(), lo.size() * 4, cudaMemcpyHostToDevice, st));
      DG_CUDA(cudaMemcpyAsync(b->ls_q.p, lq.data(), lq.size() * 4, cudaMemcpyHostToDevice, st));
      DG_CUDA(cudaMemcpyAsync(b->ls_code.p, lcode.data(), lcode.size() * 4, cudaMemcpyHostToDevice, st));
    }
    DG_CUDA(sync_stream(st));   // the host vectors above go out of scope
    return DG_OK;
  } catch (CudaFail& e) {
    return e.code;
  } catch (std::bad_alloc&) {
    set_error("out of host memory");
    return DG_ERR_NOMEM;
  }
}

static thread_local double g_host_mark[8];
static double g_pipe_t0 = 0;   // DG_TRACE: start of the current dg_hunt_batch call (host clock, ms)
static void prof_mark(dg_index* ix, int i, cudaStream_t st = nullptr) {
  static const bool trace = getenv("DG_TRACE") != nullptr;
  if (trace) g_host_mark[i] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  if (st && st != ix->stream) return;   // chunk-pipeline batches on the other streams are not profiled
  if (!ix->prof.enabled) return;   // (stage timings are taken on the index stream only)
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
    cudaStream_t st = b->st;
    const unsigned B = 256;
    const uint32_t nq = b->nq;
    uint64_t launches = 0;
    b->nhits = 0;
    b->ncand = 0;
    BatchDev bd = batch_dev(b);
    const uint32_t* tb = b->tabs->blob.p;
    UnitTabs ut{tb + 1280, tb + 256, tb + 768, tb};
    IndexView v = ix->view();
    ABuf<uint8_t> tmp;
    size_t tmp_cap = 0;
    auto ensure_tmp = [&](size_t bytes) -> void* {
      if (bytes > tmp_cap) { tmp.alloc(bytes + (bytes >> 3) + 256, st); tmp_cap = tmp.count; }
      return tmp.p;
    };
    prof_mark(ix, 0, st);
    // ---- prepare
    DG_CUDA(cudaMemsetAsync(b->units.p, 0, ((size_t)nq + 1) * 8, st));
    DG_CUDA(cudaMemsetAsync(b->irregular.p, 0, 4, st));
    if (nq) { k_prepare<<<grid_for(nq, B), B, 0, st>>>(b->raw.p, bd, b->fwd.p, b->rc.p, ut, b->units.p); ++launches; }
    {
      size_t tb = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb, b->units.p, b->unit_off.p, (int)(nq + 1), st);
      cub::DeviceScan::ExclusiveSum(ensure_tmp(tb), tb, b->units.p, b->unit_off.p, (int)(nq + 1), st);
      launches += 2;
    }
    if (b->maybe_capped && nq) {
      b->n_trunc = 0;
      b->n_listed = 0;
      if (b->par.indel) { k_nbr_bound<<<grid_for(nq, B), B, 0, st>>>(bd); ++launches; }
      const int rc_t = resolve_truncation(b, st);
      if (rc_t) return rc_t;
      bd = batch_dev(b);   // (picks the lists up)
    }
    prof_mark(ix, 1, st);
    // ---- search
    uint64_t cap64 = 32ULL * nq + (1ULL << 20);
    if (b->par.distance >= 2) cap64 = 256ULL * nq + (1ULL << 20);
    if (const char* e = getenv("DG_CAND_CAP")) cap64 = strtoull(e, nullptr, 10);
    if (cap64 > (1ULL << 30)) cap64 = 1ULL << 30;
    ABuf<unsigned int> ctr;     // [0] n_cand, [1] overflow
    ABuf<unsigned long long> nscripts;
    ctr.alloc(2, st);
    nscripts.alloc(1, st);
    struct Geometry { int nsm, general_per_sm; };
    static const Geometry geo = [&] {   // (every device of the box is the same part; magic statics are thread-safe)
      Geometry g{148, 1};
      cudaDeviceGetAttribute(&g.nsm, cudaDevAttrMultiProcessorCount, ix->device);
      int v2 = 1;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v2, k_search, 256, 0) == cudaSuccess) g.general_per_sm = std::max(v2, 1);
      return g;
    }();
    const int nsm = geo.nsm, general_per_sm = geo.general_per_sm;
    unsigned int hc[2] = {0, 0};
    unsigned long long h_scripts = 0;
    bool probe_timed = false;
    for (int attempt = 0; attempt < 2; ++attempt) {
      b->cands.alloc(cap64, st);
      DG_CUDA(cudaMemsetAsync(ctr.p, 0, 8, st));
      DG_CUDA(cudaMemsetAsync(nscripts.p, 0, 8, st));
      SearchOut so{b->cands.p, (uint32_t)cap64, ctr.p, ctr.p + 1, nscripts.p};
      if (nq) {
        // packed (ACGT-only, <= 31 bases) queries: presence-bitmap filter + compacted slow path
        const uint64_t npairs = b->par.reverse ? 2ULL * nq : (uint64_t)nq;
        static const int ppw_env = getenv("DG_PAIRS_PER_WARP") ? atoi(getenv("DG_PAIRS_PER_WARP")) : 0;
        uint32_t ppw = ppw_env > 0 ? (uint32_t)ppw_env : 16u;  // groups of eight 20-mers fill five passes of 32 lanes exactly
        // small batches: fewer pairs per warp so that the grid still covers every SM
        while (ppw > 1 && npairs / (8ull * ppw) < (uint64_t)nsm * 6) ppw >>= 1;
        const unsigned blocks = (unsigned)((npairs + 8ull * ppw - 1) / (8ull * ppw));
        PackedArgs pa;
        pa.qcode = bd.qcode; pa.qflag = bd.qflag; pa.dist = bd.dist; pa.off = bd.off;
        pa.nq = nq; pa.seed_len = bd.seed_len; pa.reverse = bd.reverse;
        pa.KB = (int)v.KB;
        // windows by length class, fallbacks resolved here: without the KB + 1 bitmap longer strings use KB
        pa.win_r[0] = v.present_lo; pa.win_l[0] = nullptr; pa.win_k[0] = (int)v.KB - 1;
        pa.win_r[1] = v.present_kb; pa.win_l[1] = v.present_kb_l; pa.win_k[1] = (int)v.KB;
        // The KB + 1 bitmaps filter four times better but are four times larger (34 GB each at 3 Gb): a probe
        // into them costs more (TLB reach, fewer siblings per DRAM row).  Measured on the 3 Gb index: at
        // distance <= 1 (160 strings per primer and strand) the KB windows alone are 6 % faster; at distance 2
        // (12.8 k strings) the better filter wins by 10 %.  DG_USE_HI=0/1 overrides.
        bool use_hi = v.present_hi != nullptr && b->par.distance >= 2;
        if (const char* e = getenv("DG_USE_HI")) use_hi = v.present_hi != nullptr && atoi(e) != 0;
        if (use_hi) { pa.win_r[2] = v.present_hi; pa.win_l[2] = v.present_hi_l; pa.win_k[2] = (int)v.KB + 1; }
        else { pa.win_r[2] = v.present_kb; pa.win_l[2] = v.present_kb_l; pa.win_k[2] = (int)v.KB; }
        // the regular batch: probes and slow path in separate kernels (k_probe_*, k_resolve)
        const int um = (int)b->uniform_len;
        const int ud = (int)std::min<uint32_t>(b->par.distance, um > 0 ? (uint32_t)um - 1 : 0);
        const bool no_split = getenv("DG_NO_SPLIT") != nullptr;
        const bool split = !no_split && um > 0 && !b->par.seed_len && ud >= 1 && um + ud <= kMaxPacked && v.KB != 0;
        const uint32_t* skip = split ? b->irregular.p : nullptr;
        if (split) {
          const bool indel = b->par.indel != 0;
          const int S = indel ? 8 : 3;
          ProbeShape shp;
          shp.m = um; shp.dq = ud; shp.npairs = npairs; shp.irregular = b->irregular.p;
          shp.n2 = (uint32_t)(S * um * (um + 1) / 2);
          shp.n_scripts = nscripts.p;
          shp.scripts_per_pair = (unsigned long long)(S * um + (indel ? 0 : 1));
          if (ud >= 2)
            for (int p1 = 0; p1 < um; ++p1)
              shp.scripts_per_pair += indel ? (unsigned long long)S * (4ull * (um - 1 - p1) + 4ull * (um - p1))
                                            : (unsigned long long)S * 3ull * (um - 1 - p1);
          const uint64_t n1 = npairs * (uint64_t)um;
          ABuf<uint8_t> m1;
          m1.alloc(n1 + 8, st);
          // tiles: the slow path of tile i (side stream) overlaps the probes of tile i + 1 (batch stream)
          const bool no_overlap = getenv("DG_NO_OVERLAP") != nullptr;
          AuxStream* aux = no_overlap ? nullptr : aux_of(st);
          cudaStream_t rs = aux ? aux->s : st;
          auto after = [&](cudaStream_t from, cudaStream_t to) {   // `to` continues once `from` got this far
            if (!aux || from == to) return;
            cudaEvent_t e = aux->event();
            DG_CUDA(cudaEventRecord(e, from));
            DG_CUDA(cudaStreamWaitEvent(to, e, 0));
          };
          after(st, rs);   // (the side stream starts behind everything the batch stream has queued)
          const uint64_t ntile1 = n1 >= (1ull << 22) ? 4 : 1;
          if (attempt == 0) prof_mark(ix, 6, st);
          for (uint64_t t = 0; t < ntile1; ++t) {
            const uint64_t s0 = (n1 * t / ntile1) & ~3ull, s1 = t + 1 == ntile1 ? n1 : ((n1 * (t + 1) / ntile1) & ~3ull);
            const uint64_t ns = s1 - s0;
            if (!ns) continue;
            if (indel) k_probe_singles<true><<<grid_for(ns, 256), 256, 0, st>>>(pa, shp, s0, ns, m1.p + s0);
            else k_probe_singles<false><<<grid_for(ns, 256), 256, 0, st>>>(pa, shp, s0, ns, m1.p + s0);
            if (attempt == 0 && t + 1 == ntile1) prof_mark(ix, 7, st);
            after(st, rs);
            const unsigned rb = (unsigned)std::min<uint64_t>((uint64_t)nsm * 8, (ns / 4 + 255) / 256 + 1);
            if (indel) k_resolve<true><<<rb, 256, 0, rs>>>(v, pa, so, shp, m1.p + s0, ns, 0, 0, s0);
            else k_resolve<false><<<rb, 256, 0, rs>>>(v, pa, so, shp, m1.p + s0, ns, 0, 0, s0);
            launches += 2;
          }
          probe_timed = true;
          ABuf<uint8_t> m2[2];
          if (ud >= 2) {
            // pairs of events: ~13 k scripts per string; tiles of pairs keep the result bytes at <= 128 MB each
            const uint64_t tile = std::max<uint64_t>(1, (128ull << 20) / shp.n2);
            for (auto& mb : m2) mb.alloc(std::min(tile, npairs) * (uint64_t)shp.n2 + 8, st);
            cudaEvent_t freed[2] = {nullptr, nullptr};
            uint64_t ti = 0;
            for (uint64_t p0 = 0; p0 < npairs; p0 += tile, ++ti) {
              const uint64_t np = std::min(tile, npairs - p0), nb = np * (uint64_t)shp.n2;
              const unsigned rb = (unsigned)std::min<uint64_t>((uint64_t)nsm * 8, (nb / 4 + 255) / 256 + 1);
              uint8_t* mp = m2[ti & 1].p;
              if (aux && freed[ti & 1]) DG_CUDA(cudaStreamWaitEvent(st, freed[ti & 1], 0));   // its previous reader is done
              if (indel) k_probe_pairs<true><<<grid_for(nb, 256), 256, 0, st>>>(pa, shp, p0, np, mp);
              else k_probe_pairs<false><<<grid_for(nb, 256), 256, 0, st>>>(pa, shp, p0, np, mp);
              after(st, rs);
              if (indel) k_resolve<true><<<rb, 256, 0, rs>>>(v, pa, so, shp, mp, nb, p0, 1, 0);
              else k_resolve<false><<<rb, 256, 0, rs>>>(v, pa, so, shp, mp, nb, p0, 1, 0);
              if (aux) { freed[ti & 1] = aux->event(); DG_CUDA(cudaEventRecord(freed[ti & 1], rs)); }
              launches += 2;
            }
          }
          after(rs, st);   // join: everything behind this point sees every candidate (and may reuse the mask buffers)
        }
        if (b->par.indel) k_search_packed<true><<<blocks, 256, 0, st>>>(v, pa, so, ppw, skip);
        else k_search_packed<false><<<blocks, 256, 0, st>>>(v, pa, so, ppw, skip);
        // everything else (queries holding 'N', longer than 31 bases): the byte-wise general path
        k_search<<<nsm * general_per_sm, 256, 0, st>>>(v, bd, ut, b->unit_off.p, b->uniform_units, so);
        launches += 2;
        if (b->n_listed) {   // neighbourhoods that came as lists from the host replay (distance 3)
          ListedStrings ls{b->ls_chars.p, b->ls_off.p, b->ls_q.p, b->ls_code.p, b->n_listed};
          k_search_listed<<<grid_for(b->n_listed, 128), 128, 0, st>>>(v, ls, so);
          ++launches;
        }
      }
      if (attempt == 0) prof_mark(ix, 2, st);
      DG_CUDA(cudaMemcpyAsync(hc, ctr.p, 8, cudaMemcpyDeviceToHost, st));
      DG_CUDA(cudaMemcpyAsync(&h_scripts, nscripts.p, 8, cudaMemcpyDeviceToHost, st));
      DG_CUDA(cudaMemcpyAsync(&b->h_irregular, b->irregular.p, 4, cudaMemcpyDeviceToHost, st));
      DG_CUDA(sync_stream(st));
      DG_CUDA(cudaGetLastError());
      if (!hc[1]) break;
      // more neighbour strings matched than the buffer holds (repeat-rich queries): the counter
      // kept counting, so the second attempt is sized exactly
      if (attempt == 1 || hc[0] >= 0x7FFFFFF0u) {
        set_error("candidate buffer overflow (" + std::to_string(hc[0]) + " neighbour strings matched); split the batch");
        return DG_ERR_OVERFLOW;
      }
      cap64 = (uint64_t)hc[0] + 1024;
    }
    uint32_t n = hc[0];
    uint64_t n_candidates = n;
    ABuf<Cand> c2;
    ABuf<uint8_t> keep;
    ABuf<uint32_t> nsel;
    nsel.alloc(1, st);
    Cand* cur = b->cands.p;
    // ---- antichain rule (edit mode), lexicographic order, de-duplication
    const int maxq_str = b->par.seed_len ? (int)b->par.seed_len : std::min(b->max_len, kMaxQuery);
    const bool key_sort = maxq_str + (int)b->par.distance <= kKeyChars && !getenv("DG_MERGE_SORT");
    if (n && key_sort) {
      // every string fits the 126-bit key: one kernel (antichain flag + key), one radix sort,
      // one adjacent-duplicate pass, one gather
      ABuf<CandKey> ka, kb;
      ka.alloc(n, st);
      kb.alloc(n, st);
      keep.alloc(n, st);
      const uint32_t sentinel = nq >= 0x7FFFFFFFu ? 0xFFFFFFFFu : 2u * nq;
      int qbits = 1;
      while (qbits < 32 && (1ULL << qbits) <= (uint64_t)sentinel) ++qbits;
      ABuf<unsigned int> big;
      big.alloc(1, st);
      DG_CUDA(cudaMemsetAsync(big.p, 0, 4, st));
      const int begin_bit = (maxq_str + (int)b->par.distance <= 21) ? 64 : 0;
      const bool full_sort_always = getenv("DG_FULL_KEY_SORT") != nullptr;
      uint32_t n3 = 0;
      for (int attempt = 0; attempt < 2; ++attempt) {
        const bool by_group = attempt == 0 && !full_sort_always;
        const int slow_keys = getenv("DG_SLOW_KEYS") ? 1 : 0;   // test knob: the byte-wise key builder for every query
        size_t tb = 0;
        const bool radix_groups = getenv("DG_GROUP_RADIX") != nullptr;   // test knob: the former 3 radix passes
        if (by_group && !radix_groups && sentinel != 0xFFFFFFFFu) {
          // counting sort by (query, strand): groups hold 1-2 candidates, so a count per group taken while
          // the keys are made, one scan and one scatter replace three radix passes; then a rank count
          // inside each small group
          ABuf<CandKey> kc;
          ABuf<uint32_t> gcnt, goff, grank;
          kc.alloc(n, st);
          gcnt.alloc((size_t)sentinel + 2, st);
          goff.alloc((size_t)sentinel + 2, st);
          grank.alloc(n, st);
          DG_CUDA(cudaMemsetAsync(gcnt.p, 0, ((size_t)sentinel + 2) * 4, st));
          k_cand_keys<<<grid_for(n, 128), 128, 0, st>>>(bd, cur, n, sentinel, ka.p, slow_keys, gcnt.p, grank.p);
          cub::DeviceScan::ExclusiveSum(nullptr, tb, gcnt.p, goff.p, (int)(sentinel + 1), st);
          cub::DeviceScan::ExclusiveSum(ensure_tmp(tb), tb, gcnt.p, goff.p, (int)(sentinel + 1), st);
          k_scatter_keys<<<grid_for(n, B), B, 0, st>>>(ka.p, n, goff.p, grank.p, kc.p);
          k_group_order<<<grid_for(n, B), B, 0, st>>>(kc.p, n, sentinel, kb.p, big.p);
        } else if (by_group) {
          // 3 radix passes on (query, strand), then a rank count inside each small group
          ABuf<CandKey> kc;
          kc.alloc(n, st);
          k_cand_keys<<<grid_for(n, 128), 128, 0, st>>>(bd, cur, n, sentinel, ka.p, slow_keys, nullptr, nullptr);
          cub::DeviceRadixSort::SortKeys(nullptr, tb, ka.p, kc.p, (int)n, CandKeyDecomposer{}, 128, 128 + qbits, st);
          cub::DeviceRadixSort::SortKeys(ensure_tmp(tb), tb, ka.p, kc.p, (int)n, CandKeyDecomposer{}, 128, 128 + qbits, st);
          k_group_order<<<grid_for(n, B), B, 0, st>>>(kc.p, n, sentinel, kb.p, big.p);
        } else {
          k_cand_keys<<<grid_for(n, 128), 128, 0, st>>>(bd, cur, n, sentinel, ka.p, slow_keys, nullptr, nullptr);
          cub::DeviceRadixSort::SortKeys(nullptr, tb, ka.p, kb.p, (int)n, CandKeyDecomposer{}, begin_bit, 128 + qbits, st);
          cub::DeviceRadixSort::SortKeys(ensure_tmp(tb), tb, ka.p, kb.p, (int)n, CandKeyDecomposer{}, begin_bit, 128 + qbits, st);
        }
        k_unique_keys<<<grid_for(n, B), B, 0, st>>>(kb.p, n, sentinel, keep.p);
        cub::DeviceSelect::Flagged(nullptr, tb, kb.p, keep.p, ka.p, nsel.p, (int)n, st);
        cub::DeviceSelect::Flagged(ensure_tmp(tb), tb, kb.p, keep.p, ka.p, nsel.p, (int)n, st);
        unsigned int h_big = 0;
        DG_CUDA(cudaMemcpyAsync(&n3, nsel.p, 4, cudaMemcpyDeviceToHost, st));
        DG_CUDA(cudaMemcpyAsync(&h_big, big.p, 4, cudaMemcpyDeviceToHost, st));
        DG_CUDA(sync_stream(st));
        launches += 9;
        if (!(by_group && h_big)) break;   // a group beyond kGroupMax: once more with the full key
      }
      ABuf<Cand> sorted;
      sorted.alloc(n3 ? n3 : 1, st);
      if (n3) k_gather_cands<<<grid_for(n3, B), B, 0, st>>>(cur, ka.p, n3, sorted.p);
      b->cands.swap(sorted);  // the unsorted buffer is released with `sorted`
      cur = b->cands.p;
      n = n3;
      launches += 1;
    } else {
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
        DG_CUDA(sync_stream(st));
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
        DG_CUDA(sync_stream(st));
        if (other == c3.p) {
          // final list must outlive this scope: move it into b->cands
          DG_CUDA(cudaMemcpyAsync(b->cands.p, c3.p, (size_t)n3 * sizeof(Cand), cudaMemcpyDeviceToDevice, st));
          cur = b->cands.p;
        } else {
          cur = other;  // == b->cands.p
        }
        n = n3;
      }
    }
    b->ncand = n;
    prof_mark(ix, 3, st);
    // ---- count-only mode (padlock.h:381-427, silica.h:365-394)
    if (b->counts_only) {
      b->counts.alloc(nq ? nq : 1, st);
      DG_CUDA(cudaMemsetAsync(b->counts.p, 0, (size_t)(nq ? nq : 1) * 8, st));
      if (n) { k_count<<<grid_for(n, B), B, 0, st>>>(cur, n, b->counts.p); ++launches; }
      prof_mark(ix, 4, st);
      prof_mark(ix, 5, st);
      b->ran = true;
      if (st == ix->stream) {
        ix->prof.launches = launches;
        ix->prof.last.scripts = h_scripts;
        ix->prof.last.candidates = n_candidates;
      }
      return DG_OK;
    }
    // ---- hit budget
    b->qhits.alloc((size_t)nq + 1, st);
    b->qoff.alloc((size_t)nq + 1, st);
    DG_CUDA(cudaMemsetAsync(b->qhits.p, 0, ((size_t)nq + 1) * 8, st));
    ABuf<uint64_t> take, before, ntake, nloc, hit_off, loc_off, keys, keys2;
    ABuf<uint32_t> qkey;
    uint64_t nhits = 0, nlocate = 0;
    unsigned long long h_max_rows = 0;
    ABuf<unsigned long long> max_rows;
    if (n) {
      take.alloc(n, st); before.alloc(n, st); ntake.alloc((size_t)n + 1, st); nloc.alloc((size_t)n + 1, st);
      hit_off.alloc((size_t)n + 1, st); loc_off.alloc((size_t)n + 1, st); qkey.alloc(n, st);
      DG_CUDA(cudaMemsetAsync(ntake.p + n, 0, 8, st));
      DG_CUDA(cudaMemsetAsync(nloc.p + n, 0, 8, st));
      k_take_in<<<grid_for(n, B), B, 0, st>>>(cur, n, b->par.max_locations, take.p, qkey.p);
      size_t tb = 0;
      cub::DeviceScan::ExclusiveSumByKey(nullptr, tb, qkey.p, take.p, before.p, (int)n, cub::Equality(), st);
      cub::DeviceScan::ExclusiveSumByKey(ensure_tmp(tb), tb, qkey.p, take.p, before.p, (int)n, cub::Equality(), st);
      max_rows.alloc(1, st);
      DG_CUDA(cudaMemsetAsync(max_rows.p, 0, 8, st));
      k_take_out<<<grid_for(n, B), B, 0, st>>>(cur, n, b->par.max_locations, before.p, take.p, ntake.p, nloc.p, b->qhits.p,
                                              b->status.p, max_rows.p);
      cub::DeviceScan::ExclusiveSum(nullptr, tb, ntake.p, hit_off.p, (int)(n + 1), st);
      cub::DeviceScan::ExclusiveSum(ensure_tmp(tb), tb, ntake.p, hit_off.p, (int)(n + 1), st);
      cub::DeviceScan::ExclusiveSum(nullptr, tb, nloc.p, loc_off.p, (int)(n + 1), st);
      cub::DeviceScan::ExclusiveSum(ensure_tmp(tb), tb, nloc.p, loc_off.p, (int)(n + 1), st);
      launches += 8;
      DG_CUDA(cudaMemcpyAsync(&nhits, hit_off.p + n, 8, cudaMemcpyDeviceToHost, st));
      DG_CUDA(cudaMemcpyAsync(&nlocate, loc_off.p + n, 8, cudaMemcpyDeviceToHost, st));
      DG_CUDA(cudaMemcpyAsync(&h_max_rows, max_rows.p, 8, cudaMemcpyDeviceToHost, st));
    }
    {
      size_t tb = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb, (uint64_t*)b->qhits.p, b->qoff.p, (int)(nq + 1), st);
      cub::DeviceScan::ExclusiveSum(ensure_tmp(tb), tb, (uint64_t*)b->qhits.p, b->qoff.p, (int)(nq + 1), st);
      launches += 2;
    }
    DG_CUDA(sync_stream(st));
    uint64_t loc_cap = 1ULL << 28;   // located rows held at once (8-byte keys, two buffers)
    if (const char* e = getenv("DG_LOCATE_CAP")) loc_cap = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
    // Batches whose candidates hold more rows than that (repeat-rich queries: the reference locates
    // every occurrence before it keeps the first max_locations) are located and verified in slices
    // of consecutive candidates.
    std::vector<uint64_t> h_loc, h_hit;
    std::vector<uint32_t> slice_end;    // candidate index one past each slice
    if (nlocate > loc_cap) {
      h_loc.resize((size_t)n + 1);
      h_hit.resize((size_t)n + 1);
      DG_CUDA(cudaMemcpyAsync(h_loc.data(), loc_off.p, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, st));
      DG_CUDA(cudaMemcpyAsync(h_hit.data(), hit_off.p, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, st));
      DG_CUDA(sync_stream(st));
      uint32_t c0 = 0;
      while (c0 < n) {
        uint32_t c1 = c0 + 1;
        while (c1 < n && h_loc[c1 + 1] - h_loc[c0] <= loc_cap) ++c1;
        if (h_loc[c1] - h_loc[c0] >= (1ULL << 31)) {
          set_error("a single neighbour string occurs " + std::to_string(h_loc[c1] - h_loc[c0]) + " times: too many to locate");
          return DG_ERR_OVERFLOW;
        }
        slice_end.push_back(c1);
        c0 = c1;
      }
    } else {
      slice_end.push_back(n);
    }
    prof_mark(ix, 4, st);   // (locate and verify alternate per slice: both are reported under "verify" when sliced)
    // ---- locate + per-candidate ascending order + verify, slice by slice
    b->nhits = nhits;
    int maxq = b->par.seed_len ? (int)b->par.seed_len : std::min(b->max_len, kMaxQuery);
    int maxg = std::min(b->max_len, kMaxQuery) + 2 * (int)b->par.distance;
    if (!b->par.indel && !b->par.seed_len) maxg = maxq;
    uint32_t aln_max = (uint32_t)(maxg + maxq);
    b->pool_stride = b->par.seed_len ? (uint32_t)maxg : (b->par.indel ? 2 * aln_max : (uint32_t)(maxg + maxq));
    // hunt results travel as compact records (dg_rec); search results keep dg_hit + genomic contexts
    static const bool full_records = getenv("DG_FULL_RECORDS") != nullptr;
    b->compact = !b->par.seed_len && !full_records && b->par.distance <= (uint32_t)kMaxCompactDist;
    b->wire.alloc(nhits ? nhits : 1, st);
    if (b->compact) {
      b->recs.alloc(nhits ? nhits : 1, st);
      b->qmeta.alloc(nq ? nq : 1, st);
    } else {
      b->hits.alloc(nhits ? nhits : 1, st);
      b->pool.alloc(nhits ? nhits * b->pool_stride : 1, st);   // upper bound; pool_bytes of it are used
    }
    b->pool_bytes = 0;
    ABuf<unsigned long long> cursor;
    cursor.alloc(1, st);
    DG_CUDA(cudaMemsetAsync(cursor.p, 0, 8, st));
    uint32_t c0 = 0;
    for (size_t sl = 0; sl < slice_end.size() && nlocate; ++sl) {
      const uint32_t c1 = slice_end[sl];
      const bool whole = slice_end.size() == 1;
      const uint64_t row0 = whole ? 0 : h_loc[c0], row1 = whole ? nlocate : h_loc[c1];
      const uint64_t hit0 = whole ? 0 : h_hit[c0], hit1 = whole ? nhits : h_hit[c1];
      const uint64_t nrows = row1 - row0;
      if (nrows) {
        keys.alloc(nrows, st);
        keys2.alloc(nrows, st);
        k_locate<<<grid_for(nrows, 128), 128, 0, st>>>(v, cur, n, loc_off.p, row0, nrows, keys.p);
        if (h_max_rows <= (unsigned long long)kGroupMax && !getenv("DG_LOCATE_RADIX")) {
          k_locate_order<<<grid_for(nrows, 256), 256, 0, st>>>(keys.p, loc_off.p, row0, nrows, keys2.p);
          launches += 2;
        } else {
          int cbits = 1;
          while ((1ULL << cbits) < (uint64_t)n + 1 && cbits < 32) ++cbits;
          size_t tb = 0;
          cub::DeviceRadixSort::SortKeys(nullptr, tb, keys.p, keys2.p, (int)nrows, 0, 32 + cbits, st);
          cub::DeviceRadixSort::SortKeys(ensure_tmp(tb), tb, keys.p, keys2.p, (int)nrows, 0, 32 + cbits, st);
          launches += 4;
        }
      }
      if (whole) prof_mark(ix, 4, st);
      if (hit1 > hit0) {
        VerifyArgs a;
        a.cands = cur; a.ncand = n; a.hit_off = hit_off.p; a.loc_off = loc_off.p; a.keys = keys2.p; a.key_base = row0; a.nhits = nhits;
        a.hits = b->hits.p; a.recs = b->recs.p; a.wire = b->wire.p; a.pool = b->pool.p; a.pool_cursor = cursor.p;
        a.peer.nranks = 0;
        if (b->peer_from_verify) comm_peer_out(ix, b, &a.peer);
        a.srow_ints = (uint32_t)(maxq + 2);
        a.trace_bytes = (uint32_t)(((maxg + 1) * (maxq + 1) + 3) / 4 + 4);
        a.scratch_stride = a.srow_ints * 4 + a.trace_bytes + 3 * (aln_max + 8);
        a.scratch_stride = (a.scratch_stride + 15) & ~15u;
        bool need_scratch = (b->par.indel || b->par.seed_len) && (maxq > kLocalQ || maxg > kLocalG);
        const uint64_t span = hit1 - hit0;
        uint64_t chunk = span;
        if (need_scratch) {
          uint64_t budget = 1ULL << 30;
          chunk = std::max<uint64_t>(1, std::min<uint64_t>(span, budget / a.scratch_stride));
        }
        ABuf<uint8_t> scratch;
        scratch.alloc(need_scratch ? chunk * a.scratch_stride : 1, st);
        a.scratch = scratch.p;
        for (uint64_t first = hit0; first < hit1; first += chunk) {
          a.first_hit = first;
          a.chunk = std::min<uint64_t>(chunk, hit1 - first);
          if (b->compact) k_verify<true><<<grid_for(a.chunk, kVerifyBlock), kVerifyBlock, 0, st>>>(v, bd, a);
          else k_verify<false><<<grid_for(a.chunk, kVerifyBlock), kVerifyBlock, 0, st>>>(v, bd, a);
          ++launches;
        }
      }
      c0 = c1;
    }
    if (b->compact) {
      // (no alignment pool, hence no size to read back: the run ends without a host round trip)
      b->qpack.alloc(nq ? nq : 1, st);
      if (nq) { k_pack_qmeta<<<grid_for(nq, B), B, 0, st>>>(b->status.p, b->dist.p, nq, b->qmeta.p, b->qcode.p, b->qpack.p); ++launches; }
    } else {
      unsigned long long used = 0;
      DG_CUDA(cudaMemcpyAsync(&used, cursor.p, 8, cudaMemcpyDeviceToHost, st));
      DG_CUDA(sync_stream(st));
      b->pool_bytes = used;
    }
    prof_mark(ix, 5, st);
    if (getenv("DG_TRACE"))
      fprintf(stderr, "[dg_batch_run] host ms: prepare %.3f search %.3f filter %.3f locate %.3f verify %.3f (nq %u, cands %u, hits %llu); "
              "marks at %.3f %.3f %.3f %.3f %.3f %.3f ms of the call\n",
              g_host_mark[1] - g_host_mark[0], g_host_mark[2] - g_host_mark[1], g_host_mark[3] - g_host_mark[2],
              g_host_mark[4] - g_host_mark[3], g_host_mark[5] - g_host_mark[4], nq, n, (unsigned long long)nhits,
              g_host_mark[0] - g_pipe_t0, g_host_mark[1] - g_pipe_t0, g_host_mark[2] - g_pipe_t0, g_host_mark[3] - g_pipe_t0,
              g_host_mark[4] - g_pipe_t0, g_host_mark[5] - g_pipe_t0);
    DG_CUDA(cudaGetLastError());
    b->ran = true;
    if (st == ix->stream) {
      ix->prof.launches = launches;
      ix->prof.last.scripts = h_scripts;
      ix->prof.last.candidates = n_candidates;
      ix->prof.last.located = nlocate;
      ix->prof.last.hits = nhits;
      ix->prof.probe_timed = probe_timed;
    }
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
  ix->prof.last.ms_probe = 0;
  if (ix->prof.probe_timed) cudaEventElapsedTime(&ix->prof.last.ms_probe, ix->prof.ev[6], ix->prof.ev[7]);
  ix->prof.last.launches = ix->prof.launches;
}

static int fetch_impl(dg_batch* b, dg_result** out) {
  if (!b->ran) { set_error("dg_batch_fetch before dg_batch_run"); return DG_ERR_ARG; }
  dg_index* ix = b->ix;
  dg_result* r = nullptr;
  try {
    DG_CUDA(cudaSetDevice(ix->device));
    cudaStream_t st = b->st;
    r = new dg_result();
    uint32_t nq = b->nq;
    r->nq = nq;
    r->nhits = b->nhits;
    r->compact = b->compact;
    r->cum = ix->h_cum;
    r->seqs.alloc(b->nbytes, true);
    if (b->compact) {
      r->recs.alloc(b->nhits * sizeof(dg_rec), true);
      r->qmeta.alloc((size_t)nq * 2, true);
      if (b->nhits) DG_CUDA(cudaMemcpyAsync(r->recs.p, b->recs.p, b->nhits * sizeof(dg_rec), cudaMemcpyDeviceToHost, st));
      if (nq) DG_CUDA(cudaMemcpyAsync(r->qmeta.p, b->qmeta.p, (size_t)nq * 2, cudaMemcpyDeviceToHost, st));
      r->transfer_bytes = b->nhits * sizeof(dg_rec) + (size_t)nq * 2 + b->nbytes;
    } else {
      r->hits.alloc(b->nhits * sizeof(dg_hit), true);
      r->qoff.alloc(((size_t)nq + 1) * 8, true);
      r->status.alloc((size_t)nq * 4, true);
      r->dist.alloc((size_t)nq * 4, true);
      r->pool.alloc(b->pool_bytes, true);
      if (b->nhits) {
        DG_CUDA(cudaMemcpyAsync(r->hits.p, b->hits.p, b->nhits * sizeof(dg_hit), cudaMemcpyDeviceToHost, st));
        if (b->pool_bytes) DG_CUDA(cudaMemcpyAsync(r->pool.p, b->pool.p, b->pool_bytes, cudaMemcpyDeviceToHost, st));
      }
      DG_CUDA(cudaMemcpyAsync(r->qoff.p, b->qoff.p, ((size_t)nq + 1) * 8, cudaMemcpyDeviceToHost, st));
      if (nq) {
        DG_CUDA(cudaMemcpyAsync(r->status.p, b->status.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
        DG_CUDA(cudaMemcpyAsync(r->dist.p, b->dist.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
      }
      r->transfer_bytes = b->nhits * sizeof(dg_hit) + b->pool_bytes + ((size_t)nq + 1) * 8 + (size_t)nq * 8 + b->nbytes;
    }
    if (b->nbytes) DG_CUDA(cudaMemcpyAsync(r->seqs.p, b->fwd.p, b->nbytes, cudaMemcpyDeviceToHost, st));
    DG_CUDA(sync_stream(st));
    if (b->host_off.size() == (size_t)nq + 1) keep_offsets(r, b->host_off.data(), nq);
    else if (b->uniform_host_len) r->uniform_len = b->uniform_host_len;
    prof_collect(ix);
    *out = r;
    return DG_OK;
  } catch (CudaFail& e) {
    delete r;
    return e.code;
  } catch (std::bad_alloc&) {
    delete r;
    set_error("out of host memory");
    return DG_ERR_NOMEM;
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
  b->peer_from_verify = true;
  return run_impl(b);
}
int dg_batch_fetch(dg_batch* b, dg_result** out) {
  if (!b || !out) { set_error("null argument"); return DG_ERR_ARG; }
  return fetch_impl(b, out);
}
int dg_batch_summary(dg_batch* b, uint64_t* n_hits, uint64_t* n_candidates) {
  if (!b || !b->ran) { set_error("batch has not run"); return DG_ERR_ARG; }
  cudaStreamSynchronize(b->st);
  prof_collect(b->ix);
  if (n_hits) *n_hits = b->nhits;
  if (n_candidates) *n_candidates = b->ncand;
  return DG_OK;
}
int dg_index_wire_records(dg_index* idx, const void** device_ptr, uint64_t* n) {
  if (!idx || !device_ptr || !n) { set_error("null argument"); return DG_ERR_ARG; }
  cudaSetDevice(idx->device);
  cudaStreamSynchronize(idx->stream);
  for (auto s2 : idx->xstream) if (s2) cudaStreamSynchronize(s2);
  *device_ptr = idx->wire_n ? (const void*)idx->wire.p : nullptr;
  *n = idx->wire_n;
  return DG_OK;
}
void dg_batch_free(dg_batch* b) {
  if (!b) return;
  cudaSetDevice(b->ix->device);
  delete b;
}

// Large batches are cut into chunks that flow through a pipeline: three host workers, each with its
// own compute stream, take the chunks in turn (stage -> search -> verify), so the host-side gaps
// of one chunk (size read-backs, allocations, launches) are filled by the other chunk's kernels;
// finished chunks are committed in query order: ids and offsets are rebased on the device and the
// records travel to the host on the copy stream, straight into their final place in the result,
// while later chunks are still being searched.
namespace {
struct ChunkPipe {
  dg_index* idx;
  const char* seqs;
  const uint64_t* offsets;
  uint32_t nq, nchunks;
  std::vector<uint32_t> bounds;   // chunk c = queries [bounds[c], bounds[c + 1])
  const dg_params* params;
  dg_result* r;
  std::mutex mu;
  std::condition_variable cv;
  int nworkers = 2;
  uint32_t next_commit = 0;      // chunks are committed (final offsets assigned) in order
  const uint8_t* d_up = nullptr; // the caller's sequences in device memory, chunk c valid once uploaded > c
  std::vector<cudaEvent_t> up_ev, copied_ev, rebased_ev;   // per chunk, from the index's event pool
  uint32_t uploaded = 0;
  std::vector<uint64_t> chunk_len;   // per chunk: the common query length, or 0 (set before `uploaded` passes the chunk)
  PeerOut peer;                      // peer mode of a bound communicator (compact results only)
  uint64_t hit_base = 0, pool_base = 0;
  int rc = DG_OK;
  std::string err;
  double t_begin = 0;
  bool trace = false;

  void fail(int code, const std::string& msg) {
    std::lock_guard<std::mutex> g(mu);
    if (rc == DG_OK) { rc = code; err = msg; }
    cv.notify_all();
  }

  void worker(int w) {
    struct Live { dg_batch* b; cudaEvent_t copied; };
    std::vector<Live> live;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    try {
      DG_CUDA(cudaSetDevice(idx->device));
      cudaStream_t cs = idx->copy_stream;
      std::vector<uint64_t> so;
      for (uint32_t c = (uint32_t)w; c < nchunks; c += (uint32_t)nworkers) {
        { std::lock_guard<std::mutex> g(mu); if (rc != DG_OK) break; }
        double tm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        tm[0] = now() - t_begin;
        // chunk c runs on a stream of higher priority than chunks c + 1 and c + 2, which are in flight
        // next to it: its pending blocks are dispatched first, so chunks finish -- and their records
        // start leaving -- one after another instead of all together.  (The six priority levels wrap:
        // every sixth chunk starts again at the highest one and briefly outranks the two chunks before it;
        // batches of up to ~1.4 M queries never get there.)  A stream belongs to one worker at a time
        // as long as kXStreams is a multiple of the worker count; the stream pools are locked anyway.
        cudaStream_t st = idx->xstream[c % (uint32_t)dg_index::kXStreams];
        // release chunks whose records have reached the host (keeps at most 2 per worker alive)
        while (!live.empty() && (live.size() >= 2 || cudaEventQuery(live.front().copied) == cudaSuccess)) {
          DG_CUDA(cudaEventSynchronize(live.front().copied));
          dg_batch_free(live.front().b);
          live.erase(live.begin());
        }
        const uint32_t q0 = bounds[c], q1 = bounds[c + 1];
        const uint32_t cn = q1 - q0;
        if (offsets[q1] < offsets[q0]) { fail(DG_ERR_ARG, "offsets must be non-decreasing"); break; }
        tm[1] = now() - t_begin;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return uploaded > c || rc != DG_OK; });
          if (rc != DG_OK) break;
        }
        // equal-length chunks (the common case; checked by the uploader) need no chunk-local offsets
        const uint64_t uniform_L = chunk_len[c];
        if (!uniform_L) {
          so.resize((size_t)cn + 1);
          for (uint32_t k = 0; k <= cn; ++k) so[k] = offsets[q0 + k] - offsets[q0];
        }
        tm[2] = now() - t_begin;
        dg_batch* b = nullptr;
        int rc1 = stage_impl(idx, seqs + offsets[q0], so.data(), cn, params, &b, st, d_up + offsets[q0], up_ev[c], false, uniform_L);
        if (rc1) { fail(rc1, last_error_ref()); break; }
        cudaEvent_t copied = copied_ev[c];
        live.push_back(Live{b, copied});
        tm[3] = now() - t_begin;
        rc1 = run_impl(b);
        if (rc1) { fail(rc1, last_error_ref()); break; }
        tm[4] = now() - t_begin;
        // commit in query order
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return next_commit == c || rc != DG_OK; });
        if (rc != DG_OK) break;
        tm[5] = now() - t_begin;
        const bool compact = b->compact;
        r->compact = compact;
        const size_t rec_size = compact ? sizeof(dg_rec) : sizeof(dg_hit);
        HostBuf& rbuf = compact ? r->recs : r->hits;
        const uint64_t need_hits = (hit_base + b->nhits) * rec_size, need_pool = pool_base + b->pool_bytes;
        if (need_hits > rbuf.cap || need_pool > r->pool.cap) {
          // first call with this volume (later calls get right-sized blocks from the cache): size for
          // the whole batch from what the chunks so far produced
          DG_CUDA(cudaStreamSynchronize(cs));
          const double scale = 1.15 * (double)nq / (double)q1;  // q1 = queries committed so far
          rbuf.reserve(std::max<size_t>(need_hits, (size_t)(scale * need_hits)), hit_base * rec_size, true);
          r->pool.reserve(std::max<size_t>(need_pool, (size_t)(scale * need_pool)), pool_base, true);
        }
        if (hit_base + b->nhits > idx->wire.count) {
          // grow (first call with this volume): keep what earlier chunks wrote
          for (cudaStream_t s2 : idx->xstream) if (s2) DG_CUDA(cudaStreamSynchronize(s2));
          DevBuf<int4> bigger;
          bigger.alloc((size_t)(1.15 * (double)nq / (double)q1 * (double)(hit_base + b->nhits)) + 1024);
          if (hit_base) DG_CUDA(cudaMemcpy(bigger.p, idx->wire.p, hit_base * sizeof(int4), cudaMemcpyDeviceToDevice));
          std::swap(idx->wire.p, bigger.p);
          std::swap(idx->wire.count, bigger.count);
        }
        if (compact) {
          if (b->nhits)
            k_rebase_recs<<<grid_for(b->nhits, 256), 256, 0, st>>>(b->recs.p, b->nhits, q0, idx->wire.p + hit_base, peer, hit_base);
        } else {
          k_rebase<<<grid_for(std::max<uint64_t>(b->nhits, (uint64_t)cn + 1), 256), 256, 0, st>>>(
              b->hits.p, b->nhits, b->qoff.p, cn + 1, q0, pool_base, hit_base, idx->wire.p + hit_base);
        }
        cudaEvent_t ev = rebased_ev[c];
        DG_CUDA(cudaEventRecord(ev, st));
        DG_CUDA(cudaStreamWaitEvent(cs, ev, 0));
        if (compact) {
          if (b->nhits) DG_CUDA(cudaMemcpyAsync((uint8_t*)r->recs.p + hit_base * sizeof(dg_rec), b->recs.p, b->nhits * sizeof(dg_rec), cudaMemcpyDeviceToHost, cs));
          if (cn) DG_CUDA(cudaMemcpyAsync((uint16_t*)r->qmeta.p + q0, b->qmeta.p, (size_t)cn * 2, cudaMemcpyDeviceToHost, cs));
          r->transfer_bytes += b->nhits * sizeof(dg_rec) + (size_t)cn * 2;
        } else {
          if (b->nhits) {
            DG_CUDA(cudaMemcpyAsync((uint8_t*)r->hits.p + hit_base * sizeof(dg_hit), b->hits.p, b->nhits * sizeof(dg_hit), cudaMemcpyDeviceToHost, cs));
            if (b->pool_bytes) DG_CUDA(cudaMemcpyAsync((uint8_t*)r->pool.p + pool_base, b->pool.p, b->pool_bytes, cudaMemcpyDeviceToHost, cs));
          }
          DG_CUDA(cudaMemcpyAsync((uint64_t*)r->qoff.p + q0, b->qoff.p, ((size_t)cn + (c + 1 == nchunks ? 1 : 0)) * 8, cudaMemcpyDeviceToHost, cs));
          if (cn) {
            DG_CUDA(cudaMemcpyAsync((uint32_t*)r->status.p + q0, b->status.p, (size_t)cn * 4, cudaMemcpyDeviceToHost, cs));
            DG_CUDA(cudaMemcpyAsync((uint32_t*)r->dist.p + q0, b->dist.p, (size_t)cn * 4, cudaMemcpyDeviceToHost, cs));
          }
          r->transfer_bytes += b->nhits * sizeof(dg_hit) + b->pool_bytes + (size_t)cn * 16;
        }
        // the normalised sequences: a regular chunk (every query ACGT-only, one length <= 31) sends the 2-bit
        // codes k_prepare made (8 bytes per query, decoded on the host when someone asks); any other chunk
        // sends the characters
        if (cn && b->nbytes) {
          const bool as_codes = compact && !(b->h_irregular & 1u) && b->uniform_len > 0 && b->uniform_len <= 31 && b->qpack.p;
          if (as_codes) {
            DG_CUDA(cudaMemcpyAsync((uint64_t*)r->qpack.p + q0, b->qpack.p, (size_t)cn * 8, cudaMemcpyDeviceToHost, cs));
            r->packed_seqs.push_back(dg_result::PackedSeqs{q0, q1, b->uniform_len});
            r->have_seqs = false;
            r->transfer_bytes += (size_t)cn * 8;
          } else {
            DG_CUDA(cudaMemcpyAsync((uint8_t*)r->seqs.p + offsets[q0], b->fwd.p, b->nbytes, cudaMemcpyDeviceToHost, cs));
            r->transfer_bytes += b->nbytes;
          }
        }
        DG_CUDA(cudaEventRecord(copied, cs));
        hit_base += b->nhits;
        pool_base += b->pool_bytes;
        ++next_commit;
        if (trace) fprintf(stderr, "[dg_hunt_batch] worker %d committed chunk %u/%u: %u queries, %llu hits, t = %.3f ms "
                           "(loop top %.3f, offsets %.3f, upload %.3f, staged %.3f, ran %.3f, turn %.3f)\n", w, c + 1,
                           nchunks, cn, (unsigned long long)b->nhits, now() - t_begin, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5]);
        lk.unlock();
        cv.notify_all();
      }
    } catch (CudaFail& e) {
      fail(e.code, last_error_ref());
    } catch (std::bad_alloc&) {
      fail(DG_ERR_NOMEM, "out of host memory");
    }
    // batches whose records may still be travelling are handed to the caller of the pipeline, which
    // frees them after its final wait on the copy stream (a worker waiting here would sit inside a
    // blocking CUDA call while the other workers still launch: their calls stall behind it)
    {
      std::lock_guard<std::mutex> g(mu);
      for (auto& l : live) leftover.push_back(std::make_pair(l.b, l.copied));
    }
  }
  std::vector<std::pair<dg_batch*, cudaEvent_t>> leftover;
};
}  // namespace

static int hunt_chunked(dg_index* idx, const char* seqs, const uint64_t* offsets, uint32_t nq, const dg_params* params,
                        uint32_t chunk, dg_result** out) {
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  ChunkPipe p;
  // the first chunk is half-sized so that the two workers run out of phase: one worker's host gaps
  // (size read-backs between kernels) then fall into the other worker's long kernels
  // (shrinking the last chunks to shorten the device -> host tail was measured and costs more in
  // per-chunk fixed overhead than it saves)
  p.bounds.push_back(0);
  if (getenv("DG_UNEVEN")) {
    for (uint64_t q = getenv("DG_NO_STAGGER") ? chunk : chunk / 2; q < nq; q += chunk) p.bounds.push_back((uint32_t)q);
    if (p.bounds.size() > 1 && nq - p.bounds.back() < chunk / 4) p.bounds.pop_back();  // no tiny tail chunk
  } else {
    // a half-sized first chunk (the GPU starts early, the workers run out of phase), then equal chunks:
    // a small last chunk is latency-bound and finishes long after the one before it
    const uint64_t first = std::min<uint64_t>(chunk / 2, nq);
    const uint64_t rest = nq - first;
    const uint64_t nrest = (rest + chunk - 1) / chunk;
    for (uint64_t k = 0; k < nrest; ++k) p.bounds.push_back((uint32_t)(first + rest * k / nrest));
  }
  p.bounds.push_back(nq);
  const uint32_t nchunks = (uint32_t)p.bounds.size() - 1;
  p.idx = idx; p.seqs = seqs; p.offsets = offsets; p.nq = nq; p.nchunks = nchunks; p.params = params;
  p.trace = getenv("DG_TRACE") != nullptr;
  p.t_begin = now();
  g_pipe_t0 = p.t_begin;
  dg_result* r = nullptr;
  try {
    DG_CUDA(cudaSetDevice(idx->device));
    if (!idx->copy_stream) DG_CUDA(cudaStreamCreateWithFlags(&idx->copy_stream, cudaStreamNonBlocking));
    p.nworkers = 3;
    if (const char* e = getenv("DG_WORKERS")) p.nworkers = std::min(4, std::max(1, atoi(e)));
    p.nworkers = (int)std::min<uint32_t>((uint32_t)p.nworkers, nchunks);
    if (!idx->xstream[0]) {
      int least = 0, greatest = 0;
      DG_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));   // numerically lower = higher priority
      const bool flat = getenv("DG_NO_PRIORITY") != nullptr;
      for (int i = 0; i < dg_index::kXStreams; ++i)
        DG_CUDA(cudaStreamCreateWithPriority(&idx->xstream[i], cudaStreamNonBlocking, flat ? least : std::min(least, greatest + i)));
    }
    r = new dg_result();
    r->nq = nq;
    r->cum = idx->h_cum;
    static const bool full_records = getenv("DG_FULL_RECORDS") != nullptr;
    p.peer.nranks = 0;
    if (!params->seed_len && !full_records && params->distance <= (uint32_t)kMaxCompactDist) {
      comm_peer_out(idx, idx, &p.peer);
      r->qmeta.alloc((size_t)nq * 2, true);
      r->qpack.alloc((size_t)nq * 8, true);
    } else {
      r->qoff.alloc(((size_t)nq + 1) * 8, true);
      r->status.alloc((size_t)nq * 4, true);
      r->dist.alloc((size_t)nq * 4, true);
    }
    r->seqs.alloc(offsets[nq], true);
    p.r = r;
    // The calling thread uploads the sequences chunk by chunk, in order, on a stream of its own (a
    // copy from the caller's ordinary memory occupies the issuing thread; done per chunk by the
    // workers, those copies queue behind each other and behind the result copies); the workers
    // pick a chunk up as soon as its upload has been issued.
    if (!idx->up_stream) DG_CUDA(cudaStreamCreateWithFlags(&idx->up_stream, cudaStreamNonBlocking));
    if (idx->upload.count < offsets[nq] + 1) idx->upload.alloc(offsets[nq] + (offsets[nq] >> 3) + 4096);
    p.d_up = idx->upload.p;
    while (idx->ev_pool.size() < 3 * (size_t)nchunks) {
      cudaEvent_t e;
      DG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      idx->ev_pool.push_back(e);
    }
    p.chunk_len.assign(nchunks, 0);
    p.up_ev.assign(idx->ev_pool.begin(), idx->ev_pool.begin() + nchunks);
    p.copied_ev.assign(idx->ev_pool.begin() + nchunks, idx->ev_pool.begin() + 2 * (size_t)nchunks);
    p.rebased_ev.assign(idx->ev_pool.begin() + 2 * (size_t)nchunks, idx->ev_pool.begin() + 3 * (size_t)nchunks);
    std::vector<std::thread> others;
    for (int w = 0; w < p.nworkers; ++w) others.emplace_back([&p, w] { p.worker(w); });
    try {
      // Page-locked caller memory (cudaHostAlloc / cudaHostRegister, e.g. a pinned torch tensor): one
      // asynchronous copy per chunk.  Ordinary memory: such a copy occupies its caller and, issued
      // next to kernels and result copies, stalls for milliseconds (and the CUDA calls of the other
      // threads with it), so it is done in two pieces only -- the first chunk alone, so that the
      // GPU starts at once, then all the others while little else is in flight.
      cudaPointerAttributes attr;
      const bool pinned_in = cudaPointerGetAttributes(&attr, seqs) == cudaSuccess && attr.type == cudaMemoryTypeHost;
      cudaGetLastError();
      for (uint32_t c0 = 0; c0 < nchunks;) {
        const uint32_t c1 = (pinned_in || c0 == 0) ? c0 + 1 : nchunks;
        const uint64_t b0 = offsets[p.bounds[c0]], b1 = offsets[p.bounds[c1]];
        if (b1 < b0) { p.fail(DG_ERR_ARG, "offsets must be non-decreasing"); break; }
        if (b1 > b0) DG_CUDA(cudaMemcpyAsync(idx->upload.p + b0, seqs + b0, b1 - b0, cudaMemcpyHostToDevice, idx->up_stream));
        for (uint32_t c = c0; c < c1; ++c) DG_CUDA(cudaEventRecord(p.up_ev[c], idx->up_stream));
        for (uint32_t c = c0; c < c1; ++c) {
          // is the chunk of one query length?  (this thread has time; the workers are launching kernels)
          const uint32_t q0 = p.bounds[c], cn = p.bounds[c + 1] - q0;
          uint64_t L = cn ? offsets[q0 + 1] - offsets[q0] : 0, bad = 0;
          const uint64_t o0 = offsets[q0];
          const uint64_t* po = offsets + q0;
          uint64_t want = o0;
          for (uint32_t k = 0; k <= cn; ++k, want += L) bad |= po[k] ^ want;
          { std::lock_guard<std::mutex> g(p.mu); p.chunk_len[c] = bad ? 0 : L; p.uploaded = c + 1; }
          p.cv.notify_all();
        }
        c0 = c1;
      }
    } catch (CudaFail& e) {
      p.fail(e.code, last_error_ref());
    }
    const double tt0 = now();
    keep_offsets(r, offsets, nq);   // (the calling thread has nothing else to do while the workers run)
    const double tt1 = now();
    for (auto& t : others) t.join();
    const double tt2 = now();
    cudaStreamSynchronize(idx->copy_stream);
    const double tt3 = now();
    for (auto& l : p.leftover) dg_batch_free(l.first);
    p.leftover.clear();
    cudaStreamSynchronize(idx->up_stream);
    if (p.trace) fprintf(stderr, "[dg_hunt_batch] tail: offsets kept %.3f..%.3f, workers joined %.3f, copies drained %.3f, freed %.3f ms\n",
                         tt0 - p.t_begin, tt1 - p.t_begin, tt2 - p.t_begin, tt3 - p.t_begin, now() - p.t_begin);
    p.up_ev.clear();
    if (p.rc == DG_OK) {
      DG_CUDA(cudaStreamSynchronize(idx->copy_stream));
      r->nhits = p.hit_base;
      if (r->compact) r->recs.bytes = p.hit_base * sizeof(dg_rec);
      else r->hits.bytes = p.hit_base * sizeof(dg_hit);
      r->pool.bytes = p.pool_base;
      idx->wire_n = p.hit_base;
    }
  } catch (CudaFail& e) {
    p.rc = e.code;
    p.err = last_error_ref();
  } catch (std::bad_alloc&) {
    p.rc = DG_ERR_NOMEM;
    p.err = "out of host memory";
  }
  if (p.rc != DG_OK) { delete r; set_error(p.err); return p.rc; }
  if (p.trace) fprintf(stderr, "[dg_hunt_batch] nq=%u in %u chunks: %.3f ms\n", nq, nchunks, now() - p.t_begin);
  *out = r;
  return DG_OK;
}

int dg_hunt_batch(dg_index* idx, const char* seqs, const uint64_t* offsets, uint32_t nq, const dg_params* params,
                  dg_result** out) {
  static const bool trace = getenv("DG_TRACE") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  if (!idx || !offsets || !params || !out || (nq && !seqs)) { set_error("null argument"); return DG_ERR_ARG; }
  uint32_t chunk = 262144;
  if (const char* e = getenv("DG_CHUNK")) chunk = (uint32_t)std::max<long long>(1024, atoll(e));
  if (nq > chunk + chunk / 2 && offsets[0] == 0) return hunt_chunked(idx, seqs, offsets, nq, params, chunk, out);
  double t0 = now();
  dg_batch* b = nullptr;
  int rc = stage_impl(idx, seqs, offsets, nq, params, &b);
  if (rc) return rc;
  double t1 = now();
  rc = run_impl(b);
  if (!rc) {
    try {
      if (b->nhits > idx->wire.count) idx->wire.alloc(b->nhits + (b->nhits >> 3) + 1024);
      PeerOut peer;
      peer.nranks = 0;
      if (b->compact) comm_peer_out(idx, idx, &peer);
      if (b->nhits && b->compact)
        k_rebase_recs<<<grid_for(b->nhits, 256), 256, 0, b->st>>>(b->recs.p, b->nhits, 0, idx->wire.p, peer, 0);
      else if (b->nhits)
        k_rebase<<<grid_for(b->nhits, 256), 256, 0, b->st>>>(b->hits.p, b->nhits, b->qoff.p, 0, 0, 0, 0, idx->wire.p);
      idx->wire_n = b->nhits;
    } catch (CudaFail& e) { rc = e.code; }
  }
  double t2 = now();
  if (!rc) rc = fetch_impl(b, out);
  double t3 = now();
  dg_batch_free(b);
  if (trace) fprintf(stderr, "[dg_hunt_batch] nq=%u stage %.3f ms, run %.3f ms, fetch %.3f ms, free %.3f ms\n", nq, t1 - t0, t2 - t1, t3 - t2, now() - t3);
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
    DG_CUDA(sync_stream(st));
    DG_CUDA(cudaGetLastError());
    return DG_OK;
  } catch (CudaFail& e) {
    return e.code;
  }
}

const dg_rec* dg_result_records(const dg_result* r, uint64_t* n) {
  const bool have = r && r->compact;
  if (n) *n = have ? r->nhits : 0;
  return have ? (const dg_rec*)r->recs.p : nullptr;
}
int dg_rec_alignment(const dg_rec* rec, const char* query, uint32_t qlen, char* refalign, char* queryalign) {
  if (!rec || !query || !refalign || !queryalign) { set_error("null argument"); return DG_ERR_ARG; }
  if (rec->nops > kRecOps) { set_error("record holds more edit operations than the compact form carries"); return DG_ERR_FORMAT; }
  return rec_expand_rows(rec->ops, rec->nops, (const uint8_t*)query, (int)qlen, (uint8_t*)refalign, (uint8_t*)queryalign);
}
int dg_result_alignment(const dg_result* r, uint64_t i, char* refalign, char* queryalign) {
  if (!r || !refalign || !queryalign || i >= r->nhits) { set_error("bad argument"); return DG_ERR_ARG; }
  if (!r->compact) {
    const dg_hit& h = ((const dg_hit*)r->hits.p)[i];
    memcpy(refalign, (const char*)r->pool.p + h.aln_off, h.aln_len);
    memcpy(queryalign, (const char*)r->pool.p + h.aln_off + h.aln_len, h.aln_len);
    return (int)h.aln_len;
  }
  const dg_rec& c = ((const dg_rec*)r->recs.p)[i];
  if (c.query >= r->nq) { set_error("record with a query index outside the batch"); return DG_ERR_FORMAT; }
  std::vector<uint8_t> sq;
  strand_query(r, c.query, c.strand == '-', sq);
  return dg_rec_alignment(&c, (const char*)sq.data(), (uint32_t)sq.size(), refalign, queryalign);
}
void dg_recs_sort(dg_rec* recs, uint64_t n) {
  if (!recs || n < 2) return;
  std::sort(recs, recs + n, [](const dg_rec& a, const dg_rec& b) {  // DnaHit::operator< hunter.h:63-65
    return (a.score > b.score) || ((a.score == b.score) && (a.chr < b.chr)) ||
           ((a.score == b.score) && (a.chr == b.chr) && (a.start < b.start));
  });
}
uint64_t dg_result_transfer_bytes(const dg_result* r) { return r ? r->transfer_bytes : 0; }
const dg_hit* dg_result_hits(const dg_result* r, uint64_t* n) {
  if (n) *n = r ? r->nhits : 0;
  if (r) ensure_hits(r);
  return r ? (const dg_hit*)r->hits.p : nullptr;
}
const uint64_t* dg_result_query_offsets(const dg_result* r, uint32_t* nq) {
  if (nq) *nq = r ? r->nq : 0;
  if (r) ensure_qinfo(r);
  return r ? (const uint64_t*)r->qoff.p : nullptr;
}
const uint32_t* dg_result_query_status(const dg_result* r) {
  if (r) ensure_qinfo(r);
  return r ? (const uint32_t*)r->status.p : nullptr;
}
const uint32_t* dg_result_query_distance(const dg_result* r) {
  if (r) ensure_qinfo(r);
  return r ? (const uint32_t*)r->dist.p : nullptr;
}
const char* dg_result_pool(const dg_result* r, uint64_t* bytes) {
  if (r) ensure_hits(r);
  if (bytes) *bytes = r ? r->pool.bytes : 0;
  return r ? (const char*)r->pool.p : nullptr;
}
const char* dg_result_sequences(const dg_result* r, uint64_t* bytes) {
  if (r) ensure_seqs(r);
  if (bytes) *bytes = r ? r->seqs.bytes : 0;
  return r ? (const char*)r->seqs.p : nullptr;
}
void dg_result_free(dg_result* r) { delete r; }

// wire format of the hit all-gather: u64 nq, u64 nhits, u64 pool bytes, u64 seq bytes, then
// qoff[nq+1], status[nq], dist[nq], hits[nhits], pool, seqs
int dg_result_pack(const dg_result* r, void* buf, uint64_t* bytes) {
  if (!r || !bytes) { set_error("null argument"); return DG_ERR_ARG; }
  ensure_qinfo(r);   // the packed form is the full one: a compact result is expanded first
  ensure_hits(r);
  ensure_seqs(r);
  uint64_t nq = r->nq, nh = r->nhits, np = r->pool.bytes, ns = r->seqs.bytes;
  uint64_t need = 32 + (nq + 1) * 8 + nq * 8 + nh * sizeof(dg_hit) + np + ns;
  if (!buf) { *bytes = need; return DG_OK; }
  if (*bytes < need) { set_error("buffer too small"); return DG_ERR_ARG; }
  uint8_t* p = (uint8_t*)buf;
  uint64_t hdr[4] = {nq, nh, np, ns};
  auto put = [&](const void* src, uint64_t n) { if (n) memcpy(p, src, n); p += n; };
  put(hdr, 32);
  put(r->qoff.p, (nq + 1) * 8);
  put(r->status.p, nq * 4);
  put(r->dist.p, nq * 4);
  put(r->hits.p, nh * sizeof(dg_hit));
  put(r->pool.p, np);
  put(r->seqs.p, ns);
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
  try {
    r->nq = (uint32_t)nq;
    r->nhits = nh;
    auto get = [&](HostBuf& hb, uint64_t n) { hb.alloc(n, false); if (n) memcpy(hb.p, p, n); p += n; };
    get(r->qoff, (nq + 1) * 8);
    get(r->status, nq * 4);
    get(r->dist, nq * 4);
    get(r->hits, nh * sizeof(dg_hit));
    get(r->pool, np);
    get(r->seqs, ns);
  } catch (std::bad_alloc&) {
    delete r;
    set_error("out of host memory");
    return DG_ERR_NOMEM;
  }
  *out = r;
  return DG_OK;
}

}  // extern "C"
