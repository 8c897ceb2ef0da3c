// dg_build.cu -- builds the device-resident index (DESIGN.md "Data layout in HBM") either from a
// `.fm9` written by `dicey index` (SDSL csa_wt<> serialization, parsed by fm9.hpp) or from a text
// in dicey's dump format (reference src/index.h:96-123), entirely on the GPU.
//
//   .fm9 path : wavelet-tree bits -> BWT bytes (k_decode_bwt, the arithmetic of wt_pc::operator[]
//               wt_pc.hpp:294-311 over rank_support_v blocks rank_support_v.hpp:104-115)
//               -> occurrence blocks + exception lists -> K-mer interval table
//               -> text rebuilt from the ISA samples by 64-step LF chains (the walk of
//               suffix_array_algorithm.hpp:588-609, run for every sample at once).
//   text path : suffix array by bucketed radix sort of 21-symbol prefixes with tie refinement
//               (replaces divsufsort in construct_sa.hpp:96-120), BWT = T[SA-1]
//               (construct_bwt.hpp:35-71), SA samples every 32 rows, ISA samples every 64 positions
//               (csa_sampling_strategy.hpp:70-93, 669-706).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <cub/cub.cuh>
#include <deque>
#include <queue>
#include <thrust/iterator/counting_iterator.h>

#include "dg_common.cuh"
#include "fm9.hpp"
#include "fm9_select.hpp"

namespace dg {

namespace {

struct Temp {  // growable CUB scratch
  void* p = nullptr;
  size_t cap = 0;
  ~Temp() { if (p) cudaFree(p); }
  void* ensure(size_t bytes) {
    if (bytes > cap) {
      if (p) cudaFree(p);
      p = nullptr;
      cap = bytes + (bytes >> 2) + 256;
      DG_CUDA(cudaMalloc(&p, cap));
    }
    return p;
  }
};

inline unsigned grid_for(uint64_t items, unsigned block) { return (unsigned)((items + block - 1) / block); }

// ---------------------------------------------------------------- wavelet tree -> BWT bytes
struct WtView {
  const uint64_t* bv;
  const uint64_t* bb;       // rank_support_v basic blocks
  const uint64_t* bv_pos;   // per node
  const uint64_t* bv_rank;  // per node (leaves: the symbol)
  const uint16_t* child0;
  const uint16_t* child1;
};

__device__ __forceinline__ uint64_t rank1_v(const WtView& w, uint64_t idx) {
  // rank_support_v<1>::rank (rank_support_v.hpp:104-115)
  const uint64_t* p = w.bb + ((idx >> 8) & 0xFFFFFFFFFFFFFFFEULL);
  uint64_t r = p[0] + ((p[1] >> (63 - 9 * ((idx & 0x1FF) >> 6))) & 0x1FF);
  if (idx & 0x3F) r += __popcll(w.bv[idx >> 6] & ((1ULL << (idx & 0x3F)) - 1));
  return r;
}

// int_vector<0> -> u32: entry i occupies bits [i * width, (i + 1) * width) of the word stream
// (int_vector.hpp:813-842; one zero word follows the payload)
__global__ void k_unpack_ints(const uint64_t* __restrict__ words, uint64_t count, uint32_t width, uint32_t* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint64_t bit = i * width, word = bit >> 6;
  const uint32_t off = (uint32_t)(bit & 63);
  uint64_t v = words[word] >> off;
  if (off + width > 64) v |= words[word + 1] << (64 - off);
  if (width < 64) v &= (1ULL << width) - 1;
  out[i] = (uint32_t)v;
}

__global__ void k_decode_bwt(WtView w, uint64_t n, uint8_t* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t v = 0;
  uint64_t pos = i;
  while (w.child0[v] != 0xFFFF) {  // wt_pc::operator[] (wt_pc.hpp:294-311)
    uint64_t at = w.bv_pos[v] + pos;
    uint64_t r = rank1_v(w, at) - w.bv_rank[v];
    if ((w.bv[at >> 6] >> (at & 63)) & 1) { pos = r; v = w.child1[v]; }
    else { pos -= r; v = w.child0[v]; }
  }
  out[i] = (uint8_t)w.bv_rank[v];
}

// ---------------------------------------------------------------- BWT bytes -> occ blocks
__global__ void k_pack_blocks(const uint8_t* __restrict__ bwt, uint64_t n, uint64_t nblocks, OccBlock* __restrict__ occ,
                              uint32_t* __restrict__ ca, uint32_t* __restrict__ cc, uint32_t* __restrict__ cg,
                              uint32_t* __restrict__ ct, unsigned long long* __restrict__ n_exc) {
  uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  uint64_t lo = 0, hi = 0;
  uint32_t c[5] = {0, 0, 0, 0, 0};
  uint64_t base = b << 6;
  for (int o = 0; o < 64; ++o) {
    uint64_t i = base + o;
    if (i >= n) break;
    int code = base_code(bwt[i]);
    c[code]++;
    if (code < 4) {
      lo |= (uint64_t)(code & 1) << o;
      hi |= (uint64_t)(code >> 1) << o;
    }
  }
  occ[b].lo = lo;
  occ[b].hi = hi;
  ca[b] = c[0]; cc[b] = c[1]; cg[b] = c[2]; ct[b] = c[3];
  if (c[4]) atomicAdd(n_exc, (unsigned long long)c[4]);
}

__global__ void k_fill_counts(uint64_t nblocks, OccBlock* __restrict__ occ, const uint32_t* __restrict__ ca,
                              const uint32_t* __restrict__ cc, const uint32_t* __restrict__ cg,
                              const uint32_t* __restrict__ ct) {
  uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  occ[b].cnt[0] = ca[b]; occ[b].cnt[1] = cc[b]; occ[b].cnt[2] = cg[b]; occ[b].cnt[3] = ct[b];
}

struct IsException {
  const uint8_t* bwt;
  __device__ bool operator()(uint32_t i) const { return base_code(bwt[i]) == 4; }
};

__global__ void k_exc_fill(const uint8_t* __restrict__ bwt, const uint32_t* __restrict__ pos, uint32_t n_exc,
                           uint8_t* __restrict__ sym, uint32_t* __restrict__ flags, uint32_t* __restrict__ hist) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_exc) return;
  uint32_t p = pos[k];
  uint8_t s = bwt[p];
  sym[k] = s;
  uint32_t g = p >> kFlagShift;
  atomicOr(&flags[g >> 5], 1u << (g & 31));
  atomicAdd(&hist[s], 1u);
}

// ---------------------------------------------------------------- K-mer interval table
// level k+1 from level k: the interval of c.S is one backward step from the interval of S.
__global__ void k_kmer_level(IndexView ix, const uint2* __restrict__ prev, uint2* __restrict__ cur, uint32_t k) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t cnt = 1ULL << (2 * (k + 1));
  if (t >= cnt) return;
  uint32_t c = (uint32_t)(t >> (2 * k));
  uint2 iv = prev[t & ((1ULL << (2 * k)) - 1)];
  uint2 o = make_uint2(0, 0);
  if (iv.x < iv.y) {
    uint32_t l = iv.x, r = iv.y;
    backward_step(ix, l, r, code_base((int)c));
    if (l < r) o = make_uint2(l, r);
  }
  cur[t] = o;
}

// ---------------------------------------------------------------- KB-mer presence bitmap
// bit (packed code of T[i, i+KB), last base in the low bits) is set for every ACGT-only window of
// the text: a neighbour string whose last KB bases have a clear bit has an empty SA interval, so
// k_search_packed drops it after ONE DRAM access instead of a table lookup + backward steps.
constexpr int kPresenceChunk = 256;
__global__ void k_presence(const uint8_t* __restrict__ text, uint64_t n, uint32_t KB, uint32_t* __restrict__ bits,
                           uint32_t* __restrict__ bits_left) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t first = t * kPresenceChunk;
  if (first >= n) return;
  uint64_t end = first + kPresenceChunk;  // windows starting in [first, end)
  uint64_t last = end + KB - 1 < n ? end + KB - 1 : n;
  const uint64_t mask = (KB >= 32) ? ~0ULL : ((1ULL << (2 * KB)) - 1);
  uint64_t code = 0;
  uint32_t valid = 0;
  for (uint64_t j = first; j < last; ++j) {
    int c = base_code(text[j]);
    if (c < 4) { code = ((code << 2) | (uint64_t)c) & mask; ++valid; } else { code = 0; valid = 0; }
    if (valid >= KB) {
      const uint64_t bit = presence_bit(code, (int)KB);
      atomicOr(&bits[bit >> 5], 1u << (bit & 31));
      if (bits_left) {
        const uint64_t bl = presence_bit_left(code, (int)KB);
        atomicOr(&bits_left[bl >> 5], 1u << (bl & 31));
      }
    }
  }
}

// ---------------------------------------------------------------- text from ISA samples
// The same walk visits ISA[p] for every p, so it also fills the full suffix array (sa != null).
__global__ void k_rebuild_text(IndexView ix, const uint32_t* __restrict__ isa, uint64_t nchains, uint8_t* __restrict__ text,
                               uint32_t* __restrict__ sa) {
  uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchains) return;
  uint64_t n = ix.n, first = c << 6, end = (c + 1) << 6;
  uint32_t row;
  int64_t p;
  if (end <= n - 1) { row = isa[c + 1]; p = (int64_t)end - 1; }
  else { text[n - 1] = 0; row = 0; p = (int64_t)n - 2; if (sa) sa[0] = (uint32_t)(n - 1); }  // suffix n-1 (the sentinel) is row 0
  for (; p >= (int64_t)first; --p) {
    uint8_t s;
    row = lf_step(ix, row, &s);
    text[p] = s;
    if (sa) sa[row] = (uint32_t)p;
  }
}

// ---------------------------------------------------------------- synthetic text (dicey_b200/synth.py)
__global__ void k_synth_text(uint64_t seed, uint64_t reclen, uint64_t n, uint8_t* __restrict__ text) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == n - 1) { text[i] = 0; return; }
  uint64_t rec = i / (reclen + 1), o = i - rec * (reclen + 1);
  if (o == reclen) { text[i] = '\n'; return; }
  uint64_t z = seed + (rec * reclen + o) * 0x9E3779B97F4A7C15ULL;
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
  z ^= z >> 27; z *= 0x94D049BB133111EBULL;
  z ^= z >> 31;
  text[i] = (uint8_t)("ACGT"[z >> 62]);
}

// ---------------------------------------------------------------- suffix array on the GPU
struct CodeMap {
  uint8_t code[256];
  uint8_t bits;    // bits per symbol code: 3 (up to 8 symbols with the sentinel), 4 or 5 (IUPAC texts, up to 32)
  uint8_t syms;    // symbols per 63-bit sort key: 21, 15 or 12
};

__global__ void k_byte_hist(const uint8_t* __restrict__ t, uint64_t n, unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(&sh[t[i]], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

constexpr int kKeySyms = 21;  // 3-bit codes per 63-bit key (the most a key holds; fewer with wider codes)

__device__ __forceinline__ uint64_t suffix_key(const uint8_t* __restrict__ t, uint64_t n, uint64_t i, const CodeMap& cm) {
  uint64_t key = 0;
  const int bits = cm.bits, syms = cm.syms;
#pragma unroll
  for (int j = 0; j < kKeySyms; ++j) {
    if (j < syms) {
      uint64_t p = i + j;
      uint64_t c = p < n ? cm.code[t[p]] : 0;
      key = (key << bits) | c;
    }
  }
  return key;
}

struct InBucket {
  const uint8_t* t;
  uint64_t n;
  CodeMap cm;
  uint32_t bucket;
  __device__ bool operator()(uint32_t i) const {
    uint32_t c0 = cm.code[t[i]];
    uint32_t c1 = (uint64_t)i + 1 < n ? cm.code[t[(uint64_t)i + 1]] : 0;
    return ((c0 << cm.bits) | c1) == bucket;
  }
};

__global__ void k_pair_hist(const uint8_t* __restrict__ t, uint64_t n, CodeMap cm, unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[1024];
  const int nb = 1 << (2 * cm.bits);
  for (int i = threadIdx.x; i < nb; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint32_t c0 = cm.code[t[i]];
    uint32_t c1 = i + 1 < n ? cm.code[t[i + 1]] : 0;
    atomicAdd(&sh[(c0 << cm.bits) | c1], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

__global__ void k_make_keys(const uint8_t* __restrict__ t, uint64_t n, CodeMap cm, const uint32_t* __restrict__ suf,
                            uint64_t cnt, uint64_t depth, uint64_t* __restrict__ keys) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cnt) return;
  keys[j] = suffix_key(t, n, (uint64_t)suf[j] + depth, cm);
}

// tie flag of sorted position j: shares its key with a neighbour
__global__ void k_tie_flags(const uint64_t* __restrict__ keys, uint64_t cnt, uint8_t* __restrict__ flag) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cnt) return;
  uint64_t k = keys[j];
  bool tie = (j > 0 && keys[j - 1] == k) || (j + 1 < cnt && keys[j + 1] == k);
  flag[j] = tie ? 1 : 0;
}
// compacted tied element t at slot s: group id = compacted index of the first element of its run
__global__ void k_group_heads(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ slots, uint32_t nt,
                              uint32_t* __restrict__ head) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  uint32_t s = slots[t];
  bool is_head = (s == 0) || keys[s - 1] != keys[s];
  head[t] = is_head ? t : 0;
}
__global__ void k_gather_suffix(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ slots, uint32_t nt,
                                uint32_t* __restrict__ suf) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nt) suf[t] = sa[slots[t]];
}
__global__ void k_iota(uint32_t* __restrict__ a, uint32_t n) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) a[t] = t;
}
__global__ void k_gather_u32(const uint32_t* __restrict__ src, const uint32_t* __restrict__ perm, uint32_t n,
                             uint32_t* __restrict__ dst) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) dst[t] = src[perm[t]];
}
__global__ void k_gather_u64(const uint64_t* __restrict__ src, const uint32_t* __restrict__ perm, uint32_t n,
                             uint64_t* __restrict__ dst) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) dst[t] = src[perm[t]];
}
__global__ void k_scatter_sa(uint32_t* __restrict__ sa, const uint32_t* __restrict__ slots, const uint32_t* __restrict__ suf,
                             uint32_t nt) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nt) sa[slots[t]] = suf[t];
}
// after refinement: element t still tied iff (gid, key2) equals a neighbour's
__global__ void k_retie(const uint32_t* __restrict__ gid, const uint64_t* __restrict__ key2, uint32_t nt,
                        uint8_t* __restrict__ flag, uint32_t* __restrict__ head) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  bool same_prev = t > 0 && gid[t - 1] == gid[t] && key2[t - 1] == key2[t];
  bool same_next = t + 1 < nt && gid[t + 1] == gid[t] && key2[t + 1] == key2[t];
  flag[t] = (same_prev || same_next) ? 1 : 0;
  head[t] = same_prev ? 0 : t;  // new group heads (meaningful for the tied ones)
}
// re-number group ids of the surviving tied elements after compaction: head index -> rank among survivors
__global__ void k_renumber(const uint32_t* __restrict__ headscan, const uint32_t* __restrict__ keep_idx, uint32_t nk,
                           uint32_t* __restrict__ gid_out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nk) gid_out[t] = headscan[keep_idx[t]];
}

__global__ void k_bwt_from_sa(const uint8_t* __restrict__ text, const uint32_t* __restrict__ sa, uint64_t n,
                              uint8_t* __restrict__ bwt, uint32_t* __restrict__ sa_s, uint32_t* __restrict__ isa_s) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  uint32_t p = sa[j];
  bwt[j] = p ? text[p - 1] : text[n - 1];
  if ((j & (kSaSample - 1)) == 0) sa_s[j / kSaSample] = p;
  if ((p & 63) == 0) isa_s[p >> 6] = (uint32_t)j;
}

struct MaxOp {
  __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

// Sorts the suffixes listed in suf[0..cnt) (all starting with the same 2 symbols) into sa_out.
void sort_bucket(const uint8_t* text, uint64_t n, const CodeMap& cm, uint32_t* suf, uint64_t cnt, uint32_t* sa_out,
                 uint64_t* keys_a, uint64_t* keys_b, uint32_t* suf_b, uint8_t* flag, Temp& tmp, cudaStream_t st) {
  if (cnt == 0) return;
  if (cnt >= (1ULL << 31)) { set_error("suffix bucket too large for the GPU builder"); throw CudaFail{DG_ERR_UNSUPPORTED}; }
  const unsigned B = 256;
  k_make_keys<<<grid_for(cnt, B), B, 0, st>>>(text, n, cm, suf, cnt, 0, keys_a);
  size_t tb = 0;
  const int key_bits = cm.bits * cm.syms, first_bits = cm.bits * (cm.syms - 2);   // (the first two symbols are the bucket's)
  cub::DeviceRadixSort::SortPairs(nullptr, tb, keys_a, keys_b, suf, suf_b, (int)cnt, 0, first_bits, st);
  cub::DeviceRadixSort::SortPairs(tmp.ensure(tb), tb, keys_a, keys_b, suf, suf_b, (int)cnt, 0, first_bits, st);
  DG_CUDA(cudaMemcpyAsync(sa_out, suf_b, cnt * 4, cudaMemcpyDeviceToDevice, st));
  // ---- ties
  k_tie_flags<<<grid_for(cnt, B), B, 0, st>>>(keys_b, cnt, flag);
  DevBuf<uint32_t> slots, nsel;
  nsel.alloc(1);
  // upper bound for tied elements is cnt; allocate lazily after counting
  {
    size_t tb2 = 0;
    thrust::counting_iterator<uint32_t> it(0);
    // count first
    DevBuf<uint32_t> dummy;
    cub::DeviceReduce::Sum(nullptr, tb2, flag, nsel.p, (int)cnt, st);
    cub::DeviceReduce::Sum(tmp.ensure(tb2), tb2, flag, nsel.p, (int)cnt, st);
  }
  uint32_t nt = 0;
  DG_CUDA(cudaMemcpyAsync(&nt, nsel.p, 4, cudaMemcpyDeviceToHost, st));
  DG_CUDA(cudaStreamSynchronize(st));
  if (nt == 0) return;
  slots.alloc(nt);
  {
    size_t tb2 = 0;
    thrust::counting_iterator<uint32_t> it(0);
    cub::DeviceSelect::Flagged(nullptr, tb2, it, flag, slots.p, nsel.p, (int)cnt, st);
    cub::DeviceSelect::Flagged(tmp.ensure(tb2), tb2, it, flag, slots.p, nsel.p, (int)cnt, st);
  }
  DevBuf<uint32_t> gid, gid2, sufc, sufc2, perm, perm2, head, keep;
  DevBuf<uint64_t> key2, key2b;
  DevBuf<uint8_t> tflag;
  gid.alloc(nt); gid2.alloc(nt); sufc.alloc(nt); sufc2.alloc(nt); perm.alloc(nt); perm2.alloc(nt); head.alloc(nt);
  keep.alloc(nt); key2.alloc(nt); key2b.alloc(nt); tflag.alloc(nt);
  DevBuf<uint32_t> slots2;
  slots2.alloc(nt);
  // group ids = compacted index of the run head (inclusive max-scan of head markers)
  k_group_heads<<<grid_for(nt, B), B, 0, st>>>(keys_b, slots.p, nt, head.p);
  {
    size_t tb2 = 0;
    cub::DeviceScan::InclusiveScan(nullptr, tb2, head.p, gid.p, MaxOp(), (int)nt, st);
    cub::DeviceScan::InclusiveScan(tmp.ensure(tb2), tb2, head.p, gid.p, MaxOp(), (int)nt, st);
  }
  k_gather_suffix<<<grid_for(nt, B), B, 0, st>>>(sa_out, slots.p, nt, sufc.p);
  uint64_t depth = cm.syms;
  uint32_t* cur_slots = slots.p;
  uint32_t* alt_slots = slots2.p;
  for (int round = 0; nt > 0; ++round) {
    if (round > 100000) { set_error("text too repetitive for the GPU suffix sorter"); throw CudaFail{DG_ERR_UNSUPPORTED}; }
    // key2 = next 21 symbols; order by (gid, key2): stable sort by key2, then stable sort by gid
    k_make_keys<<<grid_for(nt, B), B, 0, st>>>(text, n, cm, sufc.p, nt, depth, key2.p);
    k_iota<<<grid_for(nt, B), B, 0, st>>>(perm.p, nt);
    size_t tb2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb2, key2.p, key2b.p, perm.p, perm2.p, (int)nt, 0, key_bits, st);
    cub::DeviceRadixSort::SortPairs(tmp.ensure(tb2), tb2, key2.p, key2b.p, perm.p, perm2.p, (int)nt, 0, key_bits, st);
    k_gather_u32<<<grid_for(nt, B), B, 0, st>>>(gid.p, perm2.p, nt, gid2.p);
    cub::DeviceRadixSort::SortPairs(nullptr, tb2, gid2.p, gid.p, perm2.p, perm.p, (int)nt, 0, 32, st);
    cub::DeviceRadixSort::SortPairs(tmp.ensure(tb2), tb2, gid2.p, gid.p, perm2.p, perm.p, (int)nt, 0, 32, st);
    // now gid.p = sorted gids (same sequence as before), perm.p = element order
    k_gather_u32<<<grid_for(nt, B), B, 0, st>>>(sufc.p, perm.p, nt, sufc2.p);
    k_gather_u64<<<grid_for(nt, B), B, 0, st>>>(key2.p, perm.p, nt, key2b.p);
    k_scatter_sa<<<grid_for(nt, B), B, 0, st>>>(sa_out, cur_slots, sufc2.p, nt);
    // which are still tied?
    k_retie<<<grid_for(nt, B), B, 0, st>>>(gid.p, key2b.p, nt, tflag.p, head.p);
    cub::DeviceScan::InclusiveScan(nullptr, tb2, head.p, gid2.p, MaxOp(), (int)nt, st);
    cub::DeviceScan::InclusiveScan(tmp.ensure(tb2), tb2, head.p, gid2.p, MaxOp(), (int)nt, st);
    thrust::counting_iterator<uint32_t> it(0);
    cub::DeviceSelect::Flagged(nullptr, tb2, it, tflag.p, keep.p, nsel.p, (int)nt, st);
    cub::DeviceSelect::Flagged(tmp.ensure(tb2), tb2, it, tflag.p, keep.p, nsel.p, (int)nt, st);
    uint32_t nk = 0;
    DG_CUDA(cudaMemcpyAsync(&nk, nsel.p, 4, cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaStreamSynchronize(st));
    if (nk == 0) break;
    // compact survivors: slots, suffixes, group ids (head index is unique per group; keep it as id)
    k_gather_u32<<<grid_for(nk, B), B, 0, st>>>(cur_slots, keep.p, nk, alt_slots);
    k_gather_u32<<<grid_for(nk, B), B, 0, st>>>(sufc2.p, keep.p, nk, sufc.p);
    k_renumber<<<grid_for(nk, B), B, 0, st>>>(gid2.p, keep.p, nk, gid.p);
    std::swap(cur_slots, alt_slots);
    nt = nk;
    depth += cm.syms;
  }
}

}  // namespace

// ---------------------------------------------------------------- shared tail of both paths
static void finish_from_bwt(dg_index* ix, const uint8_t* d_bwt, const std::vector<uint32_t>& Cb,
                            const std::vector<uint8_t>& present) {
  cudaStream_t st = ix->stream;
  const unsigned B = 256;
  uint64_t n = ix->n;
  uint64_t nblocks = (n >> 6) + 1;
  Temp tmp;
  ix->occ.alloc(nblocks);
  {
    DevBuf<uint32_t> ca, cc, cg, ct;
    DevBuf<unsigned long long> nexc;
    ca.alloc(nblocks); cc.alloc(nblocks); cg.alloc(nblocks); ct.alloc(nblocks); nexc.alloc(1);
    DG_CUDA(cudaMemsetAsync(nexc.p, 0, 8, st));
    k_pack_blocks<<<grid_for(nblocks, B), B, 0, st>>>(d_bwt, n, nblocks, ix->occ.p, ca.p, cc.p, cg.p, ct.p, nexc.p);
    uint32_t* arrs[4] = {ca.p, cc.p, cg.p, ct.p};
    for (int c = 0; c < 4; ++c) {
      size_t tb = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb, arrs[c], arrs[c], (int)nblocks, st);
      cub::DeviceScan::ExclusiveSum(tmp.ensure(tb), tb, arrs[c], arrs[c], (int)nblocks, st);
    }
    k_fill_counts<<<grid_for(nblocks, B), B, 0, st>>>(nblocks, ix->occ.p, ca.p, cc.p, cg.p, ct.p);
    unsigned long long ne = 0;
    DG_CUDA(cudaMemcpyAsync(&ne, nexc.p, 8, cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaStreamSynchronize(st));
    if (ne >= (1ULL << 32)) { set_error("too many non-ACGT BWT symbols"); throw CudaFail{DG_ERR_UNSUPPORTED}; }
    ix->n_exc = (uint32_t)ne;
  }
  // exception lists
  uint64_t nflagw = (((n >> kFlagShift) + 1) + 31) / 32;
  ix->excflag.alloc(nflagw);
  DG_CUDA(cudaMemsetAsync(ix->excflag.p, 0, nflagw * 4, st));
  ix->exc_pos.alloc(ix->n_exc ? ix->n_exc : 1);
  ix->exc_sym.alloc(ix->n_exc ? ix->n_exc : 1);
  ix->rare_pos.alloc(ix->n_exc ? ix->n_exc : 1);
  ix->rare_off.alloc(257);
  std::vector<uint32_t> h_off(257, 0);
  if (ix->n_exc) {
    DevBuf<uint32_t> nsel, hist;
    nsel.alloc(1); hist.alloc(256);
    DG_CUDA(cudaMemsetAsync(hist.p, 0, 256 * 4, st));
    IsException pred{d_bwt};
    uint64_t done = 0;
    for (uint64_t start = 0; start < n; start += (1ULL << 30)) {
      uint64_t cnt = std::min<uint64_t>(1ULL << 30, n - start);
      thrust::counting_iterator<uint32_t> it((uint32_t)start);
      size_t tb = 0;
      cub::DeviceSelect::If(nullptr, tb, it, ix->exc_pos.p + done, nsel.p, (int)cnt, pred, st);
      cub::DeviceSelect::If(tmp.ensure(tb), tb, it, ix->exc_pos.p + done, nsel.p, (int)cnt, pred, st);
      uint32_t got = 0;
      DG_CUDA(cudaMemcpyAsync(&got, nsel.p, 4, cudaMemcpyDeviceToHost, st));
      DG_CUDA(cudaStreamSynchronize(st));
      done += got;
    }
    if (done != ix->n_exc) { set_error("internal: exception count mismatch"); throw CudaFail{DG_ERR_CUDA}; }
    k_exc_fill<<<grid_for(ix->n_exc, B), B, 0, st>>>(d_bwt, ix->exc_pos.p, ix->n_exc, ix->exc_sym.p, ix->excflag.p, hist.p);
    // group by symbol (stable radix sort on the byte keeps rows ascending inside a group)
    DevBuf<uint8_t> symo;
    symo.alloc(ix->n_exc);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, ix->exc_sym.p, symo.p, ix->exc_pos.p, ix->rare_pos.p, (int)ix->n_exc, 0, 8, st);
    cub::DeviceRadixSort::SortPairs(tmp.ensure(tb), tb, ix->exc_sym.p, symo.p, ix->exc_pos.p, ix->rare_pos.p, (int)ix->n_exc, 0, 8, st);
    std::vector<uint32_t> h_hist(256);
    DG_CUDA(cudaMemcpyAsync(h_hist.data(), hist.p, 256 * 4, cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaStreamSynchronize(st));
    for (int s = 0; s < 256; ++s) h_off[s + 1] = h_off[s] + h_hist[s];
  }
  DG_CUDA(cudaMemcpyAsync(ix->rare_off.p, h_off.data(), 257 * 4, cudaMemcpyHostToDevice, st));
  ix->Cb.alloc(256);
  ix->present.alloc(256);
  ix->h_Cb = Cb;
  ix->h_present = present;
  DG_CUDA(cudaMemcpyAsync(ix->Cb.p, Cb.data(), 256 * 4, cudaMemcpyHostToDevice, st));
  DG_CUDA(cudaMemcpyAsync(ix->present.p, present.data(), 256, cudaMemcpyHostToDevice, st));
  ix->C4[0] = Cb['A']; ix->C4[1] = Cb['C']; ix->C4[2] = Cb['G']; ix->C4[3] = Cb['T'];
  DG_CUDA(cudaStreamSynchronize(st));
}

static void build_kmer_table(dg_index* ix) {
  cudaStream_t st = ix->stream;
  uint32_t K = 0;
  if (const char* e = getenv("DG_KMER")) K = (uint32_t)atoi(e);
  if (K == 0) {
    // about one text suffix per table cell, capped at 15 (8 GiB of uint2 at 3 Gb: HBM capacity
    // traded for one fewer dependent DRAM access per neighbour string)
    K = 1;
    while (K < 15 && (1ULL << (2 * (K + 1))) <= ix->n) ++K;
  }
  if (K > 16) K = 16;
  ix->K = K;
  DevBuf<uint2> a, b;
  ix->kmer.alloc(1ULL << (2 * K));
  if (K >= 1) a.alloc(1ULL << (2 * (K - 1)));
  if (K >= 2) b.alloc(1ULL << (2 * (K - 2)));
  // levels alternate between a and b so that level K lands in ix->kmer
  uint2 root = make_uint2(0, (uint32_t)ix->n);
  std::vector<uint2*> lv(K + 1);
  lv[K] = ix->kmer.p;
  for (int k = (int)K - 1; k >= 0; --k) lv[k] = ((K - 1 - k) % 2 == 0) ? a.p : b.p;
  DG_CUDA(cudaMemcpyAsync(lv[0], &root, sizeof(uint2), cudaMemcpyHostToDevice, st));
  IndexView v = ix->view();
  for (uint32_t k = 0; k < K; ++k) {
    uint64_t cnt = 1ULL << (2 * (k + 1));
    k_kmer_level<<<grid_for(cnt, 256), 256, 0, st>>>(v, lv[k], lv[k + 1], k);
  }
  DG_CUDA(cudaGetLastError());
  DG_CUDA(cudaStreamSynchronize(st));
}

// HBM capacity traded for locate latency: the whole suffix array (4 n bytes) unless DG_FULL_SA=0
static bool want_full_sa() {
  const char* e = getenv("DG_FULL_SA");
  return !(e && atoi(e) == 0);
}

// needs ix->text
static void build_presence_bitmap(dg_index* ix) {
  cudaStream_t st = ix->stream;
  uint32_t KB = 0;
  if (const char* e = getenv("DG_BITMAP_K")) {
    KB = (uint32_t)atoi(e);  // 0 switches the filter off
  } else {
    // about 1 set bit in 16 at most, capped at 18 (8 GiB of bits at 3 Gb)
    KB = 8;
    while (KB < 18 && (1ULL << (2 * KB)) < 16 * ix->n) ++KB;
  }
  if (KB > 19) KB = 19;
  if (KB && KB < 7) KB = 7;
  ix->KB = KB;
  if (!KB) return;
  uint64_t nthreads = (ix->n + kPresenceChunk - 1) / kPresenceChunk;
  // the left-anchored twin of a bitmap (DG_BITMAP_LEFT=0: none) when it fits comfortably
  const char* lf = getenv("DG_BITMAP_LEFT");
  const bool want_left = !(lf && atoi(lf) == 0);
  auto build = [&](DevBuf<uint32_t>& buf, uint32_t kb, DevBuf<uint32_t>* left) {
    uint64_t words = (1ULL << (2 * kb)) >> 5;
    buf.alloc(words);
    DG_CUDA(cudaMemsetAsync(buf.p, 0, words * 4, st));
    if (left) {
      size_t free_b = 0, total_b = 0;
      DG_CUDA(cudaMemGetInfo(&free_b, &total_b));
      if (want_left && words * 4 * 2 < free_b) {
        left->alloc(words);
        DG_CUDA(cudaMemsetAsync(left->p, 0, words * 4, st));
      } else {
        left = nullptr;
      }
    }
    k_presence<<<grid_for(nthreads, 128), 128, 0, st>>>(ix->text.p, ix->n, kb, buf.p, left ? left->p : nullptr);
    DG_CUDA(cudaGetLastError());
  };
  build(ix->present_kb, KB, &ix->present_kb_l);
  // neighbours: KB - 1 always (a quarter of the size); KB + 1 (four times the size) when it fits
  // comfortably in what is left of the HBM (DG_BITMAP_EXTRA=0 switches both off)
  const char* ex = getenv("DG_BITMAP_EXTRA");
  if (!(ex && atoi(ex) == 0)) {
    build(ix->present_lo, KB - 1, nullptr);
    size_t free_b = 0, total_b = 0;
    DG_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t hi_bytes = (1ULL << (2 * (KB + 1))) >> 3;
    if (KB + 1 <= 19 && hi_bytes * 3 < free_b) build(ix->present_hi, KB + 1, &ix->present_hi_l);
  }
  DG_CUDA(cudaStreamSynchronize(st));
}

static dg_index* new_index(int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available (the dicey_b200 library has no CPU path)");
    throw CudaFail{DG_ERR_CUDA};
  }
  if (device < 0 || device >= ndev) { set_error("bad device ordinal"); throw CudaFail{DG_ERR_ARG}; }
  DG_CUDA(cudaSetDevice(device));
  dg_index* ix = new dg_index();
  ix->device = device;
  DG_CUDA(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
  {
    // batch temporaries come from the stream-ordered pool: keep freed memory in the pool instead
    // of returning it to the driver at every synchronisation
    cudaMemPool_t pool;
    DG_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = ~0ULL;
    DG_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  }
  // every hot access is a random 32-byte sector: ask L2 not to widen DRAM fetches beyond that
  {
    size_t gran = 32;
    if (const char* e = getenv("DG_L2_FETCH")) gran = (size_t)atoi(e);
    if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    cudaGetLastError();
  }
  ix->cum.alloc(1);
  DG_CUDA(cudaMemset(ix->cum.p, 0, 8));
  return ix;
}

int build_from_fm9(const char* path, int device, dg_index** out) {
  const bool trace = getenv("DG_TRACE") != nullptr;   // wall time of every load stage on stderr
  auto t_last = std::chrono::steady_clock::now();
  auto stage = [&](const char* what) {
    if (!trace) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[load] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  Fm9 f;
  std::string err;
  int rc = fm9_parse(path, f, err);
  stage("read + parse .fm9");
  if (rc) { set_error(err); return rc; }
  if (f.n >= (1ULL << 32) - 64) { set_error("text longer than 2^32 - 64 symbols is outside the device path"); return DG_ERR_UNSUPPORTED; }
  if (f.n < 2 || f.nodes.empty()) { set_error("empty index"); return DG_ERR_FORMAT; }
  dg_index* ix = nullptr;
  try {
    ix = new_index(device);
    cudaStream_t st = ix->stream;
    ix->n = f.n;
    ix->sigma = f.sigma;
    // alphabet: C per byte value (csa_alphabet_strategy.hpp: char2comp / C)
    std::vector<uint32_t> Cb(256, 0);
    std::vector<uint8_t> present(256, 0);
    for (int s = 0; s < 256; ++s) {
      uint8_t cc = f.char2comp[s];
      if (cc != 0 || s == 0) { present[s] = 1; Cb[s] = (uint32_t)f.C[cc]; }
    }
    // wavelet tree -> BWT bytes
    DevBuf<uint8_t> bwt;
    {
      size_t nn = f.nodes.size();
      std::vector<uint64_t> bvp(nn), bvr(nn);
      std::vector<uint16_t> c0(nn), c1(nn);
      for (size_t i = 0; i < nn; ++i) {
        bvp[i] = f.nodes[i].bv_pos; bvr[i] = f.nodes[i].bv_pos_rank;
        c0[i] = f.nodes[i].child[0]; c1[i] = f.nodes[i].child[1];
      }
      DevBuf<uint64_t> d_bv, d_bb, d_bvp, d_bvr;
      DevBuf<uint16_t> d_c0, d_c1;
      d_bv.alloc(f.bv.words + 1); d_bb.alloc(f.rank_bb.words ? f.rank_bb.words : 1); d_bvp.alloc(nn); d_bvr.alloc(nn); d_c0.alloc(nn); d_c1.alloc(nn);
      if (f.bv.words) DG_CUDA(cudaMemcpyAsync(d_bv.p, f.bv.p, f.bv.bytes(), cudaMemcpyHostToDevice, st));
      DG_CUDA(cudaMemsetAsync(d_bv.p + f.bv.words, 0, 8, st));
      if (f.rank_bb.words) DG_CUDA(cudaMemcpyAsync(d_bb.p, f.rank_bb.p, f.rank_bb.bytes(), cudaMemcpyHostToDevice, st));
      DG_CUDA(cudaMemcpyAsync(d_bvp.p, bvp.data(), nn * 8, cudaMemcpyHostToDevice, st));
      DG_CUDA(cudaMemcpyAsync(d_bvr.p, bvr.data(), nn * 8, cudaMemcpyHostToDevice, st));
      DG_CUDA(cudaMemcpyAsync(d_c0.p, c0.data(), nn * 2, cudaMemcpyHostToDevice, st));
      DG_CUDA(cudaMemcpyAsync(d_c1.p, c1.data(), nn * 2, cudaMemcpyHostToDevice, st));
      bwt.alloc(f.n);
      WtView w{d_bv.p, d_bb.p, d_bvp.p, d_bvr.p, d_c0.p, d_c1.p};
      k_decode_bwt<<<grid_for(f.n, 256), 256, 0, st>>>(w, f.n, bwt.p);
      DG_CUDA(cudaGetLastError());
      DG_CUDA(cudaStreamSynchronize(st));
    }
    stage("context, wavelet tree -> BWT");
    finish_from_bwt(ix, bwt.p, Cb, present);
    bwt.release();
    stage("occ blocks, exceptions");
    // SA / ISA samples: the packed int_vector<0> payloads go up as they are and are widened to u32 there
    {
      DevBuf<uint64_t> packed;
      packed.alloc(std::max(f.sa_words.words, f.isa_words.words) + 1);
      auto unpack = [&](const Fm9Span& sp, uint64_t count, uint8_t width, uint32_t* dst) {
        if (sp.words) DG_CUDA(cudaMemcpyAsync(packed.p, sp.p, sp.bytes(), cudaMemcpyHostToDevice, st));
        DG_CUDA(cudaMemsetAsync(packed.p + sp.words, 0, 8, st));
        k_unpack_ints<<<grid_for(count, 256), 256, 0, st>>>(packed.p, count, width, dst);
        DG_CUDA(cudaGetLastError());
      };
      ix->sa_samples.alloc(f.sa_count);
      unpack(f.sa_words, f.sa_count, f.sa_width, ix->sa_samples.p);
      ix->isa_samples.alloc(f.isa_count + 1);
      DG_CUDA(cudaMemsetAsync(ix->isa_samples.p + f.isa_count, 0, 4, st));
      unpack(f.isa_words, f.isa_count, f.isa_width, ix->isa_samples.p);
      DG_CUDA(cudaStreamSynchronize(st));
    }
    f.unmap();
    stage("SA / ISA samples");
    build_kmer_table(ix);
    stage("k-mer interval table");
    // text from the ISA samples
    ix->text.alloc(f.n + 64);
    DG_CUDA(cudaMemsetAsync(ix->text.p, 0, f.n + 64, st));
    uint64_t nchains = (f.n - 1) / 64 + 1;
    if (want_full_sa()) ix->sa_full.alloc(f.n);
    IndexView wv = ix->view();
    wv.sa_full = nullptr;  // the walk itself must use the samples
    k_rebuild_text<<<grid_for(nchains, 128), 128, 0, st>>>(wv, ix->isa_samples.p, nchains, ix->text.p, ix->sa_full.p);
    DG_CUDA(cudaGetLastError());
    DG_CUDA(cudaStreamSynchronize(st));
    stage("text + full suffix array");
    build_presence_bitmap(ix);
    stage("presence bitmaps");
    *out = ix;
    return DG_OK;
  } catch (CudaFail& e) {
    if (ix) dg_index_close(ix);
    return e.code;
  } catch (std::bad_alloc&) {
    if (ix) dg_index_close(ix);
    set_error("out of host memory");
    return DG_ERR_NOMEM;
  }
}

// Builds everything from ix->text (n bytes incl. sentinel, already on the device).
static void build_from_device_text(dg_index* ix) {
  cudaStream_t st = ix->stream;
  const unsigned B = 256;
  uint64_t n = ix->n;
  const uint8_t* text = ix->text.p;
  Temp tmp;
  // alphabet
  DevBuf<unsigned long long> d_hist;
  d_hist.alloc(256);
  DG_CUDA(cudaMemsetAsync(d_hist.p, 0, 256 * 8, st));
  k_byte_hist<<<148 * 8, 256, 0, st>>>(text, n, d_hist.p);
  std::vector<unsigned long long> hist(256);
  DG_CUDA(cudaMemcpyAsync(hist.data(), d_hist.p, 256 * 8, cudaMemcpyDeviceToHost, st));
  DG_CUDA(cudaStreamSynchronize(st));
  if (hist[0] != 1) { set_error("text must not contain NUL bytes"); throw CudaFail{DG_ERR_ARG}; }
  CodeMap cm;
  memset(cm.code, 0, sizeof(cm.code));
  std::vector<uint32_t> Cb(256, 0);
  std::vector<uint8_t> present(256, 0);
  uint32_t sigma = 0;
  uint64_t run = 0;
  for (int s = 0; s < 256; ++s) {
    Cb[s] = (uint32_t)run;
    if (hist[s]) {
      present[s] = 1;
      if (sigma >= 32) { set_error("more than 31 distinct symbols: outside the GPU builder (use `dicey index`)"); throw CudaFail{DG_ERR_UNSUPPORTED}; }
      cm.code[s] = (uint8_t)sigma++;
      run += hist[s];
    }
  }
  for (int s = 0; s < 256; ++s) if (!present[s]) Cb[s] = 0;
  ix->sigma = sigma;
  cm.bits = sigma <= 8 ? 3 : (sigma <= 16 ? 4 : 5);   // DNA texts (A C G T N, separator, sentinel) keep 21 symbols per key
  cm.syms = (uint8_t)(63 / cm.bits);
  if (cm.syms > kKeySyms) cm.syms = kKeySyms;
  const uint32_t nbuckets = 1u << (2 * cm.bits);
  // bucket sizes by the first two symbols
  DevBuf<unsigned long long> d_ph;
  d_ph.alloc(nbuckets);
  DG_CUDA(cudaMemsetAsync(d_ph.p, 0, nbuckets * 8, st));
  k_pair_hist<<<148 * 8, 256, 0, st>>>(text, n, cm, d_ph.p);
  std::vector<unsigned long long> ph(nbuckets);
  DG_CUDA(cudaMemcpyAsync(ph.data(), d_ph.p, nbuckets * 8, cudaMemcpyDeviceToHost, st));
  DG_CUDA(cudaStreamSynchronize(st));
  uint64_t maxb = 0;
  for (auto v : ph) maxb = std::max<uint64_t>(maxb, v);
  DevBuf<uint32_t> sa, suf, suf_b, nsel;
  DevBuf<uint64_t> keys_a, keys_b;
  DevBuf<uint8_t> flag;
  sa.alloc(n); suf.alloc(maxb); suf_b.alloc(maxb); keys_a.alloc(maxb); keys_b.alloc(maxb); flag.alloc(maxb); nsel.alloc(1);
  uint64_t out_pos = 0;
  for (uint32_t b = 0; b < nbuckets; ++b) {
    uint64_t cnt = ph[b];
    if (!cnt) continue;
    InBucket pred{text, n, cm, b};
    uint64_t done = 0;
    for (uint64_t start = 0; start < n; start += (1ULL << 30)) {
      uint64_t c = std::min<uint64_t>(1ULL << 30, n - start);
      thrust::counting_iterator<uint32_t> it((uint32_t)start);
      size_t tb = 0;
      cub::DeviceSelect::If(nullptr, tb, it, suf.p + done, nsel.p, (int)c, pred, st);
      cub::DeviceSelect::If(tmp.ensure(tb), tb, it, suf.p + done, nsel.p, (int)c, pred, st);
      uint32_t got = 0;
      DG_CUDA(cudaMemcpyAsync(&got, nsel.p, 4, cudaMemcpyDeviceToHost, st));
      DG_CUDA(cudaStreamSynchronize(st));
      done += got;
    }
    if (done != cnt) { set_error("internal: bucket count mismatch"); throw CudaFail{DG_ERR_CUDA}; }
    sort_bucket(text, n, cm, suf.p, cnt, sa.p + out_pos, keys_a.p, keys_b.p, suf_b.p, flag.p, tmp, st);
    out_pos += cnt;
  }
  if (out_pos != n) { set_error("internal: suffix count mismatch"); throw CudaFail{DG_ERR_CUDA}; }
  suf.release(); suf_b.release(); keys_a.release(); keys_b.release(); flag.release();
  // BWT, SA samples, ISA samples
  DevBuf<uint8_t> bwt;
  bwt.alloc(n);
  uint64_t nsa = (n + kSaSample - 1) / kSaSample, nisa = (n - 1) / 64 + 1;
  ix->sa_samples.alloc(nsa);
  ix->isa_samples.alloc(nisa + 1);
  k_bwt_from_sa<<<grid_for(n, B), B, 0, st>>>(text, sa.p, n, bwt.p, ix->sa_samples.p, ix->isa_samples.p);
  DG_CUDA(cudaGetLastError());
  DG_CUDA(cudaStreamSynchronize(st));
  if (want_full_sa()) { ix->sa_full.p = sa.p; ix->sa_full.count = sa.count; sa.p = nullptr; sa.count = 0; }
  else sa.release();
  finish_from_bwt(ix, bwt.p, Cb, present);
  bwt.release();
  build_kmer_table(ix);
  build_presence_bitmap(ix);
}

int build_from_text_host(const uint8_t* text, uint64_t len, int device, dg_index** out) {
  if (!text || !out) { set_error("null argument"); return DG_ERR_ARG; }
  if (len + 1 >= (1ULL << 32) - 64) { set_error("text longer than 2^32 - 64 symbols is outside the device path"); return DG_ERR_UNSUPPORTED; }
  dg_index* ix = nullptr;
  try {
    ix = new_index(device);
    ix->n = len + 1;
    ix->text.alloc(ix->n + 64);
    DG_CUDA(cudaMemsetAsync(ix->text.p, 0, ix->n + 64, ix->stream));
    DG_CUDA(cudaMemcpyAsync(ix->text.p, text, len, cudaMemcpyHostToDevice, ix->stream));
    DG_CUDA(cudaStreamSynchronize(ix->stream));
    build_from_device_text(ix);
    *out = ix;
    return DG_OK;
  } catch (CudaFail& e) {
    if (ix) dg_index_close(ix);
    return e.code;
  }
}

int build_synthetic(uint64_t seed, uint32_t nrec, uint64_t reclen, int device, dg_index** out) {
  if (!out || nrec == 0 || reclen == 0) { set_error("bad argument"); return DG_ERR_ARG; }
  uint64_t n = (uint64_t)nrec * (reclen + 1) + 1;
  if (n >= (1ULL << 32) - 64) { set_error("text longer than 2^32 - 64 symbols is outside the device path"); return DG_ERR_UNSUPPORTED; }
  dg_index* ix = nullptr;
  try {
    ix = new_index(device);
    ix->n = n;
    ix->text.alloc(n + 64);
    DG_CUDA(cudaMemsetAsync(ix->text.p, 0, n + 64, ix->stream));
    k_synth_text<<<grid_for(n, 256), 256, 0, ix->stream>>>(seed, reclen, n, ix->text.p);
    DG_CUDA(cudaGetLastError());
    build_from_device_text(ix);
    // records: nrec x (reclen + 1), as util.h:201 would report them
    std::vector<uint32_t> sl(nrec, (uint32_t)(reclen + 1));
    int rc = dg_index_set_records(ix, sl.data(), nrec);
    if (rc) { dg_index_close(ix); return rc; }
    *out = ix;
    return DG_OK;
  } catch (CudaFail& e) {
    if (ix) dg_index_close(ix);
    return e.code;
  }
}

}  // namespace dg

// ================================================================================================
// .fm9 writer: store_to_checked_file (index.h:122 -> io.hpp:814-828) of a csa_wt<> whose wavelet
// tree, rank blocks, samples and alphabet are rebuilt from the device index.  The Huffman shape
// follows _huff_shape::construct_tree (wt_huff.hpp:72-100) and _byte_tree's BFS numbering
// (wt_helper.hpp:199-269), and the two select_support_mcl sections are rebuilt by fm9_select.hpp,
// so that the written bytes equal SDSL's.
namespace dg {
namespace {

constexpr int kWtInternal = 32;   // internal nodes of the Huffman-shaped tree (alphabets of up to 32 symbols)
constexpr int kWtClasses = 32;    // symbol classes: 0..3 = A C G T, 4.. = the other symbols of the text
constexpr int kWtDepth = 32;      // longest root-to-leaf path
struct WtShape {            // device-side description of the tree for k_wt_bits (lives in device memory)
  int n_internal;           // internal nodes, numbered 0..n_internal-1 in BFS order among internals
  int n_rare;
  uint64_t bv_pos[kWtInternal];        // start of each internal node's bit vector
  uint8_t path_len[256];               // per byte symbol
  uint8_t path_node[256][kWtDepth];    // internal node visited at depth d
  uint8_t path_bit[256][kWtDepth];     // branch taken at depth d
  uint8_t member[kWtInternal][kWtClasses];   // member[v][k]: k-th symbol class lies under v
  uint8_t rare_sym[kWtClasses];        // byte values of the classes 4..
};

// one thread per occurrence block: appends the block's 64 symbols to every internal node on
// their paths; bits are staged per node and flushed with one atomicOr per touched word
__global__ void k_wt_bits(IndexView ix, const WtShape* __restrict__ shp, uint64_t nblocks, unsigned long long* __restrict__ bv) {
  uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  uint64_t base = b << 6;
  if (base >= ix.n) return;
  const WtShape& sh = *shp;
  OccBlock blk = load_block(ix.occ + b);
  // class counts before the block
  uint64_t cls[kWtClasses];
  for (int c = 0; c < 4; ++c) cls[c] = blk.cnt[c];
  for (int k = 0; k < sh.n_rare; ++k) {
    uint8_t s = sh.rare_sym[k];
    uint32_t a = ix.rare_off[s], e = ix.rare_off[s + 1];
    cls[4 + k] = lower_bound_u32(ix.rare_pos, a, e, (uint32_t)base) - a;
  }
  uint64_t at[kWtInternal];          // next bit position per internal node
  unsigned long long acc[kWtInternal];
  for (int v = 0; v < sh.n_internal; ++v) {
    uint64_t cnt = 0;
    for (int k = 0; k < 4 + sh.n_rare; ++k) if (sh.member[v][k]) cnt += cls[k];
    at[v] = sh.bv_pos[v] + cnt;
    acc[v] = 0;
  }
  bool flagged = region_flag(ix, base);
  uint32_t ek = flagged ? lower_bound_u32(ix.exc_pos, 0, ix.n_exc, (uint32_t)base) : 0;
  for (int o = 0; o < 64; ++o) {
    uint64_t i = base + o;
    if (i >= ix.n) break;
    uint8_t s = code_base((int)(((blk.hi >> o) & 1) << 1 | ((blk.lo >> o) & 1)));
    if (flagged && ek < ix.n_exc && ix.exc_pos[ek] == (uint32_t)i) { s = ix.exc_sym[ek]; ++ek; }
    int len = sh.path_len[s];
    for (int d = 0; d < len; ++d) {
      int v = sh.path_node[s][d];
      uint64_t pos = at[v]++;
      if (sh.path_bit[s][d]) acc[v] |= 1ULL << (pos & 63);
      if ((pos & 63) == 63) {  // word complete: flush
        if (acc[v]) atomicOr(&bv[pos >> 6], acc[v]);
        acc[v] = 0;
      }
    }
  }
  for (int v = 0; v < sh.n_internal; ++v)
    if (acc[v]) atomicOr(&bv[(at[v] - 1) >> 6], acc[v]);
}

// rank_support_v<1> (rank_support_v.hpp:57-96): per 512-bit superblock the absolute count and the
// seven 9-bit relative counts
__global__ void k_sb_popc(const unsigned long long* __restrict__ bv, uint64_t nwords, uint64_t nsb, uint64_t* __restrict__ sbsum) {
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nsb) return;
  uint64_t s = 0;
  for (int t = 0; t < 8; ++t) { uint64_t w = 8 * k + t; if (w < nwords) s += __popcll(bv[w]); }
  sbsum[k] = s;
}
__global__ void k_rank_blocks(const unsigned long long* __restrict__ bv, uint64_t nwords, uint64_t nsb,
                              const uint64_t* __restrict__ sbabs, uint64_t* __restrict__ bb) {
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nsb) return;
  uint64_t rel = 0, s = 0;
  for (int t = 1; t <= 7; ++t) {
    uint64_t w = 8 * k + t - 1;
    if (w < nwords) s += __popcll(bv[w]);
    if (8 * k + t <= nwords) rel |= s << (63 - 9 * t);
  }
  bb[2 * k] = sbabs[k];
  bb[2 * k + 1] = rel;
}

struct PcNode { uint64_t freq, sym; uint64_t parent, child[2]; };
const uint64_t kUndef = 0xFFFFFFFFFFFFFFFFULL;

bool put(FILE* f, const void* p, size_t n) { return n == 0 || fwrite(p, 1, n, f) == n; }
bool put_u64(FILE* f, uint64_t v) { return put(f, &v, 8); }
bool put_int_vector(FILE* f, const uint64_t* words, uint64_t bits, uint8_t width) {
  return put_u64(f, ((uint64_t)width << 56) | bits) && put(f, words, ((bits + 63) >> 6) * 8);
}
void pack_ints(const std::vector<uint32_t>& v, uint8_t width, std::vector<uint64_t>& out) {
  uint64_t bits = (uint64_t)v.size() * width;
  out.assign((bits + 63) >> 6, 0);
  for (uint64_t i = 0; i < v.size(); ++i) {
    uint64_t bit = i * width, w = bit >> 6, o = bit & 63;
    out[w] |= (uint64_t)v[i] << o;
    if (o + width > 64) out[w + 1] |= (uint64_t)v[i] >> (64 - o);
  }
}

}  // namespace

int write_fm9(dg_index* ix, const char* path) {
  try {
    DG_CUDA(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    uint64_t n = ix->n;
    // ---- symbol frequencies from C
    std::vector<int> syms;
    for (int s = 0; s < 256; ++s) if (ix->h_present[s]) syms.push_back(s);
    int sigma = (int)syms.size();
    std::vector<uint64_t> C(sigma + 1, 0), freq(256, 0);
    for (int i = 0; i < sigma; ++i) C[i] = ix->h_Cb[syms[i]];
    C[sigma] = n;
    for (int i = 0; i < sigma; ++i) freq[syms[i]] = C[i + 1] - C[i];
    if (sigma < 2) { set_error("alphabet too small to write"); return DG_ERR_UNSUPPORTED; }
    // ---- Huffman shape (wt_huff.hpp:72-100)
    std::vector<PcNode> tmp;
    typedef std::pair<uint64_t, uint64_t> P;
    std::priority_queue<P, std::vector<P>, std::greater<P>> pq;
    for (int s = 0; s < 256; ++s)
      if (freq[s] > 0) { pq.push(P(freq[s], tmp.size())); tmp.push_back(PcNode{freq[s], (uint64_t)s, kUndef, {kUndef, kUndef}}); }
    while (pq.size() > 1) {
      P v1 = pq.top(); pq.pop();
      P v2 = pq.top(); pq.pop();
      tmp[v1.second].parent = tmp.size();
      tmp[v2.second].parent = tmp.size();
      pq.push(P(v1.first + v2.first, tmp.size()));
      tmp.push_back(PcNode{v1.first + v2.first, 0, kUndef, {v1.second, v2.second}});
    }
    // ---- BFS numbering (wt_helper.hpp:199-236)
    size_t nn = tmp.size();
    std::vector<Fm9Node> nodes(nn);
    std::vector<uint64_t> nfreq(nn);
    std::vector<uint64_t> tchild0(nn), tchild1(nn);
    auto setnode = [&](size_t dst, const PcNode& src, uint16_t parent) {
      nfreq[dst] = src.freq;
      nodes[dst].bv_pos = 0;
      nodes[dst].bv_pos_rank = src.sym;
      nodes[dst].parent = parent;
      tchild0[dst] = src.child[0];
      tchild1[dst] = src.child[1];
      nodes[dst].child[0] = nodes[dst].child[1] = 0xFFFF;
    };
    setnode(0, tmp.back(), 0xFFFF);
    uint64_t bv_size = 0;
    size_t node_cnt = 1;
    std::deque<size_t> q;
    q.push_back(0);
    while (!q.empty()) {
      size_t idx = q.front();
      q.pop_front();
      nodes[idx].bv_pos = bv_size;
      bool internal = tchild0[idx] != kUndef;
      if (internal) {
        bv_size += nfreq[idx];
        uint64_t ch[2] = {tchild0[idx], tchild1[idx]};
        for (int k = 0; k < 2; ++k) {
          setnode(node_cnt, tmp[ch[k]], (uint16_t)idx);
          q.push_back(node_cnt);
          nodes[idx].child[k] = (uint16_t)node_cnt++;
        }
      }
    }
    uint16_t c_to_leaf[256];
    uint64_t pathv[256];
    for (int i = 0; i < 256; ++i) c_to_leaf[i] = 0xFFFF;
    for (size_t v = 0; v < nn; ++v) if (nodes[v].child[0] == 0xFFFF) c_to_leaf[(uint8_t)nodes[v].bv_pos_rank] = (uint16_t)v;
    for (uint32_t c = 0, prev_c = 0; c < 256; ++c) {
      if (c_to_leaf[c] != 0xFFFF) {
        uint16_t v = c_to_leaf[c];
        uint64_t pw = 0, pl = 0;
        while (v != 0) {
          pw <<= 1;
          if (nodes[nodes[v].parent].child[1] == v) pw |= 1ULL;
          ++pl;
          v = nodes[v].parent;
        }
        pathv[c] = pw | (pl << 56);
        prev_c = c;
      } else {
        pathv[c] = prev_c;
      }
    }
    // ---- device description of the paths
    WtShape sh;
    memset(&sh, 0, sizeof(sh));
    std::vector<int> internal_id(nn, -1);
    for (size_t v = 0; v < nn; ++v)
      if (nodes[v].child[0] != 0xFFFF) {
        if (sh.n_internal >= kWtInternal) { set_error("alphabet too large for the .fm9 writer"); return DG_ERR_UNSUPPORTED; }
        sh.bv_pos[sh.n_internal] = nodes[v].bv_pos;
        internal_id[v] = sh.n_internal++;
      }
    int cls_of[256];
    for (int i = 0; i < 256; ++i) cls_of[i] = -1;
    cls_of['A'] = 0; cls_of['C'] = 1; cls_of['G'] = 2; cls_of['T'] = 3;
    for (int s : syms)
      if (cls_of[s] < 0) {
        if (4 + sh.n_rare >= kWtClasses) { set_error("alphabet too large for the .fm9 writer"); return DG_ERR_UNSUPPORTED; }
        sh.rare_sym[sh.n_rare] = (uint8_t)s;
        cls_of[s] = 4 + sh.n_rare++;
      }
    for (int s : syms) {
      uint64_t pw = pathv[s] & ((1ULL << 56) - 1), pl = pathv[s] >> 56;
      if (pl > (uint64_t)kWtDepth) { set_error("wavelet tree too deep for the .fm9 writer"); return DG_ERR_UNSUPPORTED; }
      sh.path_len[s] = (uint8_t)pl;
      uint16_t v = 0;
      for (uint64_t d = 0; d < pl; ++d) {
        int bit = (int)((pw >> d) & 1);
        sh.path_node[s][d] = (uint8_t)internal_id[v];
        sh.path_bit[s][d] = (uint8_t)bit;
        sh.member[internal_id[v]][cls_of[s]] = 1;
        v = nodes[v].child[bit];
      }
    }
    // ---- wavelet tree bits + rank blocks on the device
    uint64_t nwords = (bv_size + 63) >> 6;
    uint64_t nsb = ((bv_size + 63) >> 9) + 1;
    DevBuf<unsigned long long> d_bv;
    DevBuf<uint64_t> d_sbsum, d_sbabs, d_bb;
    d_bv.alloc(nwords + 8);
    d_sbsum.alloc(nsb); d_sbabs.alloc(nsb); d_bb.alloc(2 * nsb);
    DG_CUDA(cudaMemsetAsync(d_bv.p, 0, (nwords + 8) * 8, st));
    uint64_t nblocks = (n + 63) >> 6;
    DevBuf<WtShape> d_sh;
    d_sh.alloc(1);
    DG_CUDA(cudaMemcpyAsync(d_sh.p, &sh, sizeof(WtShape), cudaMemcpyHostToDevice, st));
    k_wt_bits<<<grid_for(nblocks, 128), 128, 0, st>>>(ix->view(), d_sh.p, nblocks, d_bv.p);
    k_sb_popc<<<grid_for(nsb, 256), 256, 0, st>>>(d_bv.p, nwords, nsb, d_sbsum.p);
    {
      Temp tmpb;
      size_t tb = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb, d_sbsum.p, d_sbabs.p, (int)nsb, st);
      cub::DeviceScan::ExclusiveSum(tmpb.ensure(tb), tb, d_sbsum.p, d_sbabs.p, (int)nsb, st);
      k_rank_blocks<<<grid_for(nsb, 256), 256, 0, st>>>(d_bv.p, nwords, nsb, d_sbabs.p, d_bb.p);
      DG_CUDA(cudaGetLastError());
      DG_CUDA(cudaStreamSynchronize(st));
    }
    std::vector<uint64_t> bv(nwords), bb(2 * nsb);
    DG_CUDA(cudaMemcpy(bv.data(), d_bv.p, nwords * 8, cudaMemcpyDeviceToHost));
    DG_CUDA(cudaMemcpy(bb.data(), d_bb.p, 2 * nsb * 8, cudaMemcpyDeviceToHost));
    d_bv.release();
    // bv_pos_rank of internal nodes (wt_helper.hpp:272-279): rank1(bv_pos)
    auto rank1 = [&](uint64_t idx) -> uint64_t {
      const uint64_t* p = bb.data() + ((idx >> 8) & 0xFFFFFFFFFFFFFFFEULL);
      uint64_t r = p[0] + ((p[1] >> (63 - 9 * ((idx & 0x1FF) >> 6))) & 0x1FF);
      if (idx & 0x3F) r += (uint64_t)__builtin_popcountll(bv[idx >> 6] & ((1ULL << (idx & 0x3F)) - 1));
      return r;
    };
    for (size_t v = 0; v < nn; ++v) if (nodes[v].child[0] != 0xFFFF) nodes[v].bv_pos_rank = rank1(nodes[v].bv_pos);
    // ---- samples
    uint64_t nsa = (n + kSaSample - 1) / kSaSample, nisa = (n - 1) / 64 + 1;
    std::vector<uint32_t> sa(nsa), isa(nisa);
    DG_CUDA(cudaMemcpy(sa.data(), ix->sa_samples.p, nsa * 4, cudaMemcpyDeviceToHost));
    DG_CUDA(cudaMemcpy(isa.data(), ix->isa_samples.p, nisa * 4, cudaMemcpyDeviceToHost));
    uint8_t width = 1;
    while ((n >> width) != 0) ++width;  // bits::hi(n) + 1
    std::vector<uint64_t> saw, isaw;
    pack_ints(sa, width, saw);
    sa.clear(); sa.shrink_to_fit();
    pack_ints(isa, width, isaw);
    // ---- file
    FILE* f = fopen(path, "wb");
    if (!f) { set_error(std::string("cannot create ") + path); return DG_ERR_IO; }
    bool ok = put_u64(f, n) && put_u64(f, (uint64_t)sigma);
    ok = ok && put_int_vector(f, bv.data(), bv_size, 1);
    ok = ok && put_int_vector(f, bb.data(), (uint64_t)bb.size() * 64, 64);
    {
      // select_support_mcl<1>, <0> over m_bv (never read by count / locate / extract, written so that
      // the file equals SDSL's byte for byte); DG_FM9_NO_SELECT=1 writes them empty (arg_cnt = 0)
      if (getenv("DG_FM9_NO_SELECT")) {
        ok = ok && put_u64(f, 0) && put_u64(f, 0);
      } else {
        for (int one = 1; one >= 0 && ok; --one) {
          SelectMclWriter sel(bv.data(), bv_size, one == 1);
          ok = put(f, sel.bytes().data(), sel.bytes().size());
        }
      }
    }
    ok = ok && put_u64(f, (uint64_t)nn);
    for (size_t v = 0; ok && v < nn; ++v) {
      uint8_t rec[22];
      memcpy(rec, &nodes[v].bv_pos, 8);
      memcpy(rec + 8, &nodes[v].bv_pos_rank, 8);
      memcpy(rec + 16, &nodes[v].parent, 2);
      memcpy(rec + 18, &nodes[v].child[0], 2);
      memcpy(rec + 20, &nodes[v].child[1], 2);
      ok = put(f, rec, 22);
    }
    ok = ok && put(f, c_to_leaf, sizeof(c_to_leaf)) && put(f, pathv, sizeof(pathv));
    ok = ok && put_int_vector(f, saw.data(), (uint64_t)nsa * width, width);
    ok = ok && put_int_vector(f, isaw.data(), (uint64_t)nisa * width, width);
    // byte_alphabet (csa_alphabet_strategy.hpp:233-244)
    uint64_t c2c[32];
    memset(c2c, 0, sizeof(c2c));
    for (int i = 0; i < sigma; ++i) ((uint8_t*)c2c)[syms[i]] = (uint8_t)i;
    ok = ok && put_int_vector(f, c2c, 256 * 8, 8);
    std::vector<uint64_t> comp((sigma * 8 + 63) / 64, 0);
    for (int i = 0; i < sigma; ++i) ((uint8_t*)comp.data())[i] = (uint8_t)syms[i];
    ok = ok && put_int_vector(f, comp.data(), (uint64_t)sigma * 8, 8);
    ok = ok && put_int_vector(f, C.data(), (uint64_t)(sigma + 1) * 64, 64);
    uint16_t sg = (uint16_t)sigma;
    ok = ok && put(f, &sg, 2);
    ok = (fclose(f) == 0) && ok;
    if (!ok) { set_error(std::string("write failed: ") + path); return DG_ERR_IO; }
    FILE* c = fopen((std::string(path) + "_check").c_str(), "wb");
    uint64_t h = fm9_type_hash();
    if (!c || fwrite(&h, 8, 1, c) != 1) { if (c) fclose(c); set_error("cannot write the _check sidecar"); return DG_ERR_IO; }
    fclose(c);
    return DG_OK;
  } catch (CudaFail& e) {
    return e.code;
  }
}
}  // namespace dg
