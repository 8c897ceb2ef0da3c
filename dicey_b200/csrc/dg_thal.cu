// dg_thal.cu -- batched melting temperatures on the GPU: the thal() gate of `dicey search`
// (reference src/silica.h:508-519; the arithmetic is dg_thal.cuh) for every candidate site at once.
//
// k_thal_warp (the product path) gives every pair to one warp: the DP table, the list of paired
// cells and one row of end terms live in shared memory, the lanes evaluate the loop partners of a
// cell side by side and an arg-min shuffle picks the one the reference's sequential scan would
// have kept (thal_end1_tm_lanes in dg_thal.cuh has the argument).  k_thal is the sequential form,
// one pair per thread with the tables of a launch interleaved in a global scratch slab; it serves
// the pairs the warp form declines (none within primer3's length limit, but the condition is
// checked) and DG_THAL_SEQ=1 routes everything through it.  k_thal_wide takes the pairs with one
// side longer than THAL_MAX_ALIGN = 60 (thal.h:58, :2440-2451; up to THAL_MAX_SEQ = 10 000): one warp
// per pair again, the table (up to 9.6 MB) in a global work area of the warp (thal_end1_tm_wide).
// Built with -fmad=false: results equal the reference's bit for bit.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "dg_common.cuh"
#include "dg_thal.cuh"
#include "thal_params.hpp"

using namespace dg;

struct dg_thal {
  int device = 0;
  cudaStream_t st = nullptr;
  ThalParams* d_params = nullptr;
  ThalParams h_params;
  int sms = 148;
};

namespace {

struct ThalWarp {   // the Warp concept of dg_thal.cuh on 32 lanes
  static constexpr int n = 32;
  int lane;
  __device__ void sync() const { __syncwarp(); }
  __device__ unsigned ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
  __device__ unsigned lanemask_lt() const { return (1u << lane) - 1u; }
  __device__ bool all(bool p) const { return __all_sync(0xffffffffu, p) != 0; }
  __device__ bool any(bool p) const { return __any_sync(0xffffffffu, p) != 0; }
  __device__ unsigned group_mask(int first_lane, int lanes, bool member) const {
    if (!member) return 1u << lane;
    return lanes >= 32 ? 0xffffffffu : (((1u << lanes) - 1u) << first_lane);
  }
  // arg-min over the lanes of `mask` by (g, o): three integer min-reductions on an order-preserving
  // image of the double (callers pass g + 0.0, so there is one zero), then the winner's payload
  __device__ void argmin(unsigned mask, double& g, uint32_t& o, double& S, double& H) const {
    const long long b = __double_as_longlong(g);
    const unsigned long long key = b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ULL);
    const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
    const uint32_t mhi = __reduce_min_sync(mask, hi);
    bool in = hi == mhi;
    const uint32_t mlo = __reduce_min_sync(mask, in ? lo : 0xffffffffu);
    in = in && lo == mlo;
    const uint32_t mo = __reduce_min_sync(mask, in ? o : 0xffffffffu);
    in = in && o == mo;
    const unsigned winners = __ballot_sync(mask, in);   // never empty: the lanes holding (mhi, mlo) include one with o == mo
    const int who = __ffs(winners) - 1;
    g = __shfl_sync(mask, g, who);
    o = mo;
    S = __shfl_sync(mask, S, who);
    H = __shfl_sync(mask, H, who);
  }
};

// shared memory of one warp: table (S, H interleaved), the paired-cell list, the row starts, the
// two encoded sequences
__host__ __device__ inline size_t thal_warp_bytes(uint64_t cells) {
  return (size_t)cells * 16 + (((size_t)cells * 2 + 15) & ~(size_t)15) + 64 * 2 + 64 + 64 + 256;
}

__global__ void __launch_bounds__(128) k_thal_warp(const ThalParams* __restrict__ p, const uint8_t* __restrict__ s1,
                                                   const uint64_t* __restrict__ off1, const uint8_t* __restrict__ s2,
                                                   const uint64_t* __restrict__ off2, const uint32_t* __restrict__ ids, uint32_t count,
                                                   uint64_t cells, double* __restrict__ tm, uint8_t* __restrict__ ok) {
  extern __shared__ __align__(16) unsigned char smem[];
  const uint32_t wpb = blockDim.x >> 5, warp = threadIdx.x >> 5;
  unsigned char* base = smem + (size_t)warp * thal_warp_bytes(cells);
  double* tab = (double*)base;
  uint16_t* plist = (uint16_t*)(tab + 2 * cells);
  uint16_t* rstart = (uint16_t*)((unsigned char*)plist + (((size_t)cells * 2 + 15) & ~(size_t)15));
  uint8_t* n1 = (uint8_t*)(rstart + 64);
  uint8_t* n2 = n1 + 64;
  uint8_t* codes = n2 + 64;
  ThalWarp wp;
  wp.lane = threadIdx.x & 31;
  // the warps of the grid share the pairs round-robin (the grid is sized to what is resident at once)
  for (uint32_t t = blockIdx.x * wpb + warp; t < count; t += gridDim.x * wpb) {
    const uint32_t q = ids[t];
    const int len1 = (int)(off1[q + 1] - off1[q]), len2 = (int)(off2[q + 1] - off2[q]);
    double out = -kThalInf;
    int rc = 0;
    __syncwarp();
    if (len1 <= kThalMaxLen && len2 <= kThalMaxLen && (uint64_t)(len1 > 0 ? len1 : 0) * (uint64_t)(len2 > 0 ? len2 : 0) <= cells)
      rc = thal_end1_tm_lanes(wp, p, s1 + off1[q], len1, s2 + off2[q], len2, n1, n2, tab, plist, rstart, codes, &out);
    else if (len1 <= 0 || len2 <= 0)
      out = 0.0;
    if (wp.lane == 0) {
      tm[q] = out;
      ok[q] = (uint8_t)rc;   // 2: the sequential kernel recomputes this pair
    }
  }
}

__global__ void __launch_bounds__(128) k_thal(const ThalParams* __restrict__ p, const uint8_t* __restrict__ s1,
                                              const uint64_t* __restrict__ off1, const uint8_t* __restrict__ s2,
                                              const uint64_t* __restrict__ off2, const uint32_t* __restrict__ ids, uint32_t first,
                                              uint32_t count, double* __restrict__ scratch, uint64_t cells, double* __restrict__ tm,
                                              uint8_t* __restrict__ ok) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const uint32_t q = ids ? ids[first + t] : first + t;
  const int len1 = (int)(off1[q + 1] - off1[q]), len2 = (int)(off2[q + 1] - off2[q]);
  uint8_t n1[kThalMaxLen + 2], n2[kThalMaxLen + 2];
  double out = -kThalInf;
  bool good = false;
  if (len1 <= kThalMaxLen && len2 <= kThalMaxLen && (uint64_t)len1 * (uint64_t)len2 <= cells)
    good = thal_end1_tm(p, s1 + off1[q], len1, s2 + off2[q], len2, n1, n2, scratch + t, scratch + cells * count + t, &out, (long)count);
  else if (len1 <= 0 || len2 <= 0)
    out = 0.0;
  tm[q] = out;
  ok[q] = good ? 1 : 0;
}

// global work area of one warp of k_thal_wide: the table, then num1 / a1 / ra and num2 / b / rb
__host__ __device__ inline size_t thal_wide_bytes(uint64_t cells, uint32_t max1, uint32_t max2) {
  return ((size_t)cells * 16 + 3 * ((size_t)max1 + 2) + 3 * ((size_t)max2 + 2) + 255) & ~(size_t)255;
}

__global__ void __launch_bounds__(128) k_thal_wide(const ThalParams* __restrict__ p, const uint8_t* __restrict__ s1,
                                                   const uint64_t* __restrict__ off1, const uint8_t* __restrict__ s2,
                                                   const uint64_t* __restrict__ off2, const uint32_t* __restrict__ ids, uint32_t count,
                                                   uint64_t cells, uint32_t max1, uint32_t max2, unsigned char* slab,
                                                   double* __restrict__ tm, uint8_t* __restrict__ ok) {
  const uint32_t wpb = blockDim.x >> 5, warp = threadIdx.x >> 5;
  unsigned char* base = slab + (size_t)(blockIdx.x * wpb + warp) * thal_wide_bytes(cells, max1, max2);
  double* tab = (double*)base;
  uint8_t* n1 = (uint8_t*)(tab + 2 * cells);
  uint8_t* a1 = n1 + (max1 + 2);
  uint8_t* ra = a1 + (max1 + 2);
  uint8_t* n2 = ra + (max1 + 2);
  uint8_t* b = n2 + (max2 + 2);
  uint8_t* rb = b + (max2 + 2);
  ThalWarp wp;
  wp.lane = threadIdx.x & 31;
  for (uint32_t t = blockIdx.x * wpb + warp; t < count; t += gridDim.x * wpb) {
    const uint32_t q = ids[t];
    const int64_t l1 = (int64_t)(off1[q + 1] - off1[q]), l2 = (int64_t)(off2[q + 1] - off2[q]);
    double out = -kThalInf;
    int rc = 0;
    __syncwarp();
    if (l1 > 0 && l2 > 0 && l1 <= (int64_t)max1 && l2 <= (int64_t)max2 && thal_lengths_ok((int)l1, (int)l2) &&
        (uint64_t)l1 * (uint64_t)l2 <= cells) {
      const int len1 = (int)l1, len2 = (int)l2;
      rc = thal_end1_tm_wide(wp, p, s1 + off1[q], len1, s2 + off2[q], len2, n1, n2, tab, a1, ra, b, rb, &out);
      if (rc == 2) {   // the entropy cutoff: the sequential form on one lane, in the same work area
        __syncwarp();
        if (wp.lane == 0) rc = thal_end1_tm_any(p, s1 + off1[q], len1, s2 + off2[q], len2, n1, n2, tab, tab + 1, &out, 2) ? 1 : 0;
        __syncwarp();
      }
    } else if (l1 <= 0 || l2 <= 0) {
      out = 0.0;
    }
    if (wp.lane == 0) {
      tm[q] = out;
      ok[q] = (uint8_t)rc;
    }
  }
}

int finish_open(dg_thal* t, int device, dg_thal** out) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available (the dicey_b200 library has no CPU path)");
    delete t;
    return DG_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { set_error("bad device ordinal"); delete t; return DG_ERR_ARG; }
  try {
    DG_CUDA(cudaSetDevice(device));
    t->device = device;
    DG_CUDA(cudaStreamCreateWithFlags(&t->st, cudaStreamNonBlocking));
    DG_CUDA(cudaDeviceGetAttribute(&t->sms, cudaDevAttrMultiProcessorCount, device));
    DG_CUDA(cudaMalloc((void**)&t->d_params, sizeof(ThalParams)));
    DG_CUDA(cudaMemcpy(t->d_params, &t->h_params, sizeof(ThalParams), cudaMemcpyHostToDevice));
  } catch (CudaFail& e) {
    dg_thal_close(t);
    return e.code;
  }
  *out = t;
  return DG_OK;
}

}  // namespace

extern "C" {

int dg_thal_open(const char* primer3_config_dir, double mv, double dv, double dntp, double dna_conc, int device, dg_thal** out) {
  if (!primer3_config_dir || !out) { set_error("null argument"); return DG_ERR_ARG; }
  dg_thal* t = new dg_thal();
  std::string err;
  if (!thal_params_from_config(primer3_config_dir, mv, dv, dntp, dna_conc, t->h_params, err)) {
    set_error(err);
    delete t;
    return DG_ERR_IO;
  }
  return finish_open(t, device, out);
}

int dg_thal_open_tables(const char* table_dump_path, int device, dg_thal** out) {
  if (!table_dump_path || !out) { set_error("null argument"); return DG_ERR_ARG; }
  dg_thal* t = new dg_thal();
  std::string err;
  if (!thal_params_from_dump(table_dump_path, t->h_params, err)) {
    set_error(err);
    delete t;
    return DG_ERR_IO;
  }
  return finish_open(t, device, out);
}

void dg_thal_close(dg_thal* t) {
  if (!t) return;
  cudaSetDevice(t->device);
  if (t->st) { cudaStreamSynchronize(t->st); cudaStreamDestroy(t->st); }
  if (t->d_params) cudaFree(t->d_params);
  delete t;
}

int dg_thal_batch(dg_thal* t, const char* seq1, const uint64_t* off1, const char* seq2, const uint64_t* off2, uint32_t n,
                  double* tm, uint8_t* ok) {
  if (!t || !off1 || !off2 || !tm || !ok || (n && (!seq1 || !seq2))) { set_error("null argument"); return DG_ERR_ARG; }
  if (n == 0) return DG_OK;
  try {
    DG_CUDA(cudaSetDevice(t->device));
    cudaStream_t st = t->st;
    const bool trace = getenv("DG_TRACE") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto stage = [&](const char* what) {   // DG_TRACE=1: host wall time per stage (synchronises the stream)
      if (!trace) return;
      cudaStreamSynchronize(st);
      auto now = std::chrono::steady_clock::now();
      fprintf(stderr, "[thal] %-22s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
      t_last = now;
    };
    const uint64_t nb1 = off1[n], nb2 = off2[n];
    uint64_t cells = 1, wide_cells = 1;
    uint32_t wide_max1 = 1, wide_max2 = 1;
    std::vector<uint32_t> wide;
    for (uint32_t q = 0; q < n; ++q) {
      if (off1[q + 1] < off1[q] || off2[q + 1] < off2[q]) { set_error("offsets must be non-decreasing"); return DG_ERR_ARG; }
      const uint64_t a = off1[q + 1] - off1[q], b = off2[q + 1] - off2[q];
      if (a <= (uint64_t)kThalMaxLen && b <= (uint64_t)kThalMaxLen) cells = std::max(cells, a * b);
      // one side longer than the shared-memory forms take, and a pair the reference accepts: k_thal_wide
      if (a && b && (a > (uint64_t)kThalMaxLen) != (b > (uint64_t)kThalMaxLen) && a <= (uint64_t)kThalMaxSeq && b <= (uint64_t)kThalMaxSeq) {
        wide.push_back(q);
        wide_cells = std::max(wide_cells, a * b);
        wide_max1 = std::max(wide_max1, (uint32_t)a);
        wide_max2 = std::max(wide_max2, (uint32_t)b);
      }
    }
    DevBuf<uint8_t> d_s1, d_s2, d_ok;
    DevBuf<uint64_t> d_o1, d_o2;
    DevBuf<double> d_tm, scratch;
    DevBuf<uint32_t> d_ids;
    d_s1.alloc(nb1 + 1); d_s2.alloc(nb2 + 1); d_o1.alloc((size_t)n + 1); d_o2.alloc((size_t)n + 1); d_tm.alloc(n); d_ok.alloc(n);
    DG_CUDA(cudaMemcpyAsync(d_s1.p, seq1, nb1, cudaMemcpyHostToDevice, st));
    DG_CUDA(cudaMemcpyAsync(d_s2.p, seq2, nb2, cudaMemcpyHostToDevice, st));
    DG_CUDA(cudaMemcpyAsync(d_o1.p, off1, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
    DG_CUDA(cudaMemcpyAsync(d_o2.p, off2, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
    stage("alloc + H2D");
    // the sequential kernel over `ids` (null: all pairs), at most ~2 GiB of scratch per launch
    auto sequential = [&](const uint32_t* ids, uint64_t m) {
      uint64_t per = std::max<uint64_t>(1024, std::min<uint64_t>(m, (2ULL << 30) / (16 * cells)));
      per = std::min<uint64_t>(per, 1u << 20);
      scratch.alloc(2 * cells * per);
      for (uint64_t first = 0; first < m; first += per) {
        const uint32_t count = (uint32_t)std::min<uint64_t>(per, m - first);
        k_thal<<<(count + 127) / 128, 128, 0, st>>>(t->d_params, d_s1.p, d_o1.p, d_s2.p, d_o2.p, ids, (uint32_t)first, count, scratch.p,
                                                     cells, d_tm.p, d_ok.p);
      }
      DG_CUDA(cudaGetLastError());
    };
    const char* seq_env = getenv("DG_THAL_SEQ");
    if (seq_env && atoi(seq_env) == 1) {
      sequential(nullptr, n);
    } else {
      // size classes (by table cells), so that long pairs do not take the shared memory -- and with it
      // the resident warps -- of all the others
      static const uint64_t kClass[] = {400, 480, 576, 704, 896, 1280, 2048, 3600};
      constexpr int kClasses = (int)(sizeof(kClass) / sizeof(kClass[0]));
      std::vector<uint32_t> order(n), cls(n);
      uint64_t count[kClasses] = {0}, cap[kClasses] = {0}, start[kClasses + 1] = {0};
      for (uint32_t q = 0; q < n; ++q) {
        const uint64_t a = off1[q + 1] - off1[q], b = off2[q + 1] - off2[q];
        const bool valid = a <= (uint64_t)kThalMaxLen && b <= (uint64_t)kThalMaxLen;
        int c = 0;
        while (valid && c + 1 < kClasses && a * b > kClass[c]) ++c;
        cls[q] = (uint32_t)c;
        ++count[c];
        if (valid) cap[c] = std::max(cap[c], a * b);
      }
      for (int c = 0; c < kClasses; ++c) start[c + 1] = start[c] + count[c];
      {
        uint64_t at[kClasses];
        for (int c = 0; c < kClasses; ++c) at[c] = start[c];
        for (uint32_t q = 0; q < n; ++q) order[at[cls[q]]++] = q;
      }
      d_ids.alloc(n);
      DG_CUDA(cudaMemcpyAsync(d_ids.p, order.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
      stage("size classes");
      DG_CUDA(cudaFuncSetAttribute(k_thal_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      for (int c = 0; c < kClasses; ++c) {
        if (!count[c]) continue;
        const uint64_t cells_c = std::max<uint64_t>(cap[c], 1);
        const size_t per_warp = thal_warp_bytes(cells_c);
        const uint32_t wpb = (uint32_t)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / per_warp));
        int resident = 1;
        DG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_thal_warp, (int)(wpb * 32), wpb * per_warp));
        const uint64_t grid = std::min<uint64_t>((count[c] + wpb - 1) / wpb, (uint64_t)std::max(resident, 1) * (uint64_t)t->sms);
        k_thal_warp<<<(unsigned)grid, wpb * 32, wpb * per_warp, st>>>(
            t->d_params, d_s1.p, d_o1.p, d_s2.p, d_o2.p, d_ids.p + start[c], (uint32_t)count[c], cells_c, d_tm.p, d_ok.p);
      }
      DG_CUDA(cudaGetLastError());
      stage("k_thal_warp");
      const bool force_redo = seq_env && atoi(seq_env) == 2;   // test hook: treat every pair as declined
      std::vector<uint8_t> flags(n);
      DG_CUDA(cudaMemcpyAsync(flags.data(), d_ok.p, n, cudaMemcpyDeviceToHost, st));
      DG_CUDA(cudaStreamSynchronize(st));
      std::vector<uint32_t> redo;
      for (uint32_t q = 0; q < n; ++q)
        if (flags[q] == 2 || force_redo) redo.push_back(q);
      if (!redo.empty()) {
        DG_CUDA(cudaMemcpyAsync(d_ids.p, redo.data(), redo.size() * 4, cudaMemcpyHostToDevice, st));
        sequential(d_ids.p, redo.size());
      }
    }
    DG_CUDA(cudaGetLastError());
    DevBuf<uint32_t> d_wide;
    DevBuf<unsigned char> slab;
    if (!wide.empty()) {
      // last on the stream: the kernels above have written "failed" for these pairs (too long for them)
      const size_t per_warp = thal_wide_bytes(wide_cells, wide_max1, wide_max2);
      constexpr uint32_t wpb = 4;
      int resident = 1;
      DG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_thal_wide, (int)(wpb * 32), 0));
      uint64_t warps = std::min<uint64_t>(wide.size(), (uint64_t)std::max(resident, 1) * (uint64_t)t->sms * wpb);
      warps = std::min<uint64_t>(warps, std::max<uint64_t>(1, (4ULL << 30) / per_warp));   // <= 4 GiB of work areas
      uint64_t blocks = (warps + wpb - 1) / wpb;
      d_wide.alloc(wide.size());
      for (;;) {   // fewer warps side by side when the device is short of memory (the index may hold most of it)
        try {
          slab.alloc((size_t)blocks * wpb * per_warp);
          break;
        } catch (CudaFail&) {
          cudaGetLastError();
          if (blocks == 1) throw;
          blocks = (blocks + 1) / 2;
        }
      }
      DG_CUDA(cudaMemcpyAsync(d_wide.p, wide.data(), wide.size() * 4, cudaMemcpyHostToDevice, st));
      k_thal_wide<<<(unsigned)blocks, wpb * 32, 0, st>>>(t->d_params, d_s1.p, d_o1.p, d_s2.p, d_o2.p, d_wide.p, (uint32_t)wide.size(),
                                                         wide_cells, wide_max1, wide_max2, slab.p, d_tm.p, d_ok.p);
      DG_CUDA(cudaGetLastError());
      stage("k_thal_wide");
    }
    DG_CUDA(cudaMemcpyAsync(tm, d_tm.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaMemcpyAsync(ok, d_ok.p, n, cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaStreamSynchronize(st));
    stage("D2H");
    return DG_OK;
  } catch (CudaFail& e) {
    return e.code;
  } catch (std::bad_alloc&) {
    set_error("out of host memory");
    return DG_ERR_NOMEM;
  }
}

}  // extern "C"
