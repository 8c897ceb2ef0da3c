// dg_thal.cu -- batched melting temperatures on the GPU: the thal() gate of `dicey search`
// (reference src/silica.h:508-519; the arithmetic is dg_thal.cuh) for every candidate site at once.
//
// k_thal runs one pair per thread.  The two DP tables of a thread (|primer| x |site| doubles each)
// live in a scratch slab in which the tables of the threads of a launch are interleaved cell by
// cell, so the lanes of a warp -- which walk the same (i, j) order -- read and write consecutive
// addresses.  Built with -fmad=false: results equal the reference's bit for bit.
#include <algorithm>
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
};

namespace {

__global__ void __launch_bounds__(128) k_thal(const ThalParams* __restrict__ p, const uint8_t* __restrict__ s1,
                                              const uint64_t* __restrict__ off1, const uint8_t* __restrict__ s2,
                                              const uint64_t* __restrict__ off2, uint32_t first, uint32_t count, double* __restrict__ scratch,
                                              uint64_t cells, double* __restrict__ tm, uint8_t* __restrict__ ok) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const uint32_t q = first + t;
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
    const uint64_t nb1 = off1[n], nb2 = off2[n];
    uint64_t cells = 1;
    for (uint32_t q = 0; q < n; ++q) {
      if (off1[q + 1] < off1[q] || off2[q + 1] < off2[q]) { set_error("offsets must be non-decreasing"); return DG_ERR_ARG; }
      const uint64_t a = off1[q + 1] - off1[q], b = off2[q + 1] - off2[q];
      if (a <= (uint64_t)kThalMaxLen && b <= (uint64_t)kThalMaxLen) cells = std::max(cells, a * b);
    }
    DevBuf<uint8_t> d_s1, d_s2, d_ok;
    DevBuf<uint64_t> d_o1, d_o2;
    DevBuf<double> d_tm, scratch;
    d_s1.alloc(nb1 + 1); d_s2.alloc(nb2 + 1); d_o1.alloc((size_t)n + 1); d_o2.alloc((size_t)n + 1); d_tm.alloc(n); d_ok.alloc(n);
    DG_CUDA(cudaMemcpyAsync(d_s1.p, seq1, nb1, cudaMemcpyHostToDevice, st));
    DG_CUDA(cudaMemcpyAsync(d_s2.p, seq2, nb2, cudaMemcpyHostToDevice, st));
    DG_CUDA(cudaMemcpyAsync(d_o1.p, off1, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
    DG_CUDA(cudaMemcpyAsync(d_o2.p, off2, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
    // pairs per launch: two tables of `cells` doubles each, at most ~2 GiB of scratch
    uint64_t per = std::max<uint64_t>(1024, std::min<uint64_t>(n, (2ULL << 30) / (16 * cells)));
    per = std::min<uint64_t>(per, 1u << 20);
    scratch.alloc(2 * cells * per);
    for (uint64_t first = 0; first < n; first += per) {
      const uint32_t count = (uint32_t)std::min<uint64_t>(per, n - first);
      k_thal<<<(count + 127) / 128, 128, 0, st>>>(t->d_params, d_s1.p, d_o1.p, d_s2.p, d_o2.p, (uint32_t)first, count, scratch.p, cells,
                                                   d_tm.p, d_ok.p);
    }
    DG_CUDA(cudaGetLastError());
    DG_CUDA(cudaMemcpyAsync(tm, d_tm.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaMemcpyAsync(ok, d_ok.p, n, cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaStreamSynchronize(st));
    return DG_OK;
  } catch (CudaFail& e) {
    return e.code;
  } catch (std::bad_alloc&) {
    set_error("out of host memory");
    return DG_ERR_NOMEM;
  }
}

}  // extern "C"
