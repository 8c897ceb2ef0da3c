// dg_api.cu -- index life cycle and diagnostics of the C ABI (include/dicey_b200.h).
#include <algorithm>
#include <cstring>

#include "dg_common.cuh"
#include "fm9.hpp"

namespace dg {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
std::string& last_error_ref() { return g_err; }
}  // namespace dg

using namespace dg;

extern "C" {

const char* dg_last_error(void) { return g_err.c_str(); }
const char* dg_version(void) { return "dicey_b200 0.1 (sm_100a)"; }

int dg_index_open(const char* fm9_path, int device, dg_index** out) {
  if (!fm9_path || !out) { set_error("null argument"); return DG_ERR_ARG; }
  return build_from_fm9(fm9_path, device, out);
}
int dg_index_build_text(const uint8_t* text, uint64_t len, int device, dg_index** out) {
  return build_from_text_host(text, len, device, out);
}
int dg_index_build_synthetic(uint64_t seed, uint32_t nrec, uint64_t reclen, int device, dg_index** out) {
  return build_synthetic(seed, nrec, reclen, device, out);
}
int dg_index_write_fm9(dg_index* idx, const char* fm9_path) {
  if (!idx || !fm9_path) { set_error("null argument"); return DG_ERR_ARG; }
  return write_fm9(idx, fm9_path);
}

int dg_fm9_check(const char* fm9_path) {
  if (!fm9_path) { set_error("null argument"); return DG_ERR_ARG; }
  Fm9 f;
  std::string err;
  const int rc = fm9_parse(fm9_path, f, err);
  if (rc) set_error(err);
  return rc;
}

void dg_index_close(dg_index* idx) {
  if (!idx) return;
  cudaSetDevice(idx->device);
  if (idx->stream) cudaStreamSynchronize(idx->stream);
  if (idx->prof.created) for (auto& e : idx->prof.ev) cudaEventDestroy(e);
  cudaStream_t st = idx->stream, cs = idx->copy_stream;
  cudaStream_t xs[dg_index::kXStreams];
  for (int i = 0; i < dg_index::kXStreams; ++i) xs[i] = idx->xstream[i];
  for (auto s2 : xs) if (s2) cudaStreamSynchronize(s2);
  if (cs) cudaStreamSynchronize(cs);
  cudaStream_t us = idx->up_stream;
  if (us) cudaStreamSynchronize(us);
  {
    release_stream_pools(&st, 1);
    release_stream_pools(xs, dg_index::kXStreams);
  }
  for (auto e : idx->ev_pool) cudaEventDestroy(e);
  delete idx;
  if (st) cudaStreamDestroy(st);
  for (auto s2 : xs) if (s2) cudaStreamDestroy(s2);
  if (cs) cudaStreamDestroy(cs);
  if (us) cudaStreamDestroy(us);
}

uint64_t dg_index_size(const dg_index* idx) { return idx ? idx->n : 0; }

int dg_index_set_records(dg_index* idx, const uint32_t* seqlen_plus1, uint32_t nseq) {
  if (!idx || (nseq && !seqlen_plus1)) { set_error("null argument"); return DG_ERR_ARG; }
  try {
    DG_CUDA(cudaSetDevice(idx->device));
    std::vector<uint64_t> cum((size_t)nseq + 1, 0);
    for (uint32_t i = 0; i < nseq; ++i) cum[i + 1] = cum[i] + seqlen_plus1[i];
    idx->cum.alloc((size_t)nseq + 1);
    DG_CUDA(cudaMemcpy(idx->cum.p, cum.data(), cum.size() * 8, cudaMemcpyHostToDevice));
    idx->nseq = nseq;
    idx->h_cum = cum;
    return DG_OK;
  } catch (CudaFail& e) {
    return e.code;
  }
}

int dg_index_get_info(const dg_index* idx, dg_index_info* info) {
  if (!idx || !info) { set_error("null argument"); return DG_ERR_ARG; }
  info->n = idx->n;
  info->sigma = idx->sigma;
  info->kmer = idx->K;
  info->n_exceptions = idx->n_exc;
  info->device_bytes = idx->device_bytes();
  info->sa_sample = kSaSample;
  info->nseq = idx->nseq;
  info->bitmap_k = idx->KB;
  info->reserved = 0;
  return DG_OK;
}

int dg_index_fetch_text(dg_index* idx, const uint64_t* pos, const uint64_t* len, uint32_t n, char* buf) {
  if (!idx || (n && (!pos || !len || !buf))) { set_error("null argument"); return DG_ERR_ARG; }
  try {
    DG_CUDA(cudaSetDevice(idx->device));
    uint64_t at = 0;
    for (uint32_t i = 0; i < n; ++i) {
      if (pos[i] > idx->n || len[i] > idx->n - pos[i]) { set_error("text range out of bounds"); return DG_ERR_ARG; }
      if (len[i]) DG_CUDA(cudaMemcpyAsync(buf + at, idx->text.p + pos[i], len[i], cudaMemcpyDeviceToHost, idx->stream));
      at += len[i];
    }
    DG_CUDA(cudaStreamSynchronize(idx->stream));
    return DG_OK;
  } catch (CudaFail& e) {
    return e.code;
  }
}

void* dg_index_stream(const dg_index* idx) { return idx ? (void*)idx->stream : nullptr; }

int dg_index_debug_copy(dg_index* idx, const char* what, void* buf, uint64_t* bytes) {
  if (!idx || !what || !bytes) { set_error("null argument"); return DG_ERR_ARG; }
  const void* src = nullptr;
  uint64_t nb = 0;
  std::string w(what);
  if (w == "text") { src = idx->text.p; nb = idx->n; }
  else if (w == "sa_samples") { src = idx->sa_samples.p; nb = idx->sa_samples.count * 4; }
  else if (w == "isa_samples") { src = idx->isa_samples.p; nb = ((idx->n - 1) / 64 + 1) * 4; }
  else if (w == "occ") { src = idx->occ.p; nb = idx->occ.bytes(); }
  else if (w == "kmer") { src = idx->kmer.p; nb = idx->kmer.bytes(); }
  else if (w == "C") { src = idx->Cb.p; nb = 256 * 4; }
  else if (w == "exc_pos") { src = idx->exc_pos.p; nb = (uint64_t)idx->n_exc * 4; }
  else if (w == "exc_sym") { src = idx->exc_sym.p; nb = idx->n_exc; }
  else if (w == "sa_full") { src = idx->sa_full.p; nb = idx->sa_full.bytes(); }
  else if (w == "present_kb") { src = idx->present_kb.p; nb = idx->present_kb.bytes(); }
  else if (w == "present_hi") { src = idx->present_hi.p; nb = idx->present_hi.bytes(); }
  else if (w == "present_lo") { src = idx->present_lo.p; nb = idx->present_lo.bytes(); }
  else if (w == "present_kb_l") { src = idx->present_kb_l.p; nb = idx->present_kb_l.bytes(); }
  else if (w == "present_hi_l") { src = idx->present_hi_l.p; nb = idx->present_hi_l.bytes(); }
  else { set_error("unknown array name"); return DG_ERR_ARG; }
  if (!buf) { *bytes = nb; return DG_OK; }
  if (*bytes < nb) { set_error("buffer too small"); return DG_ERR_ARG; }
  try {
    DG_CUDA(cudaSetDevice(idx->device));
    DG_CUDA(cudaStreamSynchronize(idx->stream));
    if (nb) DG_CUDA(cudaMemcpy(buf, src, nb, cudaMemcpyDeviceToHost));
    *bytes = nb;
    return DG_OK;
  } catch (CudaFail& e) {
    return e.code;
  }
}

void dg_hits_sort(dg_hit* hits, uint64_t n) {
  if (!hits || n < 2) return;
  std::sort(hits, hits + n, [](const dg_hit& a, const dg_hit& b) {  // DnaHit::operator< hunter.h:63-65
    return (a.score > b.score) || ((a.score == b.score) && (a.chr < b.chr)) ||
           ((a.score == b.score) && (a.chr == b.chr) && (a.start < b.start));
  });
}

int dg_profile_enable(dg_index* idx, int on) {
  if (!idx) { set_error("null argument"); return DG_ERR_ARG; }
  idx->prof.enabled = on != 0;
  return DG_OK;
}
int dg_profile_get(dg_index* idx, dg_profile* out) {
  if (!idx || !out) { set_error("null argument"); return DG_ERR_ARG; }
  *out = idx->prof.last;
  out->launches = idx->prof.launches;
  return DG_OK;
}

}  // extern "C"
