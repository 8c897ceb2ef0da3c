// dg_common.cuh -- host-side state of one device-resident index and small CUDA helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/dicey_b200.h"
#include "dg_core.cuh"

namespace dg {

void set_error(const std::string& msg);
std::string& last_error_ref();

struct CudaFail { int code; };

#define DG_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::dg::set_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                      std::to_string(__LINE__) + " (" #expr ")");                              \
      throw ::dg::CudaFail{DG_ERR_CUDA};                                                       \
    }                                                                                          \
  } while (0)

// Device buffer owned by an Index / Batch (plain cudaMalloc: these live as long as the owner).
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t count = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void alloc(size_t n) {
    release();
    count = n;
    if (n) DG_CUDA(cudaMalloc((void**)&p, n * sizeof(T)));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    count = 0;
  }
  size_t bytes() const { return count * sizeof(T); }
};

// Where the 16-byte wire records of this rank go on EVERY GPU of a communicator whose tables are
// mapped into each other's address space (dg_comm.cu, peer mode): k_verify / k_rebase_recs store each
// record straight into the tables of all ranks over NVLink, so the exchange step that follows is a
// barrier, not a transfer.
constexpr int kMaxPeers = 16;
struct PeerOut {
  int4* tab[kMaxPeers];    // this rank's record area (header excluded) in the table of rank r
  uint32_t nranks;         // 0 = peer mode off
  uint32_t query_base;     // added to the query index of every record
  uint64_t cap;            // record places per rank
};

// Unit tables of one (distance, mode, set of query lengths): built once per index and shared by
// every batch of that shape (script_ub[256] | tab_off[512] | tab_cnt[512] | tab[...] in one block).
struct TabEntry {
  DevBuf<uint32_t> blob;
  std::vector<uint32_t> tab_cnt;   // host copy of tab_cnt
  std::vector<uint32_t> script_ub; // host copy of script_ub
};

struct ProfileState {
  bool enabled = false;
  cudaEvent_t ev[8] = {};
  bool created = false;
  dg_profile last = {};
  uint64_t launches = 0;
  bool probe_timed = false;
};

}  // namespace dg

// The opaque handle of include/dicey_b200.h.
struct dg_index {
  int device = 0;
  cudaStream_t stream = nullptr;
  static constexpr int kXStreams = 6;
  cudaStream_t xstream[kXStreams] = {};                     // compute streams of the chunk pipeline (dg_hunt_batch), descending priority
  cudaStream_t up_stream = nullptr;                         // its upload stream and the device copy of the caller's sequences
  dg::DevBuf<uint8_t> upload;
  std::vector<cudaEvent_t> ev_pool;                         // events of the chunk pipeline, created once and reused (creating or
                                                            // destroying one while copies are in flight can stall the caller)
  cudaStream_t copy_stream = nullptr;   // device -> host copies of finished chunks
  uint64_t n = 0;
  uint32_t sigma = 0;
  uint32_t K = 0;
  dg::DevBuf<dg::OccBlock> occ;
  dg::DevBuf<uint32_t> excflag, exc_pos, rare_pos, rare_off, Cb, sa_samples, isa_samples, present_kb, present_hi, present_lo, present_kb_l, present_hi_l, sa_full;
  uint32_t KB = 0;
  dg::DevBuf<uint8_t> exc_sym, present, text;
  dg::DevBuf<uint2> kmer;
  dg::DevBuf<int4> wire;        // 16-byte wire records of the last dg_hunt_batch (dg_index_wire_records)
  uint64_t wire_n = 0;
  dg::DevBuf<uint64_t> cum;
  std::vector<uint64_t> h_cum;     // host copy (record starts; a compact result rebuilds text_pos from it)
  uint32_t n_exc = 0;
  uint32_t nseq = 0;
  uint32_t C4[4] = {0, 0, 0, 0};
  std::vector<uint32_t> h_Cb;      // host copies for the .fm9 writer / info
  std::vector<uint8_t> h_present;
  dg::ProfileState prof;
  struct dg_comm* bound_comm = nullptr;   // the communicator bound to this index (dg_comm_init), if any
  std::mutex tab_mu;
  std::map<std::string, std::shared_ptr<dg::TabEntry>> tab_cache;

  dg::IndexView view() const {
    dg::IndexView v;
    v.occ = occ.p; v.n = n; v.excflag = excflag.p; v.exc_pos = exc_pos.p; v.exc_sym = exc_sym.p;
    v.n_exc = n_exc; v.rare_pos = rare_pos.p; v.rare_off = rare_off.p; v.Cb = Cb.p; v.present = present.p;
    for (int i = 0; i < 4; ++i) v.C4[i] = C4[i];
    v.kmer = kmer.p; v.K = K; v.sa_samples = sa_samples.p; v.text = text.p; v.cum = cum.p; v.nseq = nseq;
    v.present_kb = present_kb.p; v.KB = KB; v.sa_full = sa_full.p;
    v.present_hi = present_hi.p; v.present_lo = present_lo.p;
    v.present_kb_l = present_kb_l.p; v.present_hi_l = present_hi_l.p;
    return v;
  }
  uint64_t device_bytes() const {
    return occ.bytes() + excflag.bytes() + exc_pos.bytes() + rare_pos.bytes() + rare_off.bytes() + Cb.bytes() +
           sa_samples.bytes() + isa_samples.bytes() + exc_sym.bytes() + present.bytes() + text.bytes() +
           kmer.bytes() + cum.bytes() + present_kb.bytes() + present_hi.bytes() + present_lo.bytes() + present_kb_l.bytes() + present_hi_l.bytes() + sa_full.bytes();
  }
};

namespace dg {
// dg_build.cu
int build_from_fm9(const char* path, int device, dg_index** out);
int build_from_text_host(const uint8_t* text, uint64_t len, int device, dg_index** out);
int build_synthetic(uint64_t seed, uint32_t nrec, uint64_t reclen, int device, dg_index** out);
int write_fm9(dg_index* idx, const char* path);
// dg_search.cu
void release_stream_pools(const cudaStream_t* streams, int n);
// dg_comm.cu: peer-mode destination of the records `producer` (a batch, or the index for dg_hunt_batch)
// is about to write; false = peer mode off (the records are all-gathered by NCCL afterwards)
bool comm_peer_out(dg_index* idx, const void* producer, PeerOut* out);
}  // namespace dg
