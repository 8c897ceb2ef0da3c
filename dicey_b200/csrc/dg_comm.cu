// dg_comm.cu -- the multi-GPU exchange step of the path: the index is replicated
// on every GPU, the per-query loop of hunter.h:291 / silica.h:429 is sharded by rank, and the hit
// records are collected by ONE all-gather.  Two transports sit behind dg_comm:
//   NCCL   ncclAllGather on the index stream, straight from HBM over NVLink / NVSwitch.  libnccl.so.2
//          is opened at run time (a process that already loaded torch's copy shares it; the C++
//          program finds the system one), so the library has no link-time dependency on it.
//   host   a caller-supplied all-gather of host buffers (gloo in the CPU tests, MPI, ...).
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "dg_common.cuh"

namespace dg {
// dg_search.cu: the 16-byte wire records of a batch that has run (written by k_verify), or of the
// last dg_hunt_batch on the index (written by k_rebase)
int batch_wire_view(dg_batch* b, const int4** wire, uint64_t* n, cudaStream_t* st, dg_index** ix);
}

using namespace dg;

namespace {

// ---- the handful of NCCL entry points, bound at run time ------------------------------------------
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId { char internal[DG_COMM_ID_BYTES]; };
constexpr int kNcclInt8 = 0;   // ncclInt8 / ncclChar (nccl.h: ncclDataType_t)

struct Nccl {
  void* so = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  std::string why;
};

Nccl* nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("DG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      n.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (n.so) break;
      n.why = dlerror();
    }
    if (!n.so) return;
    auto sym = [&](const char* s) { return dlsym(n.so, s); };
    n.GetUniqueId = (int (*)(ncclUniqueId*))sym("ncclGetUniqueId");
    n.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))sym("ncclCommInitRank");
    n.CommDestroy = (int (*)(ncclComm_t))sym("ncclCommDestroy");
    n.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))sym("ncclAllGather");
    n.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    n.GetVersion = (int (*)(int*))sym("ncclGetVersion");
    if (!n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.AllGather) {
      n.why = "libnccl lacks a required symbol";
      n.so = nullptr;
    }
  });
  return &n;
}

int nccl_fail(const char* what, int rc) {
  Nccl* n = nccl();
  set_error(std::string(what) + ": " + (n->GetErrorString ? n->GetErrorString(rc) : "NCCL error") + " (" + std::to_string(rc) + ")");
  return DG_ERR_CUDA;
}

constexpr uint64_t kSlotHeader = 1;   // records: the first 16 bytes of a slot hold the count

// local wire records -> this rank's slot: header, then the records with global query ids
__global__ void k_wire_slot(const int4* __restrict__ wire, uint64_t n, uint64_t cap, uint32_t query_base, int4* __restrict__ slot) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) slot[0] = make_int4((int)(uint32_t)n, (int)(uint32_t)(n >> 32), 0x64676831 /* "dgh1" */, 0);
  if (i < n && i < cap) {
    int4 w = wire[i];
    w.x = (int)((uint32_t)w.x + query_base);
    slot[kSlotHeader + i] = w;
  }
}
__global__ void k_read_counts(const int4* __restrict__ table, uint64_t slot_records, int nranks, uint64_t* __restrict__ counts) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < nranks) {
    const int4 h = table[(uint64_t)r * slot_records];
    counts[r] = ((uint64_t)(uint32_t)h.y << 32) | (uint32_t)h.x;
  }
}
__global__ void k_compact_table(const int4* __restrict__ table, uint64_t slot_records, const uint64_t* __restrict__ off, int nranks,
                                int4* __restrict__ out) {
  const int r = blockIdx.y;
  const uint64_t n = off[r + 1] - off[r];
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    out[off[r] + i] = table[(uint64_t)r * slot_records + kSlotHeader + i];
}

// peer mode: this rank's hit count into the header of its slot on every rank
struct PeerHeads { int4* head[kMaxPeers]; uint32_t nranks; };
__global__ void k_peer_header(PeerHeads h, uint64_t n, uint64_t* __restrict__ token) {
  const uint32_t r = threadIdx.x;
  if (r < h.nranks) h.head[r][0] = make_int4((int)(uint32_t)n, (int)(uint32_t)(n >> 32), 0x64676831, 0);
  if (r == 0) token[0] = n;
  __threadfence_system();
}

}  // namespace

struct dg_comm {
  int nranks = 1, rank = 0;
  // NCCL transport
  ncclComm_t nc = nullptr;
  dg_index* idx = nullptr;
  DevBuf<int4> table;            // nranks slots
  DevBuf<int4> compact;          // dg_comm_fetch_table staging
  DevBuf<uint64_t> d_counts;     // nranks + (nranks + 1) offsets
  uint64_t slot_records = 0;     // records per slot, header included; identical on every rank
  uint64_t last_slot = 0;        // slot size the table currently holds (slot_records may already have grown)
  const int4* last_table = nullptr;
  uint64_t* h_counts = nullptr;  // pinned, nranks entries
  DevBuf<uint8_t> stage_send, stage_recv;   // dg_allgather_result over NCCL
  // peer mode: every rank's table (two of them, used alternately) is mapped on every rank
  bool p2p = false;
  uint64_t p2p_slot = 0;                    // records per slot, header included (fixed at init)
  DevBuf<int4> ptable[2];                   // this rank's tables
  int4* peer_tab[2][kMaxPeers] = {};        // rank r's tables as seen from this rank
  std::vector<void*> ipc_opened;
  uint64_t epoch = 0;                       // exchanges done; the next one uses table epoch & 1
  const void* writer = nullptr;             // who wrote peer records for the coming exchange, with which base
  uint64_t writer_epoch = ~0ull;
  uint32_t query_base = 0;
  DevBuf<uint64_t> d_token;                 // send / recv of the barrier all-gather
  // host transport
  dg_host_allgather_fn fn = nullptr;
  void* fn_ctx = nullptr;

  int allgather_host_bytes(const void* send, void* recv, uint64_t bytes) {
    if (fn) {
      int rc = fn(fn_ctx, send, recv, bytes);
      if (rc) { set_error("host all-gather callback failed (" + std::to_string(rc) + ")"); return DG_ERR_IO; }
      return DG_OK;
    }
    try {
      DG_CUDA(cudaSetDevice(idx->device));
      if (stage_send.count < bytes) stage_send.alloc(bytes + (bytes >> 2) + 4096);
      if (stage_recv.count < bytes * (uint64_t)nranks) stage_recv.alloc((bytes + (bytes >> 2) + 4096) * (uint64_t)nranks);
      cudaStream_t st = idx->stream;
      DG_CUDA(cudaMemcpyAsync(stage_send.p, send, bytes, cudaMemcpyHostToDevice, st));
      int rc = nccl()->AllGather(stage_send.p, stage_recv.p, bytes, kNcclInt8, nc, st);
      if (rc) return nccl_fail("ncclAllGather", rc);
      DG_CUDA(cudaMemcpyAsync(recv, stage_recv.p, bytes * (uint64_t)nranks, cudaMemcpyDeviceToHost, st));
      DG_CUDA(cudaStreamSynchronize(st));
      return DG_OK;
    } catch (CudaFail& e) {
      return e.code;
    }
  }
};

bool dg::comm_peer_out(dg_index* idx, const void* producer, PeerOut* out) {
  dg_comm* c = idx ? idx->bound_comm : nullptr;
  out->nranks = 0;
  if (!c || !c->p2p) return false;
  const int t = (int)(c->epoch & 1);
  for (int r = 0; r < c->nranks; ++r) out->tab[r] = c->peer_tab[t][r] + c->p2p_slot * (uint64_t)c->rank + kSlotHeader;
  out->nranks = (uint32_t)c->nranks;
  out->query_base = c->query_base;
  out->cap = c->p2p_slot - kSlotHeader;
  c->writer = producer;
  c->writer_epoch = c->epoch;
  return true;
}

extern "C" {

int dg_comm_set_query_base(dg_comm* c, uint64_t query_base) {
  if (!c) { set_error("null argument"); return DG_ERR_ARG; }
  c->query_base = (uint32_t)query_base;
  return DG_OK;
}

int dg_comm_get_unique_id(void* id) {
  if (!id) { set_error("null argument"); return DG_ERR_ARG; }
  Nccl* n = nccl();
  if (!n->so) { set_error("NCCL is not available: " + n->why); return DG_ERR_UNSUPPORTED; }
  ncclUniqueId u;
  int rc = n->GetUniqueId(&u);
  if (rc) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(id, u.internal, DG_COMM_ID_BYTES);
  return DG_OK;
}

int dg_comm_init(int nranks, int rank, const void* id, dg_index* idx, dg_comm** out) {
  if (!id || !idx || !out || nranks < 1 || rank < 0 || rank >= nranks) { set_error("bad argument"); return DG_ERR_ARG; }
  Nccl* n = nccl();
  if (!n->so) { set_error("NCCL is not available: " + n->why); return DG_ERR_UNSUPPORTED; }
  dg_comm* c = new dg_comm();
  try {
    DG_CUDA(cudaSetDevice(idx->device));
    c->nranks = nranks; c->rank = rank; c->idx = idx;
    ncclUniqueId u;
    memcpy(u.internal, id, DG_COMM_ID_BYTES);
    int rc = n->CommInitRank(&c->nc, nranks, u, rank);
    if (rc) { delete c; return nccl_fail("ncclCommInitRank", rc); }
    DG_CUDA(cudaHostAlloc((void**)&c->h_counts, sizeof(uint64_t) * (size_t)nranks, cudaHostAllocDefault));
    memset(c->h_counts, 0, sizeof(uint64_t) * (size_t)nranks);
    c->d_counts.alloc(2 * (size_t)nranks + 1);
    c->slot_records = 1ull << 20;   // 16 MB per rank to start with; grows in step on every rank
    if (const char* e = getenv("DG_COMM_SLOT")) c->slot_records = std::max<uint64_t>(2, strtoull(e, nullptr, 10));
    c->d_token.alloc(1 + (size_t)nranks);
    idx->bound_comm = c;
    // ---- peer mode: map every rank's tables on every rank (cudaIpc between processes, peer access
    // between the threads of one process).  Any rank that cannot makes all of them stay with NCCL.
    const char* pe = getenv("DG_COMM_P2P");
    bool want = nranks <= kMaxPeers && !(pe && atoi(pe) == 0);
    struct Card { cudaIpcMemHandle_t h[2]; uint64_t ptr[2]; int64_t pid; int32_t dev; int32_t ok; };
    Card mine;
    memset(&mine, 0, sizeof(mine));
    c->p2p_slot = 3ull << 20;       // 3 M record places per rank and table (48 MB); DG_COMM_P2P_SLOT
    if (const char* e = getenv("DG_COMM_P2P_SLOT")) c->p2p_slot = std::max<uint64_t>(2, strtoull(e, nullptr, 10));
    if (want) {
      for (int t = 0; t < 2 && want; ++t) {
        if (cudaMalloc((void**)&c->ptable[t].p, c->p2p_slot * (uint64_t)nranks * sizeof(int4)) != cudaSuccess) { cudaGetLastError(); want = false; break; }
        c->ptable[t].count = c->p2p_slot * (uint64_t)nranks;
        cudaMemset(c->ptable[t].p, 0, c->ptable[t].bytes());
        if (cudaIpcGetMemHandle(&mine.h[t], c->ptable[t].p) != cudaSuccess) { cudaGetLastError(); want = false; }
        mine.ptr[t] = (uint64_t)(uintptr_t)c->ptable[t].p;
      }
    }
    mine.pid = (int64_t)getpid();
    mine.dev = idx->device;
    mine.ok = want ? 1 : 0;
    std::vector<Card> cards((size_t)nranks);
    int rc2 = c->allgather_host_bytes(&mine, cards.data(), sizeof(Card));
    if (rc2) { idx->bound_comm = nullptr; delete c; return rc2; }
    bool all_ok = true;
    for (auto& k : cards) all_ok = all_ok && k.ok;
    int32_t mapped = all_ok ? 1 : 0;
    if (all_ok) {
      for (int r = 0; r < nranks && mapped; ++r) {
        for (int t = 0; t < 2 && mapped; ++t) {
          if (r == rank) { c->peer_tab[t][r] = c->ptable[t].p; continue; }
          if (cards[r].pid == mine.pid) {   // another thread of this process: plain peer access
            if (t == 0 && cards[r].dev != idx->device) {
              cudaError_t e = cudaDeviceEnablePeerAccess(cards[r].dev, 0);
              if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) mapped = 0;
              cudaGetLastError();
            }
            c->peer_tab[t][r] = (int4*)(uintptr_t)cards[r].ptr[t];
          } else {
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, cards[r].h[t], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mapped = 0; break; }
            c->ipc_opened.push_back(p);
            c->peer_tab[t][r] = (int4*)p;
          }
        }
      }
    }
    std::vector<int32_t> flags((size_t)nranks, 0);
    rc2 = c->allgather_host_bytes(&mapped, flags.data(), sizeof(int32_t));
    if (rc2) { idx->bound_comm = nullptr; delete c; return rc2; }
    c->p2p = true;
    for (int32_t f : flags) c->p2p = c->p2p && f;
    if (!c->p2p) { c->ptable[0].release(); c->ptable[1].release(); }
    if (getenv("DG_TRACE")) fprintf(stderr, "[dg_comm] rank %d of %d: %s\n", rank, nranks, c->p2p ? "peer mode (records stored into every rank's table by the producing kernel)" : "NCCL all-gather of the records");
    *out = c;
    return DG_OK;
  } catch (CudaFail& e) {
    delete c;
    return e.code;
  }
}

int dg_comm_init_host(int nranks, int rank, dg_host_allgather_fn fn, void* ctx, dg_comm** out) {
  if (!fn || !out || nranks < 1 || rank < 0 || rank >= nranks) { set_error("bad argument"); return DG_ERR_ARG; }
  dg_comm* c = new dg_comm();
  c->nranks = nranks; c->rank = rank; c->fn = fn; c->fn_ctx = ctx;
  *out = c;
  return DG_OK;
}

int dg_comm_rank(const dg_comm* c) { return c ? c->rank : -1; }
int dg_comm_size(const dg_comm* c) { return c ? c->nranks : 0; }

void dg_comm_destroy(dg_comm* c) {
  if (!c) return;
  if (c->idx) {
    cudaSetDevice(c->idx->device);
    cudaStreamSynchronize(c->idx->stream);
  }
  if (c->idx && c->idx->bound_comm == c) c->idx->bound_comm = nullptr;
  for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
  if (c->nc) nccl()->CommDestroy(c->nc);
  if (c->h_counts) cudaFreeHost(c->h_counts);
  delete c;
}

int dg_allgather_hits(dg_comm* c, dg_batch* b, uint64_t query_base, const dg_wire** table, uint64_t* slot_records,
                      const uint64_t** counts) {
  if (!c || !table || !slot_records || !counts) { set_error("null argument"); return DG_ERR_ARG; }
  if (!c->nc) { set_error("dg_allgather_hits needs the NCCL transport (dg_comm_init)"); return DG_ERR_UNSUPPORTED; }
  try {
    dg_index* idx = c->idx;
    DG_CUDA(cudaSetDevice(idx->device));
    const int4* wire = nullptr;
    uint64_t n = 0;
    cudaStream_t st = idx->stream;
    if (b) {
      dg_index* bi = nullptr;
      int rc = batch_wire_view(b, &wire, &n, &st, &bi);
      if (rc) return rc;
      if (bi != idx) { set_error("the batch belongs to another index than the communicator"); return DG_ERR_ARG; }
    } else {
      wire = idx->wire.p;
      n = idx->wire_n;
    }
    // peer mode: the producing kernel already stored this rank's records into every rank's table; what is
    // left is the count in the slot headers and one small all-gather that doubles as the barrier (NCCL
    // orders it after every rank's producing kernels: they precede it on each rank's stream)
    const void* producer = b ? (const void*)b : (const void*)idx;
    if (c->p2p && c->writer == producer && c->writer_epoch == c->epoch && c->query_base == (uint32_t)query_base) {
      const int t = (int)(c->epoch & 1);
      PeerHeads ph;
      ph.nranks = (uint32_t)c->nranks;
      for (int r = 0; r < c->nranks; ++r) ph.head[r] = c->peer_tab[t][r] + c->p2p_slot * (uint64_t)c->rank;
      k_peer_header<<<1, 32, 0, st>>>(ph, n, c->d_token.p);
      int rc = nccl()->AllGather(c->d_token.p, c->d_token.p + 1, sizeof(uint64_t), kNcclInt8, c->nc, st);
      if (rc) return nccl_fail("ncclAllGather", rc);
      DG_CUDA(cudaMemcpyAsync(c->h_counts, c->d_token.p + 1, sizeof(uint64_t) * (size_t)c->nranks, cudaMemcpyDeviceToHost, st));
      DG_CUDA(cudaStreamSynchronize(st));
      DG_CUDA(cudaGetLastError());
      ++c->epoch;
      uint64_t mx = 0;
      for (int r = 0; r < c->nranks; ++r) mx = std::max(mx, c->h_counts[r]);
      if (mx <= c->p2p_slot - kSlotHeader) {
        c->last_slot = c->p2p_slot;
        c->last_table = c->ptable[t].p;
        *table = reinterpret_cast<const dg_wire*>(c->ptable[t].p);
        *slot_records = c->p2p_slot;
        *counts = c->h_counts;
        return DG_OK;
      }
      // a rank holds more hits than a slot: every rank sees that and repeats the exchange through NCCL
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
      const uint64_t slot = c->slot_records;
      if (c->table.count < slot * (uint64_t)c->nranks) c->table.alloc(slot * (uint64_t)c->nranks);
      int4* mine = c->table.p + slot * (uint64_t)c->rank;   // in place: this rank's slot of the receive buffer
      const uint64_t cap = slot - kSlotHeader;
      const uint64_t work = std::max<uint64_t>(1, std::min(n, cap));
      k_wire_slot<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(wire, n, cap, (uint32_t)query_base, mine);
      int rc = nccl()->AllGather(mine, c->table.p, slot * sizeof(int4), kNcclInt8, c->nc, st);
      if (rc) return nccl_fail("ncclAllGather", rc);
      k_read_counts<<<1, 256, 0, st>>>(c->table.p, slot, c->nranks, c->d_counts.p);
      DG_CUDA(cudaMemcpyAsync(c->h_counts, c->d_counts.p, sizeof(uint64_t) * (size_t)c->nranks, cudaMemcpyDeviceToHost, st));
      DG_CUDA(cudaStreamSynchronize(st));
      DG_CUDA(cudaGetLastError());
      // every rank sees the same counts, so every rank takes the same decision
      uint64_t mx = 0;
      for (int r = 0; r < c->nranks; ++r) mx = std::max(mx, c->h_counts[r]);
      const bool fits = mx <= cap;
      if (mx + (mx >> 3) > cap) {   // grow ahead of the next call as well
        uint64_t want = mx + (mx >> 2) + kSlotHeader + 1024;
        c->slot_records = std::max(c->slot_records, want);
      }
      if (fits) {
        c->last_slot = slot;
        c->last_table = c->table.p;
        *table = reinterpret_cast<const dg_wire*>(c->table.p);
        *slot_records = slot;
        *counts = c->h_counts;
        return DG_OK;
      }
    }
    set_error("hit all-gather: slot still too small after growing");
    return DG_ERR_OVERFLOW;
  } catch (CudaFail& e) {
    return e.code;
  }
}

int dg_comm_fetch_table(dg_comm* c, dg_wire* out, uint64_t capacity, uint64_t* n) {
  if (!c || !n || !c->nc || !c->h_counts) { set_error("bad argument"); return DG_ERR_ARG; }
  try {
    DG_CUDA(cudaSetDevice(c->idx->device));
    std::vector<uint64_t> off((size_t)c->nranks + 1, 0);
    uint64_t mx = 0;
    for (int r = 0; r < c->nranks; ++r) { off[r + 1] = off[r] + c->h_counts[r]; mx = std::max(mx, c->h_counts[r]); }
    *n = off[c->nranks];
    if (!out) return DG_OK;
    if (capacity < *n) { set_error("buffer too small"); return DG_ERR_ARG; }
    if (!*n) return DG_OK;
    cudaStream_t st = c->idx->stream;
    if (c->compact.count < *n) c->compact.alloc(*n + (*n >> 3) + 1024);
    uint64_t* d_off = c->d_counts.p + c->nranks;
    DG_CUDA(cudaMemcpyAsync(d_off, off.data(), sizeof(uint64_t) * off.size(), cudaMemcpyHostToDevice, st));
    dim3 grid((unsigned)std::min<uint64_t>(1024, (mx + 255) / 256 + 1), (unsigned)c->nranks);
    k_compact_table<<<grid, 256, 0, st>>>(c->last_table, c->last_slot, d_off, c->nranks, c->compact.p);
    DG_CUDA(cudaMemcpyAsync(out, c->compact.p, *n * sizeof(int4), cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaStreamSynchronize(st));
    return DG_OK;
  } catch (CudaFail& e) {
    return e.code;
  }
}

// ---- complete results ------------------------------------------------------------------------------
// Packed form (dg_result_pack): u64 nq, nhits, pool bytes, seq bytes; qoff[nq+1] u64; status[nq] u32;
// dist[nq] u32; hits[nhits] dg_hit; pool; seqs.  The merged result is assembled in the same form and
// handed to dg_result_unpack.
int dg_allgather_result(dg_comm* c, const dg_result* local, dg_result** global) {
  if (!c || !local || !global) { set_error("null argument"); return DG_ERR_ARG; }
  uint64_t mine = 0;
  int rc = dg_result_pack(local, nullptr, &mine);
  if (rc) return rc;
  const int W = c->nranks;
  // sizes first, then the buffers padded to the largest
  std::vector<uint64_t> sizes((size_t)W, 0);
  rc = c->allgather_host_bytes(&mine, sizes.data(), sizeof(uint64_t));
  if (rc) return rc;
  uint64_t mx = 0;
  for (uint64_t s : sizes) mx = std::max(mx, s);
  mx = (mx + 15) & ~15ull;
  std::vector<uint8_t> send((size_t)mx, 0), recv((size_t)mx * (size_t)W);
  uint64_t nb = mx;
  rc = dg_result_pack(local, send.data(), &nb);
  if (rc) return rc;
  rc = c->allgather_host_bytes(send.data(), recv.data(), mx);
  if (rc) return rc;
  struct Part { uint64_t nq, nh, np, ns; const uint8_t *qoff, *status, *dist, *hits, *pool, *seqs; };
  std::vector<Part> parts((size_t)W);
  uint64_t NQ = 0, NH = 0, NP = 0, NS = 0;
  for (int r = 0; r < W; ++r) {
    const uint8_t* p = recv.data() + (size_t)mx * (size_t)r;
    if (sizes[r] < 32) { set_error("truncated result from rank " + std::to_string(r)); return DG_ERR_FORMAT; }
    uint64_t hdr[4];
    memcpy(hdr, p, 32);
    Part& t = parts[r];
    t.nq = hdr[0]; t.nh = hdr[1]; t.np = hdr[2]; t.ns = hdr[3];
    const uint64_t need = 32 + (t.nq + 1) * 8 + t.nq * 8 + t.nh * sizeof(dg_hit) + t.np + t.ns;
    if (need > sizes[r]) { set_error("inconsistent result from rank " + std::to_string(r)); return DG_ERR_FORMAT; }
    t.qoff = p + 32;
    t.status = t.qoff + (t.nq + 1) * 8;
    t.dist = t.status + t.nq * 4;
    t.hits = t.dist + t.nq * 4;
    t.pool = t.hits + t.nh * sizeof(dg_hit);
    t.seqs = t.pool + t.np;
    NQ += t.nq; NH += t.nh; NP += t.np; NS += t.ns;
  }
  if (NQ > 0xFFFFFFFFull) { set_error("more than 2^32 queries in the gathered batch"); return DG_ERR_OVERFLOW; }
  std::vector<uint8_t> all((size_t)(32 + (NQ + 1) * 8 + NQ * 8 + NH * sizeof(dg_hit) + NP + NS));
  uint64_t hdr[4] = {NQ, NH, NP, NS};
  memcpy(all.data(), hdr, 32);
  uint64_t* qoff = reinterpret_cast<uint64_t*>(all.data() + 32);
  uint8_t* status = all.data() + 32 + (NQ + 1) * 8;
  uint8_t* dist = status + NQ * 4;
  uint8_t* hits = dist + NQ * 4;
  uint8_t* pool = hits + NH * sizeof(dg_hit);
  uint8_t* seqs = pool + NP;
  uint64_t qb = 0, hb = 0, pb = 0, sb = 0;
  qoff[0] = 0;
  for (int r = 0; r < W; ++r) {
    const Part& t = parts[r];
    for (uint64_t q = 0; q < t.nq; ++q) {
      uint64_t v;
      memcpy(&v, t.qoff + (q + 1) * 8, 8);
      qoff[qb + q + 1] = v + hb;
    }
    if (t.nq) { memcpy(status + qb * 4, t.status, t.nq * 4); memcpy(dist + qb * 4, t.dist, t.nq * 4); }
    for (uint64_t i = 0; i < t.nh; ++i) {
      dg_hit h;
      memcpy(&h, t.hits + i * sizeof(dg_hit), sizeof(dg_hit));
      h.query += (uint32_t)qb;
      h.aln_off += pb;
      memcpy(hits + (hb + i) * sizeof(dg_hit), &h, sizeof(dg_hit));
    }
    if (t.np) memcpy(pool + pb, t.pool, t.np);
    if (t.ns) memcpy(seqs + sb, t.seqs, t.ns);
    qb += t.nq; hb += t.nh; pb += t.np; sb += t.ns;
  }
  return dg_result_unpack(all.data(), all.size(), global);
}

}  // extern "C"
