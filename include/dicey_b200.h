/* dicey_b200.h -- C ABI of the B200-native FM-index primer-matching engine.
 *
 * The reference (gear-genomics/dicey) has no plugin / FFI seam; the boundary is the set of
 * C++ calls its three driver loops make into SDSL and its own neighbors.h / needle.h
 * (SURVEY.md section 8b).  Each entry point below names the reference call sites it replaces.
 * Plain pointers and sizes only; no C++ or torch types; no exceptions cross this boundary.
 *
 * Ownership: the caller owns every input buffer (borrowed for the duration of the call);
 * the library owns dg_index / dg_batch / dg_result objects until the matching *_close/_free.
 * Threading: one dg_index is bound to one CUDA device and one internal stream; calls on one
 * index must be serialised by the caller; different indexes (one per GPU) are independent.
 * Every function returns DG_OK (0) or a negative dg_status; dg_last_error() gives the text.
 */
#ifndef DICEY_B200_H
#define DICEY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dg_index dg_index;
typedef struct dg_batch dg_batch;
typedef struct dg_result dg_result;

typedef enum {
  DG_OK = 0,
  DG_ERR_ARG = -1,         /* bad argument */
  DG_ERR_IO = -2,          /* file cannot be read / written */
  DG_ERR_FORMAT = -3,      /* not a csa_wt<> .fm9, or .fm9_check mismatch (SDSL io.hpp:917-936) */
  DG_ERR_CUDA = -4,        /* no usable device, or a CUDA call failed (never falls back to CPU) */
  DG_ERR_UNSUPPORTED = -5, /* outside the device path's limits (see DESIGN.md "Limits") */
  DG_ERR_OVERFLOW = -6,    /* a bounded device buffer would overflow (too many occurrences) */
  DG_ERR_NOMEM = -7
} dg_status;

/* Per-query status bits (dg_result_query_status). */
enum {
  DG_Q_TOO_SHORT = 1u << 0,       /* hunter.h:299-303: |seq| < 10 -> error, no search */
  DG_Q_DIST_ADJUSTED = 1u << 1,   /* hunter.h:312-315: distance clamped to |seq|-1 */
  DG_Q_HIT_CAP = 1u << 2,         /* hunter.h:434-437: hits >= max_locations */
  DG_Q_NBR_CAP = 1u << 3,         /* hunter.h:342-345: neighbourhood size >= max_neighborhood */
  DG_Q_NBR_UNVERIFIED = 1u << 4,  /* the neighbourhood may reach max_neighborhood but its truncated form could
                                     not be replayed (strings longer than 42 characters): searched untruncated */
  DG_Q_SKIPPED = 1u << 5,         /* silica.h:363,388: primer not longer than the seed k-mer */
  DG_Q_UNSUPPORTED = 1u << 6      /* outside the device path's limits (length > 255; distance > 2 with length + distance > 42): not searched */
};

/* hunter.h:37-50 (HunterConfig) / silica.h:38-67 (SilicaConfig), the fields the hot path reads. */
typedef struct {
  uint32_t distance;          /* -d; clamped per query to |seq|-1                      */
  uint32_t max_neighborhood;  /* -x, default 10000                                     */
  uint32_t max_locations;     /* -m, default 1000 (hunt) / 10000 (search)              */
  uint8_t indel;              /* 1 = edit distance, 0 = Hamming (-n)                   */
  uint8_t reverse;            /* 1 = also search the reverse complement (0 with -f)    */
  uint8_t reserved[2];
  uint32_t seed_len;          /* 0 = whole query (hunt); k (search, silica.h:449-451)  */
} dg_params;

/* One DnaHit (hunter.h:53-66) in the reference's push order.  refalign/queryalign are
 * aln_len bytes each at pool + aln_off and pool + aln_off + aln_len.                   */
typedef struct {
  uint32_t query;     /* index of the query in the batch                               */
  int32_t score;      /* -(edit or Hamming distance)                                   */
  uint32_t chr;       /* refIndex                                                      */
  uint32_t start;     /* DnaHit.start (1-based, after leading-gap stripping)           */
  uint64_t text_pos;  /* occurrence position of the neighbour string in the text       */
  uint64_t aln_off;   /* offset of refalign in the alignment pool                      */
  uint32_t aln_len;   /* alignment columns                                             */
  uint32_t alignpos;  /* search only: chrpos + leading gap columns (silica.h:522-532)  */
  uint8_t strand;     /* '+' or '-'                                                    */
  uint8_t pad[7];
} dg_hit;

/* The same DnaHit in the compact form `hunt` results travel in (24 bytes instead of 48 + two
 * alignment strings): coordinates, score, strand, and the at most three columns in which the
 * alignment differs from "the query copied into both rows" (every such column costs one unit of the
 * score, so a hit of distance <= 3 never has more).  dg_rec_alignment rebuilds refalign / queryalign
 * from the record and the query; dg_result_hits still hands out dg_hit + pool (expanded on the host
 * on first use).  ops: three 20-bit fields from bit 0, ascending by column: column (10 bits) |
 * type << 10 (1 mismatch, 2 '-' in refalign, 3 '-' in queryalign) | genomic byte << 12; bits 60-63:
 * (start - 1) - (text_pos - start of the record) + 8.                                          */
typedef struct {
  uint32_t query;
  uint32_t chr;       /* refIndex                                                      */
  uint32_t start;     /* DnaHit.start (1-based, after leading-gap stripping)           */
  int16_t score;      /* -(edit or Hamming distance)                                   */
  uint8_t strand;     /* '+' or '-'                                                    */
  uint8_t nops;       /* number of ops (0..3)                                          */
  uint64_t ops;
} dg_rec;

typedef struct {
  uint64_t n;             /* csa.size(): text length including the sentinel            */
  uint32_t sigma;         /* alphabet size including the sentinel                      */
  uint32_t kmer;          /* K of the K-mer -> SA interval table                       */
  uint64_t n_exceptions;  /* BWT symbols outside ACGT                                  */
  uint64_t device_bytes;  /* HBM held by this index                                    */
  uint32_t sa_sample;     /* suffix-array sampling rate (csa_wt<>: 32)                 */
  uint32_t nseq;          /* records set by dg_index_set_records                       */
  uint32_t bitmap_k;      /* KB of the KB-mer presence bitmap (0 = none)               */
  uint32_t reserved;
} dg_index_info;

/* Stage timings of the last dg_batch_run with profiling enabled (CUDA events on the index
 * stream), and the launch count claimed for bench.py's "gpu_launches".                */
typedef struct {
  float ms_prepare;   /* normalise + reverse-complement + script counting              */
  float ms_search;    /* k_search: neighbour enumeration x backward search (dominant)  */
  float ms_filter;    /* minimality filter, lexicographic sort, dedupe, cap scan       */
  float ms_locate;    /* SA-interval expansion + per-neighbour position sort           */
  float ms_verify;    /* chromosome lookup, context fetch, NW traceback                */
  float ms_total;     /* first to last event                                           */
  uint64_t launches;  /* kernels launched by the run (ours + CUB's)                    */
  uint64_t scripts;   /* edit scripts evaluated by k_search                            */
  uint64_t candidates;/* neighbour strings with a non-empty SA interval                */
  uint64_t located;   /* text positions located                                        */
  uint64_t hits;      /* hits emitted                                                  */
  float ms_probe;     /* k_probe_singles alone (the dominant kernel of a regular batch; part of ms_search) */
  float reserved;
} dg_profile;

/* ---- index ------------------------------------------------------------------------ */

/* load_from_checked_file(fm_index, file): hunter.h:253-260, silica.h:340-347,
 * padlock.h:259-265 -> SDSL io.hpp:917-936 -> csa_wt::load csa_wt.hpp:381-388.
 * Parses the .fm9 written by `dicey index` as-is (+ .fm9_check) and transcodes it to the
 * device layout (DESIGN.md "Data layout in HBM").  device >= 0 selects the CUDA device.  */
int dg_index_open(const char* fm9_path, int device, dg_index** out);

/* What load_from_checked_file (index.h:94, SDSL io.hpp:917-936) tests before `dicey index` keeps an
 * existing file: the .fm9_check sidecar holds the csa_wt<> type hash and the file parses to its end.
 * Host only (no device is touched).  DG_OK, or DG_ERR_IO / DG_ERR_FORMAT with the reason.      */
int dg_fm9_check(const char* fm9_path);

/* index.h:96-123 (dump -> construct) for a text already in dump format (records upper-cased,
 * joined by '\n', trailing '\n'; no sentinel): suffix array, BWT and device layout are built
 * on the GPU.  Alphabet: at most 31 distinct byte values besides the sentinel (DNA with N keeps
 * the 3-bit sort keys; IUPAC texts take 4- or 5-bit ones).                                     */
int dg_index_build_text(const uint8_t* text, uint64_t len, int device, dg_index** out);

/* Same, for the seeded synthetic reference of SURVEY.md 8(d) generated on the device
 * (dicey_b200/synth.py documents the generator).                                          */
int dg_index_build_synthetic(uint64_t seed, uint32_t nrec, uint64_t reclen, int device,
                             dg_index** out);

/* store_to_checked_file (index.h:122): writes the csa_wt<> .fm9 + .fm9_check from the device
 * index, byte for byte what SDSL writes for the same text (csa_wt.hpp:362-373: Huffman-shaped
 * wavelet tree, rank_support_v, both select_support_mcl, SA / ISA samples, byte_alphabet).      */
int dg_index_write_fm9(dg_index* idx, const char* fm9_path);

void dg_index_close(dg_index* idx);

/* fm_index.size(): hunter.h:368-369, silica.h:487-488. */
uint64_t dg_index_size(const dg_index* idx);

/* getSeqLenName (util.h:183-206): seqlen[i] = faidx length + 1, used for the text position
 * -> (refIndex, chrpos) mapping of hunter.h:358-362 / silica.h:475-479.                   */
int dg_index_set_records(dg_index* idx, const uint32_t* seqlen_plus1, uint32_t nseq);

int dg_index_get_info(const dg_index* idx, dg_index_info* info);

/* Text substrings from the device-resident text (what `dicey search` reads back from the FASTA
 * with faidx_fetch_seq for its amplicon sequences, silica.h:170-173; the index text is the
 * upper-cased FASTA): n ranges [pos[i], pos[i] + len[i]) of text positions, written back to back
 * into buf (sum of len bytes).                                                                */
int dg_index_fetch_text(dg_index* idx, const uint64_t* pos, const uint64_t* len, uint32_t n, char* buf);

/* The CUDA stream (cudaStream_t) every kernel of this index is launched on. */
void* dg_index_stream(const dg_index* idx);

/* Copies of device-resident index arrays for tests (what = "text", "sa_samples", "occ",
 * "kmer", "C", "present_kb"); returns the byte count through *bytes; buf may be NULL to query the size. */
int dg_index_debug_copy(dg_index* idx, const char* what, void* buf, uint64_t* bytes);

/* ---- batched queries ---------------------------------------------------------------- */

/* The per-query loop of `dicey hunt` (hunter.h:289-433): neighbors() on both strands
 * (neighbors.h:86-92), sdsl::count / locate / extract per neighbour
 * (suffix_array_algorithm.hpp:447-454, 521-535, 643-657), needle() / needleScore()
 * (needle.h:59-138, hunter.h:79-88), DnaHit push.  Hits come back per query in the
 * reference's push order (strand, neighbour lexicographic, position ascending), capped at
 * max_locations; the caller applies hunter.h:440 std::sort and the JSON writer.
 * seqs holds the nq raw query sequences back to back; offsets has nq+1 entries.
 * With params->seed_len = k it is the FM / NW part of `dicey search` (silica.h:449-573):
 * the last k bases are the seed, contexts are extended by |primer|-k on the 5' side, and
 * each hit additionally carries alignpos; every candidate is returned (the thal Tm gate of
 * silica.h:508-519 is dg_thal_batch).
 * seqs / offsets may live in ordinary or page-locked host memory (cudaHostAlloc,
 * cudaHostRegister, a pinned torch tensor): large batches are cut into chunks that are
 * uploaded, searched and read back as a pipeline, and page-locked input uploads without
 * occupying the calling thread.  The result's buffers are page-locked when the driver
 * grants it.                                                                              */
int dg_hunt_batch(dg_index* idx, const char* seqs, const uint64_t* offsets, uint32_t nq,
                  const dg_params* params, dg_result** out);

/* The same call split into its three phases, so a caller (bench.py) can keep inputs resident
 * in HBM: stage = host -> device copy of the queries; run = every kernel, asynchronous on the
 * index stream; fetch = device -> host copy of the hit records (synchronises).             */
int dg_batch_stage(dg_index* idx, const char* seqs, const uint64_t* offsets, uint32_t nq,
                   const dg_params* params, dg_batch** out);
int dg_batch_run(dg_batch* b);
int dg_batch_fetch(dg_batch* b, dg_result** out);
int dg_batch_summary(dg_batch* b, uint64_t* n_hits, uint64_t* n_candidates); /* synchronises */
void dg_batch_free(dg_batch* b);

/* sdsl::count over a neighbourhood: padlock.h:381-427 (exact count of each string when
 * params->distance == 0, else the sum over neighbors() of both strands) and the prune of
 * silica.h:365-394.  counts[i] receives the total for query i.                           */
int dg_count_batch(dg_index* idx, const char* seqs, const uint64_t* offsets, uint32_t nq,
                   const dg_params* params, uint64_t* counts);

/* sdsl::backward_search for literal patterns (suffix_array_algorithm.hpp:207-226): the
 * closed interval [l[i], r[i]] (r = l - 1 when empty), as SDSL reports it.              */
int dg_backward_search_batch(dg_index* idx, const char* seqs, const uint64_t* offsets,
                             uint32_t nq, uint64_t* l, uint64_t* r);

/* The hit records of the last dg_hunt_batch on this index as 16-byte wire records in HBM
 * (int32 x 4: query, chr, start, (score & 0xFFFF) | strand << 16 -- coordinates, strand and
 * distance; alignments stay in the dg_result), valid until the next dg_hunt_batch / close.  The
 * multi-GPU layer all-gathers them over NVLink from device memory.              */
int dg_index_wire_records(dg_index* idx, const void** device_ptr, uint64_t* n);

/* ---- results ------------------------------------------------------------------------ */
/* Compact records of a `hunt` result in push order (NULL with *n = 0 for a `search` result, whose
 * hits carry genomic contexts: use dg_result_hits).                                           */
const dg_rec* dg_result_records(const dg_result* r, uint64_t* n);
/* refalign / queryalign of record i (each needs |query| + 3 bytes; not NUL-terminated); returns
 * the number of alignment columns, or a negative dg_status.                                   */
int dg_result_alignment(const dg_result* r, uint64_t i, char* refalign, char* queryalign);
/* The same for a record on its own: query = the normalised sequence the hit's strand searched
 * (the query for '+', its reverse complement for '-'), qlen its length.                       */
int dg_rec_alignment(const dg_rec* rec, const char* query, uint32_t qlen, char* refalign, char* queryalign);
/* hunter.h:440 std::sort for compact records (same order as dg_hits_sort).                    */
void dg_recs_sort(dg_rec* recs, uint64_t n);
/* Bytes that crossed PCIe for this result (device -> host), for bench.py's accounting.        */
uint64_t dg_result_transfer_bytes(const dg_result* r);
const dg_hit* dg_result_hits(const dg_result* r, uint64_t* n);
const uint64_t* dg_result_query_offsets(const dg_result* r, uint32_t* nq); /* nq+1 entries */
const uint32_t* dg_result_query_status(const dg_result* r);               /* nq entries   */
const uint32_t* dg_result_query_distance(const dg_result* r);             /* clamped d    */
const char* dg_result_pool(const dg_result* r, uint64_t* bytes);
const char* dg_result_sequences(const dg_result* r, uint64_t* bytes);     /* normalised queries, same offsets as the input */
void dg_result_free(dg_result* r);

/* hunter.h:440 std::sort(ht.begin(), ht.end()) with DnaHit::operator< (hunter.h:63-65: score
 * descending, then chr, then start).  The order of ties is libstdc++'s introsort applied to the
 * push order, exactly as in the reference; sorts the n records in place.                    */
void dg_hits_sort(dg_hit* hits, uint64_t n);

/* ---- multi-GPU ---------------------------------------------------------------------- */
/* The path shards by independent queries (hunter.h:291, silica.h:429: every iteration of the
 * per-query loop reads the index only): the index is replicated on every GPU, rank r of nranks
 * (one process or one host thread per GPU) owns a contiguous shard of the query batch, and ONE
 * exchange step collects the hit records -- an all-gather.
 *
 * dg_comm_init binds rank `rank` of `nranks` to the device and stream of `idx`; every rank passes
 * the 128-byte id rank 0 got from dg_comm_get_unique_id (sent through whatever channel the
 * launcher has: torch.distributed broadcast, MPI, a file).  The transport is NCCL (libnccl.so.2 is
 * loaded at run time, so a process that already holds torch's copy shares it).
 * dg_comm_init_host builds a communicator over a caller-supplied all-gather of HOST buffers
 * (send: bytes_per_rank bytes, recv: nranks * bytes_per_rank, rank order) -- for fabrics other
 * than NCCL and for the CPU tests (gloo); it supports dg_allgather_result only.            */
typedef struct dg_comm dg_comm;
#define DG_COMM_ID_BYTES 128
typedef int (*dg_host_allgather_fn)(void* ctx, const void* send, void* recv, uint64_t bytes_per_rank);
int dg_comm_get_unique_id(void* id);
int dg_comm_init(int nranks, int rank, const void* id, dg_index* idx, dg_comm** out);
int dg_comm_init_host(int nranks, int rank, dg_host_allgather_fn fn, void* ctx, dg_comm** out);
int dg_comm_rank(const dg_comm* c);
int dg_comm_size(const dg_comm* c);
void dg_comm_destroy(dg_comm* c);

/* One hit on the wire: coordinates, strand and distance (alignment strings stay with the rank
 * that owns the query).  query is the GLOBAL query index (local index + query_base).        */
typedef struct {
  uint32_t query;
  uint32_t chr;        /* refIndex                                                          */
  uint32_t start;      /* DnaHit.start                                                      */
  int16_t score;       /* -(edit or Hamming distance)                                       */
  uint8_t strand;      /* '+' or '-'                                                        */
  uint8_t reserved;
} dg_wire;

/* The exchange step on device memory: the hit records of `b` (a batch that has run; NULL = the
 * last dg_hunt_batch on the communicator's index) are all-gathered straight from HBM with ONE
 * ncclAllGather issued on the index stream behind the batch's kernels -- no host round trip
 * between the kernels and the collective.  Every rank sends a fixed-size slot (a 16-byte header
 * holding its hit count, then `slot_records - 1` record places); the slot size is agreed once per
 * communicator and grows, on every rank alike, only when a rank's count no longer fits (the
 * all-gather is then repeated once).  On return *table points at nranks slots in DEVICE memory
 * (rank r's records start at table + r * slot_records + 1), counts[r] (HOST memory, owned by the
 * communicator) is the number of hits of rank r; both stay valid until the next call.       */
int dg_allgather_hits(dg_comm* c, dg_batch* b, uint64_t query_base, const dg_wire** table,
                      uint64_t* slot_records, const uint64_t** counts);
/* Peer mode.  When every rank's tables could be mapped on every rank at dg_comm_init (cudaIpc between
 * processes, peer access between threads; DG_COMM_P2P=0 turns it off), the kernels that PRODUCE the
 * records (k_verify of a staged batch, the commit kernel of dg_hunt_batch) store each record straight
 * into the tables of all ranks over NVLink while they run, and dg_allgather_hits shrinks to the slot
 * headers plus one 8-byte ncclAllGather that doubles as the barrier.  The records carry the query
 * base that was set BEFORE they were produced: call dg_comm_set_query_base first, then run / hunt,
 * then dg_allgather_hits with the same base (any mismatch, or a rank with more hits than a slot
 * holds, silently takes the NCCL path above).  Tables alternate between two buffers per exchange.  */
int dg_comm_set_query_base(dg_comm* c, uint64_t query_base);
/* Device -> host copy of the gathered records, compacted in rank order (sum of counts entries). */
int dg_comm_fetch_table(dg_comm* c, dg_wire* out, uint64_t capacity, uint64_t* n);

/* The same exchange for complete results (records, alignment strings, per-query status and the
 * normalised queries): every rank passes its local dg_result and receives the result of the whole
 * batch, query ids / offsets rebased in rank order -- what rank 0 of a multi-process `hunt` needs to
 * print the JSON of every query.  Works on either transport.                                  */
int dg_allgather_result(dg_comm* c, const dg_result* local, dg_result** global);

/* Flat wire format of one result (what dg_allgather_result moves).                          */
int dg_result_pack(const dg_result* r, void* buf, uint64_t* bytes);   /* buf NULL -> size  */
int dg_result_unpack(const void* buf, uint64_t bytes, dg_result** out);

/* ---- melting temperatures (the Tm gate of `dicey search`) ---------------------------- */
/* primer3thal::thal(oligo1, oligo2, &a, &o) with a.type = thal_end1 and a.temponly = 1 as
 * silica.h:316-329,508-519 and padlock.h call it, for n pairs at once on the GPU, bit for bit
 * (o.temp as a double; ok[i] = the bool thal() returns).  dg_thal_open reads the primer3_config
 * directory a dicey installation ships (-i, silica.h:216) as get_thermodynamic_values does
 * (thal.h:2368-2393); mv / dv / dntp in mM, dna_conc in nM (silica.h:243-246).
 * dg_thal_open_tables reads the table dump `oracle/_ref/dicey_ref thal` writes (tests).
 * Lengths as in the reference (thal.h:58, :2440-2451): at most one side longer than
 * THAL_MAX_ALIGN = 60, neither longer than THAL_MAX_SEQ = 10 000; any other pair gets ok = 0
 * and THAL_ERROR_SCORE (-999999), an empty side ok = 0 and 0.0 -- what thal() leaves in o.temp. */
typedef struct dg_thal dg_thal;
int dg_thal_open(const char* primer3_config_dir, double mv, double dv, double dntp, double dna_conc, int device, dg_thal** out);
int dg_thal_open_tables(const char* table_dump_path, int device, dg_thal** out);
int dg_thal_batch(dg_thal* t, const char* seq1, const uint64_t* off1, const char* seq2, const uint64_t* off2, uint32_t n,
                  double* tm, uint8_t* ok);
void dg_thal_close(dg_thal* t);

/* ---- diagnostics -------------------------------------------------------------------- */
int dg_profile_enable(dg_index* idx, int on);
int dg_profile_get(dg_index* idx, dg_profile* out);
const char* dg_last_error(void);
const char* dg_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DICEY_B200_H */
