"""ctypes binding of ``include/dicey_b200.h`` and the host-side mirror of dicey's ``hunt`` driver.

The functions here follow the reference's driver code, with the hot loops replaced by one
batched library call:

* :class:`Index`            ``load_from_checked_file`` (hunter.h:253-260) / ``getSeqLenName`` (util.h:183-206)
* :meth:`Index.hunt`        the per-query loop of hunter.h:289-433 for a whole batch
* :meth:`HuntResult.sorted_hits`  ``std::sort(ht)`` hunter.h:440
* :func:`hunt_json`         ``writeJsonDnaHitOut`` hunter.h:99-160 (byte-identical JSON lines)
"""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def library_path() -> str:
    # DICEY_B200_LIB: another build of the same library (A/B runs of kernel variants)
    return os.environ.get("DICEY_B200_LIB") or os.path.join(_HERE, "libdicey_b200.so")


class DiceyB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dicey_b200 error {code}: {msg}")
        self.code = code


class _Params(C.Structure):
    _fields_ = [("distance", C.c_uint32), ("max_neighborhood", C.c_uint32), ("max_locations", C.c_uint32),
                ("indel", C.c_uint8), ("reverse", C.c_uint8), ("reserved", C.c_uint8 * 2), ("seed_len", C.c_uint32)]


class _Info(C.Structure):
    _fields_ = [("n", C.c_uint64), ("sigma", C.c_uint32), ("kmer", C.c_uint32), ("n_exceptions", C.c_uint64),
                ("device_bytes", C.c_uint64), ("sa_sample", C.c_uint32), ("nseq", C.c_uint32), ("bitmap_k", C.c_uint32),
                ("reserved", C.c_uint32)]


class _Profile(C.Structure):
    _fields_ = [("ms_prepare", C.c_float), ("ms_search", C.c_float), ("ms_filter", C.c_float),
                ("ms_locate", C.c_float), ("ms_verify", C.c_float), ("ms_total", C.c_float),
                ("launches", C.c_uint64), ("scripts", C.c_uint64), ("candidates", C.c_uint64),
                ("located", C.c_uint64), ("hits", C.c_uint64), ("ms_probe", C.c_float), ("reserved", C.c_float)]


HIT_DTYPE = np.dtype([("query", "<u4"), ("score", "<i4"), ("chr", "<u4"), ("start", "<u4"), ("text_pos", "<u8"),
                      ("aln_off", "<u8"), ("aln_len", "<u4"), ("alignpos", "<u4"), ("strand", "u1"),
                      ("pad", "u1", (7,))])
assert HIT_DTYPE.itemsize == 48

Q_TOO_SHORT, Q_DIST_ADJUSTED, Q_HIT_CAP, Q_NBR_CAP, Q_NBR_UNVERIFIED, Q_SKIPPED, Q_UNSUPPORTED = (1 << i for i in range(7))

_lib = None

# every symbol include/dicey_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "dg_index_open", "dg_index_build_text", "dg_index_build_synthetic", "dg_index_write_fm9", "dg_index_close",
    "dg_index_size", "dg_index_set_records", "dg_index_get_info", "dg_index_stream", "dg_index_debug_copy",
    "dg_hunt_batch", "dg_batch_stage", "dg_batch_run", "dg_batch_fetch", "dg_batch_summary",
    "dg_batch_free", "dg_index_wire_records", "dg_index_fetch_text", "dg_thal_open", "dg_thal_open_tables", "dg_thal_batch", "dg_thal_close",
    "dg_count_batch", "dg_backward_search_batch", "dg_result_hits", "dg_result_query_offsets",
    "dg_result_query_status", "dg_result_query_distance", "dg_result_pool", "dg_result_sequences",
    "dg_result_free", "dg_hits_sort", "dg_result_pack", "dg_result_unpack", "dg_profile_enable",
    "dg_profile_get", "dg_last_error", "dg_version",
    "dg_comm_get_unique_id", "dg_comm_init", "dg_comm_init_host", "dg_comm_rank", "dg_comm_size", "dg_comm_destroy",
    "dg_allgather_hits", "dg_comm_fetch_table", "dg_allgather_result", "dg_comm_set_query_base",
    "dg_fm9_check", "dg_result_records", "dg_result_alignment", "dg_rec_alignment", "dg_recs_sort", "dg_result_transfer_bytes",
]

REC_DTYPE = np.dtype([("query", "<u4"), ("chr", "<u4"), ("start", "<u4"), ("score", "<i2"), ("strand", "u1"), ("nops", "u1"),
                      ("ops", "<u8")])
assert REC_DTYPE.itemsize == 24
WIRE_DTYPE = np.dtype([("query", "<u4"), ("chr", "<u4"), ("start", "<u4"), ("score", "<i2"), ("strand", "u1"), ("reserved", "u1")])
assert WIRE_DTYPE.itemsize == 16
HOST_ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)


def library() -> C.CDLL:
    """Loads ``libdicey_b200.so``; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C dicey_b200/csrc` (dicey_b200 has no CPU path)")
    lib = C.CDLL(path)
    vp, u64p, u32p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    lib.dg_last_error.restype = C.c_char_p
    lib.dg_version.restype = C.c_char_p
    lib.dg_index_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    lib.dg_index_build_text.argtypes = [vp, C.c_uint64, C.c_int, C.POINTER(vp)]
    lib.dg_index_build_synthetic.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(vp)]
    lib.dg_index_write_fm9.argtypes = [vp, C.c_char_p]
    lib.dg_fm9_check.argtypes = [C.c_char_p]
    lib.dg_index_close.argtypes = [vp]
    lib.dg_index_close.restype = None
    lib.dg_index_size.argtypes = [vp]
    lib.dg_index_size.restype = C.c_uint64
    lib.dg_index_set_records.argtypes = [vp, vp, C.c_uint32]
    lib.dg_index_get_info.argtypes = [vp, C.POINTER(_Info)]
    lib.dg_index_stream.argtypes = [vp]
    lib.dg_index_stream.restype = vp
    lib.dg_index_debug_copy.argtypes = [vp, C.c_char_p, vp, u64p]
    for name in ("dg_hunt_batch", "dg_batch_stage"):
        getattr(lib, name).argtypes = [vp, vp, vp, C.c_uint32, C.POINTER(_Params), C.POINTER(vp)]
    lib.dg_batch_run.argtypes = [vp]
    lib.dg_batch_fetch.argtypes = [vp, C.POINTER(vp)]
    lib.dg_batch_summary.argtypes = [vp, u64p, u64p]
    lib.dg_index_wire_records.argtypes = [vp, C.POINTER(vp), u64p]
    lib.dg_index_fetch_text.argtypes = [vp, vp, vp, C.c_uint32, vp]
    lib.dg_thal_open.argtypes = [C.c_char_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(vp)]
    lib.dg_thal_open_tables.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    lib.dg_thal_batch.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, vp, vp]
    lib.dg_thal_close.argtypes = [vp]
    lib.dg_thal_close.restype = None
    lib.dg_batch_free.argtypes = [vp]
    lib.dg_batch_free.restype = None
    lib.dg_count_batch.argtypes = [vp, vp, vp, C.c_uint32, C.POINTER(_Params), vp]
    lib.dg_backward_search_batch.argtypes = [vp, vp, vp, C.c_uint32, vp, vp]
    lib.dg_result_hits.argtypes = [vp, u64p]
    lib.dg_result_hits.restype = vp
    lib.dg_result_records.argtypes = [vp, u64p]
    lib.dg_result_records.restype = vp
    lib.dg_result_alignment.argtypes = [vp, C.c_uint64, vp, vp]
    lib.dg_rec_alignment.argtypes = [vp, vp, C.c_uint32, vp, vp]
    lib.dg_recs_sort.argtypes = [vp, C.c_uint64]
    lib.dg_recs_sort.restype = None
    lib.dg_result_transfer_bytes.argtypes = [vp]
    lib.dg_result_transfer_bytes.restype = C.c_uint64
    lib.dg_result_query_offsets.argtypes = [vp, u32p]
    lib.dg_result_query_offsets.restype = vp
    lib.dg_result_query_status.argtypes = [vp]
    lib.dg_result_query_status.restype = vp
    lib.dg_result_query_distance.argtypes = [vp]
    lib.dg_result_query_distance.restype = vp
    lib.dg_result_pool.argtypes = [vp, u64p]
    lib.dg_result_pool.restype = vp
    lib.dg_result_sequences.argtypes = [vp, u64p]
    lib.dg_result_sequences.restype = vp
    lib.dg_result_free.argtypes = [vp]
    lib.dg_result_free.restype = None
    lib.dg_hits_sort.argtypes = [vp, C.c_uint64]
    lib.dg_hits_sort.restype = None
    lib.dg_result_pack.argtypes = [vp, vp, u64p]
    lib.dg_result_unpack.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
    lib.dg_profile_enable.argtypes = [vp, C.c_int]
    lib.dg_profile_get.argtypes = [vp, C.POINTER(_Profile)]
    lib.dg_comm_get_unique_id.argtypes = [vp]
    lib.dg_comm_init.argtypes = [C.c_int, C.c_int, vp, vp, C.POINTER(vp)]
    lib.dg_comm_init_host.argtypes = [C.c_int, C.c_int, HOST_ALLGATHER_FN, vp, C.POINTER(vp)]
    lib.dg_comm_rank.argtypes = [vp]
    lib.dg_comm_size.argtypes = [vp]
    lib.dg_comm_destroy.argtypes = [vp]
    lib.dg_comm_destroy.restype = None
    lib.dg_allgather_hits.argtypes = [vp, vp, C.c_uint64, C.POINTER(vp), u64p, C.POINTER(vp)]
    lib.dg_comm_fetch_table.argtypes = [vp, vp, C.c_uint64, u64p]
    lib.dg_comm_set_query_base.argtypes = [vp, C.c_uint64]
    lib.dg_allgather_result.argtypes = [vp, vp, C.POINTER(vp)]
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != 0:
        raise DiceyB200Error(rc, library().dg_last_error().decode())


def _from_ptr(ptr, nbytes: int, dtype) -> np.ndarray:
    """Zero-copy view of library-owned memory (the owner must outlive the array)."""
    if not ptr or nbytes == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype)


def pack_sequences(seqs) -> tuple[np.ndarray, np.ndarray]:
    """list of bytes/str -> (concatenated uint8 buffer, uint64 offsets[nq+1])."""
    if isinstance(seqs, tuple) and len(seqs) == 2:
        return np.ascontiguousarray(seqs[0], dtype=np.uint8), np.ascontiguousarray(seqs[1], dtype=np.uint64)
    if isinstance(seqs, np.ndarray) and seqs.ndim == 2:  # rectangular (nq, L) uint8
        nq, L = seqs.shape
        return np.ascontiguousarray(seqs, dtype=np.uint8).reshape(-1), np.arange(nq + 1, dtype=np.uint64) * np.uint64(L)
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    return np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8), off


@dataclass
class HuntParams:
    """HunterConfig (hunter.h:37-50): the fields the hot path reads, with dicey's defaults."""
    distance: int = 1            # -d
    hamming: bool = False        # -n  (indel = not hamming)
    forward_only: bool = False   # -f  (reverse = not forward_only)
    maxmatches: int = 1000       # -m
    max_neighborhood: int = 10000  # -x
    seed_len: int = 0            # search: -k

    def to_c(self) -> _Params:
        return _Params(self.distance, self.max_neighborhood, self.maxmatches, 0 if self.hamming else 1,
                       0 if self.forward_only else 1, (C.c_uint8 * 2)(0, 0), self.seed_len)


class HuntResult:
    """Hits of one batch in the reference's push order (hunter.h:349-433).

    `hunt` results come back from the device in compact form (``records``: REC_DTYPE, 24 bytes per
    hit, alignments as edit operations); ``hits`` / ``pool`` / ``qoff`` / ``status`` / ``dist`` are the
    expanded views the library derives on the host the first time one of them is read."""

    _LAZY = ("hits", "qoff", "status", "dist", "pool", "seqs")

    def __init__(self, hits=None, qoff=None, status=None, dist=None, pool=None, seqs=None, seq_off=None, handle=None):
        if hits is not None:
            self.hits, self.qoff, self.status, self.dist, self.pool, self.seqs = hits, qoff, status, dist, pool, seqs
        self.seq_off = seq_off
        self._res = handle  # the arrays are views into this dg_result

    def __getattr__(self, name):
        if name in HuntResult._LAZY and self.__dict__.get("_res"):
            self._load()
            return self.__dict__[name]
        raise AttributeError(name)

    def _load(self) -> None:
        lib, res = library(), self._res
        n = C.c_uint64(0)
        hp = lib.dg_result_hits(res, C.byref(n))
        self.hits = _from_ptr(hp, n.value * HIT_DTYPE.itemsize, HIT_DTYPE)
        nq = C.c_uint32(0)
        qp = lib.dg_result_query_offsets(res, C.byref(nq))
        self.qoff = _from_ptr(qp, (nq.value + 1) * 8, np.uint64)
        self.status = _from_ptr(lib.dg_result_query_status(res), nq.value * 4, np.uint32)
        self.dist = _from_ptr(lib.dg_result_query_distance(res), nq.value * 4, np.uint32)
        nb = C.c_uint64(0)
        pp = lib.dg_result_pool(res, C.byref(nb))
        self.pool = _from_ptr(pp, nb.value, np.uint8)
        sp = lib.dg_result_sequences(res, C.byref(nb))
        self.seqs = _from_ptr(sp, nb.value, np.uint8)

    @property
    def records(self) -> np.ndarray:
        """The compact records as they crossed PCIe (empty for a `search` result)."""
        n = C.c_uint64(0)
        p = library().dg_result_records(self._res, C.byref(n)) if self._res else None
        return _from_ptr(p, n.value * REC_DTYPE.itemsize, REC_DTYPE)

    def alignment(self, i: int) -> tuple[str, str]:
        """(refalign, queryalign) of hit i, rebuilt from its compact record (dg_result_alignment)."""
        ra, qa = C.create_string_buffer(300), C.create_string_buffer(300)
        n = library().dg_result_alignment(self._res, i, ra, qa)
        if n < 0:
            _check(n)
        return ra.raw[:n].decode(), qa.raw[:n].decode()

    @property
    def transfer_bytes(self) -> int:
        return int(library().dg_result_transfer_bytes(self._res)) if self._res else 0

    def close(self) -> None:
        if self._res:
            self._load()
            h, self._res = self._res, None
            self.hits = self.hits.copy(); self.qoff = self.qoff.copy(); self.status = self.status.copy()
            self.dist = self.dist.copy(); self.pool = self.pool.copy(); self.seqs = self.seqs.copy()
            library().dg_result_free(h)

    def __del__(self):
        if self.__dict__.get("_res"):
            try:
                library().dg_result_free(self._res)
            except Exception:
                pass
            self._res = None

    @property
    def nq(self) -> int:
        return len(self.qoff) - 1

    def sequence(self, q: int) -> bytes:
        return self.seqs[int(self.seq_off[q]):int(self.seq_off[q + 1])].tobytes()

    def _rec(self, h, search: bool = False):
        o, n = int(h["aln_off"]), int(h["aln_len"])
        ra = self.pool[o:o + n].tobytes().decode()
        if search:
            return (int(h["chr"]), int(h["start"]), int(h["alignpos"]), chr(int(h["strand"])), ra)
        qa = self.pool[o + n:o + 2 * n].tobytes().decode()
        return (int(h["score"]), int(h["chr"]), int(h["start"]), chr(int(h["strand"])), ra, qa)

    def push_hits(self, q: int):
        """(score, chr, start, strand, refalign, queryalign) in push order."""
        return [self._rec(h) for h in self.hits[int(self.qoff[q]):int(self.qoff[q + 1])]]

    def sorted_hits(self, q: int):
        """After std::sort(ht) (hunter.h:440), ties resolved exactly as libstdc++ does."""
        h = self.hits[int(self.qoff[q]):int(self.qoff[q + 1])].copy()
        if len(h) > 1:
            library().dg_hits_sort(h.ctypes.data, len(h))
        return [self._rec(x) for x in h]

    def seed_hits(self, q: int):
        """search: (refIndex, chrpos, alignpos, strand, genomicseq) per candidate, push order."""
        return [self._rec(h, True) for h in self.hits[int(self.qoff[q]):int(self.qoff[q + 1])]]

    def records_tsv(self, params: HuntParams, raws=None, q_lo: int = 0, q_hi: int | None = None) -> str:
        """Canonical dump of queries [q_lo, q_hi): per query one Q line (index, normalised sequence,
        clamped distance, #messages, #hits), its messages (M), the hits in push order (P) and after
        std::sort (S) -- the format of `dicey_ref hunt --records` (oracle/ref_driver.cpp), used for
        whole-batch parity checks (bench.py, tests)."""
        q_hi = self.nq if q_hi is None else q_hi
        out = []
        for q in range(q_lo, q_hi):
            raw = None if raws is None else bytes(raws[q])
            msgs = self.messages(q, params, raw)
            if msgs and msgs[0].startswith("Error"):
                seq, dist, push, srt = ("" if raw is None else raw.decode()), params.distance, [], []
            else:
                seq, dist, push, srt = self.sequence(q).decode(), int(self.dist[q]), self.push_hits(q), self.sorted_hits(q)
            out.append(f"Q\t{q - q_lo}\t{seq}\t{dist}\t{len(msgs)}\t{len(push)}\n")
            out.extend(f"M\t{m}\n" for m in msgs)
            out.extend("P\t%d\t%d\t%d\t%s\t%s\t%s\n" % h for h in push)
            out.extend("S\t%d\t%d\t%d\t%s\t%s\t%s\n" % h for h in srt)
        return "".join(out)

    def messages(self, q: int, params: HuntParams, raw: bytes | None = None):
        """The msg vector of hunter.h for query q, in the reference's order."""
        msg = []
        st = int(self.status[q])
        if st & Q_TOO_SHORT:
            return ["Error: Input sequence is shorter than 10 nucleotides!"]
        if raw is not None:
            up = raw.upper()
            msg += ["Warning: Non-DNA character in nucleotide sequence detected and replaced by 'N'!"
                    for ch in up if ch not in b"ACGT"]
        if st & Q_DIST_ADJUSTED:
            msg.append("Warning: Distance was adjusted to sequence length!")
        if st & Q_NBR_CAP:
            x = params.max_neighborhood
            msg.append(f"Warning: Neighborhood size exceeds {x} candidates. Only first {x} neighbors are searched, "
                       "results are likely incomplete!")
        if st & Q_NBR_UNVERIFIED:
            msg.append(f"Warning: Neighborhood may exceed {params.max_neighborhood} candidates; the reference's truncation is not "
                       "reproduced for sequences longer than 40 nucleotides, all neighbors were searched!")
        if st & Q_HIT_CAP:
            m = params.maxmatches
            msg.append(f"Warning: More than {m} matches found. Only first {m} matches are reported, results are "
                       "likely incomplete!")
        return msg


def collect_result(res, seq_off) -> HuntResult:
    """A HuntResult over a library-owned dg_result (freed with the HuntResult); the expanded arrays
    are fetched on first use."""
    return HuntResult(seq_off=seq_off, handle=res)


def pack_result(res: HuntResult) -> np.ndarray:
    """dg_result_pack: the flat wire format of one result (a uint8 array)."""
    if res._res is None:
        raise ValueError("result has been closed")
    lib = library()
    nb = C.c_uint64(0)
    _check(lib.dg_result_pack(res._res, None, C.byref(nb)))
    buf = np.empty(nb.value, dtype=np.uint8)
    _check(lib.dg_result_pack(res._res, buf.ctypes.data, C.byref(nb)))
    return buf


def unpack_result(buf: np.ndarray, seq_off=None) -> HuntResult:
    """dg_result_unpack: a HuntResult from the wire format."""
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    h = C.c_void_p()
    _check(library().dg_result_unpack(buf.ctypes.data, buf.size, C.byref(h)))
    return collect_result(h.value, seq_off)


class Index:
    """The device-resident FM-index (one per GPU)."""

    def __init__(self, handle: int, names=None):
        self._h = handle
        self.names = list(names) if names else []

    # -- construction --------------------------------------------------------------------
    @classmethod
    def open(cls, fm9_path: str, device: int = 0) -> "Index":
        h = C.c_void_p()
        _check(library().dg_index_open(os.fsencode(fm9_path), device, C.byref(h)))
        return cls(h.value)

    @classmethod
    def build_text(cls, text: bytes | np.ndarray, device: int = 0) -> "Index":
        a = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else np.ascontiguousarray(text, np.uint8)
        h = C.c_void_p()
        _check(library().dg_index_build_text(a.ctypes.data, a.size, device, C.byref(h)))
        return cls(h.value)

    @classmethod
    def build_synthetic(cls, seed: int, nrec: int, reclen: int, device: int = 0) -> "Index":
        h = C.c_void_p()
        _check(library().dg_index_build_synthetic(seed, nrec, reclen, device, C.byref(h)))
        return cls(h.value, [f"chr{i + 1}" for i in range(nrec)])

    def close(self) -> None:
        if self._h:
            library().dg_index_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- metadata ------------------------------------------------------------------------
    def set_records(self, names, seqlen_plus1) -> None:
        """getSeqLenName (util.h:183-206): names and faidx lengths + 1."""
        a = np.ascontiguousarray(seqlen_plus1, dtype=np.uint32)
        _check(library().dg_index_set_records(self._h, a.ctypes.data, a.size))
        self.names = list(names)

    def size(self) -> int:
        return int(library().dg_index_size(self._h))

    def info(self) -> dict:
        i = _Info()
        _check(library().dg_index_get_info(self._h, C.byref(i)))
        return {k: int(getattr(i, k)) for k, _ in _Info._fields_}

    def stream(self) -> int:
        return int(library().dg_index_stream(self._h) or 0)

    def wire_records(self) -> tuple[int, int]:
        """(device address, count) of the 16-byte wire records of the last hunt() on this index."""
        p, n = C.c_void_p(), C.c_uint64(0)
        _check(library().dg_index_wire_records(self._h, C.byref(p), C.byref(n)))
        return int(p.value or 0), int(n.value)

    def debug_array(self, what: str, dtype=np.uint8) -> np.ndarray:
        nb = C.c_uint64(0)
        _check(library().dg_index_debug_copy(self._h, what.encode(), None, C.byref(nb)))
        buf = np.empty(nb.value, dtype=np.uint8)
        _check(library().dg_index_debug_copy(self._h, what.encode(), buf.ctypes.data, C.byref(nb)))
        return buf.view(dtype)

    def write_fm9(self, path: str) -> None:
        _check(library().dg_index_write_fm9(self._h, os.fsencode(path)))

    def profile(self, on: bool) -> None:
        _check(library().dg_profile_enable(self._h, 1 if on else 0))

    def last_profile(self) -> dict:
        p = _Profile()
        _check(library().dg_profile_get(self._h, C.byref(p)))
        return {k: getattr(p, k) for k, _ in _Profile._fields_}

    # -- queries -------------------------------------------------------------------------
    def _collect(self, res, seq_off) -> HuntResult:
        return collect_result(res, seq_off)

    def hunt(self, seqs, params: HuntParams | None = None) -> HuntResult:
        """hunter.h:289-433 for every query of the batch (one library call, host buffers)."""
        params = params or HuntParams()
        buf, off = pack_sequences(seqs)
        res = C.c_void_p()
        p = params.to_c()
        _check(library().dg_hunt_batch(self._h, buf.ctypes.data, off.ctypes.data, len(off) - 1, C.byref(p), C.byref(res)))
        return self._collect(res.value, off)

    def stage(self, seqs, params: HuntParams | None = None) -> "Batch":
        params = params or HuntParams()
        buf, off = pack_sequences(seqs)
        b = C.c_void_p()
        p = params.to_c()
        _check(library().dg_batch_stage(self._h, buf.ctypes.data, off.ctypes.data, len(off) - 1, C.byref(p), C.byref(b)))
        return Batch(self, b.value, off)

    def count(self, seqs, params: HuntParams | None = None) -> np.ndarray:
        """Neighbourhood counts (padlock.h:381-427, silica.h:365-394)."""
        params = params or HuntParams()
        buf, off = pack_sequences(seqs)
        out = np.zeros(len(off) - 1, dtype=np.uint64)
        p = params.to_c()
        _check(library().dg_count_batch(self._h, buf.ctypes.data, off.ctypes.data, len(off) - 1, C.byref(p), out.ctypes.data))
        return out

    def backward_search(self, seqs) -> tuple[np.ndarray, np.ndarray]:
        """sdsl::backward_search closed intervals [l, r] of literal patterns."""
        buf, off = pack_sequences(seqs)
        nq = len(off) - 1
        l = np.zeros(nq, dtype=np.uint64)
        r = np.zeros(nq, dtype=np.uint64)
        _check(library().dg_backward_search_batch(self._h, buf.ctypes.data, off.ctypes.data, nq, l.ctypes.data, r.ctypes.data))
        return l, r


class Batch:
    """A staged batch: queries resident in HBM (dg_batch_stage / run / fetch)."""

    def __init__(self, index: Index, handle: int, off):
        self.index, self._h, self.off = index, handle, off

    def run(self) -> None:
        _check(library().dg_batch_run(self._h))

    def summary(self) -> tuple[int, int]:
        a, b = C.c_uint64(0), C.c_uint64(0)
        _check(library().dg_batch_summary(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def fetch(self) -> HuntResult:
        res = C.c_void_p()
        _check(library().dg_batch_fetch(self._h, C.byref(res)))
        return self.index._collect(res.value, self.off)

    def free(self) -> None:
        if self._h:
            library().dg_batch_free(self._h)
            self._h = None


class Comm:
    """One rank of the multi-GPU exchange (include/dicey_b200.h, "multi-GPU"): NCCL over the index
    stream (``Comm.init``) or a caller-supplied host all-gather (``Comm.init_host``, CPU tests)."""

    ID_BYTES = 128

    def __init__(self, handle: int, nranks: int, rank: int, keep=None):
        self._h, self.nranks, self.rank, self._keep = handle, nranks, rank, keep

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * Comm.ID_BYTES)()
        _check(library().dg_comm_get_unique_id(buf))
        return bytes(buf)

    @classmethod
    def init(cls, nranks: int, rank: int, uid: bytes, index: "Index") -> "Comm":
        h = C.c_void_p()
        idb = (C.c_uint8 * Comm.ID_BYTES).from_buffer_copy(uid)
        _check(library().dg_comm_init(nranks, rank, idb, index._h, C.byref(h)))
        return cls(h.value, nranks, rank)

    @classmethod
    def init_host(cls, nranks: int, rank: int, allgather) -> "Comm":
        """allgather(send: np.ndarray[uint8, n]) -> np.ndarray[uint8, nranks * n] in rank order."""
        def _cb(ctx, send, recv, nbytes):
            try:
                a = np.frombuffer((C.c_uint8 * nbytes).from_address(send), dtype=np.uint8) if nbytes else np.zeros(0, np.uint8)
                out = np.ascontiguousarray(allgather(a), dtype=np.uint8)
                if out.size != nbytes * nranks:
                    return 2
                if out.size:
                    C.memmove(recv, out.ctypes.data, out.size)
                return 0
            except Exception:
                return 1
        fn = HOST_ALLGATHER_FN(_cb)
        h = C.c_void_p()
        _check(library().dg_comm_init_host(nranks, rank, fn, None, C.byref(h)))
        return cls(h.value, nranks, rank, keep=fn)

    def set_query_base(self, query_base: int) -> None:
        """Peer mode: the global index of this rank's first query, set before the records are produced."""
        _check(library().dg_comm_set_query_base(self._h, query_base))

    def allgather_hits(self, batch: "Batch | None" = None, query_base: int = 0):
        """The exchange step on device memory: (device address of the slot table, records per slot,
        per-rank hit counts).  Rank r's records start at table + (r * slot + 1) * 16 bytes."""
        t, slot, cn = C.c_void_p(), C.c_uint64(0), C.c_void_p()
        _check(library().dg_allgather_hits(self._h, batch._h if batch is not None else None, query_base, C.byref(t), C.byref(slot),
                                           C.byref(cn)))
        counts = _from_ptr(cn.value, 8 * self.nranks, np.uint64).copy()
        return int(t.value or 0), int(slot.value), counts

    def fetch_table(self) -> np.ndarray:
        """The gathered wire records on the host, compacted in rank order."""
        n = C.c_uint64(0)
        _check(library().dg_comm_fetch_table(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=WIRE_DTYPE)
        if n.value:
            _check(library().dg_comm_fetch_table(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def allgather_result(self, res: "HuntResult", seq_off=None) -> "HuntResult":
        """Complete results of every rank, merged in rank order (query ids and offsets rebased)."""
        if res._res is None:
            raise ValueError("result has been closed")
        g = C.c_void_p()
        _check(library().dg_allgather_result(self._h, res._res, C.byref(g)))
        return collect_result(g.value, seq_off)

    def close(self) -> None:
        if self._h:
            library().dg_comm_destroy(self._h)
            self._h = None


class Thal:
    """Batched melting temperatures (primer3 thal, thal_end1) on the GPU: the gate of silica.h:508-519."""

    def __init__(self, handle: int):
        self._h = handle

    @classmethod
    def open(cls, primer3_config_dir: str, mv: float = 50.0, dv: float = 1.5, dntp: float = 0.6, dna_conc: float = 50.0,
             device: int = 0) -> "Thal":
        h = C.c_void_p()
        _check(library().dg_thal_open(os.fsencode(primer3_config_dir), mv, dv, dntp, dna_conc, device, C.byref(h)))
        return cls(h.value)

    @classmethod
    def open_tables(cls, dump_path: str, device: int = 0) -> "Thal":
        h = C.c_void_p()
        _check(library().dg_thal_open_tables(os.fsencode(dump_path), device, C.byref(h)))
        return cls(h.value)

    def tm(self, oligos1, oligos2) -> tuple[np.ndarray, np.ndarray]:
        """(Tm in Celsius as float64, ok flags) for the pairs (oligos1[i], oligos2[i])."""
        b1, o1 = pack_sequences(oligos1)
        b2, o2 = pack_sequences(oligos2)
        n = len(o1) - 1
        assert n == len(o2) - 1
        tm = np.zeros(n, dtype=np.float64)
        ok = np.zeros(n, dtype=np.uint8)
        _check(library().dg_thal_batch(self._h, b1.ctypes.data, o1.ctypes.data, b2.ctypes.data, o2.ctypes.data, n, tm.ctypes.data,
                                       ok.ctypes.data))
        return tm, ok

    def close(self) -> None:
        if self._h:
            library().dg_thal_close(self._h)
            self._h = None


DICEY_VERSION = "0.5.1"  # reference src/version.h:8 (meta.version of the JSON envelope)


def _dump(obj) -> str:
    # nlohmann::json::dump(): keys in std::map order, no whitespace
    return json.dumps(obj, sort_keys=True, separators=(",", ":"), ensure_ascii=False)


def hunt_json(result: HuntResult, q: int, params: HuntParams, names, genome: str, outfile: str = "",
              qname: str = "", raw: bytes | None = None) -> str:
    """writeJsonDnaHitOut (hunter.h:99-160) for query q: one JSON object, byte for byte."""
    msgs = result.messages(q, params, raw)
    out = ['{"errors": [']
    errors = False
    for i, m in enumerate(msgs):
        t = "warning"
        if m.startswith("Error"):
            errors, t = True, "error"
        out.append(("," if i else "") + _dump({"type": t, "title": m}))
    out.append("]")
    if not errors:
        meta = {"version": DICEY_VERSION, "subcommand": "hunt", "distance": int(result.dist[q]),
                "sequence": result.sequence(q).decode(), "genome": genome, "outfile": outfile,
                "maxmatches": params.maxmatches, "hamming": params.hamming, "forwardonly": params.forward_only}
        if qname:
            meta["name"] = qname
        out.append(',"meta":' + _dump(meta) + ',"data":[')
        oldchr, oldstart, first = 999999, 0, True
        for (score, chrom, start, strand, ra, qa) in result.sorted_hits(q):
            if oldchr != chrom or oldstart != start:
                if not first:
                    out.append(",")
                first = False
                end = start + sum(1 for ch in ra if ch != "-") - 1
                out.append(_dump({"distance": abs(score), "chr": names[chrom], "start": start, "end": end,
                                  "strand": strand, "refalign": ra, "queryalign": qa}))
            oldchr, oldstart = chrom, start
        out.append("]")
    out.append("}")
    return "".join(out)
