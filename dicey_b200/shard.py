"""Multi-GPU plumbing of the hot path: the index is replicated on every GPU, the
query batch is split into contiguous equal shards by rank (so concatenating the per-rank results
in rank order restores input order), every rank runs independently, and ONE exchange step
collects the hit records: an all-gather of the packed per-rank results over torch.distributed
(NCCL over NVLink on the GPU box; gloo in the CPU tests).  There is no other collective on the
data path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .api import HIT_DTYPE, HuntResult, _check, _from_ptr, library


def shard_bounds(n: int, world: int) -> list[int]:
    """Contiguous, equal-count shards: rank r owns queries [b[r], b[r+1])."""
    return [(n * r) // world for r in range(world + 1)]


def pack_result(res: HuntResult) -> np.ndarray:
    """dg_result_pack: the wire format of one rank's hits (a flat uint8 array)."""
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
    lib = library()
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    h = C.c_void_p()
    _check(lib.dg_result_unpack(buf.ctypes.data, buf.size, C.byref(h)))
    n = C.c_uint64(0)
    hp = lib.dg_result_hits(h, C.byref(n))
    hits = _from_ptr(hp, n.value * HIT_DTYPE.itemsize, HIT_DTYPE)
    nq = C.c_uint32(0)
    qp = lib.dg_result_query_offsets(h, C.byref(nq))
    qoff = _from_ptr(qp, (nq.value + 1) * 8, np.uint64)
    status = _from_ptr(lib.dg_result_query_status(h), nq.value * 4, np.uint32)
    dist = _from_ptr(lib.dg_result_query_distance(h), nq.value * 4, np.uint32)
    nb = C.c_uint64(0)
    pp = lib.dg_result_pool(h, C.byref(nb))
    pool = _from_ptr(pp, nb.value, np.uint8)
    sp = lib.dg_result_sequences(h, C.byref(nb))
    seqs = _from_ptr(sp, nb.value, np.uint8)
    return HuntResult(hits, qoff, status, dist, pool, seqs, seq_off, h.value)


def merge_results(parts: list[HuntResult], seq_offs: list[np.ndarray]) -> HuntResult:
    """Concatenates per-rank results (rank order = query order): query ids, hit offsets, pool
    offsets and sequence offsets are rebased; the records themselves are untouched."""
    hits, qoff, status, dist, pool, seqs, soff = [], [np.zeros(1, np.uint64)], [], [], [], [], [np.zeros(1, np.uint64)]
    qbase = hbase = pbase = sbase = 0
    for r, so in zip(parts, seq_offs):
        h = r.hits.copy()
        h["query"] += np.uint32(qbase)
        h["aln_off"] += np.uint64(pbase)
        hits.append(h)
        qoff.append(r.qoff[1:] + np.uint64(hbase))
        status.append(r.status)
        dist.append(r.dist)
        pool.append(r.pool)
        seqs.append(r.seqs)
        so = np.asarray(so, dtype=np.uint64)
        soff.append(so[1:] - so[0] + np.uint64(sbase))
        qbase += r.nq
        hbase += len(r.hits)
        pbase += r.pool.size
        sbase += int(so[-1] - so[0])
    cat = np.concatenate
    return HuntResult(cat(hits) if hits else np.zeros(0, HIT_DTYPE), cat(qoff), cat(status) if status else np.zeros(0, np.uint32),
                      cat(dist) if dist else np.zeros(0, np.uint32), cat(pool) if pool else np.zeros(0, np.uint8),
                      cat(seqs) if seqs else np.zeros(0, np.uint8), cat(soff), None)


def allgather_bytes(buf: np.ndarray, device=None) -> list[np.ndarray]:
    """All-gather of variable-length byte buffers: sizes first, then buffers padded to the
    largest.  Works on any initialised torch.distributed backend."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    size = torch.tensor([buf.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    mine = torch.zeros(mx, dtype=torch.uint8, device=dev)
    if buf.size:
        mine[:buf.size].copy_(torch.from_numpy(buf))
    out = torch.empty(world * mx, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, mine)
    host = out.cpu().numpy().reshape(world, mx)
    return [host[r, :sizes[r]].copy() for r in range(world)]


class _DeviceBytes:
    """A raw device address dressed up for torch.as_tensor (zero-copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


WIRE_INTS = 4  # int32 words per hit on the wire: query, chr, start, score (low 16 bits) | strand (bits 16-23)


def _wire_records(raw_u8):
    """dg_hit records (uint8 tensor, 48 bytes each) -> the 16-byte wire record the ranks exchange:
    hit coordinates, strand and edit distance.  Alignment strings stay with the owning rank."""
    import torch
    w = raw_u8.view(torch.int32).view(-1, HIT_DTYPE.itemsize // 4)
    query, score, chrom, start = w[:, 0], w[:, 1], w[:, 2], w[:, 3]
    strand = w[:, 10] & 0xFF                      # byte 40 of dg_hit
    return torch.stack((query, chrom, start, (score & 0xFFFF) | (strand << 16)), dim=1).contiguous()


def unwire_records(wire: np.ndarray) -> np.ndarray:
    """The inverse for the host: structured array (query, chr, start, score, strand)."""
    wire = np.ascontiguousarray(wire, dtype=np.int32).reshape(-1, WIRE_INTS)
    out = np.zeros(len(wire), dtype=[("query", "<u4"), ("chr", "<u4"), ("start", "<u4"), ("score", "<i4"), ("strand", "u1")])
    out["query"], out["chr"], out["start"] = wire[:, 0].view(np.uint32), wire[:, 1].view(np.uint32), wire[:, 2].view(np.uint32)
    out["score"] = (wire[:, 3] & 0xFFFF).astype(np.int16).astype(np.int32)
    out["strand"] = ((wire[:, 3] >> 16) & 0xFF).astype(np.uint8)
    return out


def _allgather_wire(mine):
    """counts first, then the wire records padded to the largest count (one NCCL all-gather each)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    n = mine.shape[0]
    cnt = torch.tensor([n], dtype=torch.int64, device=mine.device)
    counts = torch.empty(world, dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(counts, cnt)
    mx = max(int(counts.max().item()), 1)
    padded = torch.zeros((mx, WIRE_INTS), dtype=torch.int32, device=mine.device)
    padded[:n] = mine
    out = torch.empty((world * mx, WIRE_INTS), dtype=torch.int32, device=mine.device)
    dist.all_gather_into_tensor(out, padded)
    return out.view(world, mx, WIRE_INTS), counts


def allgather_hits_device(batch):
    """The exchange step of the path on device memory: the hit records of a batch
    that has run are turned into 16-byte wire records and all-gathered over NCCL / NVLink straight
    from HBM.  Returns (int32 tensor [world, max_count, 4] on the device, int64 counts [world])."""
    import torch
    ptr, n = batch.device_hits()
    isz = HIT_DTYPE.itemsize
    if n:
        mine = _wire_records(torch.as_tensor(_DeviceBytes(ptr, n * isz), device="cuda"))
    else:
        mine = torch.zeros((0, WIRE_INTS), dtype=torch.int32, device="cuda")
    return _allgather_wire(mine)


def allgather_hits_index(index):
    """The exchange after Index.hunt(): the library keeps the wire records of the last call in HBM
    (dg_index_wire_records), so nothing is uploaded again."""
    import torch
    ptr, n = index.wire_records()
    if n:
        mine = torch.as_tensor(_DeviceBytes(ptr, n * 4 * WIRE_INTS), device="cuda").view(torch.int32).view(-1, WIRE_INTS)
    else:
        mine = torch.zeros((0, WIRE_INTS), dtype=torch.int32, device="cuda")
    return _allgather_wire(mine)


def allgather_hits_host(res: HuntResult):
    """The same exchange starting from a host-resident result (dg_hunt_batch): the records go back
    to the device as wire records (16 of their 48 bytes) and are all-gathered there."""
    import torch
    if len(res.hits):
        raw = torch.from_numpy(res.hits.view(np.uint8).reshape(-1)).cuda(non_blocking=True)
        mine = _wire_records(raw)
    else:
        mine = torch.zeros((0, WIRE_INTS), dtype=torch.int32, device="cuda")
    return _allgather_wire(mine)


def hunt_sharded(index, seqs, params, rank: int | None = None, world: int | None = None) -> HuntResult:
    """The whole multi-GPU call: this rank hunts its shard on its own GPU, then the packed hit
    records are all-gathered and merged; every rank returns the global result."""
    import torch.distributed as dist
    from .api import pack_sequences
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    buf, off = pack_sequences(seqs)
    nq = len(off) - 1
    b = shard_bounds(nq, world)
    lo, hi = b[rank], b[rank + 1]
    sub_off = off[lo:hi + 1] - off[lo]
    sub_buf = buf[int(off[lo]):int(off[hi])]
    local = index.hunt((sub_buf, sub_off), params)
    gathered = allgather_bytes(pack_result(local))
    parts = [unpack_result(g) for g in gathered]
    return merge_results(parts, [off[b[r]:b[r + 1] + 1] for r in range(world)])
