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


def allgather_hits_device(batch):
    """The exchange step of the path on device memory: the dg_hit records of a batch
    that has run are all-gathered over NCCL / NVLink straight from HBM -- counts first, then the
    records padded to the largest count.  Returns (uint8 tensor [world, max_count * 48] on the
    device, int64 counts [world] on the device)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    ptr, n = batch.device_hits()
    isz = HIT_DTYPE.itemsize
    cnt = torch.tensor([n], dtype=torch.int64, device="cuda")
    counts = torch.empty(world, dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(counts, cnt)
    mx = max(int(counts.max().item()), 1)
    mine = torch.zeros(mx * isz, dtype=torch.uint8, device="cuda")
    if n:
        mine[:n * isz].copy_(torch.as_tensor(_DeviceBytes(ptr, n * isz), device="cuda"))
    out = torch.empty(world * mx * isz, dtype=torch.uint8, device="cuda")
    dist.all_gather_into_tensor(out, mine)
    return out.view(world, mx * isz), counts


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
