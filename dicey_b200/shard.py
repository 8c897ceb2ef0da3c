"""Multi-GPU plumbing of the hot path: the index is replicated on every GPU, the
query batch is split into contiguous equal shards by rank (so concatenating the per-rank results
in rank order restores input order), every rank runs independently, and ONE exchange step
collects the hit records.  The exchange itself lives behind the C ABI (``dg_comm_*``,
``dg_allgather_hits``, ``dg_allgather_result`` in include/dicey_b200.h): NCCL on the index stream
on the GPU box, a host all-gather callback (gloo) in the CPU tests.  This module only wires a
``torch.distributed`` process group to it: the NCCL unique id travels by ``broadcast``, and the
host transport calls ``all_gather_into_tensor``.
"""
from __future__ import annotations

import numpy as np

from .api import Comm, HuntResult, pack_result, pack_sequences, unpack_result  # noqa: F401  (re-exported)


def shard_bounds(n: int, world: int) -> list[int]:
    """Contiguous, equal-count shards: rank r owns queries [b[r], b[r+1])."""
    return [(n * r) // world for r in range(world + 1)]


def comm_from_process_group(index=None) -> Comm:
    """A dg_comm for the default torch.distributed group: NCCL when the group is NCCL and an index
    (bound to this rank's GPU) is given, otherwise the host transport over the group."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    if index is not None and dist.get_backend() == "nccl":
        uid = torch.zeros(Comm.ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        return Comm.init(world, rank, bytes(uid.cpu().numpy()), index)

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"

    def allgather(send: np.ndarray) -> np.ndarray:
        mine = torch.from_numpy(send.copy()).to(dev)
        out = torch.empty(world * send.size, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, mine)
        return out.cpu().numpy()

    return Comm.init_host(world, rank, allgather)


def global_offsets(off: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(off, dtype=np.uint64)


def hunt_sharded(index, seqs, params, comm: Comm) -> HuntResult:
    """The whole multi-GPU call: this rank hunts its shard on its own GPU, then the complete results
    are all-gathered (dg_allgather_result) and every rank returns the result of the whole batch."""
    buf, off = pack_sequences(seqs)
    nq = len(off) - 1
    b = shard_bounds(nq, comm.nranks)
    lo, hi = b[comm.rank], b[comm.rank + 1]
    sub_off = off[lo:hi + 1] - off[lo]
    sub_buf = buf[int(off[lo]):int(off[hi])]
    local = index.hunt((sub_buf, sub_off), params)
    return comm.allgather_result(local, global_offsets(off))
