"""world_size-2 gloo test of the multi-GPU plumbing (dicey_b200/shard.py): contiguous primer
shards, one all-gather of the packed hit records, merge in rank order."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dicey_b200 import shard
    from test_library import fake_result
    nq_total = 7
    b = shard.shard_bounds(nq_total, world)
    mine = shard.unpack_result(fake_result(b[rank + 1] - b[rank], rank + 1, 10 + rank))
    gathered = shard.allgather_bytes(shard.pack_result(mine))
    parts = [shard.unpack_result(g) for g in gathered]
    offs = [np.arange(b[r], b[r + 1] + 1, dtype=np.uint64) * 10 for r in range(world)]
    m = shard.merge_results(parts, offs)
    # the 16-byte wire records of the device-side exchange (counts, then padded records)
    import torch
    wire = shard._wire_records(torch.from_numpy(mine.hits.view(np.uint8).reshape(-1).copy()))
    allw, counts = shard._allgather_wire(wire)
    got = [shard.unwire_records(allw[r, :int(counts[r])].numpy()) for r in range(world)]
    wire_ok = all(len(got[r]) == (b[r + 1] - b[r]) * (r + 1) and (got[r]["chr"] == 10 + r).all() and (got[r]["score"] == -1).all()
                  and (got[r]["strand"] == ord("+")).all() for r in range(world))
    q.put((rank, m.nq, len(m.hits), [int(x) for x in m.qoff], [int(x) for x in m.hits["chr"]],
           [int(x) for x in m.hits["query"]], wire_ok))
    dist.destroy_process_group()


def test_gloo_allgather_of_hit_records():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # rank 0 owns queries 0..2 (1 hit each), rank 1 owns 3..6 (2 hits each); both ranks see the same merge
    for rank, nq, nh, qoff, chrs, qs, wire_ok in out:
        assert wire_ok
        assert nq == 7 and nh == 3 * 1 + 4 * 2
        assert qoff == [0, 1, 2, 3, 5, 7, 9, 11]
        assert chrs == [10] * 3 + [11] * 8
        assert qs == [0, 1, 2, 3, 3, 4, 4, 5, 5, 6, 6]
