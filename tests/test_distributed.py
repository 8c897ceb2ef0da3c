"""world_size-4 (and 2) gloo tests of the multi-GPU entry points of the C ABI (dg_comm_init_host,
dg_allgather_result) wired to a torch.distributed process group by dicey_b200/shard.py: contiguous
primer shards, one all-gather of the packed results, merge in rank order."""
import os
import sys

import numpy as np
import pytest
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
    nq_total = 4 * world - 1
    b = shard.shard_bounds(nq_total, world)
    comm = shard.comm_from_process_group()          # gloo group -> host transport (dg_comm_init_host)
    assert (comm.rank, comm.nranks) == (rank, world)
    mine = shard.unpack_result(fake_result(b[rank + 1] - b[rank], rank + 1, 10 + rank))
    m = comm.allgather_result(mine, np.arange(nq_total + 1, dtype=np.uint64) * 10)
    # a second exchange on the same communicator, with one empty shard
    empty = shard.unpack_result(fake_result(0 if rank == 1 else 2, 1, 50 + rank))
    m2 = comm.allgather_result(empty)
    q.put((rank, m.nq, len(m.hits), [int(x) for x in m.qoff], [int(x) for x in m.hits["chr"]],
           [int(x) for x in m.hits["query"]], m.push_hits(m.nq - 1)[-1], m.sequence(m.nq - 1),
           m2.nq, [int(x) for x in m2.hits["chr"]]))
    comm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_allgather_of_hit_records(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nq_total = 4 * world - 1
    b = [(nq_total * r) // world for r in range(world + 1)]
    per_rank = [b[r + 1] - b[r] for r in range(world)]
    want_qoff = [0]
    want_chr, want_q = [], []
    for r in range(world):
        for k in range(per_rank[r]):
            want_qoff.append(want_qoff[-1] + r + 1)
            want_q += [b[r] + k] * (r + 1)
        want_chr += [10 + r] * (per_rank[r] * (r + 1))
    # every rank sees the same merge
    for rank, nq, nh, qoff, chrs, qs, last, lastseq, nq2, chr2 in out:
        assert nq == nq_total and nh == len(want_chr)
        assert qoff == want_qoff
        assert chrs == want_chr
        assert qs == want_q
        assert last == (-1, 10 + world - 1, per_rank[-1] * world, "+", "ACGT", "ACGA")
        assert lastseq == b"ACGTACGTAC"
        assert nq2 == 2 * (world - 1)
        assert chr2 == [c for r in range(world) if r != 1 for c in [50 + r] * 2]
