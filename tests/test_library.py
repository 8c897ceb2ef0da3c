"""CPU tests of the boundary: the C ABI library loads, exports every symbol the header declares,
fails loudly without a GPU, and its host-only entry points work."""
import os
import re

import numpy as np
import pytest

import dicey_b200.api as api
from dicey_b200 import shard, synth
from util import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dicey_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = api.library()
    syms = header_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert missing == []
    assert sorted(api.EXPORTS) == syms


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.DiceyB200Error) as e:
        api.Index.open(os.path.join(GOLDEN, "t1m.fm9"), 0)
    assert e.value.code == -4 and "no CPU path" in str(e.value)
    with pytest.raises(api.DiceyB200Error):
        api.Index.build_synthetic(42, 2, 1000, 0)


def test_open_rejects_bad_files(tmp_path):
    bad = tmp_path / "x.fm9"
    bad.write_bytes(b"\0" * 100)
    with pytest.raises(api.DiceyB200Error) as e:
        api.Index.open(str(bad), 0)
    assert e.value.code == -2          # no _check sidecar (load_from_checked_file refuses)
    (tmp_path / "x.fm9_check").write_bytes(b"\1" * 8)
    with pytest.raises(api.DiceyB200Error) as e:
        api.Index.open(str(bad), 0)
    assert e.value.code == -3          # wrong type hash


def test_open_checks_the_fm9_framing(tmp_path):
    """The .fm9 parser walks a read-only mapping of the file: a truncated copy of a real index, a copy
    with trailing bytes and an empty file are format errors; the intact copy gets past the parser and
    fails only for want of a GPU (this suite runs without one)."""
    import shutil
    src = os.path.join(GOLDEN, "t1m.fm9")
    data = open(src, "rb").read()
    check = open(src + "_check", "rb").read()

    def try_open(payload):
        f = tmp_path / "g.fm9"
        f.write_bytes(payload)
        (tmp_path / "g.fm9_check").write_bytes(check)
        with pytest.raises(api.DiceyB200Error) as e:
            api.Index.open(str(f), 0)
        return e.value.code, str(e.value)

    for cut in (0, 7, 16, len(data) // 3, len(data) - 2000, len(data) - 1):
        code, msg = try_open(data[:cut])
        assert code == -3 and "malformed .fm9" in msg, (cut, msg)
    code, msg = try_open(data + b"\0")
    assert code == -3 and "trailing bytes" in msg
    import torch
    if not torch.cuda.is_available():
        code, msg = try_open(data)
        assert code == -4 and "no CPU path" in msg   # parsed; no device to put it on


def test_hits_sort_matches_dnahit_order():
    h = np.zeros(5, dtype=api.HIT_DTYPE)
    h["score"] = [-1, 0, -1, 0, -2]
    h["chr"] = [2, 1, 1, 0, 0]
    h["start"] = [5, 9, 7, 3, 1]
    api.library().dg_hits_sort(h.ctypes.data, len(h))
    assert [(int(x["score"]), int(x["chr"]), int(x["start"])) for x in h] == [(0, 0, 3), (0, 1, 9), (-1, 1, 7), (-1, 2, 5), (-2, 0, 1)]


def fake_result(nq, hits_per_q, tag):
    """A packed dg_result built by hand (wire format of dg_result_pack)."""
    nh = nq * hits_per_q
    hits = np.zeros(nh, dtype=api.HIT_DTYPE)
    hits["query"] = np.repeat(np.arange(nq), hits_per_q)
    hits["score"] = -1
    hits["chr"] = tag
    hits["start"] = np.arange(nh) + 1
    hits["aln_len"] = 4
    hits["aln_off"] = np.arange(nh) * 8
    hits["strand"] = ord("+")
    pool = np.frombuffer((b"ACGT" + b"ACGA") * nh, dtype=np.uint8)
    qoff = (np.arange(nq + 1) * hits_per_q).astype(np.uint64)
    seqs = np.frombuffer(b"ACGTACGTAC" * nq, dtype=np.uint8)
    hdr = np.array([nq, nh, pool.size, seqs.size], dtype=np.uint64)
    return np.concatenate([hdr.view(np.uint8), qoff.view(np.uint8), np.zeros(nq, np.uint32).view(np.uint8),
                           np.ones(nq, np.uint32).view(np.uint8), hits.view(np.uint8), pool, seqs])


def test_pack_unpack_roundtrip():
    a = shard.unpack_result(fake_result(3, 2, 7), np.arange(4, dtype=np.uint64) * 10)
    assert a.nq == 3 and len(a.hits) == 6 and a.push_hits(1)[0][4:] == ("ACGT", "ACGA")
    assert np.array_equal(shard.pack_result(a), fake_result(3, 2, 7))
    assert a.sequence(2) == b"ACGTACGTAC"
    assert shard.shard_bounds(10, 4) == [0, 2, 5, 7, 10]


def test_allgather_result_merges_in_rank_order():
    """dg_allgather_result over the host transport (dg_comm_init_host), three ranks played by three
    threads of this process that meet in a barrier-style all-gather."""
    import threading
    world = 3
    lock, cond = threading.Lock(), threading.Condition()
    state = {"round": 0, "parts": {}, "done": {}}

    def make_allgather(rank):
        def allgather(send):
            with cond:
                rnd = state["round"]
                state["parts"][rank] = bytes(send)
                if len(state["parts"]) == world:
                    state["done"][rnd] = b"".join(state["parts"][r] for r in range(world))
                    state["parts"] = {}
                    state["round"] += 1
                    cond.notify_all()
                else:
                    cond.wait_for(lambda: rnd in state["done"], timeout=30)
                return np.frombuffer(state["done"][rnd], dtype=np.uint8)
        return allgather

    out = {}

    def run(rank):
        comm = api.Comm.init_host(world, rank, make_allgather(rank))
        mine = shard.unpack_result(fake_result(rank + 1, rank + 1, 20 + rank))
        out[rank] = comm.allgather_result(mine, np.arange(7, dtype=np.uint64) * 10)
        assert (comm.rank, comm.nranks) == (rank, world)
        comm.close()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(60)
    for r in range(world):
        m = out[r]
        assert m.nq == 6 and len(m.hits) == 1 + 4 + 9
        assert [int(x) for x in m.qoff] == [0, 1, 3, 5, 8, 11, 14]
        assert [int(x) for x in m.hits["chr"]] == [20] + [21] * 4 + [22] * 9
        assert [int(x) for x in m.hits["query"]] == [0, 1, 1, 2, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5]
        assert m.push_hits(5)[2] == (-1, 22, 9, "+", "ACGT", "ACGA")
        assert m.sequence(4) == b"ACGTACGTAC"


def _ops(*ops):
    """(column, kind, genomic byte) triples -> the ops word of a dg_rec (include/dicey_b200.h)."""
    w = 0
    for k, (col, kind, ref) in enumerate(sorted(ops)):
        w |= (col | (kind << 10) | (ref << 12)) << (20 * k)
    return w


def test_compact_record_alignment_host_side():
    """dg_rec_alignment (host code, no device): both alignment rows rebuilt from the query and the
    record's edit operations -- mismatch, gap in the reference row, gap in the query row."""
    import ctypes as C
    lib = api.library()
    rec = np.zeros(1, dtype=api.REC_DTYPE)
    ra, qa = C.create_string_buffer(64), C.create_string_buffer(64)

    def rows(query, ops):
        rec["ops"] = _ops(*ops)
        rec["nops"] = len(ops)
        n = lib.dg_rec_alignment(rec.ctypes.data, query, len(query), ra, qa)
        assert n >= 0
        return ra.raw[:n].decode(), qa.raw[:n].decode()

    q = b"ACGTACGTAC"
    assert rows(q, []) == ("ACGTACGTAC", "ACGTACGTAC")
    assert rows(q, [(3, 1, ord("G"))]) == ("ACGGACGTAC", "ACGTACGTAC")                    # mismatch at column 3
    assert rows(q, [(4, 2, 0)]) == ("ACGT-CGTAC", "ACGTACGTAC")                           # '-' in refalign
    assert rows(q, [(4, 3, ord("T"))]) == ("ACGTTACGTAC", "ACGT-ACGTAC")                  # '-' in queryalign
    assert rows(q, [(0, 1, ord("N")), (5, 3, ord("A")), (10, 1, ord("G"))]) == ("NCGTAACGTAG", "ACGTA-CGTAC")
    rec["nops"] = 4                                                                      # more than the form carries
    assert lib.dg_rec_alignment(rec.ctypes.data, q, len(q), ra, qa) < 0


def test_record_sort_orders_like_hit_sort():
    """dg_recs_sort and dg_hits_sort (std::sort of hunter.h:440 on the two record forms) give the same
    permutation, ties included (same comparator, same introsort, same sequence of keys)."""
    rng = np.random.default_rng(3)
    lib = api.library()
    for n in (0, 1, 2, 17, 400, 5000):
        hits = np.zeros(n, dtype=api.HIT_DTYPE)
        recs = np.zeros(n, dtype=api.REC_DTYPE)
        hits["score"] = recs["score"] = -rng.integers(0, 3, n)
        hits["chr"] = recs["chr"] = rng.integers(0, 4, n)
        hits["start"] = recs["start"] = rng.integers(1, 30, n)          # many ties
        hits["query"] = recs["query"] = np.arange(n)                    # identifies the element
        if n:
            lib.dg_hits_sort(hits.ctypes.data, n)
            lib.dg_recs_sort(recs.ctypes.data, n)
        assert np.array_equal(hits["query"], recs["query"])
        key = list(zip((-hits["score"]).tolist(), hits["chr"].tolist(), hits["start"].tolist()))
        assert key == sorted(key)


def test_synth_generator_is_windowable():
    t = synth.text(42, 3, 1000)
    assert t.size == 3 * 1001 and t[1000] == 10
    assert np.array_equal(synth.bases(42, 1000 + 17, 50), t[1001 + 17:1001 + 67])
    p = synth.primers_fast(42, 3, 1000, 8, 20, 1, True)
    assert p.shape == (8, 20) and set(np.unique(p)) <= set(b"ACGT")
