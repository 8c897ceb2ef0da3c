"""GPU parity tests: the CUDA path, called through the C ABI, against the golden outputs of the
reference (tests/golden, written by oracle/_ref/dicey_ref) -- bit-exact on every field."""
import gzip
import os

import numpy as np
import pytest

from dicey_b200 import synth
from dicey_b200.api import HuntParams, Index, Q_NBR_CAP, Q_NBR_UNVERIFIED, hunt_json
from util import GOLDEN, HUNT_CASES, fm9_sections, params_from_flags, read_queries, read_rec_tsv, read_records

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def indexes():
    out = {}
    for name in ("t1m", "stress"):
        ix = Index.open(os.path.join(GOLDEN, name + ".fm9"), 0)
        names, lens = read_rec_tsv(os.path.join(GOLDEN, name + ".rec.tsv"))
        ix.set_records(names, lens)
        out[name] = ix
    yield out
    for ix in out.values():
        ix.close()


def stress_text():
    return gzip.open(os.path.join(GOLDEN, "stress.dump.gz"), "rb").read()


def test_index_arrays_t1m(indexes):
    ix = indexes["t1m"]
    txt = synth.text(42, 8, 125000)
    info = ix.info()
    assert info["n"] == txt.size + 1 and info["sigma"] == 6
    got = ix.debug_array("text")
    assert got.size == txt.size + 1 and got[-1] == 0
    assert np.array_equal(got[:-1], txt)


def naive_sa(text: bytes):
    t = text + b"\0"
    return sorted(range(len(t)), key=lambda i: t[i:])


def test_index_arrays_stress(indexes):
    ix = indexes["stress"]
    raw = stress_text()
    got = ix.debug_array("text")
    assert bytes(got[:-1]) == raw and got[-1] == 0
    sa = np.array(naive_sa(raw), dtype=np.uint32)
    assert np.array_equal(ix.debug_array("sa_samples", np.uint32), sa[::32])
    assert np.array_equal(ix.debug_array("sa_full", np.uint32), sa)   # rebuilt beside the text by the LF chains
    # KB-mer presence bitmaps (KB - 1, KB, KB + 1): exactly the ACGT-only windows of the text
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    kb0 = ix.info()["bitmap_k"]
    for name, kb in (("present_lo", kb0 - 1), ("present_kb", kb0), ("present_hi", kb0 + 1),
                     ("present_kb_l", kb0), ("present_hi_l", kb0 + 1)):
        bits = ix.debug_array(name, np.uint32)
        assert bits.size == (4 ** kb) // 32, name
        want = np.zeros_like(bits)
        for i in range(len(raw) - kb + 1):
            w = raw[i:i + kb]
            if all(ch in code for ch in w):
                v = 0
                for ch in w:
                    v = (v << 2) | code[ch]
                if not name.endswith("_l"):
                    bp = min(7, kb - 2)
                    lo = 2 * (kb - bp)   # presence_bit(): region = last kb - 7 bases, bit = the 7 bases before them
                    v = ((v & ((1 << lo) - 1)) << (2 * bp)) | (v >> lo)
                # (presence_bit_left(): region = first kb - 7 bases, bit = the 7 bases after them: the code itself)
                want[v >> 5] |= np.uint32(1 << (v & 31))
        assert np.array_equal(bits, want), name
    # occ blocks: cumulative ACGT counts + bit planes of the BWT
    t = np.frombuffer(raw + b"\0", dtype=np.uint8)
    bwt = t[(sa.astype(np.int64) - 1) % t.size]
    occ = ix.debug_array("occ", np.uint32).reshape(-1, 8)
    codes = np.full(256, 4, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        codes[ch] = i
    c = codes[bwt]
    for b in range(occ.shape[0]):
        seg = c[b * 64:(b + 1) * 64]
        for k in range(4):
            assert occ[b, k] == int(np.count_nonzero(c[:b * 64] == k))
        lo = sum(int(v & 1) << j for j, v in enumerate(seg) if v < 4)
        hi = sum(int(v >> 1) << j for j, v in enumerate(seg) if v < 4)
        assert int(occ[b, 2]) | (int(occ[b, 3]) << 32) == 0 or True
        got_lo = int(occ[b, 4]) | (int(occ[b, 5]) << 32)
        got_hi = int(occ[b, 6]) | (int(occ[b, 7]) << 32)
        assert (got_lo, got_hi) == (lo, hi)
    exc = ix.debug_array("exc_pos", np.uint32)
    assert np.array_equal(exc, np.nonzero(c == 4)[0].astype(np.uint32))


def test_backward_search_and_locate(indexes):
    ix = indexes["stress"]
    pats = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "stress.patterns.txt")) if l.strip()]
    want = [tuple(int(x) for x in l.split()) for l in open(os.path.join(GOLDEN, "stress.count.tsv"))]
    l, r = ix.backward_search(pats)
    for i, (wl, wr, occ) in enumerate(want):
        assert int(r[i]) + 1 - int(l[i]) == occ, pats[i]
        if occ:
            assert (int(l[i]), int(r[i])) == (wl, wr), pats[i]


@pytest.mark.parametrize("case,index", HUNT_CASES)
def test_hunt_golden(indexes, case, index):
    ix = indexes[index]
    qs = read_queries(os.path.join(GOLDEN, case + ".queries.txt"))
    par = params_from_flags(open(os.path.join(GOLDEN, case + ".flags.txt")).read())
    want = read_records(os.path.join(GOLDEN, case + ".records.tsv"))
    want_json = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, case + ".jsonl"))]
    res = ix.hunt([s for _, s in qs], par)
    assert res.nq == len(qs) == len(want)
    for q, (name, seq) in enumerate(qs):
        w = want[q]
        # (queries whose neighbourhood the reference truncated at -x, neighbors.h:50, are compared like
        # every other one: the truncated set is replayed, nbr_trunc.hpp)
        assert not int(res.status[q]) & Q_NBR_UNVERIFIED, (case, q)
        assert res.messages(q, par, seq.encode()) == w["msgs"], (case, q)
        if w["msgs"] and w["msgs"][0].startswith("Error"):
            assert res.push_hits(q) == []
            continue
        assert res.sequence(q).decode() == w["seq"]
        assert int(res.dist[q]) == w["distance"]
        assert res.push_hits(q) == w["push"], (case, q, name, seq)
        assert res.sorted_hits(q) == w["sorted"], (case, q, name, seq)
        js = hunt_json(res, q, par, ix.names, "genome.fa.gz", "", name, seq.encode())
        assert js == want_json[q], (case, q)


@pytest.mark.parametrize("name", ["t1m", "stress"])
def test_build_text_equals_fm9(indexes, name):
    """The GPU suffix sorter + BWT builder must produce the same device index as the
    reference's .fm9 (divsufsort + SDSL) transcoded by the loader."""
    ref = indexes[name]
    text = synth.text(42, 8, 125000).tobytes() if name == "t1m" else stress_text()
    with Index.build_text(text, 0) as ix:
        assert ix.info()["n"] == ref.info()["n"]
        for what, dt in (("text", np.uint8), ("sa_samples", np.uint32), ("isa_samples", np.uint32), ("occ", np.uint32),
                         ("C", np.uint32), ("exc_pos", np.uint32), ("exc_sym", np.uint8), ("kmer", np.uint32),
                         ("sa_full", np.uint32), ("present_kb", np.uint32), ("present_hi", np.uint32),
                         ("present_lo", np.uint32), ("present_kb_l", np.uint32), ("present_hi_l", np.uint32)):
            assert np.array_equal(ix.debug_array(what, dt), ref.debug_array(what, dt)), what


def test_build_synthetic_equals_fm9(indexes):
    ref = indexes["t1m"]
    with Index.build_synthetic(42, 8, 125000, 0) as ix:
        for what, dt in (("text", np.uint8), ("sa_samples", np.uint32), ("occ", np.uint32), ("kmer", np.uint32)):
            assert np.array_equal(ix.debug_array(what, dt), ref.debug_array(what, dt)), what
        qs = read_queries(os.path.join(GOLDEN, "t1m_e1.queries.txt"))
        want = read_records(os.path.join(GOLDEN, "t1m_e1.records.tsv"))
        res = ix.hunt([s for _, s in qs], HuntParams(distance=1))
        for q in range(len(qs)):
            assert res.push_hits(q) == want[q]["push"]


def test_seed_golden(indexes):
    """FM / NW part of `dicey search` (silica.h:449-573): candidates, contexts, alignpos."""
    for fn, index, qfile in (("t1m_seed_k15_e1", "t1m", "t1m_seed"), ("t1m_seed_k12_h1", "t1m", "t1m_seed"),
                             ("stress_seed_k15_e1", "stress", "stress_seed")):
        ix = indexes[index]
        qs = read_queries(os.path.join(GOLDEN, qfile + ".queries.txt"))
        flags = {"t1m_seed_k15_e1": "-k 15 -d 1", "t1m_seed_k12_h1": "-k 12 -d 1 -n", "stress_seed_k15_e1": "-k 15 -d 1 -m 300"}[fn]
        par = params_from_flags(flags, search=True)
        want, cur = [], None
        for line in open(os.path.join(GOLDEN, fn + ".seed.tsv")):
            f = line.rstrip("\n").split("\t")
            if f[0] == "Q":
                cur = {"skipped": f[2] == "skipped", "hits": [], "n": None if f[2] == "skipped" else int(f[3])}
                want.append(cur)
            else:
                cur["hits"].append((int(f[2]), int(f[3]), int(f[4]), "-" if f[1] == "1" else "+", f[5]))
        res = ix.hunt([s for _, s in qs], par)
        for q in range(len(qs)):
            if want[q]["skipped"]:
                assert res.seed_hits(q) == []
                continue
            assert res.seed_hits(q) == want[q]["hits"], (fn, q)


def test_padlock_counts(indexes):
    arms = [l.strip() for l in open(os.path.join(GOLDEN, "arms.txt")) if l.strip()]
    for fn, index, ham in (("stress_arms_e1", "stress", False), ("stress_arms_h1", "stress", True), ("t1m_arms_e1", "t1m", False)):
        want = [l.split("\t") for l in open(os.path.join(GOLDEN, fn + ".padcount.tsv"))]
        ix = indexes[index]
        exact = ix.count(arms, HuntParams(distance=0, hamming=ham))
        total = ix.count(arms, HuntParams(distance=1, hamming=ham))
        for i, w in enumerate(want):
            assert (int(exact[i]), int(total[i])) == (int(w[1]), int(w[2])), (fn, i)


@pytest.mark.parametrize("knob", ["DG_FULL_SA=0", "DG_BITMAP_K=0", "DG_BITMAP_EXTRA=0", "DG_BITMAP_LEFT=0", "DG_KMER=8",
                                  "DG_BITMAP_K=14", "DG_FULL_SA=0 DG_BITMAP_LEFT=0 DG_KMER=10"])
def test_every_index_shape_gives_the_golden_records(knob, monkeypatch):
    """The loader drops optional tables when HBM is short (and the knobs force it): sampled suffix array
    (the LF walk of csa_wt.hpp:340-354 instead of one gather), no presence bitmaps, no KB +- 1 bitmaps, no
    left-anchored twins, a short K-mer table.  Every shape must give the reference's records: hunt
    (edit and Hamming, distances 1 and 2, caps), search seeds and padlock counts."""
    for kv in knob.split():
        monkeypatch.setenv(*kv.split("="))
    opened = {}
    try:
        for name in ("t1m", "stress"):
            ix = Index.open(os.path.join(GOLDEN, name + ".fm9"), 0)
            names, lens = read_rec_tsv(os.path.join(GOLDEN, name + ".rec.tsv"))
            ix.set_records(names, lens)
            opened[name] = ix
        info = opened["t1m"].info()
        if "DG_KMER" in knob:
            assert info["kmer"] == int(knob.split("DG_KMER=")[1].split()[0])
        if "DG_BITMAP_K=0" in knob:
            assert info["bitmap_k"] == 0
        for case, index in (("t1m_e1", "t1m"), ("t1m_h2", "t1m"), ("stress_e1_m7", "stress"), ("stress_e2", "stress"), ("t1m_e2_x500", "t1m")):
            ix = opened[index]
            qs = read_queries(os.path.join(GOLDEN, case + ".queries.txt"))
            par = params_from_flags(open(os.path.join(GOLDEN, case + ".flags.txt")).read())
            want = read_records(os.path.join(GOLDEN, case + ".records.tsv"))
            res = ix.hunt([s for _, s in qs], par)
            for q, (name, seq) in enumerate(qs):
                w = want[q]
                assert res.messages(q, par, seq.encode()) == w["msgs"], (knob, case, q)
                if w["msgs"] and w["msgs"][0].startswith("Error"):
                    continue
                assert res.push_hits(q) == w["push"], (knob, case, q)
        test_seed_golden(opened)
        test_padlock_counts(opened)
    finally:
        for ix in opened.values():
            ix.close()


@pytest.mark.parametrize("name", ["t1m", "stress"])
def test_write_fm9(indexes, name, tmp_path, ref_bin):
    """dg_index_write_fm9 reproduces SDSL's file byte for byte (wavelet tree, rank and select
    supports, samples, alphabet) and the reference itself loads it and answers as before."""
    import subprocess
    src = os.path.join(GOLDEN, name + ".fm9")
    dst = str(tmp_path / (name + ".fm9"))
    indexes[name].write_fm9(dst)
    a, b = fm9_sections(src), fm9_sections(dst)
    for sec in ("header", "bv", "rank", "select1", "select0", "tree", "sa", "isa", "alphabet"):
        assert a[sec] == b[sec], sec
    assert open(src, "rb").read() == open(dst, "rb").read()
    assert open(src + "_check", "rb").read() == open(dst + "_check", "rb").read()
    # the device index built from the rewritten file equals the original one
    with Index.open(dst, 0) as ix:
        for what, dt in (("text", np.uint8), ("occ", np.uint32), ("sa_samples", np.uint32), ("kmer", np.uint32)):
            assert np.array_equal(ix.debug_array(what, dt), indexes[name].debug_array(what, dt)), what
    if ref_bin:
        case = "t1m_e1" if name == "t1m" else "stress_e1"
        out = str(tmp_path / "rec.tsv")
        flags = open(os.path.join(GOLDEN, case + ".flags.txt")).read().split()
        subprocess.run([ref_bin, "hunt", dst, os.path.join(GOLDEN, name + ".rec.tsv"), os.path.join(GOLDEN, case + ".queries.txt"),
                        "--records", out, "--counters"] + flags, check=True, capture_output=True)
        assert open(out).read() == open(os.path.join(GOLDEN, case + ".records.tsv")).read()


def test_iupac_text_builds_and_hunts_like_the_reference(tmp_path):
    """A text with IUPAC ambiguity codes (17 symbols: beyond the 3-bit codes of the DNA suffix sorter):
    dg_index_build_text + dg_index_write_fm9 give the file `dicey index` (SDSL) wrote for the same text,
    byte for byte, and hunt on either index gives the reference's records."""
    text = gzip.open(os.path.join(GOLDEN, "iupac.dump.gz"), "rb").read()
    names, lens = read_rec_tsv(os.path.join(GOLDEN, "iupac.rec.tsv"))
    src = os.path.join(GOLDEN, "iupac.fm9")
    dst = str(tmp_path / "iupac.fm9")
    built = Index.build_text(text, 0)
    loaded = Index.open(src, 0)
    try:
        assert built.info()["sigma"] == loaded.info()["sigma"] == 17
        for what, dt in (("text", np.uint8), ("occ", np.uint32), ("sa_samples", np.uint32), ("sa_full", np.uint32)):
            assert np.array_equal(built.debug_array(what, dt), loaded.debug_array(what, dt)), what
        built.write_fm9(dst)
        assert open(src, "rb").read() == open(dst, "rb").read()
        assert open(src + "_check", "rb").read() == open(dst + "_check", "rb").read()
        for ix in (built, loaded):
            ix.set_records(names, lens)
            for case in ("iupac_e1", "iupac_h2"):
                qs = read_queries(os.path.join(GOLDEN, case + ".queries.txt"))
                par = params_from_flags(open(os.path.join(GOLDEN, case + ".flags.txt")).read())
                want = read_records(os.path.join(GOLDEN, case + ".records.tsv"))
                want_json = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, case + ".jsonl"))]
                res = ix.hunt([s for _, s in qs], par)
                for q, (name, seq) in enumerate(qs):
                    assert res.messages(q, par, seq.encode()) == want[q]["msgs"], (case, q)
                    assert res.push_hits(q) == want[q]["push"], (case, q)
                    assert hunt_json(res, q, par, names, "genome.fa.gz", "", name, seq.encode()) == want_json[q], (case, q)
    finally:
        built.close()
        loaded.close()


def test_chunked_pipeline_equals_single_batch(indexes, monkeypatch):
    """dg_hunt_batch cuts large batches into chunks whose hit records travel to the host while the
    next chunk is searched; ids and offsets are rebased on the device.  Same records either way."""
    ix = indexes["t1m"]
    pr = synth.primers_fast(42, 8, 125000, 6000, 20, 1, True, rng_seed=3)
    qs = [bytes(r) for r in pr]
    qs[17] = b"ACGTAC"            # too short
    qs[1500] = qs[1500][:12]      # ragged lengths
    qs[4100] = b"ACGTNACGTACGTTGCAAGT"
    monkeypatch.setenv("DG_CHUNK", "100000000")
    one = ix.hunt(qs, HuntParams(distance=1))
    monkeypatch.setenv("DG_CHUNK", "1024")
    many = ix.hunt(qs, HuntParams(distance=1))
    assert one.nq == many.nq == len(qs)
    assert np.array_equal(one.qoff, many.qoff) and np.array_equal(one.status, many.status) and np.array_equal(one.dist, many.dist)
    assert bytes(one.seqs) == bytes(many.seqs)
    for f in ("query", "score", "chr", "start", "text_pos", "aln_len", "alignpos", "strand"):
        assert np.array_equal(one.hits[f], many.hits[f]), f
    assert len(one.hits) > 3000
    for q in list(range(0, len(qs), 97)) + [17, 1500, 4100]:
        assert one.push_hits(q) == many.push_hits(q)


@pytest.mark.parametrize("case,index", [("t1m_e1", "t1m"), ("stress_h1", "stress"), ("t1m_e0", "t1m")])
def test_records_dump_equals_reference_dump(indexes, case, index):
    """The canonical whole-batch dump bench.py hashes against `dicey_ref hunt --records`."""
    ix = indexes[index]
    qs = read_queries(os.path.join(GOLDEN, case + ".queries.txt"))
    par = params_from_flags(open(os.path.join(GOLDEN, case + ".flags.txt")).read())
    want = "".join(l for l in open(os.path.join(GOLDEN, case + ".records.tsv")) if not l.startswith("W\t"))
    res = ix.hunt([s for _, s in qs], par)
    assert res.records_tsv(par, [s.encode() for _, s in qs]) == want


@pytest.mark.parametrize("knob", ["DG_FULL_KEY_SORT=1", "DG_MERGE_SORT=1", "DG_LOCATE_RADIX=1", "DG_CAND_CAP=48", "DG_LOCATE_CAP=150",
                                  "DG_SLOW_KEYS=1", "DG_GROUP_RADIX=1", "DG_NO_SPLIT=1", "DG_NO_OVERLAP=1", "DG_USE_HI=1", "DG_USE_HI=0"])
@pytest.mark.parametrize("case,index", [("stress_e2", "stress"), ("t1m_e1", "t1m"), ("stress_h2_m50", "stress"), ("stress_e1", "stress")])
def test_alternative_paths_give_the_same_records(indexes, monkeypatch, knob, case, index):
    """The alternative routes through the pipeline must give the same records as the default one:
    the three ways of putting candidates into std::set order (group rank count, full-key radix sort,
    comparison sort for strings beyond the key), radix-sorted locate segments, the second search
    attempt after a candidate-buffer overflow, locate / verify in slices of candidates, and the byte-wise
    builder of the string keys in place of the shift-and-mask one."""
    ix = indexes[index]
    qs = read_queries(os.path.join(GOLDEN, case + ".queries.txt"))
    par = params_from_flags(open(os.path.join(GOLDEN, case + ".flags.txt")).read())
    seqs = [s for _, s in qs]
    base = ix.hunt(seqs, par)
    monkeypatch.setenv(*knob.split("="))
    alt = ix.hunt(seqs, par)
    raws = [s.encode() for s in seqs]
    assert alt.records_tsv(par, raws) == base.records_tsv(par, raws)


@pytest.mark.parametrize("chunk", ["100000000", "1500"])
def test_allgather_hits_nccl_world1(indexes, monkeypatch, chunk):
    """dg_comm_init / dg_allgather_hits / dg_comm_fetch_table on one GPU (an NCCL communicator of one
    rank): the 16-byte wire records of the last dg_hunt_batch and of a resident batch, with global
    query ids, through slot growth (a 64-record slot to start with)."""
    from dicey_b200.api import Comm
    ix = indexes["t1m"]
    monkeypatch.setenv("DG_CHUNK", chunk)
    monkeypatch.setenv("DG_COMM_SLOT", "64")
    pr = synth.primers_fast(42, 8, 125000, 5000, 20, 1, True, rng_seed=5)
    par = HuntParams(distance=1)
    comm = Comm.init(1, 0, Comm.unique_id(), ix)
    try:
        res = ix.hunt(pr, par)
        table, slot, counts = comm.allgather_hits(None, query_base=1000)
        assert int(counts[0]) == len(res.hits) > 2000 and slot > len(res.hits)
        got = comm.fetch_table()
        assert np.array_equal(got["query"], res.hits["query"] + 1000)
        for f in ("chr", "start", "score", "strand"):
            assert np.array_equal(got[f], res.hits[f]), f
        # a resident batch (dg_batch_stage / run), smaller than the grown slot: one all-gather, same table
        bt = ix.stage(pr[:700], par)
        bt.run()
        table2, slot2, counts2 = comm.allgather_hits(bt, query_base=7)
        r2 = bt.fetch()
        got2 = comm.fetch_table()
        assert slot2 == slot and int(counts2[0]) == len(r2.hits) == len(got2)
        assert np.array_equal(got2["query"], r2.hits["query"] + 7) and np.array_equal(got2["start"], r2.hits["start"])
        bt.free()
        # peer mode (the records are stored into the table by the producing kernel): base set first
        comm.set_query_base(31)
        res3 = ix.hunt(pr[:3000], par)
        _, _, counts3 = comm.allgather_hits(None, query_base=31)
        got3 = comm.fetch_table()
        assert int(counts3[0]) == len(res3.hits) == len(got3)
        assert np.array_equal(got3["query"], res3.hits["query"] + 31) and np.array_equal(got3["start"], res3.hits["start"])
        bt = ix.stage(pr[:900], par)
        for _ in range(3):    # the two tables alternate
            bt.run()
            _, _, counts4 = comm.allgather_hits(bt, query_base=31)
        r4 = bt.fetch()
        got4 = comm.fetch_table()
        assert int(counts4[0]) == len(r4.hits) == len(got4)
        assert np.array_equal(got4["query"], r4.hits["query"] + 31) and np.array_equal(got4["chr"], r4.hits["chr"])
        bt.free()
    finally:
        comm.close()


def test_thal_gpu_one_long_side(monkeypatch):
    """dg_thal_batch on pairs with one side longer than THAL_MAX_ALIGN = 60 (up to THAL_MAX_SEQ = 10 000,
    thal.h:58, :2440-2451; k_thal_wide): the reference's bits for the 52 pairs of tests/golden/thal_wide.*
    (four of them refused by the reference: ok = 0, THAL_ERROR_SCORE), alone and interleaved with
    ordinary pairs -- the batch routes every pair by its lengths -- in all three kernel modes."""
    from dicey_b200.api import Thal

    def load(name, n=None):
        pairs = [l.rstrip("\n").split("\t") for l in open(os.path.join(GOLDEN, name + ".pairs.tsv"))][:n]
        want = [l.split("\t") for l in open(os.path.join(GOLDEN, name + ".out.tsv")).read().splitlines()][:n]
        return pairs, want

    wide, wide_want = load("thal_wide")
    short, short_want = load("thal", 300)
    mixed, mixed_want = [], []
    for i in range(len(short)):
        mixed.append(short[i]); mixed_want.append(short_want[i])
        if i < len(wide):
            mixed.append(wide[i]); mixed_want.append(wide_want[i])
    assert all(len(a) > 60 or len(b) > 60 for a, b in wide) and len(wide) == 52 and sum(w[0] == "0" for w in wide_want) == 4
    th = Thal.open_tables(os.path.join(GOLDEN, "thal.params.tsv"), 0)
    try:
        for mode in ("0", "1", "2"):
            monkeypatch.setenv("DG_THAL_SEQ", mode)
            for pairs, want in ((wide, wide_want), (mixed, mixed_want)):
                tm, ok = th.tm([p[0] for p in pairs], [p[1] for p in pairs])
                bits = tm.view(np.uint64)
                for i, w in enumerate(want):
                    assert int(ok[i]) == int(w[0]) and int(bits[i]) == int(w[2], 16), (mode, i, len(pairs[i][0]), len(pairs[i][1]), float(tm[i]), w[1])
    finally:
        th.close()


def test_thal_gpu_is_bit_exact(monkeypatch):
    """dg_thal_batch (-fmad=false) against the 64-bit patterns of the reference's own thal() for the
    golden pairs (709 primer-like, 1500 at the limits of the DP): the warp-per-pair kernel, the
    sequential kernel alone (DG_THAL_SEQ=1), and the warp kernel with every pair handed on to the
    sequential one (DG_THAL_SEQ=2, the path of a declined pair); then the same pairs many times
    over in one batch."""
    from dicey_b200.api import Thal
    th = Thal.open_tables(os.path.join(GOLDEN, "thal.params.tsv"), 0)
    try:
        for name in ("thal", "thal_long"):
            pairs = [l.rstrip("\n").split("\t") for l in open(os.path.join(GOLDEN, name + ".pairs.tsv"))]
            want = [l.split("\t") for l in open(os.path.join(GOLDEN, name + ".out.tsv")).read().splitlines()]
            for mode in ("0", "1", "2"):
                monkeypatch.setenv("DG_THAL_SEQ", mode)
                tm, ok = th.tm([p[0] for p in pairs], [p[1] for p in pairs])
                bits = tm.view(np.uint64)
                for i, w in enumerate(want):
                    assert int(ok[i]) == int(w[0]) and int(bits[i]) == int(w[2], 16), (name, mode, i, pairs[i], float(tm[i]), w[1])
            monkeypatch.setenv("DG_THAL_SEQ", "0")
            reps = 150 if name == "thal" else 20
            tm2, ok2 = th.tm([p[0] for p in pairs] * reps, [p[1] for p in pairs] * reps)
            assert np.array_equal(tm2.view(np.uint64).reshape(reps, -1), np.tile(bits, (reps, 1))) and ok2.all()
    finally:
        th.close()
