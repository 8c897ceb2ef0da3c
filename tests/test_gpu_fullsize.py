"""BASELINE.json's configurations at FULL size (3 Gb synthetic reference, 24 x 125 Mb) on the GPU.

Bit-exact comparison with the reference is only affordable on samples at this size (the reference
needs ~1 ms per primer at edit distance 1 and ~2 s at edit distance 2), so each configuration is
checked twice:
  * size-independent properties over the whole batch: every reported alignment is re-derived from
    the seeded text generator (dicey_b200/synth.py, no index involved), its score is recounted
    from the alignment columns, every planted primer is found at its locus, and for a handful of
    primers the hits equal a brute-force scan of all 3 G text windows;
  * a sample of the same queries through oracle/_ref/dicey_ref (the reference's own SDSL /
    neighbors.h / needle.h) on the same index, written as .fm9 by dg_index_write_fm9 -- records
    compared bit for bit.  Skipped when the reference binary has not been built.
"""
import os
import subprocess
import tempfile
import time

import numpy as np
import pytest

from dicey_b200 import synth
from dicey_b200.api import HuntParams, Index

pytestmark = pytest.mark.gpu

SEED, NREC, RECLEN = 42, 24, 125_000_000
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "dicey_ref")
COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


@pytest.fixture(scope="module")
def big():
    t0 = time.time()
    ix = Index.build_synthetic(SEED, NREC, RECLEN, 0)
    print(f"[fullsize] 3 Gb index built on the GPU in {time.time() - t0:.1f} s, {ix.info()['device_bytes'] / 1e9:.1f} GB")
    yield ix
    ix.close()


@pytest.fixture(scope="module")
def fm9(big):
    """The same index as an SDSL-loadable .fm9 for the reference binary (None without the binary)."""
    if not os.path.exists(REF_BIN):
        yield None
        return
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(d, f"dicey_b200_test_{os.getpid()}.fm9")
    t0 = time.time()
    big.write_fm9(path)
    rec = path + ".rec.tsv"
    with open(rec, "w") as f:
        for i in range(NREC):
            f.write(f"chr{i + 1}\t{RECLEN}\n")
    print(f"[fullsize] .fm9 written in {time.time() - t0:.1f} s ({os.path.getsize(path) / 1e9:.2f} GB)")
    yield path, rec
    for p in (path, path + "_check", rec):
        try:
            os.remove(p)
        except OSError:
            pass


def revcomp(b: bytes) -> bytes:
    return b.translate(COMP)[::-1]


def window(chrom: int, start1: int, length: int) -> bytes:
    """Bases [start1, start1 + length) (1-based) of record `chrom`, from the seeded generator."""
    return synth.bases(SEED, chrom * RECLEN + start1 - 1, length).tobytes()


def check_alignment(rec, query: bytes, dmax: int, indel: bool):
    score, chrom, start, strand, ra, qa = rec
    ra, qa = ra.encode(), qa.encode()
    q = query if strand == "+" else revcomp(query)
    assert len(ra) == len(qa)
    g = ra.replace(b"-", b"")
    assert qa.replace(b"-", b"") == q                       # the query row is the (strand of the) query
    assert window(chrom, start, len(g)) == g                # the reference row is the genome at (chr, start)
    cost = sum(1 for a, b in zip(ra, qa) if a != b)         # mismatches + gap columns
    assert cost == -score and cost <= dmax
    if not indel:
        assert b"-" not in ra + qa
    assert not (ra[:1] == b"-" and qa[:1] == b"-")


def planted_found(res, truth, n_rows, length):
    missing = 0
    for i in range(0, n_rows, 2):
        t = i // 2
        want_chr = int(truth["rec"][t])
        lo = int(truth["off"][t]) + 1
        strand = "-" if truth["rc"][t] else "+"
        ok = False
        for (score, chrom, start, st, ra, qa) in res.push_hits(i):
            if chrom == want_chr and st == strand and abs(start - lo) <= 2:
                ok = True
        missing += 0 if ok else 1
    return missing


def run_ref(args):
    return subprocess.run([REF_BIN] + args, check=True, capture_output=True, text=True).stdout


def write_queries(path, rows):
    with open(path, "wb") as f:
        f.write(b"\n".join(bytes(r) for r in rows) + b"\n")


# ---------------------------------------------------------------------------------------------
def test_config2_hamming1_10k(big, fm9, tmp_path):
    """dicey hunt: 10k random 20-mers, hamming-distance 1, 3 Gb reference, 1 x B200."""
    n = 10_000
    pr, truth = synth.primers_fast(SEED, NREC, RECLEN, n, 20, 1, False, rng_seed=11, return_truth=True)
    par = HuntParams(distance=1, hamming=True)
    res = big.hunt(pr, par)
    assert res.nq == n and (res.status == 0).all()
    for q in range(n):
        hits = res.push_hits(q)
        assert len(set(h[1:4] for h in hits)) == len(hits)   # one record per (chr, start, strand)
        for h in hits:
            check_alignment(h, bytes(pr[q]), 1, False)
    assert planted_found(res, truth, n, 20) == 0
    # brute force over all 3 G windows for a few primers (planted and random): no index involved
    import torch
    text = torch.from_numpy(big.debug_array("text")).cuda()
    nl = (text == 10) | (text == 0)
    bad = torch.zeros(text.numel() - 19, dtype=torch.uint8, device="cuda")
    for j in range(20):
        bad |= nl[j:j + bad.numel()].to(torch.uint8)
    for q in (0, 1, 2, 3, 4, 5):
        want = set()
        for strand, s in (("+", bytes(pr[q])), ("-", revcomp(bytes(pr[q])))):
            mism = bad * 2
            for j in range(20):
                mism += (text[j:j + bad.numel()] != s[j]).to(torch.uint8)
            pos = torch.nonzero(mism <= 1).flatten().cpu().numpy()
            for p in pos:
                want.add((int(p) // (RECLEN + 1), int(p) % (RECLEN + 1) + 1, strand))
        got = set((h[1], h[2], h[3]) for h in res.push_hits(q))
        assert got == want, q
    del text, nl, bad
    torch.cuda.empty_cache()
    if fm9:
        path, rec = fm9
        m = 2000
        qf = str(tmp_path / "q.txt")
        out = str(tmp_path / "rec.tsv")
        write_queries(qf, pr[:m])
        run_ref(["hunt", path, rec, qf, "-d", "1", "-n", "--threads", str(os.cpu_count() or 1), "--records", out])
        sub = big.hunt(pr[:m], par)
        assert sub.records_tsv(par, pr[:m]) == open(out).read()


def test_config4_edit2_sample(big, fm9, tmp_path):
    """dicey hunt: 20-mers at edit-distance 2 on the 3 Gb reference (the 1 M x 8-GPU config, one shard's worth of logic)."""
    n = 20_000
    pr, truth = synth.primers_fast(SEED, NREC, RECLEN, n, 20, 2, True, rng_seed=13, return_truth=True)
    par = HuntParams(distance=2)
    t0 = time.time()
    res = big.hunt(pr, par)
    dt = time.time() - t0
    print(f"[fullsize] edit-2: {n} primers in {dt * 1e3:.0f} ms ({n / dt / 1e3:.0f} k primers/s), {len(res.hits)} hits")
    for q in range(0, n, 7):
        for h in res.push_hits(q):
            check_alignment(h, bytes(pr[q]), 2, True)
    assert planted_found(res, truth, n, 20) == 0
    if fm9:
        path, rec = fm9
        m = 32
        qf = str(tmp_path / "q.txt")
        out = str(tmp_path / "rec.tsv")
        write_queries(qf, pr[:m])
        run_ref(["hunt", path, rec, qf, "-d", "2", "--threads", str(os.cpu_count() or 1), "--records", out])
        sub = big.hunt(pr[:m], par)
        assert sub.records_tsv(par, pr[:m]) == open(out).read()


def test_config3_search_seeds(big, fm9, tmp_path):
    """dicey search (FM / NW part): 1k primer pairs, 15-mer seeds at edit-distance 1, 3 Gb reference."""
    rng = np.random.default_rng(5)
    n = 2000
    rows = []
    for i in range(n):
        L = int(rng.integers(18, 26))
        chrom, off = int(rng.integers(0, NREC)), int(rng.integers(0, RECLEN - 3000))
        s = bytearray(window(chrom, off + 1, L))
        if i % 3 == 0:                      # one substitution inside the 3' seed
            j = L - 1 - int(rng.integers(0, 15))
            s[j] = ord("ACGT"[("ACGT".index(chr(s[j])) + 1 + int(rng.integers(0, 3))) % 4])
        rows.append(bytes(s) if i % 2 == 0 else revcomp(bytes(s)))
    par = HuntParams(distance=1, seed_len=15, maxmatches=10000)
    res = big.hunt(rows, par)
    assert res.nq == n
    nfound = 0
    for q in range(n):
        for (chrom, chrpos, alignpos, strand, ctx) in res.seed_hits(q):
            assert window(chrom, chrpos + 1, len(ctx)).decode() == ctx   # the context handed to the Tm gate
            assert chrpos <= alignpos <= chrpos + len(ctx)
            nfound += 1
    assert nfound >= n
    if fm9:
        path, rec = fm9
        m = 300
        qf = str(tmp_path / "p.txt")
        with open(qf, "w") as f:
            for i, r in enumerate(rows[:m]):
                f.write(f"p{i}\t{r.decode()}\n")
        want, cur = [], None
        for line in run_ref(["seed", path, rec, qf, "-k", "15", "-d", "1"]).splitlines():
            f = line.split("\t")
            if f[0] == "Q":
                cur = []
                want.append(cur)
            else:
                cur.append((int(f[2]), int(f[3]), int(f[4]), "-" if f[1] == "1" else "+", f[5]))
        for q in range(m):
            assert res.seed_hits(q) == want[q], q


def test_config5_padlock_arm_counts(big, fm9, tmp_path):
    """dicey padlock (count-only use of the path): 20-mer arms of 100 synthetic regions, exact and edit-1 neighbourhood counts."""
    rng = np.random.default_rng(9)
    arms = []
    for i in range(100):
        chrom, off = int(rng.integers(0, NREC)), int(rng.integers(0, RECLEN - 3000))
        region = window(chrom, off + 1, 2000)
        for j in range(4):
            o = int(rng.integers(0, 1960))
            arms.append(region[o:o + 20])
    exact = big.count(arms, HuntParams(distance=0))
    total = big.count(arms, HuntParams(distance=1))
    assert (exact >= 1).all() and (total >= exact).all()
    if fm9:
        path, rec = fm9
        qf = str(tmp_path / "arms.txt")
        write_queries(qf, arms)
        lines = run_ref(["padcount", path, qf, "-d", "1"]).splitlines()
        assert len(lines) == len(arms)
        for i, line in enumerate(lines):
            f = line.split("\t")
            assert (int(exact[i]), int(total[i])) == (int(f[1]), int(f[2])), i


def test_config5_padlock_windows_tm_and_counts(big, fm9, tmp_path):
    """BASELINE config 5 at its stated size: 100 regions of 2 kb, every sliding 2 x 20 window
    (padlock.h:326-427): GC filter on the host, the three thal() temperatures of a window (arm 1, arm 2
    and the 40-nt probe against their reverse complements: dg_thal_batch), the temperature rules of
    padlock.h:342-378, exact arm counts (sdsl::count of arm + reverse complement: dg_count_batch, d = 0)
    and neighbourhood counts (d = 1) with the uniqueness rules of padlock.h:381-427 -- every window
    through the C ABI; a sample of temperatures and counts against the reference's own thal.h / SDSL /
    neighbors.h, and the surviving windows of that sample recomputed from the reference's numbers."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from util import GOLDEN, write_primer3_config
    from dicey_b200.api import Thal
    rng = np.random.default_rng(11)
    armlen, tlen = 20, 40
    arms1, arms2, probes = [], [], []
    for i in range(100):
        chrom, off = int(rng.integers(0, NREC)), int(rng.integers(0, RECLEN - 3000))
        exon = window(chrom, off + 1, 2000)
        if i % 2:
            exon = revcomp(exon)                       # '-' strand features (padlock.h:321)
        for k in range(len(exon) - tlen + 1):
            arms1.append(exon[k:k + armlen]); arms2.append(exon[k + armlen:k + tlen]); probes.append(exon[k:k + tlen])
    n = len(probes)
    assert n == 100 * 1961

    def gc(seqs):
        a = np.frombuffer(b"".join(seqs), dtype=np.uint8).reshape(len(seqs), -1)
        return ((a == ord("G")) | (a == ord("C"))).sum(axis=1) / a.shape[1]
    g1, g2, gp = gc(arms1), gc(arms2), gc(probes)
    ok = (g1 >= 0.4) & (g1 <= 0.6) & (g2 >= 0.4) & (g2 <= 0.6) & (gp >= 0.4) & (gp <= 0.6)
    idx = np.flatnonzero(ok)
    th = Thal.open_tables(os.path.join(GOLDEN, "thal.params.tsv"), 0)
    t0 = time.time()
    try:
        o1 = [arms1[i] for i in idx] + [arms2[i] for i in idx] + [probes[i] for i in idx]
        tm, tok = th.tm(o1, [revcomp(x) for x in o1])
    finally:
        th.close()
    t_thal = time.time() - t0
    assert tok.all()
    m = len(idx)
    tm1, tm2, tmp_ = tm[:m], tm[m:2 * m], tm[2 * m:]
    keep = (tm1 <= 93 + g1[idx] - 675.0 / armlen) & (tm2 <= 93 + g2[idx] - 675.0 / armlen) & (np.abs(tm1 - tm2) <= 2)
    pmin = 81.5 + gp[idx] - 675.0 / tlen
    keep &= (tmp_ >= pmin) & (tmp_ <= pmin + 10)
    sel = idx[keep]
    t0 = time.time()
    arms = [arms1[i] for i in sel] + [arms2[i] for i in sel]
    exact = big.count(arms, HuntParams(distance=0))
    total = big.count(arms, HuntParams(distance=1))
    t_count = time.time() - t0
    s = len(sel)
    uniq = (exact[:s] <= 1) & (exact[s:] <= 1) & (total[:s] <= 2) & (total[s:] <= 2)   # armMode, edit distance 1: maxNeighborHits = 2
    print(f"[fullsize] config 5: {n} windows, {m} pass GC, {s} pass the Tm rules ({3 * m} thal pairs in {t_thal:.2f} s), "
          f"{int(uniq.sum())} unique ({4 * s} arm counts in {t_count:.2f} s)")
    assert s > 100 and uniq.sum() > 50 and (exact >= 1).all() and (total >= exact).all()
    if fm9:
        path, rec = fm9
        cfg = write_primer3_config(str(tmp_path / "p3cfg"))
        # temperatures of a sample of windows (all three per window), bit for bit
        samp = idx[:: max(1, len(idx) // 200)][:200]
        pf = str(tmp_path / "pairs.tsv")
        with open(pf, "wb") as f:
            for i in samp:
                for x in (arms1[i], arms2[i], probes[i]):
                    f.write(x + b"\t" + revcomp(x) + b"\n")
        want = [l.split("\t") for l in run_ref(["thal", cfg + "/", pf, "/dev/null"]).splitlines()]
        pos = {int(i): j for j, i in enumerate(idx)}
        for r, i in enumerate(samp):
            j = pos[int(i)]
            for c, arr in enumerate((tm1, tm2, tmp_)):
                assert int(arr[j:j + 1].view(np.uint64)[0]) == int(want[3 * r + c][2], 16), (int(i), c)
        # counts of a sample of surviving windows
        ssel = list(range(0, s, max(1, s // 100)))[:100]
        qf = str(tmp_path / "arms.txt")
        write_queries(qf, [arms[k] for k in ssel] + [arms[s + k] for k in ssel])
        lines = run_ref(["padcount", path, qf, "-d", "1"]).splitlines()
        ref_exact = np.array([int(l.split("\t")[1]) for l in lines]); ref_total = np.array([int(l.split("\t")[2]) for l in lines])
        k = len(ssel)
        assert (ref_exact[:k] == exact[ssel]).all() and (ref_exact[k:] == exact[[s + x for x in ssel]]).all()
        # (the reference stops adding once a total exceeds maxNeighborHits: only the decision is comparable)
        ref_uniq = (ref_exact[:k] <= 1) & (ref_exact[k:] <= 1) & (ref_total[:k] <= 2) & (ref_total[k:] <= 2)
        assert (ref_uniq == uniq[ssel]).all()


def test_config_headline_edit1_properties(big):
    """The bench workload itself (1 M 20-mers, edit distance 1): whole-batch invariants."""
    n = 1_000_000
    pr, truth = synth.primers_fast(SEED, NREC, RECLEN, n, 20, 1, True, rng_seed=7, return_truth=True)
    par = HuntParams(distance=1)
    res = big.hunt(pr, par)
    assert res.nq == n and int(res.qoff[-1]) == len(res.hits)
    h = res.hits
    assert (np.diff(h["query"].astype(np.int64)) >= 0).all()            # hits grouped by query, in query order
    assert ((h["score"] <= 0) & (h["score"] >= -1)).all()
    assert (h["chr"] < NREC).all() and (h["start"] >= 1).all() and (h["start"] <= RECLEN).all()
    # the planted half is found (sampled: re-deriving 1 M alignments in Python would take minutes)
    sub = np.arange(0, 40_000, 2)
    for i in sub:
        t = i // 2
        hits = res.push_hits(int(i))
        assert any(c == int(truth["rec"][t]) and abs(s - int(truth["off"][t]) - 1) <= 2 for (_, c, s, _, _, _) in hits), i
        for rec in hits:
            check_alignment(rec, bytes(pr[i]), 1, True)
    # text_pos <-> (chr, start) consistency over the whole batch, vectorised
    lead = h["start"].astype(np.int64) - 1 + h["chr"].astype(np.int64) * (RECLEN + 1) - h["text_pos"].astype(np.int64)
    assert (np.abs(lead) <= 2).all()


def test_fm9_loaded_as_is_at_full_size(big, fm9):
    """The 1.6 GB .fm9 (byte-identical to what `dicey index` writes for this text) is parsed and
    transcoded to the device layout; the result must equal the index built from the text on the GPU."""
    if not fm9:
        pytest.skip("needs the .fm9 written for the reference binary")
    path, _ = fm9
    t0 = time.time()
    ix = Index.open(path, 0)
    print(f"[fullsize] dg_index_open of the 3 Gb .fm9: {time.time() - t0:.1f} s")
    try:
        ix.set_records([f"chr{i + 1}" for i in range(NREC)], [RECLEN + 1] * NREC)
        a, b = ix.info(), big.info()
        for k in ("n", "sigma", "kmer", "bitmap_k", "n_exceptions"):  # (device_bytes differs: with two replicas on one GPU the second skips the optional KB + 1 bitmap)
            assert a[k] == b[k], k
        for what, dt in (("C", np.uint32), ("sa_samples", np.uint32), ("isa_samples", np.uint32), ("exc_pos", np.uint32)):
            assert np.array_equal(ix.debug_array(what, dt), big.debug_array(what, dt)), what
        for what in ("text", "occ"):
            x = ix.debug_array(what)
            y = big.debug_array(what)
            assert x.size == y.size and np.array_equal(x, y), what
            del x, y
        pr = synth.primers_fast(SEED, NREC, RECLEN, 20_000, 20, 1, True, rng_seed=21)
        par = HuntParams(distance=1)
        r1, r2 = ix.hunt(pr, par), big.hunt(pr, par)
        assert np.array_equal(r1.qoff, r2.qoff)
        for f in ("query", "score", "chr", "start", "text_pos", "aln_len", "strand"):
            assert np.array_equal(r1.hits[f], r2.hits[f]), f
        for q in range(0, 20_000, 211):
            assert r1.push_hits(q) == r2.push_hits(q)
    finally:
        ix.close()


def test_config3_search_end_to_end(big, fm9, tmp_path):
    """dicey search at full size: `dicey-b200 search` on the 3 Gb .fm9 (loaded as-is) against the
    reference driver (SDSL + neighbors.h + needle.h + thal.h + nlohmann) for planted primer pairs:
    binding sites, melting temperatures, amplicons, penalties -- the JSON must be the same bytes."""
    if not fm9:
        pytest.skip("needs oracle/_ref/dicey_ref")
    import shutil
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from util import write_primer3_config
    path, rec = fm9
    d = str(tmp_path)
    subprocess.run(["make", "-C", os.path.join(ROOT, "dicey_b200", "host")], check=True, capture_output=True)
    os.symlink(path, os.path.join(d, "genome.fa.fm9"))
    os.symlink(path + "_check", os.path.join(d, "genome.fa.fm9_check"))
    with open(os.path.join(d, "genome.fa.gz"), "wb") as f:
        f.write(b"placeholder: search only needs the .fai and the index")
    with open(os.path.join(d, "genome.fa.gz.fai"), "w") as f:
        for i in range(NREC):
            f.write(f"chr{i + 1}\t{RECLEN}\t0\t60\t61\n")
    cfg = write_primer3_config(os.path.join(d, "p3cfg"))
    rng = np.random.default_rng(17)
    with open(os.path.join(d, "primers.fa"), "w") as f:
        npairs = int(os.environ.get("DG_SEARCH_PAIRS", "1000"))   # BASELINE config 3
        nref = int(os.environ.get("DG_SEARCH_REF_PAIRS", "40"))     # pairs of the byte-for-byte comparison (the reference needs ~0.25 s per pair)
        lines_fa = []
        for i in range(npairs):
            chrom, off, alen = int(rng.integers(0, NREC)), int(rng.integers(0, RECLEN - 5000)), int(rng.integers(150, 3000))
            L1, L2 = int(rng.integers(18, 25)), int(rng.integers(18, 25))
            fw = bytearray(window(chrom, off + 1, L1))
            rv = bytearray(revcomp(window(chrom, off + alen - L2 + 1, L2)))
            if i % 3 == 1:
                fw[1] = ord("ACGT"[("ACGT".index(chr(fw[1])) + 1) % 4])
            lines_fa.append(f">amp{i}_F\n{fw.decode()}\n>amp{i}_R\n{rv.decode()}\n")
        f.write("".join(lines_fa))
    with open(os.path.join(d, "primers_ref.fa"), "w") as f:
        f.write("".join(lines_fa[:nref]))
    import json
    exe = os.path.join(ROOT, "dicey_b200", "dicey-b200")
    # the whole configuration through the product
    t1 = time.time()
    full = subprocess.run([exe, "search", "-g", "genome.fa.gz", "-i", cfg, "primers.fa"], cwd=d, capture_output=True, text=True,
                          env=dict(os.environ, DICEY_B200_TRACE="1"))
    t2 = time.time()
    assert full.returncode == 0, full.stderr
    jf = json.loads(full.stdout)
    print(f"[fullsize] dicey-b200 search, {2 * npairs} primers: {t2 - t1:.1f} s (index load included); "
          f"{len(jf['data']['primers'])} binding sites, {len(jf['data']['amplicons'])} amplicons")
    print(full.stderr)
    assert len(jf["data"]["amplicons"]) >= npairs * 5 // 6
    # the first pairs on their own, byte for byte against the reference driver
    t0 = time.time()
    want = subprocess.run([REF_BIN, "search", path, rec, os.path.join(d, "primers_ref.fa"), cfg], check=True, capture_output=True, text=True).stdout
    t1 = time.time()
    got = subprocess.run([exe, "search", "-g", "genome.fa.gz", "-i", cfg, "primers_ref.fa"], cwd=d, capture_output=True, text=True)
    assert got.returncode == 0, got.stderr
    j = json.loads(want)
    print(f"[fullsize] search, {2 * nref} primers: reference {t1 - t0:.1f} s; {len(j['data']['primers'])} binding sites, "
          f"{len(j['data']['amplicons'])} amplicons: same bytes")
    assert got.stdout == want


def test_hunt_cli_at_full_size(big, fm9, tmp_path):
    """`dicey-b200 hunt` with a FASTA of 200 000 primers on the 3 Gb .fm9 (loaded as-is): every line
    in input order; a sample of lines equals the Python mirror of writeJsonDnaHitOut built from the
    library's records for the same primers (those records are pinned to the reference elsewhere)."""
    if not fm9:
        pytest.skip("needs the .fm9 written for the reference binary")
    from dicey_b200.api import hunt_json
    path, _ = fm9
    d = str(tmp_path)
    subprocess.run(["make", "-C", os.path.join(ROOT, "dicey_b200", "host")], check=True, capture_output=True)
    os.symlink(path, os.path.join(d, "genome.fa.fm9"))
    os.symlink(path + "_check", os.path.join(d, "genome.fa.fm9_check"))
    with open(os.path.join(d, "genome.fa.gz"), "wb") as f:
        f.write(b"placeholder: hunt only needs the .fai and the index")
    names = [f"chr{i + 1}" for i in range(NREC)]
    with open(os.path.join(d, "genome.fa.gz.fai"), "w") as f:
        for n in names:
            f.write(f"{n}\t{RECLEN}\t0\t60\t61\n")
    nq = 200_000
    pr = synth.primers_fast(SEED, NREC, RECLEN, nq, 20, 1, True, rng_seed=33)
    with open(os.path.join(d, "q.fa"), "w") as f:
        f.write("".join(f">p{q}\n{pr[q].tobytes().decode()}\n" for q in range(nq)))
    t0 = time.time()
    r = subprocess.run([os.path.join(ROOT, "dicey_b200", "dicey-b200"), "hunt", "-g", "genome.fa.gz", "q.fa"], cwd=d,
                       capture_output=True, text=True, env=dict(os.environ, DICEY_B200_TRACE="1"))
    dt = time.time() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    print(f"[fullsize] dicey-b200 hunt, {nq} primers from a FASTA: {dt:.1f} s wall, {len(r.stdout) / 1e6:.0f} MB of JSON")
    print(r.stderr)
    assert len(lines) == nq
    par = HuntParams(distance=1)
    res = big.hunt(pr, par)
    for q in list(range(0, nq, 997)) + [nq - 1]:
        assert lines[q] == hunt_json(res, q, par, names, "genome.fa.gz", qname=f"p{q}").rstrip("\n"), q
