#!/usr/bin/env python
"""Regenerates tests/golden/* by running the reference itself (oracle/_ref/dicey_ref, built from
the sources under /root/reference by oracle/Makefile) on small seeded inputs.

    python tests/golden/make_golden.py

Outputs (all committed; the .fm9 files are what `dicey index` writes -- SDSL csa_wt<>):
  t1m.fm9(+_check)      index of synth.text(seed=42, 8 x 125000)            (BASELINE config 1 scale)
  stress.dump.gz/.fm9   repeat-rich text with N runs and poly-A (caps, ties, exceptions)
  <case>.queries.txt    name<TAB>sequence per line
  <case>.records.tsv    dicey_ref hunt --records: per query Q/M/P(push order)/S(sorted) lines + W counters
  <case>.jsonl          dicey_ref hunt --json: the reference's JSON line per query
  *.count.tsv / *.locate.tsv / *.seed.tsv / *.padcount.tsv / *.neighbors.txt / *.needle.tsv
  thal*.pairs.tsv / thal*.out.tsv / thal.params.tsv   dicey_ref thal: (oligo, site) pairs, the reference's Tm bits, its tables
                        (thal: primer-like; thal_long: both sides up to 60; thal_wide: one side up to 10 000)
"""
import gzip
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from dicey_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "dicey_ref")


def run(args, **kw):
    return subprocess.run([REF] + args, check=True, capture_output=True, text=True, **kw).stdout


def stress_text(rng):
    recs = []
    unit = "".join(rng.choice("ACGT") for _ in range(600))
    for r in range(3):
        parts = []
        for c in range(25):
            u = list(unit)
            for _ in range(rng.randint(0, 12)):
                p = rng.randrange(len(u))
                op = rng.randint(0, 2)
                if op == 0:
                    u[p] = rng.choice("ACGT")
                elif op == 1:
                    del u[p]
                else:
                    u.insert(p, rng.choice("ACGT"))
            parts.append("".join(u))
            if c % 7 == 3:
                parts.append("N" * rng.randint(1, 40))
            if c % 9 == 5:
                parts.append("A" * rng.randint(15, 60))
            parts.append("".join(rng.choice("ACGT") for _ in range(rng.randint(50, 400))))
        recs.append("".join(parts))
    recs.append("ACGT" * 8 + "A" * 30 + "ACGTTGCA" * 5)  # a short record: context trimming at borders
    return recs


def write_queries(path, qs):
    with open(path, "w") as f:
        for n, s in qs:
            f.write(f"{n}\t{s}\n" if n else f"{s}\n")


def hunt_case(name, fm9, rec, qs, flags):
    qf = os.path.join(HERE, name + ".queries.txt")
    write_queries(qf, qs)
    out = run(["hunt", fm9, rec, qf, "--records", os.path.join(HERE, name + ".records.tsv"), "--json",
               os.path.join(HERE, name + ".jsonl"), "--counters", "--genome", "genome.fa.gz"] + flags)
    with open(os.path.join(HERE, name + ".flags.txt"), "w") as f:
        f.write(" ".join(flags) + "\n")
    print(name, out.strip())


def main():
    rng = random.Random(1234)
    tmp = "/tmp/dicey_golden"
    os.makedirs(tmp, exist_ok=True)
    # ---- t1m
    txt = synth.text(42, 8, 125000)
    dump = os.path.join(tmp, "t1m.dump")
    open(dump, "wb").write(txt.tobytes())
    t1m = os.path.join(HERE, "t1m.fm9")
    print(run(["index", dump, t1m, tmp]).strip())
    names, sl = synth.records(8, 125000)
    rec1 = os.path.join(HERE, "t1m.rec.tsv")
    open(rec1, "w").write("".join(f"{n}\t{l - 1}\n" for n, l in zip(names, sl)))
    # ---- stress
    recs = stress_text(rng)
    sdump = os.path.join(tmp, "stress.dump")
    stext = "".join(r + "\n" for r in recs)
    open(sdump, "w").write(stext)
    with gzip.GzipFile(os.path.join(HERE, "stress.dump.gz"), "wb", compresslevel=9, mtime=0) as f:
        f.write(stext.encode())
    sfm = os.path.join(HERE, "stress.fm9")
    print(run(["index", sdump, sfm, tmp]).strip())
    rec2 = os.path.join(HERE, "stress.rec.tsv")
    open(rec2, "w").write("".join(f"s{i + 1}\t{len(r)}\n" for i, r in enumerate(recs)))

    # ---- queries on t1m: planted / random, lengths 18..25, config-1 style 18-mer at d=0
    def planted(m, d, indel, i):
        p = synth.primers(42, 8, 125000, 1, m, d, indel, rng_seed=1000 + i, planted_frac=1.0)[0]
        return p.decode()

    q18 = [("q18", synth.bases(42, 3 * 125000 + 777, 18).tobytes().decode())]
    hunt_case("cfg1_d0", t1m, rec1, q18, ["-d", "0"])
    qs = []
    for i in range(60):
        m = rng.choice([18, 20, 20, 20, 22, 25])
        if i % 3 == 2:
            qs.append((f"r{i}", "".join(rng.choice("ACGT") for _ in range(m))))
        else:
            qs.append((f"p{i}", planted(m, 2, True, i)))
    qs.append(("short", "ACGTACG"))
    qs.append(("lower", qs[0][1].lower()))
    qs.append(("withn", qs[1][1][:7] + "N" + qs[1][1][8:]))
    qs.append(("ten", synth.bases(42, 5000, 10).tobytes().decode()))
    hunt_case("t1m_e1", t1m, rec1, qs, ["-d", "1"])
    hunt_case("t1m_h1", t1m, rec1, qs, ["-d", "1", "-n"])
    hunt_case("t1m_h2", t1m, rec1, qs, ["-d", "2", "-n"])
    hunt_case("t1m_e2", t1m, rec1, qs[:24] + qs[-4:], ["-d", "2"])
    hunt_case("t1m_e1_fwd", t1m, rec1, qs[:20], ["-d", "1", "-f"])
    hunt_case("t1m_e0", t1m, rec1, qs[:20], ["-d", "0"])

    # ---- stress queries: pieces of the repeat unit, poly-A, border pieces, N-containing
    sq = []
    for i in range(40):
        r = rng.choice(recs[:3])
        m = rng.choice([12, 15, 20, 20, 24, 30])
        p = rng.randrange(0, len(r) - m)
        s = r[p:p + m]
        if i % 4 == 1:
            s = list(s)
            s[rng.randrange(m)] = rng.choice("ACGT")
            s = "".join(s)
        if i % 5 == 3:
            s = synth.revcomp(s.encode()).decode()
        sq.append((f"s{i}", s))
    sq.append(("polyA", "A" * 20))
    sq.append(("polyA12", "A" * 12))
    sq.append(("acgt", "ACGT" * 5))
    sq.append(("border_l", recs[3][:20]))
    sq.append(("border_r", recs[3][-20:]))
    sq.append(("first", recs[0][:22]))
    sq.append(("last", recs[2][-18:]))
    sq.append(("palin", "ACGTACGTACGTACGTACGT"))
    hunt_case("stress_e1", sfm, rec2, sq, ["-d", "1"])
    hunt_case("stress_h1", sfm, rec2, sq, ["-d", "1", "-n"])
    hunt_case("stress_e1_m7", sfm, rec2, sq, ["-d", "1", "-m", "7"])
    hunt_case("stress_h2_m50", sfm, rec2, sq, ["-d", "2", "-n", "-m", "50"])
    hunt_case("stress_e2", sfm, rec2, sq[:16] + sq[-8:], ["-d", "2", "-m", "200"])

    # ---- literal patterns: count / locate
    pats = [s for _, s in sq[:30]] + ["A", "N", "NN", "ACGTN", "T" * 5, "\n".strip() or "G", recs[3][:12]]
    pf = os.path.join(HERE, "stress.patterns.txt")
    open(pf, "w").write("\n".join(pats) + "\n")
    open(os.path.join(HERE, "stress.count.tsv"), "w").write(run(["count", sfm, pf]))
    open(os.path.join(HERE, "stress.locate.tsv"), "w").write(run(["locate", sfm, pf]))
    # ---- search seeds (FM / NW part of silica.h) and padlock counts
    primers = [(f"pr{i}", planted(rng.choice([18, 20, 22, 24]), 1, True, 500 + i)) for i in range(16)]
    prf = os.path.join(HERE, "t1m_seed.queries.txt")
    write_queries(prf, primers)
    open(os.path.join(HERE, "t1m_seed_k15_e1.seed.tsv"), "w").write(run(["seed", t1m, rec1, prf, "-k", "15", "-d", "1"]))
    open(os.path.join(HERE, "t1m_seed_k12_h1.seed.tsv"), "w").write(run(["seed", t1m, rec1, prf, "-k", "12", "-d", "1", "-n"]))
    sprf = os.path.join(HERE, "stress_seed.queries.txt")
    write_queries(sprf, [(n, s) for n, s in sq[:12] if len(s) >= 20])
    open(os.path.join(HERE, "stress_seed_k15_e1.seed.tsv"), "w").write(run(["seed", sfm, rec2, sprf, "-k", "15", "-d", "1", "-m", "300"]))
    arms = [s for _, s in sq[:10] if len(s) == 20] + [planted(20, 0, False, 900 + i) for i in range(6)]
    af = os.path.join(HERE, "arms.txt")
    open(af, "w").write("\n".join(arms) + "\n")
    open(os.path.join(HERE, "stress_arms_e1.padcount.tsv"), "w").write(run(["padcount", sfm, af, "-d", "1"]))
    open(os.path.join(HERE, "stress_arms_h1.padcount.tsv"), "w").write(run(["padcount", sfm, af, "-d", "1", "-n"]))
    open(os.path.join(HERE, "t1m_arms_e1.padcount.tsv"), "w").write(run(["padcount", t1m, af, "-d", "1"]))
    # ---- neighbourhoods and alignments (unit-level vectors for tests/hostsim)
    nq = ["".join(rng.choice("ACGT") for _ in range(m)) for m in (10, 12, 15, 18, 20, 20, 25)] + \
         ["A" * 20, "AC" * 10, "AAAAACCCCCGGGGGTTTTT", "ACGTNACGTACGTACGTNNA"]
    nf = os.path.join(HERE, "neighbors.queries.txt")
    open(nf, "w").write("\n".join(nq) + "\n")
    for d, ham in ((1, False), (1, True), (2, True)):
        tag = f"{'h' if ham else 'e'}{d}"
        with gzip.GzipFile(os.path.join(HERE, f"neighbors_{tag}.txt.gz"), "wb", mtime=0) as f:
            f.write(run(["neighbors", nf, "-d", str(d), "-x", "1000000"] + (["-n"] if ham else [])).encode())
    with gzip.GzipFile(os.path.join(HERE, "neighbors_e2.txt.gz"), "wb", mtime=0) as f:
        f.write(run(["neighbors", nf, "-d", "2", "-x", "1000000"]).encode())
    pairs = []
    for _ in range(400):
        m = rng.choice([10, 15, 20, 20, 25])
        q = "".join(rng.choice(rng.choice(["ACGT", "AC", "A"])) for _ in range(m))
        g = list(q)
        for _ in range(rng.randint(0, 3)):
            p = rng.randrange(len(g))
            op = rng.randint(0, 2)
            if op == 0:
                g[p] = rng.choice("ACGT")
            elif op == 1 and len(g) > 5:
                del g[p]
            else:
                g.insert(p, rng.choice("ACGT"))
        g = "".join(rng.choice("ACGT") for _ in range(rng.randint(0, 2))) + "".join(g) + \
            "".join(rng.choice("ACGT") for _ in range(rng.randint(0, 4)))
        pairs.append((g, q))
    pfn = os.path.join(HERE, "needle.pairs.tsv")
    open(pfn, "w").write("".join(f"{g}\t{q}\n" for g, q in pairs))
    open(os.path.join(HERE, "needle.out.tsv"), "w").write(run(["needle", pfn, "-"]))
    make_thal()
    make_thal_long()
    make_thal_wide()


def make_thal():
    """The Tm gate of `dicey search` (primer3 thal, thal_end1): pairs (primer, genomic site as
    silica.h:508-511 passes them), the reference's results as 64-bit patterns, and the nearest-neighbour
    tables the reference held after reading its primer3_config directory (the GPU box has no
    /root/reference to read them from)."""
    import numpy as np
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    pairs = []
    for i in range(700):
        L = int(rng.integers(15, 31))
        primer = bytes(acgt[rng.integers(0, 4, L)])
        site = bytearray(primer.translate(comp)[::-1])
        kind = i % 7
        if kind in (1, 2, 3):
            for _ in range(kind):
                p = int(rng.integers(0, len(site)))
                site[p] = acgt[rng.integers(0, 4)]
        elif kind == 4:
            p = int(rng.integers(1, len(site) - 1))
            del site[p]
        elif kind == 5:
            p = int(rng.integers(1, len(site) - 1))
            site.insert(p, int(acgt[rng.integers(0, 4)]))
        elif kind == 6:
            site = bytearray(acgt[rng.integers(0, 4, int(rng.integers(15, 34)))])
        fl, fr = int(rng.integers(0, 4)), int(rng.integers(0, 4))
        site = bytes(acgt[rng.integers(0, 4, fl)]) + bytes(site) + bytes(acgt[rng.integers(0, 4, fr)])
        if i % 53 == 0:
            site = site[:3] + b"N" + site[4:]
        if i % 97 == 0:
            primer = primer.lower()
        pairs.append((primer, site))
    pairs += [(b"ACGTACGTACGTACGTACGT", b"ACGTACGTACGTACGTACGT"), (b"GAATTCGAATTC", b"GAATTCGAATTC"),
              (b"GGGGGGGGGGCCCCCCCCCC", b"GGGGGGGGGGCCCCCCCCCC"), (b"AAAAAAAAAAAAAAAAAAAA", b"TTTTTTTTTTTTTTTTTTTT"),
              (b"AAAAAAAAAAAAAAAAAAAA", b"CCCCCCCCCCCCCCCCCCCC"), (b"ACG", b"CGT"), (b"A", b"T"),
              (b"GCGCGCGCGCGCGCGCGCGC", b"GCGCGCGCGCGCGCGCGCGC"), (b"ATATATATATATATATATAT", b"ATATATATATATATATATAT")]
    pfn = os.path.join(HERE, "thal.pairs.tsv")
    with open(pfn, "wb") as f:
        for a, b in pairs:
            f.write(a + b"\t" + b + b"\n")
    out = run(["thal", "/root/reference/src/primer3_config/", pfn, os.path.join(HERE, "thal.params.tsv")])
    open(os.path.join(HERE, "thal.out.tsv"), "w").write(out)


def make_thal_long():
    """A second thal set at the limits of the DP: both sequences 8..60 bases, several mismatches,
    multi-base deletions / insertions (bulges and internal loops up to the 30-base loop limit),
    replaced middles, low-complexity oligos.  Results only (the tables are thal.params.tsv)."""
    import numpy as np
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    pairs = []
    for i in range(1500):
        L = int(rng.integers(8, 61))
        kind = i % 9
        if kind == 7:
            unit = bytes(acgt[rng.integers(0, 4, int(rng.integers(1, 4)))])
            primer = (unit * 60)[:L]
        else:
            primer = bytes(acgt[rng.integers(0, 4, L)])
        site = bytearray(primer.translate(comp)[::-1])
        if kind in (1, 2, 3):
            for _ in range(kind * 2):
                site[int(rng.integers(0, len(site)))] = acgt[rng.integers(0, 4)]
        elif kind == 4:
            for _ in range(int(rng.integers(1, 4))):
                if len(site) > 3:
                    del site[int(rng.integers(1, len(site) - 1))]
        elif kind == 5:
            for _ in range(int(rng.integers(1, 6))):
                site.insert(int(rng.integers(1, len(site) - 1)), int(acgt[rng.integers(0, 4)]))
        elif kind == 6:
            site = bytearray(acgt[rng.integers(0, 4, int(rng.integers(8, 61)))])
        elif kind == 8:
            a = int(rng.integers(2, max(3, L // 2)))
            site[a:a + int(rng.integers(0, 10))] = bytes(acgt[rng.integers(0, 4, int(rng.integers(1, 20)))])
        site = bytes(site)[:60] or b"A"
        pairs.append((primer, site))
    pfn = os.path.join(HERE, "thal_long.pairs.tsv")
    with open(pfn, "wb") as f:
        for a, b in pairs:
            f.write(a + b"\t" + b + b"\n")
    out = run(["thal", "/root/reference/src/primer3_config/", pfn, "/dev/null"])
    open(os.path.join(HERE, "thal_long.out.tsv"), "w").write(out)


def make_thal_wide():
    """thal() with ONE side longer than THAL_MAX_ALIGN = 60 (thal.h:58, :2440-2451 -- the other side up to
    THAL_MAX_SEQ = 10 000): an oligo against a long target holding its (exact / mismatched / bulged /
    internal-loop) binding site, with the long sequence on either side, the site at the very ends, lengths
    at both limits; and the pairs the reference refuses (both sides longer than 60, a side longer than
    10 000).  Results only (the tables are thal.params.tsv)."""
    import numpy as np
    rng = np.random.default_rng(2024)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")

    def rnd(n):
        return bytes(acgt[rng.integers(0, 4, int(n))])

    pairs = []
    longs = [61, 62, 63, 64, 65, 96, 97, 128, 200, 333, 700, 1500, 4000, 10000]
    for i, L in enumerate(longs * 3):
        k = int(rng.integers(12, 61)) if i % 5 else 60
        oligo = rnd(k)
        site = bytearray(oligo.translate(comp)[::-1])
        kind = i % 6
        if kind == 1:
            for _ in range(3):
                site[int(rng.integers(0, len(site)))] = acgt[rng.integers(0, 4)]
        elif kind == 2:
            del site[int(rng.integers(1, len(site) - 1))]
        elif kind == 3:
            q = int(rng.integers(1, len(site) - 1))
            site[q:q] = rnd(rng.integers(1, 12))
        elif kind == 4:
            q = int(rng.integers(2, len(site) - 2))
            site[q:q + int(rng.integers(1, 6))] = rnd(rng.integers(1, 9))
        site = bytes(site)[:L]
        where = i % 4
        free = L - len(site)
        at = 0 if where == 0 else free if where == 1 else int(rng.integers(0, free + 1))
        target = rnd(at) + site + rnd(free - at)
        if i % 11 == 0:
            target = target[:5] + b"N" + target[6:]
        if i % 13 == 0:
            target = target.lower()
        pairs.append((oligo, target) if (i // len(longs)) % 2 == 0 or i % 3 == 0 else (target, oligo))
    unit = b"ACGT" * 2500
    pairs += [(b"ACGTACGTACGTACGTACGT", unit), (unit, b"ACGTACGTACGTACGTACGT"), (b"A" * 30, b"T" * 9000), (b"G" * 2000, b"C" * 60),
              (b"A", rnd(500)), (rnd(500), b"T"),
              (rnd(61), rnd(61)), (rnd(60), rnd(10001)), (rnd(10001), rnd(20)), (rnd(300), rnd(61))]   # refused by the reference
    pfn = os.path.join(HERE, "thal_wide.pairs.tsv")
    with open(pfn, "wb") as f:
        for a, b in pairs:
            f.write(a + b"\t" + b + b"\n")
    out = run(["thal", "/root/reference/src/primer3_config/", pfn, "/dev/null"])
    open(os.path.join(HERE, "thal_wide.out.tsv"), "w").write(out)


def make_truncation_cases():
    """hunt cases in which the reference truncates its neighbourhood at -x (neighbors.h:50, warning of
    hunter.h:342-345): small caps in both modes, and 24-26-mers at edit distance 2 under the default cap.
    Uses the committed t1m / stress indexes; own random stream, so the other cases do not move."""
    rng = random.Random(777)
    t1m, rec1 = os.path.join(HERE, "t1m.fm9"), os.path.join(HERE, "t1m.rec.tsv")
    sfm, rec2 = os.path.join(HERE, "stress.fm9"), os.path.join(HERE, "stress.rec.tsv")

    def planted(m, d, indel, i):
        return synth.primers(42, 8, 125000, 1, m, d, indel, rng_seed=7000 + i, planted_frac=1.0)[0].decode()

    qs = [(f"p{i}", planted(rng.choice([18, 20, 20, 22]), 1, True, i)) for i in range(14)]
    qs.append(("withn", qs[0][1][:5] + "N" + qs[0][1][6:]))
    qs.append(("polyA", "A" * 20))
    hunt_case("t1m_e1_x50", t1m, rec1, qs, ["-d", "1", "-x", "50"])
    hunt_case("t1m_e1_x150", t1m, rec1, qs, ["-d", "1", "-x", "150"])
    hunt_case("t1m_e1_m0", t1m, rec1, qs[:6], ["-d", "1", "-m", "0"])
    hunt_case("t1m_h2_x300", t1m, rec1, qs, ["-d", "2", "-n", "-x", "300"])
    hunt_case("t1m_h1_x40", t1m, rec1, qs, ["-d", "1", "-n", "-x", "40", "-f"])
    q2 = [(f"e{i}", planted(rng.choice([18, 20, 21, 22]), 2, True, 100 + i)) for i in range(10)]
    hunt_case("t1m_e2_x500", t1m, rec1, q2, ["-d", "2", "-x", "500"])
    hunt_case("t1m_e2_x5000", t1m, rec1, q2, ["-d", "2", "-x", "5000"])
    q3 = [(f"l{i}", planted(m, 2, True, 200 + i)) for i, m in enumerate([24, 24, 25, 25, 26, 23, 22, 28])]
    hunt_case("t1m_e2_long", t1m, rec1, q3, ["-d", "2"])
    # distance 3 (searched from host-made neighbour lists): short primers, both modes
    q4 = [(f"t{i}", planted(m, 2, True, 300 + i)) for i, m in enumerate([10, 11, 12, 12, 13, 13, 14, 16])]
    q4.append(("nine", "ACGTACGTA"))
    hunt_case("t1m_h3", t1m, rec1, q4, ["-d", "3", "-n"])
    hunt_case("t1m_e3_x3000", t1m, rec1, q4, ["-d", "3", "-x", "3000"])
    hunt_case("t1m_d12", t1m, rec1, q4[:3], ["-d", "12", "-n", "-x", "300"])
    recs = gzip.open(os.path.join(HERE, "stress.dump.gz"), "rt").read().split("\n")
    sq = []
    for i in range(10):
        r = recs[i % 3]
        m = rng.choice([20, 22, 24, 25])
        p0 = rng.randrange(0, len(r) - m)
        sq.append((f"s{i}", r[p0:p0 + m]))
    hunt_case("stress_e2_x2000", sfm, rec2, sq, ["-d", "2", "-x", "2000", "-m", "200"])
    # the truncated sets themselves (unit-level vectors for the replay of dicey_b200/csrc/nbr_trunc.hpp)
    nf = os.path.join(HERE, "neighbors.queries.txt")
    for d, ham, x in ((1, False, 50), (2, False, 500), (2, False, 3000), (2, True, 300), (1, True, 20)):
        tag = f"{'h' if ham else 'e'}{d}_x{x}"
        with gzip.GzipFile(os.path.join(HERE, f"neighbors_{tag}.txt.gz"), "wb", mtime=0) as f:
            f.write(run(["neighbors", nf, "-d", str(d), "-x", str(x)] + (["-n"] if ham else [])).encode())


def make_iupac_case():
    """A small text with IUPAC ambiguity codes (18 symbols with the separator and the sentinel): the index the
    reference builds for it and one hunt case -- pins dg_index_build_text / dg_index_write_fm9 beyond the
    seven-symbol DNA alphabet."""
    rng = random.Random(4242)
    tmp = "/tmp/dicey_golden"
    os.makedirs(tmp, exist_ok=True)
    recs = []
    for r in range(3):
        s = [rng.choice("ACGT") for _ in range(6000 + 500 * r)]
        for _ in range(60):
            s[rng.randrange(len(s))] = rng.choice("RYKMSWBDHVN")
        for _ in range(3):
            p0 = rng.randrange(len(s) - 30)
            s[p0:p0 + rng.randint(2, 25)] = "N" * rng.randint(2, 25)
        recs.append("".join(s))
    text = "".join(r + "\n" for r in recs)
    dump = os.path.join(tmp, "iupac.dump")
    open(dump, "w").write(text)
    with gzip.GzipFile(os.path.join(HERE, "iupac.dump.gz"), "wb", compresslevel=9, mtime=0) as f:
        f.write(text.encode())
    fm = os.path.join(HERE, "iupac.fm9")
    print(run(["index", dump, fm, tmp]).strip())
    rec = os.path.join(HERE, "iupac.rec.tsv")
    open(rec, "w").write("".join(f"u{i + 1}\t{len(r)}\n" for i, r in enumerate(recs)))
    qs = []
    for i in range(24):
        r = recs[i % 3]
        m = rng.choice([16, 18, 20, 22])
        p0 = rng.randrange(0, len(r) - m)
        piece = list(r[p0:p0 + m])
        if i % 3 == 1:
            piece[rng.randrange(m)] = rng.choice("ACGT")
        qs.append((f"i{i}", "".join(piece)))
    hunt_case("iupac_e1", fm, rec, qs, ["-d", "1"])
    hunt_case("iupac_h2", fm, rec, qs, ["-d", "2", "-n"])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "iupac":
        make_iupac_case()
    elif len(sys.argv) > 1 and sys.argv[1] == "trunc":
        make_truncation_cases()
    elif len(sys.argv) > 1 and sys.argv[1] == "thalwide":
        make_thal_wide()
    else:
        main()
        make_truncation_cases()
        make_iupac_case()
