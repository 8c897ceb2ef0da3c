"""A genome-like reference (100 Mb: a diverged repeat family, microsatellites, a homopolymer, N runs,
soft-masked lower case) end to end against the reference's own code:

  * `dicey-b200 index` (GPU suffix sort, BWT, wavelet tree, rank / select supports) must write the
    SAME BYTES as SDSL's construct + store_to_checked_file on the dump of the same FASTA
    (`dicey_ref index`, i.e. divsufsort + wt_huff + rank_support_v + select_support_mcl);
  * `dicey-b200 hunt` on it must print the same JSON lines as the reference driver for primers
    placed in unique sequence, in the repeat family (hit caps, thousands of occurrences per
    neighbour), next to N runs and in low-complexity sequence, at distances 1 and 2.

Needs oracle/_ref/dicey_ref (skipped otherwise); SDSL's construction takes ~30 s on the host."""
import gzip
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "dicey_ref")
BIN = os.path.join(ROOT, "dicey_b200", "dicey-b200")
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def make_genome(rng):
    recs = []
    fam = ACGT[rng.integers(0, 4, 300)]
    for r, L in enumerate((40_000_000, 30_000_000, 20_000_000, 10_000_000)):
        s = ACGT[rng.integers(0, 4, L)].copy()
        for _ in range(60):                                   # a 300 bp element, ~5 % diverged copies
            c = fam.copy()
            mut = rng.random(300) < 0.05
            c[mut] = ACGT[rng.integers(0, 4, int(mut.sum()))]
            o = int(rng.integers(0, L - 400))
            s[o:o + 300] = c
        for _ in range(3):                                    # microsatellites and a homopolymer
            o = int(rng.integers(0, L - 30_000))
            unit = ACGT[rng.integers(0, 4, int(rng.integers(2, 5)))]
            s[o:o + 12_000] = np.resize(unit, 12_000)
        o = int(rng.integers(0, L - 50_000))
        s[o:o + 20_000] = ord("A")
        for _ in range(2):                                    # assembly gaps
            o = int(rng.integers(0, L - 120_000))
            s[o:o + int(rng.integers(10_000, 100_000))] = ord("N")
        recs.append(s)
    recs[1][:5000] = ord("N")                                  # a record that starts with a gap
    return recs, fam


def write_fasta(path, recs, rng):
    with gzip.open(path, "wb", compresslevel=1) as f:
        for i, s in enumerate(recs):
            f.write(f">chr{i + 1} synthetic test record\n".encode())
            t = s.copy()
            lo = int(rng.integers(0, len(t) - 2_000_000))
            sel = t[lo:lo + 1_000_000]
            t[lo:lo + 1_000_000] = np.where(sel != ord("N"), sel | 0x20, sel)     # soft-masked lower case
            tb = t.tobytes()
            f.write(b"\n".join(tb[o:o + 60] for o in range(0, len(tb), 60)) + b"\n")


def revcomp(b):
    return b.translate(bytes.maketrans(b"ACGTN", b"TGCAN"))[::-1]


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/dicey_ref has not been built")
def test_index_and_hunt_on_a_genome_like_reference(tmp_path):
    subprocess.run(["make", "-C", os.path.join(ROOT, "dicey_b200", "host")], check=True, capture_output=True)
    rng = np.random.default_rng(2024)
    recs, fam = make_genome(rng)
    d = str(tmp_path)
    fa = os.path.join(d, "genome.fa.gz")
    write_fasta(fa, recs, rng)
    # the dump the reference indexes (index.h:96-115): records upper-cased, joined by '\n', trailing '\n'
    with open(os.path.join(d, "ref.dump"), "wb") as f:
        for s in recs:
            f.write(s.tobytes())
            f.write(b"\n")
    ref = subprocess.Popen([REF_BIN, "index", os.path.join(d, "ref.dump"), os.path.join(d, "ref.fm9"), d],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    r = subprocess.run([BIN, "index", "genome.fa.gz"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert ref.wait() == 0, ref.stderr.read()
    mine, theirs = os.path.join(d, "genome.fa.fm9"), os.path.join(d, "ref.fm9")
    assert os.path.getsize(mine) == os.path.getsize(theirs)
    assert subprocess.run(["cmp", "-s", mine, theirs]).returncode == 0            # byte for byte
    assert open(mine + "_check", "rb").read() == open(theirs + "_check", "rb").read()
    # primers: unique loci, the repeat family (forward and reverse complement), gap flanks, low complexity
    qs = []
    for i in range(60):
        rr = int(rng.integers(0, 4))
        o = int(rng.integers(0, len(recs[rr]) - 30))
        s = bytes(recs[rr][o:o + int(rng.integers(18, 26))])
        if b"N" in s:
            continue
        qs.append((f"u{i}", s if i % 2 else revcomp(s)))
    for i in range(12):
        o = int(rng.integers(0, 270))
        s = bytearray(fam[o:o + 22].tobytes())
        if i % 3 == 0:
            s[7] = ord("ACGT"[("ACGT".index(chr(s[7])) + 1) % 4])
        qs.append((f"rep{i}", bytes(s) if i % 2 else revcomp(bytes(s))))
    qs.append(("polyA", b"A" * 20))
    qs.append(("withN", bytes(recs[0][1000:1009]) + b"N" + bytes(recs[0][1010:1021])))
    qs.append(("lower", bytes(recs[2][5000:5021]).lower()))
    with open(os.path.join(d, "q.fa"), "w") as f:
        for n, s in qs:
            f.write(f">{n}\n{s.decode()}\n")
    with open(os.path.join(d, "q.txt"), "w") as f:
        for n, s in qs:
            f.write(f"{n}\t{s.decode()}\n")
    with open(os.path.join(d, "rec.tsv"), "w") as f:
        for i, s in enumerate(recs):
            f.write(f"chr{i + 1}\t{len(s)}\n")
    for flags in (["-d", "1"], ["-d", "1", "-n", "-m", "50"], ["-d", "2", "-m", "200"], ["-d", "0", "-f"]):
        want = os.path.join(d, "want.jsonl")
        subprocess.run([REF_BIN, "hunt", theirs, os.path.join(d, "rec.tsv"), os.path.join(d, "q.txt"), "--json", want,
                        "--threads", str(os.cpu_count() or 1)] + flags, check=True, capture_output=True)
        r = subprocess.run([BIN, "hunt", "-g", "genome.fa.gz"] + flags + ["q.fa"], cwd=d, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        got, exp = r.stdout.splitlines(), open(want).read().splitlines()
        assert len(got) == len(exp) == len(qs)
        for (n, _), g, e in zip(qs, got, exp):
            if "Neighborhood size exceeds" in e:
                continue
            assert g == e, (flags, n)
