"""The C++ host program (dicey_b200/dicey-b200): dicey's `hunt` command line and JSON output over
the C ABI.  CPU tests cover argument handling and the error records that need no index; the GPU
tests compare its stdout / gzipped output byte for byte with the reference's golden JSON lines."""
import gzip
import os
import shutil
import subprocess

import pytest

from util import GOLDEN, HUNT_CASES, read_queries

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "dicey_b200", "dicey-b200")


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.run(["make", "-C", os.path.join(ROOT, "dicey_b200", "host")], check=True, capture_output=True)


def run(args, cwd=None, env=None):
    return subprocess.run([BIN] + args, capture_output=True, text=True, cwd=cwd, env=env)


def make_genome_dir(tmp_path, index):
    """<dir>/genome.fa.gz (+ .fai) and <dir>/genome.fa.fm9 (+ _check): the layout hunter.h:248-256 expects."""
    d = str(tmp_path)
    with gzip.open(os.path.join(d, "genome.fa.gz"), "wb") as f:
        f.write(b">placeholder\nACGT\n")  # hunt only checks that the genome exists; names come from the .fai
    with open(os.path.join(d, "genome.fa.gz.fai"), "w") as fai:
        for line in open(os.path.join(GOLDEN, index + ".rec.tsv")):
            n, l = line.split()
            fai.write(f"{n}\t{l}\t0\t60\t61\n")
    shutil.copy(os.path.join(GOLDEN, index + ".fm9"), os.path.join(d, "genome.fa.fm9"))
    shutil.copy(os.path.join(GOLDEN, index + ".fm9_check"), os.path.join(d, "genome.fa.fm9_check"))
    return d


def test_usage_and_unknown_command():
    r = run([])
    assert r.returncode == 0 and "Usage: dicey <command> <arguments>" in r.stdout
    r = run(["hunt"])
    assert r.returncode == 255 and r.stdout.startswith("Usage: dicey hunt [OPTIONS] -g Danio_rerio.fa.gz CATTACTAACATCAGT")
    r = run(["frobnicate"])
    assert r.returncode == 1 and "Unrecognized command frobnicate" in r.stderr
    r = run(["hunt", "--bogus", "x"])
    assert r.returncode == 1 and "unrecognised option" in r.stderr
    assert run(["version"]).stdout.startswith("Dicey version: v0.5.1")


def test_missing_genome_is_an_error_record(tmp_path):
    r = run(["hunt", "-g", str(tmp_path / "nope.fa.gz"), "ACGTACGTACGT"])
    assert r.returncode == 1
    assert r.stdout == '{"errors": [{"title":"Error: Genome does not exist!","type":"error"}]}\n'
    # -o: the same record as one gzip member of the (truncated) output file
    out = tmp_path / "o.json.gz"
    out.write_bytes(b"stale")
    r = run(["hunt", "--genome=" + str(tmp_path / "nope.fa.gz"), "-o", str(out), "ACGTACGTACGT"])
    assert r.returncode == 1 and r.stdout == ""
    assert gzip.open(out, "rt").read() == '{"errors": [{"title":"Error: Genome does not exist!","type":"error"}]}\n'


def test_no_gpu_means_index_error_not_a_cpu_search(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    d = make_genome_dir(tmp_path, "t1m")
    r = run(["hunt", "-g", "genome.fa.gz", "GAATACACCTAAAAACAC"], cwd=d)
    assert r.returncode == 1
    assert r.stdout == '{"errors": [{"title":"Error: FM-Index cannot be loaded!","type":"error"}]}\n'
    assert "no CPU path" in r.stderr


def fasta_of(case, path):
    qs = read_queries(os.path.join(GOLDEN, case + ".queries.txt"))
    with open(path, "w") as f:
        for n, s in qs:
            f.write(f">{n}\n{s}\n")
    return qs


@pytest.mark.gpu
@pytest.mark.parametrize("case,index", HUNT_CASES)
def test_hunt_cli_matches_reference_json(tmp_path, case, index):
    d = make_genome_dir(tmp_path, index)
    qs = fasta_of(case, os.path.join(d, "q.fa"))
    flags = open(os.path.join(GOLDEN, case + ".flags.txt")).read().split()
    want = [l for l in open(os.path.join(GOLDEN, case + ".jsonl"))]
    r = run(["hunt", "-g", "genome.fa.gz"] + flags + ["q.fa"], cwd=d)
    assert r.returncode == 0, r.stderr
    got = r.stdout.splitlines(keepends=True)
    assert len(got) == len(want) == len(qs)
    for q, (g, w) in enumerate(zip(got, want)):
        assert g == w, (case, q)


@pytest.mark.gpu
@pytest.mark.parametrize("case,index", [("t1m_e1", "t1m"), ("stress_e2", "stress")])
def test_hunt_cli_devices_shards_the_queries(tmp_path, case, index):
    """--devices: the FASTA is cut into contiguous shards, one replica of the index and one host thread per
    entry; the JSON comes out in input order and equals the reference's.  Two shards on one GPU always;
    two GPUs (with the NCCL exchange of the hit records: dg_comm_init / dg_allgather_hits) when the box has them."""
    import torch
    d = make_genome_dir(tmp_path, index)
    qs = fasta_of(case, os.path.join(d, "q.fa"))
    flags = open(os.path.join(GOLDEN, case + ".flags.txt")).read().split()
    want = open(os.path.join(GOLDEN, case + ".jsonl")).read()
    lists = ["0,0", "0,0,0"] + (["0,1"] if torch.cuda.device_count() >= 2 else [])
    for devs in lists:
        r = run(["hunt", "-g", "genome.fa.gz", "--devices", devs] + flags + ["q.fa"], cwd=d,
                env=dict(os.environ, DICEY_B200_TRACE="1"))
        assert r.returncode == 0, r.stderr
        assert r.stdout == want, devs
        if devs == "0,1":
            assert "dg_allgather_hits" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("case,index", [("t1m_e1", "t1m"), ("stress_e1_m7", "stress")])
def test_hunt_cli_many_queries_threaded_output(tmp_path, case, index):
    """A FASTA of several thousand queries takes the threaded JSON path (blocks of queries formatted
    side by side, written in order, one gzip member per block): the lines must still be the
    reference's golden lines in input order, on stdout and in the gzipped output file, and must
    not depend on the number of threads."""
    d = make_genome_dir(tmp_path, index)
    qs = read_queries(os.path.join(GOLDEN, case + ".queries.txt"))
    want1 = [l for l in open(os.path.join(GOLDEN, case + ".jsonl"))]
    reps = 70000 // len(qs) + 1            # more than one 65536-query block
    with open(os.path.join(d, "q.fa"), "w") as f:
        for _ in range(reps):
            for n, s in qs:
                f.write(f">{n}\n{s}\n")
    flags = open(os.path.join(GOLDEN, case + ".flags.txt")).read().split()
    outs = []
    for threads in ("1", "7"):
        r = subprocess.run([BIN, "hunt", "-g", "genome.fa.gz"] + flags + ["q.fa"], capture_output=True, text=True, cwd=d,
                           env=dict(os.environ, DICEY_B200_THREADS=threads))
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout)
    assert outs[0] == outs[1]
    got = outs[1].splitlines(keepends=True)
    assert len(got) == reps * len(qs)
    for q, g in enumerate(got):
        w = want1[q % len(qs)]
        assert g == w, (case, q)
    r = run(["hunt", "-g", "genome.fa.gz"] + flags + ["-o", "out.json.gz", "q.fa"], cwd=d)
    assert r.returncode == 0 and r.stdout == ""
    assert gzip.open(os.path.join(d, "out.json.gz"), "rt").read() == outs[1].replace('"outfile":""', '"outfile":"out.json.gz"')


@pytest.mark.gpu
def test_hunt_cli_single_sequence_and_outfile(tmp_path):
    d = make_genome_dir(tmp_path, "t1m")
    name, seq = read_queries(os.path.join(GOLDEN, "cfg1_d0.queries.txt"))[0]
    want = open(os.path.join(GOLDEN, "cfg1_d0.jsonl")).read()
    r = run(["hunt", "-g", "genome.fa.gz", "-d", "0", seq], cwd=d)
    assert r.returncode == 0, r.stderr
    assert r.stdout == want.replace(f'"name":"{name}",', "")  # a literal sequence has no name (hunter.h:286)
    r = run(["hunt", "-g", "genome.fa.gz", "-d0", "-o", "out.json.gz", seq], cwd=d)
    assert r.returncode == 0 and r.stdout == ""
    got = gzip.open(os.path.join(d, "out.json.gz"), "rt").read()
    assert got == want.replace(f'"name":"{name}",', "").replace('"outfile":""', '"outfile":"out.json.gz"')
    # a too-short sequence is an error record, the run still succeeds (hunter.h:299-303)
    r = run(["hunt", "-g", "genome.fa.gz", "ACGTACG"], cwd=d)
    assert r.returncode == 0
    assert r.stdout == '{"errors": [{"title":"Error: Input sequence is shorter than 10 nucleotides!","type":"error"}]}\n'


@pytest.mark.gpu
def test_index_cli_writes_a_loadable_fm9(tmp_path):
    """`dicey-b200 index genome.fa.gz` then `hunt` on the result reproduces the golden records."""
    import numpy as np
    from dicey_b200 import synth
    d = str(tmp_path)
    txt = synth.text(42, 8, 125000).tobytes().decode().split("\n")
    with gzip.open(os.path.join(d, "genome.fa.gz"), "wt") as f:
        for i, rec in enumerate(r for r in txt if r):
            f.write(f">chr{i + 1} synthetic\n")
            for o in range(0, len(rec), 60):
                f.write(rec[o:o + 60].lower() if i == 3 else rec[o:o + 60])
                f.write("\n")
    r = run(["index", "genome.fa.gz"], cwd=d)
    assert r.returncode == 0, r.stderr
    assert os.path.exists(os.path.join(d, "genome.fa.fm9")) and os.path.exists(os.path.join(d, "genome.fa.fm9_check"))
    fasta_of("t1m_e1", os.path.join(d, "q.fa"))
    r = run(["hunt", "-g", "genome.fa.gz", "q.fa"], cwd=d)
    assert r.returncode == 0, r.stderr
    assert r.stdout == open(os.path.join(GOLDEN, "t1m_e1.jsonl")).read()
    fai = open(os.path.join(d, "genome.fa.gz.fai")).read().splitlines()
    assert fai[0].split("\t")[:2] == ["chr1", "125000"] and len(fai) == 8


SEARCH_CASES = [("search_t1m_default", "t1m", "search_t1m"), ("search_t1m_h1_k12", "t1m", "search_t1m"),
                ("search_t1m_c55_l1000", "t1m", "search_t1m"), ("search_t1m_d0", "t1m", "search_t1m"),
                ("search_t1m_prune", "t1m", "search_t1m"), ("search_stress_l400", "stress", "search_stress"),
                ("search_stress_big", "stress", "search_stress")]


@pytest.mark.gpu
@pytest.mark.parametrize("case,index,primers", SEARCH_CASES)
def test_search_cli_matches_reference_json(tmp_path, case, index, primers):
    """`dicey-b200 search` (seeds + NW on the GPU, primer3 thal on the GPU, Tm gate, de-duplication,
    amplicon pairing, penalties, JSON with nlohmann's number format) against the output of the
    reference driver built around the reference's own SDSL / neighbors.h / needle.h / thal.h / json."""
    import hashlib
    from util import write_primer3_config
    d = make_genome_dir(tmp_path, index)
    cfg = write_primer3_config(os.path.join(d, "p3cfg"))
    shutil.copy(os.path.join(GOLDEN, primers + ".primers.fa"), os.path.join(d, "primers.fa"))
    flags = open(os.path.join(GOLDEN, case + ".flags.txt")).read().split()
    r = run(["search", "-g", "genome.fa.gz", "-i", cfg] + flags + ["primers.fa"], cwd=d)
    assert r.returncode == 0, r.stderr
    want_file = os.path.join(GOLDEN, case + ".json")
    if os.path.exists(want_file):
        assert r.stdout == open(want_file).read()
    else:   # large outputs are pinned by their SHA-256
        want = open(want_file + ".sha256").read().split()[0]
        assert hashlib.sha256(r.stdout.encode()).hexdigest() == want


@pytest.mark.gpu
def test_search_cli_errors_and_outfile(tmp_path):
    from util import write_primer3_config
    d = make_genome_dir(tmp_path, "t1m")
    cfg = write_primer3_config(os.path.join(d, "p3cfg"))
    shutil.copy(os.path.join(GOLDEN, "search_t1m.primers.fa"), os.path.join(d, "primers.fa"))
    r = run(["search", "-g", "genome.fa.gz", "-i", os.path.join(d, "nope"), "primers.fa"], cwd=d)
    assert r.returncode == 1 and r.stdout == '{"errors": [{"title":"Error: Cannot find primer3 config directory!","type":"error"}]}\n'
    r = run(["search", "-g", "genome.fa.gz", "-i", cfg, "missing.fa"], cwd=d)
    assert r.returncode == 1 and "Error: Input fasta file is missing!" in r.stdout
    r = run(["search", "-g", "genome.fa.gz", "-i", cfg, "-o", "out.json.gz", "-c", "55", "-l", "1000", "primers.fa"], cwd=d)
    assert r.returncode == 0 and r.stdout == ""
    want = open(os.path.join(GOLDEN, "search_t1m_c55_l1000.json")).read().replace('"outfile":""', '"outfile":"out.json.gz"')
    assert gzip.open(os.path.join(d, "out.json.gz"), "rt").read() == want
