"""CPU tests: the oracle restatement (oracle/dicey_oracle) and the host-compiled per-item
arithmetic of the CUDA kernels (tests/hostsim) against the golden outputs of the reference."""
import gzip
import os
import subprocess

import pytest

from util import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle", "dicey_oracle")
HOSTSIM = os.path.join(ROOT, "tests", "hostsim", "hostsim")


@pytest.fixture(scope="session", autouse=True)
def built():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostsim")], check=True, capture_output=True)


def run(binary, args):
    return subprocess.run([binary] + args, check=True, capture_output=True, text=True, cwd=GOLDEN).stdout


FAST_HUNT = ["cfg1_d0", "t1m_e1", "t1m_h1", "t1m_e0", "t1m_e1_fwd", "stress_e1", "stress_h1", "stress_e1_m7", "stress_h2_m50", "t1m_h2",
             "t1m_e1_x50", "t1m_e1_x150", "t1m_h2_x300", "t1m_h1_x40", "t1m_e2_x500", "t1m_h3", "t1m_d12"]


@pytest.mark.parametrize("case", FAST_HUNT)
def test_oracle_hunt_matches_reference(case, tmp_path):
    index = "t1m" if case.startswith(("t1m", "cfg1")) else "stress"
    flags = open(os.path.join(GOLDEN, case + ".flags.txt")).read().split()
    out = str(tmp_path / "rec.tsv")
    run(ORACLE, ["hunt", index + ".fm9", index + ".rec.tsv", case + ".queries.txt", "--records", out] + flags)
    want = "".join(l for l in open(os.path.join(GOLDEN, case + ".records.tsv")) if not l.startswith("W\t"))
    assert open(out).read() == want


def test_oracle_count_locate_seed_padcount():
    assert run(ORACLE, ["count", "stress.fm9", "stress.patterns.txt"]) == open(os.path.join(GOLDEN, "stress.count.tsv")).read()
    assert run(ORACLE, ["locate", "stress.fm9", "stress.patterns.txt"]) == open(os.path.join(GOLDEN, "stress.locate.tsv")).read()
    assert run(ORACLE, ["seed", "t1m.fm9", "t1m.rec.tsv", "t1m_seed.queries.txt", "-k", "15", "-d", "1"]) == \
        open(os.path.join(GOLDEN, "t1m_seed_k15_e1.seed.tsv")).read()
    assert run(ORACLE, ["seed", "t1m.fm9", "t1m.rec.tsv", "t1m_seed.queries.txt", "-k", "12", "-d", "1", "-n"]) == \
        open(os.path.join(GOLDEN, "t1m_seed_k12_h1.seed.tsv")).read()
    assert run(ORACLE, ["seed", "stress.fm9", "stress.rec.tsv", "stress_seed.queries.txt", "-k", "15", "-d", "1", "-m", "300"]) == \
        open(os.path.join(GOLDEN, "stress_seed_k15_e1.seed.tsv")).read()
    assert run(ORACLE, ["padcount", "stress.fm9", "arms.txt", "-d", "1"]) == open(os.path.join(GOLDEN, "stress_arms_e1.padcount.tsv")).read()
    assert run(ORACLE, ["padcount", "stress.fm9", "arms.txt", "-d", "1", "-n"]) == open(os.path.join(GOLDEN, "stress_arms_h1.padcount.tsv")).read()


def test_oracle_needle_and_neighbors():
    assert run(ORACLE, ["needle", "needle.pairs.tsv"]) == open(os.path.join(GOLDEN, "needle.out.tsv")).read()
    for tag, args in (("e1", ["-d", "1"]), ("h1", ["-d", "1", "-n"]), ("h2", ["-d", "2", "-n"])):
        want = gzip.open(os.path.join(GOLDEN, f"neighbors_{tag}.txt.gz"), "rt").read()
        assert run(ORACLE, ["neighbors", "neighbors.queries.txt", "-x", "1000000"] + args) == want


@pytest.mark.parametrize("tag,d,indel", [("e1", 1, 1), ("h1", 1, 0), ("h2", 2, 0), ("e2", 2, 1)])
def test_kernel_scripts_reproduce_neighbor_sets(tag, d, indel):
    """Edit scripts + the antichain rule of dg_core.cuh == std::set contents of neighbors();
    the packed 2-bit fast path is cross-checked against the byte path inside hostsim."""
    want = gzip.open(os.path.join(GOLDEN, f"neighbors_{tag}.txt.gz"), "rt").read()
    assert run(HOSTSIM, ["neighbors", "neighbors.queries.txt", str(d), str(indel)]) == want


@pytest.mark.parametrize("tag,d,indel,x", [("e1_x50", 1, 1, 50), ("e2_x500", 2, 1, 500), ("e2_x3000", 2, 1, 3000),
                                          ("h2_x300", 2, 0, 300), ("h1_x20", 1, 0, 20), ("e1", 1, 1, 1000000), ("h2", 2, 0, 1000000)])
def test_replay_reproduces_truncated_neighbor_sets(tag, d, indel, x):
    """nbr_trunc.hpp (the host replay behind DG_Q_NBR_CAP) == the sets neighbors() leaves when the cap -x
    stops its search (neighbors.h:50), and the untruncated ones."""
    want = gzip.open(os.path.join(GOLDEN, f"neighbors_{tag}.txt.gz"), "rt").read()
    assert run(HOSTSIM, ["replay", "neighbors.queries.txt", str(d), str(indel), str(x)]) == want


def test_pair_slot_index_arithmetic():
    """pair_slot of dg_core.cuh (slot -> first position, first kind, second position in k_probe_pairs /
    k_resolve) on every slot of every length up to 31, for 3 and 8 kinds per position."""
    assert run(HOSTSIM, ["pairslots"]).startswith("pair_slot checked on ")


def test_neighborhood_bound_is_sound():
    """nbr_upper_bound_part (the device-side certificate that the cap cannot be reached) never falls
    below the number of distinct strings, and certifies 20-mers at edit distance 2 under the default cap."""
    out = run(HOSTSIM, ["nbrbound", "400", "3"])
    assert out.startswith("bound >= distinct on 400 queries")
    assert int(out.rsplit("max ", 1)[1]) < 10000


def test_kernel_needle_reproduces_alignments():
    """needle_align of dg_core.cuh (both trace layouts) == needle() + the gap stripping of hunter.h:391-401."""
    got = run(HOSTSIM, ["needle", "needle.pairs.tsv"]).strip().split("\n")
    ref = open(os.path.join(GOLDEN, "needle.out.tsv")).read().strip().split("\n")
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        sc, r0, r1 = r.split("\t")
        last = len(r1) - 1
        for j, ch in enumerate(r1):
            if ch != "-":
                last = j
        lead, ra, qa, nlead = True, "", "", 0
        for j in range(last + 1):
            if r1[j] != "-":
                lead = False
            if not lead:
                ra += r0[j]; qa += r1[j]
            else:
                nlead += 1
        assert g == f"{sc}\t{nlead}\t{ra}\t{qa}"


@pytest.mark.parametrize("name", ["t1m", "stress"])
def test_select_supports_rebuilt_byte_for_byte(name):
    """fm9_select.hpp (used by dg_index_write_fm9) against the select_support_mcl sections SDSL wrote
    into the golden indexes: t1m takes SDSL's init_fast path, stress (< 100000 bits) init_slow."""
    out = run(HOSTSIM, ["select", name + ".fm9"])
    assert out.count("bytes identical") == 2


def test_thal_restatement_is_bit_exact():
    """dg_thal.cuh (the Tm gate of `dicey search`, primer3 thal in thal_end1 mode) compiled for the
    host: 709 (primer, site) pairs -- perfect sites, 1-3 mismatches, bulges, unrelated sequence,
    self-complementary oligos, N, lower case -- must give the 64-bit patterns the reference's own
    thal() produced (tests/golden/thal.out.tsv, written by `dicey_ref thal`)."""
    got = run(HOSTSIM, ["thal", "thal.params.tsv", "thal.pairs.tsv"])
    want = open(os.path.join(GOLDEN, "thal.out.tsv")).read()
    assert got == want
    temps = [float(l.split("\t")[1]) for l in want.splitlines()]
    assert len(temps) == 709 and sum(t > 45.0 for t in temps) > 300
    # the limits of the DP: both sides up to 60 bases, long bulges and internal loops
    assert run(HOSTSIM, ["thal", "thal.params.tsv", "thal_long.pairs.tsv"]) == open(os.path.join(GOLDEN, "thal_long.out.tsv")).read()


def test_thal_lane_cooperative_form_is_bit_exact():
    """thal_end1_tm_lanes -- the arrangement the GPU kernel runs (paired-cell list, LSH / RSH hoisted
    out of the fill, arg-min over loop partners instead of the sequential scan) -- with one lane on
    the host, against the reference's results for both golden sets (2209 pairs)."""
    for name in ("thal", "thal_long"):
        got = run(HOSTSIM, ["thal2", "thal.params.tsv", name + ".pairs.tsv"])
        assert got == open(os.path.join(GOLDEN, name + ".out.tsv")).read(), name


def test_thal_lane_groups_on_eight_concurrent_lanes():
    """The same template on eight lanes that really run side by side (threads; ballot, lane groups
    and the group arg-min built on a barrier): what the warp does on the GPU -- list compaction,
    one lane group per paired cell of a row, first-minimum selection -- must reproduce the
    reference's bits for the primer-like golden set and the first 300 pairs at the limits of the DP."""
    got = run(HOSTSIM, ["thal8", "thal.params.tsv", "thal.pairs.tsv"])
    assert got == open(os.path.join(GOLDEN, "thal.out.tsv")).read()
    # ... and on 32 lanes, the width of the warp k_thal_warp runs on
    assert run(HOSTSIM, ["thal32", "thal.params.tsv", "thal.pairs.tsv"]) == got
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".tsv", delete=False) as f:
        f.write("".join(open(os.path.join(GOLDEN, "thal_long.pairs.tsv")).readlines()[:300]))
        part = f.name
    try:
        got = run(HOSTSIM, ["thal8", "thal.params.tsv", part])
        assert got.splitlines() == open(os.path.join(GOLDEN, "thal_long.out.tsv")).read().splitlines()[:300]
    finally:
        os.unlink(part)


def test_thal_one_long_side_is_bit_exact():
    """thal() accepts one sequence longer than THAL_MAX_ALIGN = 60, up to THAL_MAX_SEQ = 10 000
    (thal.h:58, :2440-2451).  tests/golden/thal_wide.* holds 52 such pairs (oligo against a target of
    61 .. 10 000 bases on either side, exact / mismatched / bulged sites, N, lower case, and the four
    kinds of pair the reference refuses) with the reference's own results: the sequential form
    (thal_end1_tm_any) and the arrangement k_thal_wide runs (thal_end1_tm_wide: live cells by ballot,
    loop partners by coordinates, table in the caller's memory) on one lane must give those bits --
    and the wide arrangement must also reproduce all 2209 ordinary pairs."""
    want = open(os.path.join(GOLDEN, "thal_wide.out.tsv")).read()
    assert [l.split("\t")[0] for l in want.splitlines()].count("0") == 4
    assert run(HOSTSIM, ["thalany", "thal.params.tsv", "thal_wide.pairs.tsv"]) == want
    assert run(HOSTSIM, ["thalw", "thal.params.tsv", "thal_wide.pairs.tsv"]) == want
    for name in ("thal", "thal_long"):
        assert run(HOSTSIM, ["thalw", "thal.params.tsv", name + ".pairs.tsv"]) == open(os.path.join(GOLDEN, name + ".out.tsv")).read(), name


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "dicey_ref")


@pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.isdir("/root/reference/src/primer3_config")),
                    reason="needs the compiled reference and its primer3_config directory")
def test_thal_one_long_side_against_the_reference_on_fresh_pairs(tmp_path):
    """Beyond the committed fixture: 240 freshly drawn pairs with one side of 61 .. 600 bases (the long
    side first or second, sites planted or not, poly-runs, N) go through the reference's own thal()
    (oracle/_ref/dicey_ref, this container only) and through both host forms; all three must print
    the same bits."""
    import numpy as np
    rng = np.random.default_rng(99)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    lines = []
    for i in range(240):
        k, L = int(rng.integers(1, 61)), int(rng.integers(61, 601))
        oligo = bytes(acgt[rng.integers(0, 4, k)]) if i % 9 else bytes([acgt[i % 4]]) * k
        target = bytearray(acgt[rng.integers(0, 4, L)])
        if i % 3:
            site = bytearray(oligo.translate(comp)[::-1])
            for _ in range(i % 4):
                site[int(rng.integers(0, len(site)))] = acgt[rng.integers(0, 4)]
            at = int(rng.integers(0, L - len(site) + 1)) if i % 5 else L - len(site)
            target[at:at + len(site)] = site
        if i % 17 == 0:
            target[int(rng.integers(0, L))] = ord("N")
        a, b = (oligo, bytes(target)) if i % 2 else (bytes(target), oligo)
        lines.append(a + b"\t" + b + b"\n")
    pairs = tmp_path / "pairs.tsv"
    pairs.write_bytes(b"".join(lines))
    want = subprocess.run([REF_BIN, "thal", "/root/reference/src/primer3_config/", str(pairs), "/dev/null"], check=True,
                          capture_output=True, text=True).stdout
    assert len(want.splitlines()) == 240 and all(l.startswith("1\t") for l in want.splitlines())
    assert len({l.split("\t")[2] for l in want.splitlines()}) > 150           # (not all the same trivial answer)
    assert run(HOSTSIM, ["thalw", "thal.params.tsv", str(pairs)]) == want
    assert run(HOSTSIM, ["thalany", "thal.params.tsv", str(pairs)]) == want


def test_thal_wide_form_on_a_full_warp_of_concurrent_lanes():
    """thal_end1_tm_wide on 32 lanes that really run side by side (threads; the collectives on a
    barrier): ballot of the live cells of a 32-column chunk, one lane group per live cell, partners
    split over the lanes of a group, first-minimum selection.  The primer-like golden set (709
    pairs, rows of 15-40 columns: groups of every size) and the first 20 pairs with a long side."""
    got = run(HOSTSIM, ["thalw32", "thal.params.tsv", "thal.pairs.tsv"])
    assert got == open(os.path.join(GOLDEN, "thal.out.tsv")).read()
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".tsv", delete=False) as f:
        f.write("".join(open(os.path.join(GOLDEN, "thal_wide.pairs.tsv")).readlines()[:20]))
        part = f.name
    try:
        got = run(HOSTSIM, ["thalw32", "thal.params.tsv", part])
        assert got.splitlines() == open(os.path.join(GOLDEN, "thal_wide.out.tsv")).read().splitlines()[:20]
    finally:
        os.unlink(part)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/primer3_config"), reason="needs the reference's primer3_config directory")
def test_thal_config_loader_reproduces_the_reference_tables():
    """thal_params_from_config (what the product reads: dicey's -i directory) against the tables the
    reference held after get_thermodynamic_values() (the committed dump)."""
    out = run(HOSTSIM, ["thalcfg", "/root/reference/src/primer3_config/", "thal.params.tsv"])
    assert "tables identical" in out


def test_json_double_format_matches_nlohmann():
    """dicey_b200/host/jsonnum.hpp (Grisu2 + nlohmann's formatting rules) on 31 k doubles: the Tm /
    penalty values of the golden search outputs, random values over 60 orders of magnitude, random
    bit patterns, edge cases -- against nlohmann::json(double).dump() (tests/golden/jsonfloat.out.txt)."""
    assert run(HOSTSIM, ["jsonfloat", "jsonfloat.hex.txt"]) == open(os.path.join(GOLDEN, "jsonfloat.out.txt")).read()


def test_primer3_config_rebuilt_from_the_dump_loads_identically(tmp_path):
    """tests/util.write_primer3_config (the -i directory the search tests hand to dicey-b200 on the GPU
    box) -> thal_params_from_config -> the same tables, bit for bit."""
    from util import write_primer3_config
    d = write_primer3_config(str(tmp_path / "p3"))
    assert "tables identical" in run(HOSTSIM, ["thalcfg", d + "/", "thal.params.tsv"])
