#!/usr/bin/env python
"""dg_thal_batch on pairs with one side longer than THAL_MAX_ALIGN (k_thal_wide) against the 64-bit
patterns the reference's own thal() produced (tests/golden/thal_wide.*), without importing torch
(a short GPU call): the wide pairs alone, mixed into a batch of ordinary pairs, and with the
sequential redo path forced (DG_THAL_SEQ=2).  Prints one line per check and exits non-zero on a
mismatch.

    python tools/thal_wide_check.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dicey_b200.api import Thal  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def load(name):
    pairs = [l.rstrip("\n").split("\t") for l in open(os.path.join(GOLDEN, name + ".pairs.tsv"))]
    want = [l.split("\t") for l in open(os.path.join(GOLDEN, name + ".out.tsv")).read().splitlines()]
    return pairs, want


def check(th, pairs, want, label):
    t0 = time.time()
    tm, ok = th.tm([p[0] for p in pairs], [p[1] for p in pairs])
    dt = time.time() - t0
    bits = tm.view(np.uint64)
    bad = [i for i, w in enumerate(want) if int(ok[i]) != int(w[0]) or int(bits[i]) != int(w[2], 16)]
    print(f"{label}: {len(pairs)} pairs in {dt:.3f} s, {len(bad)} mismatches", flush=True)
    for i in bad[:8]:
        print(f"   pair {i} ({len(pairs[i][0])} x {len(pairs[i][1])}): got ok={int(ok[i])} tm={float(tm[i])!r}, want {want[i][0]} {want[i][1]}")
    return not bad


def main():
    wide, wide_want = load("thal_wide")
    short, short_want = load("thal")
    short, short_want = short[:300], short_want[:300]
    th = Thal.open_tables(os.path.join(GOLDEN, "thal.params.tsv"), 0)
    good = True
    try:
        good &= check(th, wide, wide_want, "wide pairs")
        mixed, mixed_want = [], []
        for i in range(max(len(wide), len(short))):   # interleaved: the dispatch must route each pair by its lengths
            if i < len(short):
                mixed.append(short[i]); mixed_want.append(short_want[i])
            if i < len(wide):
                mixed.append(wide[i]); mixed_want.append(wide_want[i])
        good &= check(th, mixed, mixed_want, "mixed batch")
        os.environ["DG_THAL_SEQ"] = "2"
        good &= check(th, mixed, mixed_want, "mixed batch, DG_THAL_SEQ=2")
        os.environ["DG_THAL_SEQ"] = "1"
        good &= check(th, mixed, mixed_want, "mixed batch, DG_THAL_SEQ=1")
        os.environ.pop("DG_THAL_SEQ")
    finally:
        th.close()
    print("THAL_WIDE_CHECK", "PASS" if good else "FAIL", flush=True)
    return 0 if good else 1


if __name__ == "__main__":
    sys.exit(main())
