"""Throughput of dg_thal_batch on packed inputs (no Python list handling in the timed region)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from dicey_b200.api import Thal, pack_sequences

pairs = [l.rstrip("\n").split("\t") for l in open("tests/golden/thal.pairs.tsv")]
pairs = [p for p in pairs if 18 <= len(p[0]) <= 25][:500]
th = Thal.open_tables("tests/golden/thal.params.tsv", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = n // len(pairs)
b1, o1 = pack_sequences([p[0] for p in pairs] * reps)
b2, o2 = pack_sequences([p[1] for p in pairs] * reps)
for _ in range(3):
    t = time.time(); tm, ok = th.tm((b1, o1), (b2, o2)); dt = time.time() - t
    print(f"{len(o1) - 1} pairs in {dt * 1e3:.1f} ms -> {(len(o1) - 1) / dt / 1e6:.3f} M pairs/s", flush=True)
