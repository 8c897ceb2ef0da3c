import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np
from dicey_b200 import synth
from dicey_b200.api import Index, HuntParams
ix = Index.build_synthetic(42, 24, 125_000_000, 0)
for n in (1000, 50000, 250000):
    pr = synth.primers_fast(42, 24, 125_000_000, n, 20, 1, True, rng_seed=7)
    for rep in range(4):
        t = time.perf_counter(); r = ix.hunt(pr, HuntParams(distance=1)); dt = time.perf_counter() - t
        print("nq", n, "ms", round(dt * 1e3, 3), "hits", len(r.hits), file=sys.stderr)
