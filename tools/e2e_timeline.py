"""Wall time of the end-to-end hunt call (host buffers in, host records out) on the headline
workload; DG_TRACE=1 python tools/e2e_timeline.py makes the pipeline print its per-chunk timeline.
DG_PKG_ROOT=<dir> loads the package (and its library) from another tree for A/B runs."""
import os, sys, time
sys.path.insert(0, os.environ.get("DG_PKG_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dicey_b200 import synth
from dicey_b200.api import Index, HuntParams
ix = Index.build_synthetic(42, 24, 125_000_000, 0)
pr = synth.primers_fast(42, 24, 125_000_000, 1_000_000, 20, 1, True, rng_seed=7)
if os.environ.get("DG_PINNED_INPUT", "1") != "0":   # what bench.py's end-to-end arm passes: page-locked host buffers
    import torch
    pin = torch.from_numpy(pr.reshape(-1)).pin_memory()
    off = torch.from_numpy((np.arange(pr.shape[0] + 1, dtype=np.uint64) * np.uint64(pr.shape[1])).view(np.int64)).pin_memory()
    pr = (pin.numpy(), off.numpy().view(np.uint64))
ms = []
for rep in range(14):
    t = time.perf_counter(); r = ix.hunt(pr, HuntParams(distance=1)); dt = time.perf_counter() - t
    ms.append(dt * 1e3)
    nh = len(r.hits)
    del r
print(sys.argv[1] if len(sys.argv) > 1 else "", "calls ms:", " ".join(f"{x:.2f}" for x in ms[4:]), "| median", round(float(np.median(ms[4:])), 2), "hits", nh, file=sys.stderr)
