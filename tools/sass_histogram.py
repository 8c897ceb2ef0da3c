#!/usr/bin/env python
"""Instruction histogram of the hot kernels from the built object (cuobjdump -sass): what the sm_100a code
of this integer / gather path is made of (no tensor-core or TMA instructions are expected: every access
goes to a different, data-dependent sector).

    python tools/sass_histogram.py dicey_b200/csrc/dg_search.o k_probe_singlesILb1 k_resolveILb1 k_verifyILb1 > profiles/r02_sass_histogram.txt
"""
import collections
import re
import subprocess
import sys

obj, wanted = sys.argv[1], sys.argv[2:]
out = subprocess.run(["cuobjdump", "-sass", obj], check=True, capture_output=True, text=True).stdout
cur, hist = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next((w for w in wanted if w in m.group(1)), None)
        if cur:
            hist.setdefault(cur, collections.Counter())
        continue
    if cur:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            hist[cur][m.group(1)] += 1
for k in wanted:
    h = hist.get(k, {})
    total = sum(h.values())
    print(f"== {k}: {total} SASS instructions")
    groups = collections.Counter()
    for op, n in h.items():
        groups[op.split(".")[0]] += n
    print("   by opcode: " + ", ".join(f"{op} {n}" for op, n in groups.most_common(24)))
    mem = {op: n for op, n in h.items() if op.startswith(("LDG", "STG", "LDS", "STS", "LDL", "STL", "ATOM", "RED", "LDC", "UTMA", "UBLKCP", "TCGEN", "UTC"))}
    print("   memory / special: " + ", ".join(f"{op} {n}" for op, n in sorted(mem.items(), key=lambda x: -x[1])))
