#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an .ncu-rep captured with --import-source on.

    python tools/ncu_hot_lines.py gpurun_out/x.ncu-rep k_verify [top]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}"],
                         check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    fname, hdr, agg, cur = "", None, {}, None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[0] == "Line No":
            hdr = r
            ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        if r[0] != "":
            cur = (fname, r[0], r[1])
            agg.setdefault(cur, [0.0, 0.0])
        elif cur is not None:
            try:
                agg[cur][0] += float(r[ie]); agg[cur][1] += float(r[sm])
            except ValueError:
                pass
    tot = sum(v[0] for v in agg.values()) or 1.0
    tots = sum(v[1] for v in agg.values()) or 1.0
    print(f"# {kern}: {tot:.3g} warp instructions, {tots:.0f} stall samples")
    for (f, ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * n / tot:5.1f}% inst {100 * s / tots:5.1f}% samples  {f}:{ln}  {src.strip()[:100]}")


if __name__ == "__main__":
    main()
