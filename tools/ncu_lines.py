#!/usr/bin/env python
"""Aggregate an ncu report's warp-stall samples and executed instructions per CUDA source line.

usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [top_n]
(needs the kernel to be compiled with -lineinfo and captured with --import-source on)
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    per = defaultdict(lambda: [0, 0, ""])
    cur_file = ""
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if "# Samples" in r and "Line No" in r:
            hdr = r
            li, sa, ie = r.index("Line No"), r.index("# Samples"), r.index("Instructions Executed")
            si = r.index("Source")
            continue
        if hdr is None or len(r) <= ie or not r[li]:
            continue  # SASS rows have an empty line number; CUDA rows carry the per-line totals
        try:
            smp, ins = int(r[sa]), int(r[ie])
        except ValueError:
            continue
        key = (cur_file, r[li])
        per[key][0] += smp
        per[key][1] += ins
        per[key][2] = r[si].strip()[:90]
    if hdr is None:
        print(out[:2000])
        return
    tot = sum(v[0] for v in per.values()) or 1
    toti = sum(v[1] for v in per.values()) or 1
    print("total samples %d, instructions %d" % (tot, toti))
    for key, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%6.2f%% samples %6.2f%% inst  %s:%s  %s" % (100.0 * v[0] / tot, 100.0 * v[1] / toti, key[0], key[1], v[2]))


if __name__ == "__main__":
    main()
