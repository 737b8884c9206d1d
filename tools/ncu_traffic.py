#!/usr/bin/env python
"""Extract per-launch DRAM traffic and duration of the kernels of an ncu --set full capture into profiles/r2_traffic.json.

usage: python tools/ncu_traffic.py report.ncu-rep n p   (n, p = the workload shape the capture was taken on)"""
import csv
import io
import json
import os
import subprocess
import sys

rep, n, p = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
res = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    name = d["Kernel Name"]
    key = "sweep_pipe_kernel" if "sweep_pipe" in name else "gram_fp4_kernel" if "gram_fp4" in name else "gram_tc_kernel" if "gram_tc" in name else "epilogue_kernel" if "epilogue" in name else None
    if key is None or key in res:
        continue

    def val(k):
        v = float(d[k].replace(",", ""))
        unit = u[k].lower()
        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
        return v * mult
    res[key] = {"n": n, "p": p, "bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                "dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum"),
                "duration_ns_under_ncu": val("gpu__time_duration.sum"), "kernel": name[:80]}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "r2_traffic.json")
json.dump(res, open(path, "w"), indent=1)
print(json.dumps(res, indent=1))
