#!/usr/bin/env python3
"""Aggregates an ncu report's per-instruction counts by CUDA source line:
   python tools/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, agg, hdr = None, [], None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 8 and r[0].strip().isdigit() and r[2] == "-":
        try:
            d = dict(zip(hdr, r))
            agg.append((int(d["Instructions Executed"]), int(d["Warp Stall Sampling (All Samples)"]), cur, int(r[0]), r[1].strip()[:100]))
        except Exception:
            pass
tot = sum(a[0] for a in agg)
ts = sum(a[1] for a in agg)
print("total warp-instructions", tot, "stall samples", ts)
for a in sorted(agg, reverse=True)[:top]:
    print("%10d %5.1f%% | samp %5.1f%% | %s:%d  %s" % (a[0], 100 * a[0] / tot, 100 * a[1] / max(ts, 1), a[2], a[3], a[4]))
