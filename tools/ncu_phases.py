#!/usr/bin/env python3
"""Per-phase instruction/stall-sample shares of k_emit from an ncu report (phases = source line ranges
delimited by marker comments):  python tools/ncu_phases.py report.ncu-rep"""
import csv
import io
import subprocess
import sys
from pathlib import Path

rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "k_emit"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = hdr = None
agg, samp = {}, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 8 and r[0].strip().isdigit() and r[2] == "-":
        d = dict(zip(hdr, r))
        k = (cur, int(r[0]))
        agg[k] = agg.get(k, 0) + int(d["Instructions Executed"])
        samp[k] = samp.get(k, 0) + int(d["Warp Stall Sampling (All Samples)"])
tot, ts = sum(agg.values()), sum(samp.values())
src = (Path(__file__).resolve().parent.parent / "isosurface_b200/csrc/isomc_kernels.cu").read_text().split("\n")
keys = ["K4: emission", "row prefixes of the region rows", "stage the sign words", "for (uint32_t bx = 0", "P1: one thread",
        "expansion: one store", "P2: one thread", "interior: creates exactly", "on a low boundary face", "triangle-list position",
        "B: one thread per tri", "vertex descriptor (written by"]
marks = [(i, k) for i, l in enumerate(src, 1) for k in keys if k in l]
f = "isomc_kernels.cu"
rng = lambda a, b, dd: sum(v for (ff, l), v in dd.items() if ff == f and a <= l <= b)
print("total warp-instr %d, samples %d" % (tot, ts))
for (a, ka), (b, kb) in zip(marks, marks[1:]):
    print("%-34s lines %4d-%4d  inst %10d %5.1f%%  samples %5.1f%%" % (ka, a, b - 1, rng(a, b - 1, agg), 100 * rng(a, b - 1, agg) / tot,
                                                                       100 * rng(a, b - 1, samp) / ts))
print("helpers (< K4): inst %.1f%% samples %.1f%%" % (100 * rng(0, marks[0][0] - 1, agg) / tot, 100 * rng(0, marks[0][0] - 1, samp) / ts))
oth = sum(v for (ff, l), v in agg.items() if ff != f)
print("other files: inst %.1f%%" % (100 * oth / tot))
