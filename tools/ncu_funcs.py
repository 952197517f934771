#!/usr/bin/env python3
"""Warp-instruction and stall-sample shares per source FUNCTION of a kernel, from an ncu report captured with
--import-source on:  python tools/ncu_funcs.py report.ncu-rep kernel_regex
(lines are attributed to the function whose definition encloses them in the CURRENT source tree; inlined frames count once)"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

rep, kern = sys.argv[1], sys.argv[2]
ROOT = Path(__file__).resolve().parent.parent / "isosurface_b200" / "csrc"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))


def func_map(path):
    m, cur = {}, "?"
    pat = re.compile(r"^(?:template.*>\s*)?(?:ISOMC_HD|__global__|__device__|static|inline|cudaError_t|template)?.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;]*$")
    for i, l in enumerate(path.read_text().split("\n"), 1):
        if l and not l[0].isspace() and "(" in l and not l.startswith(("#", "//", "/*", " *", "}")):
            g = pat.match(l)
            if g:
                cur = g.group(1)
        m[i] = cur
    return m


maps = {}
cur, hdr = None, None
inst, samp = Counter(), Counter()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 8 and r[0].strip().isdigit() and r[2] == "-":
        d = dict(zip(hdr, r))
        p = ROOT / cur
        if p.exists():
            if cur not in maps:
                maps[cur] = func_map(p)
            key = "%s:%s" % (cur, maps[cur].get(int(r[0]), "?"))
        else:
            key = cur
        inst[key] += int(d["Instructions Executed"])
        samp[key] += int(d["Warp Stall Sampling (All Samples)"])
ti, ts = sum(inst.values()), max(sum(samp.values()), 1)
print("attributed warp-instructions %d (inlined frames are listed under every enclosing line), stall samples %d" % (ti, ts))
for k, v in inst.most_common(25):
    print("%-46s inst %5.1f%%   samples %5.1f%%" % (k, 100 * v / ti, 100 * samp[k] / ts))
