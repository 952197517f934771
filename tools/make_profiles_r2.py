#!/usr/bin/env python3
"""Turns the raw outputs of tools/gpu_final_r2.sh (gpurun_out/final_r2/) into the tracked files under profiles/:
   python tools/make_profiles_r2.py"""
import csv
import io
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

SRC, DST = ROOT / "gpurun_out" / "final_r2", ROOT / "profiles"


def last_json(path):
    return json.loads([l for l in path.read_text().splitlines() if l.startswith("{")][-1])


lines = {}
for f in sorted(SRC.glob("bench_*.json")):
    try:
        lines[f.stem[len("bench_"):]] = last_json(f)
    except Exception as e:  # noqa: BLE001
        print("skip", f.name, e)
(DST / "r02_bench_fbm512.json").write_text(json.dumps(lines["fbm512"], indent=1) + "\n")
(DST / "r02_bench_all_workloads.json").write_text(json.dumps(lines, indent=1) + "\n")
shutil.copy(SRC / "launches_fbm512.csv", DST / "r02_launches_fbm512.csv")
shutil.copy(SRC / "chunks.jsonl", DST / "r02_chunks.jsonl")

traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full --clock-control none` captures "
                       "(profiles/r02_*_ncu_full_summary.txt) of the kernel sources whose hash is src_sha; bench.py copies the dominant "
                       "kernel's figure into roofline.traffic only while its own source hash matches"}
sha = bench.kernel_source_sha()
for wl in ("fbm512", "gyroid1024", "spheres2048"):
    rep = SRC / ("%s_full.ncu-rep" % wl)
    if not rep.exists():
        continue
    txt = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(rep)], capture_output=True, text=True).stdout
    (DST / ("r02_%s_ncu_full_summary.txt" % wl)).write_text(txt)
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kern = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"]
        short = "k_sign" if "k_sign" in name else "k_count_list" if "k_count_list" in name else "k_scan_rows" if "k_scan_rows" in name else \
            "k_emit_list" if "k_emit_list" in name else name[:20]

        def b(key):
            v, u = float(d[key].replace(",", "")), units[hdr.index(key)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        kern[short] = int(b("dram__bytes_read.sum") + b("dram__bytes_write.sum"))
    traffic[wl] = {"src_sha": sha, "kernels": kern}
(DST / "traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
print("src sha", sha, {k: v.get("kernels") for k, v in traffic.items() if k != "_comment"})
for k, d in lines.items():
    print("%-12s %8.4f ms  %8.2f %s  extract frac %s  e2e %s" % (k, d.get("ms_per_step", 0), d.get("value", 0), d.get("unit"),
                                                               (d.get("roofline_extract") or {}).get("frac"), (d.get("e2e") or {}).get("ms_per_step")))
