#!/usr/bin/env python3
"""Text summary of an ncu report for profiles/:  python tools/ncu_summary.py report.ncu-rep > profiles/xyz.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
print("# ncu --set full --clock-control none summary of", rep.split("/")[-1])
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("\n## " + d.get("Kernel Name", "?")[:110])
    for k in KEYS:
        if k in d:
            print("  %-72s %s %s" % (k, d[k], units[hdr.index(k)]))
    st = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
            try:
                st.append((float(r[i]), h.split("stalled_")[1].replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("  warp stall cycles per issued instruction: " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:6]))
