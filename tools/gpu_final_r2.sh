#!/bin/bash
# Round-2 record on one B200: GPU parity suite, default bench line (+ strong records), reference arm, the other workloads, the
# ncu launch list of the bench command, one full capture per benchmark lattice, chunk throughput.  Outputs: gpurun_out/final_r2/.
o=gpurun_out/final_r2; mkdir -p $o
( timeout 1500 python -m pytest tests -m gpu -q ) > $o/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $o/pytest_gpu.log
timeout 900 python bench.py > $o/bench_fbm512.json 2> $o/bench_fbm512.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $o/bench_ref.json 2> $o/bench_ref.err; echo "ref rc=$?"
for wl in gyroid1024 spheres2048 spheres512 gyroid512; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-strong > $o/bench_$wl.json 2> $o/bench_$wl.err
done
for wl in torus256 csga256 csgb256 sphere32 torus128; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --cpu-seconds 3 --no-strong > $o/bench_$wl.json 2> $o/bench_$wl.err
done
timeout 300 python tools/bench_chunks.py > $o/chunks.jsonl 2> $o/chunks.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_fbm512.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-strong > $o/bench_under_ncu.log 2>&1
for wl in fbm512 gyroid1024 spheres2048; do
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_sign|k_count_list|k_scan_rows|k_emit_list" --launch-skip 16 --launch-count 4 -f -o $o/${wl}_full python tools/prof_one.py $wl 6 > $o/prof_$wl.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/final_r2/bench_*.json")):
    try:
        txt=[l for l in open(f).read().splitlines() if l.startswith("{")][-1]; d=json.loads(txt)
        print(f.split("/")[-1], d.get("config",{}).get("workload"), round(d.get("ms_per_step",0),4), round(d.get("value",0),2), (d.get("roofline_extract") or d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("ms_per_step"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "FAILED", e)
PY
ls -la $o/*.ncu-rep
