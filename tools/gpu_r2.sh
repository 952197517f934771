#!/bin/bash
# Round-2 GPU box visit.  usage: gpu_r2.sh [check] [pytest] [bench] [ncu]   (everything under its own timeout; logs in gpurun_out/r2/)
o=gpurun_out/r2; mkdir -p $o
has() { for a in "$@"; do [ "$a" = "$want" ] && return 0; done; return 1; }
bench() { # tag workload env...
  local tag=$1 wl=$2; shift 2
  env "$@" timeout 400 python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu > $o/bench_${tag}_$wl.json 2> $o/bench_${tag}_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("$o/bench_${tag}_$wl.json")); k=d.get("kernels_ms",{}); print("$tag $wl", round(d["ms_per_step"],4), {a:round(b,4) for a,b in k.items()}, "extract frac", round(d.get("roofline_extract",{}).get("frac",0),4))
except Exception as e: print("$tag $wl FAILED", e); print(open("$o/bench_${tag}_$wl.err").read()[-1500:])
PY
}
for want in "$@"; do
case $want in
check)
  ( timeout 300 python tools/gpu_check.py ) > $o/check.log 2>&1; echo "check rc=$?"; tail -3 $o/check.log;;
pytest)
  ( timeout 1500 python -m pytest tests -m gpu -x -q ) > $o/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $o/pytest_gpu.log;;
bench)
  for wl in fbm512 gyroid1024 spheres2048; do bench tile $wl ISOMC_PATH=tile; done
  bench plain fbm512 ISOMC_PATH=tile ISOMC_FILL=plain
  for wl in fbm512 spheres2048; do bench list $wl ISOMC_PATH=list; done
  for wl in torus256 csga256 sphere32 torus128; do bench tile $wl ISOMC_PATH=tile; done;;
benchfull)
  timeout 600 python bench.py > $o/bench_default.json 2> $o/bench_default.err; tail -c 600 $o/bench_default.json;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 12 --csv --log-file $o/launches_fbm512.csv \
      python bench.py --workload fbm512 --steps 2 --warmup 3 --no-e2e --no-cpu > $o/bench_under_ncu.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tile|k_scan" -s 9 -c 3 -o $o/fbm512_full -f \
      python bench.py --workload fbm512 --steps 2 --warmup 3 --no-e2e --no-cpu >> $o/bench_under_ncu.log 2>&1
  ls -la $o/*.ncu-rep;;
ncu2048)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tile|k_scan" -s 9 -c 3 -o $o/spheres2048_full -f \
      python bench.py --workload spheres2048 --steps 2 --warmup 3 --no-e2e --no-cpu > $o/bench_under_ncu2048.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tile|k_scan" -s 9 -c 3 -o $o/gyroid1024_full -f \
      python bench.py --workload gyroid1024 --steps 2 --warmup 3 --no-e2e --no-cpu > $o/bench_under_ncu1024.log 2>&1
  ls -la $o/*.ncu-rep;;
esac
done
