#!/bin/bash
# quick perf iteration: tile benches + per-line ncu of the tile kernels at fbm512 (and optionally spheres2048)
o=gpurun_out/r2; mkdir -p $o
tag=${1:-q}
for wl in fbm512 spheres2048 gyroid1024; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu > $o/bench_${tag}_$wl.json 2> $o/bench_${tag}_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("$o/bench_${tag}_$wl.json")); k=d.get("kernels_ms",{}); print("$tag $wl", round(d["ms_per_step"],4), {a:round(b,4) for a,b in k.items()}, "extract frac", round(d.get("roofline_extract",{}).get("frac",0),4), "V", d["config"]["vertices"], "T", d["config"]["triangles"])
except Exception as e: print("$tag $wl FAILED", e); print(open("$o/bench_${tag}_$wl.err").read()[-1500:])
PY
done
timeout 300 python tools/gpu_check.py > $o/check_$tag.log 2>&1; tail -1 $o/check_$tag.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tile" -s 6 -c 2 -o $o/fbm512_$tag -f \
      python bench.py --workload fbm512 --steps 2 --warmup 3 --no-e2e --no-cpu > $o/ncu_$tag.log 2>&1
if [ -n "$2" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tile" -s 6 -c 2 -o $o/spheres2048_$tag -f \
      python bench.py --workload spheres2048 --steps 2 --warmup 3 --no-e2e --no-cpu > $o/ncu2048_$tag.log 2>&1
fi
ls -la $o/*_$tag.ncu-rep
