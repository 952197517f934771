#!/bin/bash
# perf sweep over env variants: gpu_sweep.sh TAG "wl1 wl2" "name:ENV=val,ENV2=val name2:..."
o=gpurun_out/r2; mkdir -p $o
tag=${1:-s}; wls=${2:-fbm512}; variants=${3:-base:X=0}
for v in $variants; do
  name=${v%%:*}; envs=${v#*:}
  for wl in $wls; do
    env $(echo $envs | tr ',' ' ') timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu --no-strong > $o/bench_${tag}_${name}_$wl.json 2> $o/bench_${tag}_${name}_$wl.err
    python - <<PY
import json
try:
    d=json.load(open("$o/bench_${tag}_${name}_$wl.json")); k=d.get("kernels_ms",{}); print("$tag $name $wl", round(d["ms_per_step"],4), {a:round(b,4) for a,b in k.items()}, "frac", round(d.get("roofline_extract",{}).get("frac",0),4), "V", d["config"]["vertices"], "T", d["config"]["triangles"])
except Exception as e: print("$tag $name $wl FAILED", e); print(open("$o/bench_${tag}_${name}_$wl.err").read()[-1500:])
PY
  done
done
