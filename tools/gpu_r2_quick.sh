#!/bin/bash
# quick perf iteration: tile benches + per-line ncu of the tile kernels at fbm512 (and optionally spheres2048)
# usage: gpu_r2_quick.sh TAG [ncu2048] ; extra env variants via VARIANTS="name:ENV=val,ENV2=val name2:..."
o=gpurun_out/r2; mkdir -p $o
tag=${1:-q}
one() { # tag workload env...
  local t=$1 wl=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu --no-strong > $o/bench_${t}_$wl.json 2> $o/bench_${t}_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("$o/bench_${t}_$wl.json")); k=d.get("kernels_ms",{}); print("$t $wl", round(d["ms_per_step"],4), {a:round(b,4) for a,b in k.items()}, "extract frac", round(d.get("roofline_extract",{}).get("frac",0),4), "V", d["config"]["vertices"], "T", d["config"]["triangles"])
except Exception as e: print("$t $wl FAILED", e); print(open("$o/bench_${t}_$wl.err").read()[-1500:])
PY
}
for wl in fbm512 spheres2048 gyroid1024; do one $tag $wl ISOMC_PATH=tile; done
for v in $VARIANTS; do
  name=${v%%:*}; envs=${v#*:}
  for wl in fbm512 spheres2048; do one ${tag}_$name $wl $(echo $envs | tr ',' ' '); done
done
timeout 300 python tools/gpu_check.py > $o/check_$tag.log 2>&1; tail -1 $o/check_$tag.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tile" -s 6 -c 2 -o $o/fbm512_$tag -f \
      python bench.py --workload fbm512 --steps 2 --warmup 3 --no-e2e --no-cpu --no-strong > $o/ncu_$tag.log 2>&1
if [ -n "$2" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tile" -s 6 -c 2 -o $o/spheres2048_$tag -f \
      python bench.py --workload spheres2048 --steps 2 --warmup 3 --no-e2e --no-cpu --no-strong > $o/ncu2048_$tag.log 2>&1
fi
ls -la $o/*_$tag.ncu-rep
