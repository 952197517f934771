#!/bin/bash
# One GPU box visit: ladder check + GPU parity suite on the list path, then benches.  usage: gpu_round.sh [variants...]
mkdir -p gpurun_out
export ISOMC_EMIT=list
( timeout 300 python tools/gpu_check.py ) > gpurun_out/check_list.log 2>&1; echo "check rc=$?"
tail -2 gpurun_out/check_list.log
if [ -z "$SKIP_PYTEST" ]; then
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_list.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_list.log
fi
run() { # name env...
  local tag=$1; shift
  for wl in fbm512 gyroid1024 spheres2048; do
    env "$@" timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${tag}_$wl.json 2> gpurun_out/bench_${tag}_$wl.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${tag}_$wl.json")); k=d.get("kernels_ms",{}); print("$tag $wl", round(d["ms_per_step"],4), {a:round(b,4) for a,b in k.items()}, round(d["roofline_extract"]["frac"],4))
except Exception as e: print("$tag $wl FAILED", e)
PY
  done
}
for v in "$@"; do
  case $v in
    list4) run list4 ISOMC_EMIT=list ISOMC_LIST_MINB=4;;
    list5) run list5 ISOMC_EMIT=list ISOMC_LIST_MINB=5;;
    list6) run list6 ISOMC_EMIT=list ISOMC_LIST_MINB=6;;
    brick) run brick ISOMC_EMIT=brick;;
    pipe) run pipe ISOMC_EMIT=list ISOMC_PIPELINE=1;;
    emit4) run emit4 ISOMC_EMIT=list ISOMC_EMIT_MINB=4;;
    emit5) run emit5 ISOMC_EMIT=list ISOMC_EMIT_MINB=5;;
    emit6) run emit6 ISOMC_EMIT=list ISOMC_EMIT_MINB=6;;
  esac
done
