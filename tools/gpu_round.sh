#!/bin/bash
# One GPU box visit: ladder check, GPU parity suite, benches of the three grid workloads (list path), brick path A/B.
mkdir -p gpurun_out
( timeout 300 python tools/gpu_check.py ) > gpurun_out/check_list.log 2>&1; echo "check rc=$?" | tee -a gpurun_out/check_list.log
tail -3 gpurun_out/check_list.log
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_list.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_list.log
for wl in fbm512 gyroid1024 spheres2048; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_list_$wl.json 2> gpurun_out/bench_list_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_list_$wl.json")); print("$wl list", round(d["ms_per_step"],4), d.get("kernels_ms"), d["roofline_extract"]["frac"])
except Exception as e: print("$wl list FAILED", e)
PY
done
ISOMC_EMIT=brick timeout 300 python bench.py --workload fbm512 --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_brick_fbm512.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_brick_fbm512.json')); print('fbm512 brick', round(d['ms_per_step'],4), d.get('kernels_ms'))"
