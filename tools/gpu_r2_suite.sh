#!/bin/bash
# full GPU suite visit: golden generation (CPU-heavy, in the background), pytest -m gpu, default bench line, reference arm
o=gpurun_out/r2; mkdir -p $o
if [ -n "$GOLDEN" ]; then
  ( timeout 1500 python tools/gen_golden_full.py $o/full_hashes.json $GOLDEN > $o/golden.log 2>&1; echo "golden rc=$?" >> $o/golden.log ) &
  gpid=$!
fi
( timeout 1800 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} ) > $o/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $o/pytest_gpu.log
if [ -z "$NOBENCH" ]; then
timeout 900 python bench.py > $o/bench_default.json 2> $o/bench_default.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open("$o/bench_default.json"))
    print("bench", d["config"]["workload"], round(d["ms_per_step"],4), d.get("kernels_ms"), "frac", round(d["roofline_extract"]["frac"],4), "e2e", d.get("e2e",{}).get("ms_per_step"), "cold", d.get("cold_ms"))
    for k in ("strong_2048","strong_1024"): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in d.get(k,{}).items() if a in ("ms_per_step","n1_ms_per_step","speedup","error")}, d.get(k,{}).get("roofline_extract",{}).get("frac"))
except Exception as e: print("bench FAILED", e); print(open("$o/bench_default.err").read()[-2000:])
PY
fi
if [ -n "$GOLDEN" ]; then wait $gpid; tail -8 $o/golden.log; fi
