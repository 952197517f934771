#!/bin/bash
# compute-sanitizer passes + the randomised parity stress on one GPU box; writes gpurun_out/sanitizer.txt
mkdir -p gpurun_out
out=${1:-gpurun_out/sanitizer.txt}
mkdir -p "$(dirname $out)"
: > $out
for tool in memcheck racecheck initcheck; do
  echo "== $tool" >> $out
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_cases.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^csgA|^noise|^slab|^graph|^batch|^batch grids|^exchanged|Error|error|hazard" | head -40 >> $out
done
echo "== stress" >> $out
timeout 900 python tools/gpu_stress.py 400 7 2>&1 | tail -3 >> $out
cat $out
