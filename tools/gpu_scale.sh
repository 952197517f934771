#!/bin/bash
# usage: gpu_scale.sh N  -- strong scaling (spheres2048) and the default weak-scaling line on N GPUs of one box
N=$1; o=gpurun_out/final; mkdir -p $o
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 --workload spheres2048 > $o/scale_strong_n$N.json 2> $o/scale_strong_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 10 --warmup 3 > $o/scale_weak_n$N.json 2> $o/scale_weak_n$N.err
python - <<PY
import json
for f in ("scale_strong_n$N","scale_weak_n$N"):
    try:
        d=json.loads([l for l in open("gpurun_out/final/%s.json"%f).read().splitlines() if l.startswith("{")][-1]); print(f, d["config"]["workload"], round(d["ms_per_step"],4), round(d["value"],1), d["roofline_extract"]["frac"], d["kernels_ms"])
    except Exception as e: print(f,"FAILED",e)
PY
