#!/bin/bash
# c5 at K = 8192 ("K=4096 and 8192 both reported"), one GPU, one timed step
set -u
mkdir -p gpurun_out
timeout 150 python bench.py --config c5 --feats 8192 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c30_c5_k8192_n1.json 2> gpurun_out/r2c30_c5_k8192_n1.log
tail -n 2 gpurun_out/r2c30_c5_k8192_n1.log
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c30_c5_k8192_n1.json").read().strip().splitlines()[-1])
    print("c5 K=8192", "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "rest", round(d["compact_verify_ms_per_step"]), "value", round(d["value"]), "pairs", d["config"]["pairs_per_step"])
except Exception as e:
    print("no result:", e)
PY
