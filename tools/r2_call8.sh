#!/bin/bash
# Round 2, GPU call 8: thread vs hybrid minimal E solves after reverting the ladder / packed-scoring experiments.
set -u
mkdir -p gpurun_out
for M in thread hybrid; do
B2M_E5_MINIMAL=$M B2M_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c8_$M.json 2> gpurun_out/r2c8_$M.log
echo "== $M"; grep "b2m prof" gpurun_out/r2c8_$M.log | grep -E "solve|score" | grep -v lo_score
python - <<PY
import json
d=json.load(open("gpurun_out/r2c8_$M.json"))
print("$M", "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "verify", round(d["compact_verify_ms_per_step"]), "value", round(d["value"]))
PY
done
