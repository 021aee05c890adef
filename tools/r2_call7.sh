#!/bin/bash
# Round 2, GPU call 7: A/B of the E minimal mapping (thread / hybrid) x hypothesis scoring (scalar / packed fp32x2).
set -u
mkdir -p gpurun_out
for M in thread hybrid; do for SC in scalar packed; do
B2M_E5_MINIMAL=$M B2M_SCORE=$SC B2M_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c7_${M}_$SC.json 2> gpurun_out/r2c7_${M}_$SC.log
echo "== $M $SC"; grep "b2m prof" gpurun_out/r2c7_${M}_$SC.log | grep -E "solve|score" | grep -v lo_score
python - <<PY
import json
d=json.load(open("gpurun_out/r2c7_${M}_$SC.json"))
print("$M $SC", "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "verify", round(d["compact_verify_ms_per_step"]), "value", round(d["value"]), "models", d["roofline_verify"]["models_scored"])
PY
done; done
B2M_E5_MINIMAL=hybrid timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2c7_pytest_hybrid.log 2>&1; tail -n 4 gpurun_out/r2c7_pytest_hybrid.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2c7_pytest.log 2>&1; tail -n 4 gpurun_out/r2c7_pytest.log
