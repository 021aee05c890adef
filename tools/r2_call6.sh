#!/bin/bash
# Round 2, GPU call 6: A/B of the minimal 5-point mapping (thread vs warp per hypothesis), both with the Fujiwara bound.
set -u
mkdir -p gpurun_out
for M in thread warp; do
B2M_E5_MINIMAL=$M B2M_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c6_$M.json 2> gpurun_out/r2c6_$M.log
echo "== $M"; grep "b2m prof] kind 0" gpurun_out/r2c6_$M.log | head -7
python - <<PY
import json
d=json.load(open("gpurun_out/r2c6_$M.json"))
print("$M", "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "verify", round(d["compact_verify_ms_per_step"]), "value", round(d["value"]))
PY
done
timeout 600 python -m pytest tests/test_zz_native_gpu.py tests/test_opencv_crosscheck.py -q -m gpu -p no:cacheprovider -k "hypothesis or five_point or opencv" 2>&1 | tail -5
