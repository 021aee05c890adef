#!/bin/bash
# Round 2, call 14 (1 GPU): A/B matrix on c3 -- overlapped order (drain under GEMM0) vs sequential, two models per trip
# of the scoring loop vs one, one more resident CTA per SM for the E / F / H RANSAC kernels.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "verify or overlapped or estimat or planted or two_view or golden" > gpurun_out/r2c14_pytest.log 2>&1
tail -n 3 gpurun_out/r2c14_pytest.log
B2M_RANSAC_OCC=efh timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "verify or overlapped or estimat or planted or two_view or golden" > gpurun_out/r2c14_pytest_occ.log 2>&1
tail -n 3 gpurun_out/r2c14_pytest_occ.log
run() {
  NAME=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c14_$NAME.json 2> gpurun_out/r2c14_$NAME.log
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2c14_$NAME.json").read().strip().splitlines()[-1])
print("$NAME", "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "rest", round(d["compact_verify_ms_per_step"]), "value", round(d["value"]),
      "frac", round(d["roofline"]["frac"],4), "clocks", d["clocks"]["sm_mhz"])
PY
}
run base B2M_X=0
run sequential B2M_NO_OVERLAP=1
run score_one B2M_SCORE_ONE=1
run occ_h B2M_RANSAC_OCC=h
run occ_e B2M_RANSAC_OCC=e
run occ_f B2M_RANSAC_OCC=f
run occ_efh B2M_RANSAC_OCC=efh
run base2 B2M_X=0
