#!/bin/bash
# Round 2, call 20 (1 GPU): warp-cooperative eigen-solve of the LO refits: GPU suite, phase counters and bench A/B.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2c20_pytest.log 2>&1
tail -n 3 gpurun_out/r2c20_pytest.log
for MODE in warp thread; do
  if [ "$MODE" = "thread" ]; then export B2M_LO_EIG=thread; else unset B2M_LO_EIG; fi
  B2M_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c20_prof_$MODE.json 2> gpurun_out/r2c20_prof_$MODE.log
  echo "== $MODE"; grep "lo_solve\|kind . solve" gpurun_out/r2c20_prof_$MODE.log | tail -n 6
  timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2c20_$MODE.json 2> gpurun_out/r2c20_$MODE.log
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2c20_$MODE.json").read().strip().splitlines()[-1])
print("$MODE", "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "rest", round(d["compact_verify_ms_per_step"]), "value", round(d["value"]), "clocks", d["clocks"]["sm_mhz"])
PY
done
