#!/bin/bash
# Round 2, call 23 (1 GPU): compaction early exits + 3-points-per-thread last block: GPU suite, c3 bench, phase counters.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2c23_pytest.log 2>&1
tail -n 3 gpurun_out/r2c23_pytest.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r2c23_c3.json 2> gpurun_out/r2c23_c3.log
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c23_c3.json").read().strip().splitlines()[-1])
print("c3", "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "rest", round(d["compact_verify_ms_per_step"]), "value", round(d["value"]),
      "e2e", d["e2e"] and round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],4), "clocks", d["clocks"]["sm_mhz"])
PY
B2M_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c23_prof.json 2> gpurun_out/r2c23_prof.log
grep "score" gpurun_out/r2c23_prof.log | tail -n 6
