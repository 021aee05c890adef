#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -k "guided or golden or pipeline" > gpurun_out/r2c12_pytest.log 2>&1; tail -n 3 gpurun_out/r2c12_pytest.log
timeout 1200 python bench.py --config c5 --images 2000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c12_c5_2000.json 2> gpurun_out/r2c12_c5_2000.log
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c12_c5_2000.json").read().strip().splitlines()[-1])
print("c5/2000 N=1", "value", round(d["value"]), "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "verify", round(d["compact_verify_ms_per_step"]), "pairs", d["config"]["pairs_per_step"])
PY
