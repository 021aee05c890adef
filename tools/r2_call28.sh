#!/bin/bash
# Round 2, call 28 (1 GPU): batch size scaled with the image size: GPU suite, c4, c5 (2000 images), c3.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2c28_pytest.log 2>&1
tail -n 3 gpurun_out/r2c28_pytest.log
show() {
  python - <<PY
import json
d=json.loads(open("gpurun_out/$1.json").read().strip().splitlines()[-1])
print("$1", "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "rest", round(d["compact_verify_ms_per_step"]), "value", round(d["value"]), "launches", d["gpu_launches"], "clocks", d["clocks"]["sm_mhz"])
PY
}
timeout 900 python bench.py --config c4 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c28_c4_n1.json 2> gpurun_out/r2c28_c4_n1.log; show r2c28_c4_n1
timeout 900 python bench.py --config c5 --images 2000 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c28_c5_2000.json 2> gpurun_out/r2c28_c5_2000.log; show r2c28_c5_2000
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c28_c3.json 2> gpurun_out/r2c28_c3.log; show r2c28_c3
