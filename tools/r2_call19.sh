#!/bin/bash
# Round 2, call 19 (1 GPU): guided sweep with balanced pipes + host-side result cache: GPU suite, c5 (2000 images), c3 with e2e.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2c19_pytest.log 2>&1
tail -n 3 gpurun_out/r2c19_pytest.log
show() {
  python - <<PY
import json
d=json.loads(open("gpurun_out/$1.json").read().strip().splitlines()[-1])
print("$1", "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "rest", round(d["compact_verify_ms_per_step"]), "value", round(d["value"]),
      "e2e", d["e2e"] and (round(d["e2e"]["value"]), d["e2e"].get("wall_ms_last_step_rank0")), "clocks", d["clocks"]["sm_mhz"])
PY
}
timeout 900 python bench.py --config c5 --images 2000 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2c19_c5_2000.json 2> gpurun_out/r2c19_c5_2000.log; show r2c19_c5_2000
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r2c19_c3.json 2> gpurun_out/r2c19_c3.log; show r2c19_c3
B2M_HOST_CACHE=0 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r2c19_c3_nocache.json 2> gpurun_out/r2c19_c3_nocache.log; show r2c19_c3_nocache
