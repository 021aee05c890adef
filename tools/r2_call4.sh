#!/bin/bash
# Round 2, GPU call 4: GPU suite with the warp-cooperative 5-point solver, verify phase counters, bench.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2c4_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r2c4_pytest_gpu.log
tail -n 40 gpurun_out/r2c4_pytest_gpu.log
B2M_PROF=1 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c4_prof.json 2> gpurun_out/r2c4_prof.log
grep "b2m prof" gpurun_out/r2c4_prof.log | head -24
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.log
python - <<PY
import json
for f in ("gpurun_out/r2c4_prof.json","gpurun_out/r2c4_bench.json"):
    d=json.load(open(f))
    print(f, "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "verify", round(d["compact_verify_ms_per_step"]), "value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "verify frac", d["roofline_verify"]["frac"])
PY
