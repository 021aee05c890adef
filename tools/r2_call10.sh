#!/bin/bash
# Round 2, GPU call 10 (1 GPU): guided matching after the lazy geometric test, configs c5 and c4 on one GPU.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2c10_pytest.log 2>&1; tail -n 4 gpurun_out/r2c10_pytest.log
for C in c5 c4; do
timeout 1200 python bench.py --config $C --steps 1 --warmup 1 --no-cpu > gpurun_out/r2c10_${C}_n1.json 2> gpurun_out/r2c10_${C}_n1.log
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c10_${C}_n1.json").read().strip().splitlines()[-1])
    print("$C N=1", "value", round(d["value"]), "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "verify", round(d["compact_verify_ms_per_step"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "pairs", d["config"]["pairs_per_step"], "verified", round(d["config"]["verified_pairs_fraction"],4))
except Exception as e:
    print("$C failed", e)
PY
tail -n 2 gpurun_out/r2c10_${C}_n1.log
done
