#!/bin/bash
# Round 2, GPU call 9: exact pruning in the scoring sweep + early exit of the inverse iteration: GPU suite, counters, bench.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2c9_pytest.log 2>&1; tail -n 6 gpurun_out/r2c9_pytest.log
B2M_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c9_prof.json 2> gpurun_out/r2c9_prof.log
grep "b2m prof" gpurun_out/r2c9_prof.log | grep -E "solve|score|accum"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.log
python - <<PY
import json
for f in ("r2c9_prof","r2c9_bench"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "verify", round(d["compact_verify_ms_per_step"]), "value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "cpu", d["cpu_baseline"] and (round(d["cpu_baseline"]["value"],1), d["cpu_baseline"].get("gpu_verification_agrees_on_sample")))
PY
timeout 600 python tools/bench_db.py --images 50 --feats 4096 --out gpurun_out/r2c9_db_50x4096.json | tail -n 1
timeout 900 python tools/bench_db.py --images 400 --feats 8192 --out gpurun_out/r2c9_db_400x8192.json | tail -n 1
