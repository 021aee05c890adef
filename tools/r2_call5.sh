#!/bin/bash
# Round 2, GPU call 5: gathered column direction + Fujiwara root bound: GPU suite, phase counters, bench A/B of the schedules.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2c5_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r2c5_pytest_gpu.log
tail -n 30 gpurun_out/r2c5_pytest_gpu.log
B2M_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c5_prof.json 2> gpurun_out/r2c5_prof.log
grep "b2m prof" gpurun_out/r2c5_prof.log | head -24
for MODE in gather skip; do
B2M_K1_DIR1=$MODE timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2c5_bench_$MODE.json 2> gpurun_out/r2c5_bench_$MODE.log
done
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.log
python - <<PY
import json
for f in ("r2c5_prof","r2c5_bench_gather","r2c5_bench_skip","r2c5_bench"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "verify", round(d["compact_verify_ms_per_step"]), "value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "mode", d["roofline"]["k1_dir1_mode"], "k1 frac", round(d["roofline"]["frac"],3), "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"])
    except Exception as e: print(f, "failed", e)
PY
tail -n 3 gpurun_out/r2c5_bench.log
