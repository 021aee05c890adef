#!/bin/bash
# Round 2, call 18 (1 GPU): final validation -- GPU suite, the default bench line (c3: value, e2e, cpu_baseline), pair-batch
# A/B, c5 and c4 at full size on one GPU, ncu --set full of the guided kernel.
set -u
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2c18_pytest.log 2>&1
tail -n 4 gpurun_out/r2c18_pytest.log
show() {
  python - <<PY
import json
d=json.loads(open("gpurun_out/$1.json").read().strip().splitlines()[-1])
print("$1", "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "rest", round(d["compact_verify_ms_per_step"]), "value", round(d["value"]),
      "e2e", d["e2e"] and (round(d["e2e"]["value"]), d["e2e"].get("wall_ms_last_step_rank0")), "frac", round(d["roofline"]["frac"],4), "whole", round(d["roofline"]["whole_step_frac"],4),
      "clocks", d["clocks"]["sm_mhz"], "cpu", d.get("cpu_baseline") and (d["cpu_baseline"]["value"], d["cpu_baseline"].get("gpu_verification_agrees_on_sample")))
PY
}
timeout 900 python bench.py > gpurun_out/r2c18_c3_default.json 2> gpurun_out/r2c18_c3_default.log; show r2c18_c3_default
for PB in 2048 8192; do
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --pair-batch $PB > gpurun_out/r2c18_c3_pb$PB.json 2> gpurun_out/r2c18_c3_pb$PB.log; show r2c18_c3_pb$PB
done
timeout 900 python bench.py --config c5 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2c18_c5_n1.json 2> gpurun_out/r2c18_c5_n1.log; show r2c18_c5_n1
timeout 900 python bench.py --config c4 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c18_c4_n1.json 2> gpurun_out/r2c18_c4_n1.log; show r2c18_c4_n1
# guided kernel: launches 3-5 of the kernel = row direction + gathered direction of the 2nd batch
$NCU --set full --clock-control none --import-source on -k regex:b2m_k1_guided_kernel -s 2 -c 2 -o gpurun_out/r2c18_guided \
    python bench.py --config c5 --images 2000 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2c18_ncu_guided.log 2>&1
$NCU --metrics gpu__time_duration.sum --clock-control none -k regex:b2m_ -c 3000 --csv --log-file gpurun_out/r2c18_launches_c5_2000.csv \
    python bench.py --config c5 --images 2000 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2c18_ncu_launches.log 2>&1
ls -la gpurun_out | grep r2c18 | awk '{print $5, $9}'
