#!/bin/bash
# Round 2, call 22 (1 GPU): final state -- smoke(), GPU suite, the default bench line, c5 and c4 on one GPU, sanitizer passes.
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c22_smoke.log 2>&1; tail -n 2 gpurun_out/r2c22_smoke.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2c22_pytest.log 2>&1
tail -n 3 gpurun_out/r2c22_pytest.log
show() {
  python - <<PY
import json
d=json.loads(open("gpurun_out/$1.json").read().strip().splitlines()[-1])
print("$1", "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "rest", round(d["compact_verify_ms_per_step"]), "value", round(d["value"]),
      "e2e", d["e2e"] and (round(d["e2e"]["value"]), d["e2e"].get("wall_ms_last_step_rank0")), "frac", round(d["roofline"]["frac"],4), "frac_burst", round(d["roofline"].get("frac_burst",0),4),
      "clocks", d["clocks"]["sm_mhz"], "cpu", d.get("cpu_baseline") and (round(d["cpu_baseline"]["value"],1), d["cpu_baseline"].get("gpu_verification_agrees_on_sample")))
PY
}
timeout 900 python bench.py > gpurun_out/r2c22_c3_default.json 2> gpurun_out/r2c22_c3_default.log; show r2c22_c3_default
timeout 900 python bench.py --config c5 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2c22_c5_n1.json 2> gpurun_out/r2c22_c5_n1.log; show r2c22_c5_n1
timeout 900 python bench.py --config c4 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2c22_c4_n1.json 2> gpurun_out/r2c22_c4_n1.log; show r2c22_c4_n1
tools/sanitize.sh > gpurun_out/r2c22_sanitize.log 2>&1; grep "exit" gpurun_out/r2c22_sanitize.log
