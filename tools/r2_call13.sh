#!/bin/bash
# Round 2, call 13 (1 GPU): GPU suite with the overlapped schedule, A/B of the schedule on c3, ncu --set full of a HEAVY
# batch's RANSAC + resolve kernels (batch 6 of the exhaustive order holds ~1500 verifiable pairs; most batches hold none).
set -u
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2c13_pytest.log 2>&1
tail -n 4 gpurun_out/r2c13_pytest.log
for MODE in overlap sequential; do
  if [ "$MODE" = "sequential" ]; then export B2M_NO_OVERLAP=1; else unset B2M_NO_OVERLAP; fi
  timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r2c13_$MODE.json 2> gpurun_out/r2c13_$MODE.log
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2c13_$MODE.json").read().strip().splitlines()[-1])
print("$MODE", "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "verify", round(d["compact_verify_ms_per_step"]), "value", round(d["value"]),
      "e2e", d["e2e"] and round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],4), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
unset B2M_NO_OVERLAP
# heavy-batch captures (sequential order so that launch indices are the exhaustive order's: batch 6 = launches 18..20)
B2M_NO_OVERLAP=1 $NCU --set full --clock-control none --import-source on -k regex:b2m_ransac_kernel -s 18 -c 3 -o gpurun_out/r2c13_ransac_heavy \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2c13_ncu_ransac.log 2>&1
# resolve launches: self-test 3 (gather 2 + full 1), then 2 per batch -> batch 6 = launches 15, 16
B2M_NO_OVERLAP=1 $NCU --set full --clock-control none --import-source on -k regex:b2m_k1_resolve_kernel -s 15 -c 2 -o gpurun_out/r2c13_resolve_heavy \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2c13_ncu_resolve.log 2>&1
# gathered GEMM of batch 6: filter launches: self-test 3 (2 gather + 1 full), then 2 per batch -> batch 6 = 15 (row), 16 (gathered)
B2M_NO_OVERLAP=1 $NCU --set full --clock-control none -k regex:b2m_k1_filter_kernel -s 15 -c 2 -o gpurun_out/r2c13_k1_heavy \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2c13_ncu_k1.log 2>&1
ls -la gpurun_out | grep r2c13
