#!/bin/bash
# Round 2, GPU call 2: mirror-vs-native diagnosis, initcheck, pair-batch A/B (verify occupancy), launch list of my kernels.
set -u
mkdir -p gpurun_out
python tools/diag_mirror_vs_native.py > gpurun_out/r2_diag_mirror.log 2>&1; tail -n 12 gpurun_out/r2_diag_mirror.log
S=/usr/local/cuda/bin/compute-sanitizer
DIAG_N=120 DIAG_REPS=1 timeout 900 $S --tool initcheck python tools/diag_determinism.py > gpurun_out/r2_san_initcheck.log 2>&1
echo "initcheck exit $?" | tee -a gpurun_out/r2_san_initcheck.log
tail -n 6 gpurun_out/r2_san_initcheck.log
for PB in 1024 2048 4096 8192; do
  B2M_PROF=1 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --pair-batch $PB > gpurun_out/r2_pb_$PB.json 2> gpurun_out/r2_pb_$PB.log
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_pb_$PB.json"))
print("pair_batch $PB", "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "verify", round(d["compact_verify_ms_per_step"]), "value", round(d["value"]))
PY
  grep "b2m prof" gpurun_out/r2_pb_$PB.log | head -24
done
NCU=/usr/local/cuda/bin/ncu
for PB in 1024 4096; do
$NCU --metrics gpu__time_duration.sum --clock-control none -k regex:b2m_ -c 4000 --csv --log-file gpurun_out/r2_launches_pb${PB}_300img.csv \
    python bench.py --images 300 --feats 8192 --steps 1 --warmup 1 --no-cpu --no-e2e --pair-batch $PB > gpurun_out/r2_ncu_launches_$PB.log 2>&1
done
ls -la gpurun_out | tail -n 14
