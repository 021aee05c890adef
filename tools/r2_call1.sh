#!/bin/bash
# Round 2, GPU call 1: determinism diagnosis of the distorted-camera verifier, the whole GPU suite WITHOUT -x,
# compute-sanitizer passes, bench (split and two-direction schedules), launch list + ncu --set full of the split K1.
set -u
mkdir -p gpurun_out
python tools/diag_determinism.py > gpurun_out/r2_diag.log 2>&1; echo "diag exit $?" | tee -a gpurun_out/r2_diag.log
tail -n 30 gpurun_out/r2_diag.log
python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r2_pytest_gpu.log
tail -n 40 gpurun_out/r2_pytest_gpu.log
S=/usr/local/cuda/bin/compute-sanitizer
DIAG_N=120 DIAG_REPS=1 timeout 900 $S --tool initcheck --track-unused-memory no python tools/diag_determinism.py > gpurun_out/r2_san_initcheck.log 2>&1
echo "initcheck exit $?" | tee -a gpurun_out/r2_san_initcheck.log
grep -c "Uninitialized" gpurun_out/r2_san_initcheck.log
DIAG_N=120 DIAG_REPS=1 timeout 900 $S --tool racecheck python tools/diag_determinism.py > gpurun_out/r2_san_racecheck.log 2>&1
echo "racecheck exit $?" | tee -a gpurun_out/r2_san_racecheck.log
tail -n 12 gpurun_out/r2_san_initcheck.log gpurun_out/r2_san_racecheck.log
python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_split.json 2> gpurun_out/r2_bench_split.log
tail -n 2 gpurun_out/r2_bench_split.json
NCU=/usr/local/cuda/bin/ncu
B2M_K1_DIR1=skip $NCU --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_split_300img.csv \
    python bench.py --images 300 --feats 8192 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_ncu_launches.log 2>&1
B2M_K1_DIR1=skip $NCU --set full --clock-control none --import-source on -k regex:b2m_k1_filter_kernel -s 6 -c 2 \
    -o gpurun_out/r2_k1_split python bench.py --images 1000 --feats 8192 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out | tail -n 14
