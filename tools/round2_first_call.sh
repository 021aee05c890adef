#!/bin/bash
# The first GPU call of the next round:  gpurun --timeout 2400 -- tools/round2_first_call.sh
# 1. the whole GPU suite (validated files first, tests/test_zz_native_gpu.py last), WITHOUT -x so that one failure
#    does not hide the rest;  2. bench with the split K1 schedule and, for the A/B, with the two-direction launch;
# 3. launch list + one ncu --set full capture of the GEMM kernel under the split schedule.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r2_pytest_gpu.log
tail -n 15 gpurun_out/r2_pytest_gpu.log
python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_split.json 2> gpurun_out/r2_bench_split.log
B2M_K1_DIR1=full python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.log
tail -n 2 gpurun_out/r2_bench_split.json gpurun_out/r2_bench_full.json
NCU=/usr/local/cuda/bin/ncu
B2M_K1_DIR1=skip $NCU --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_split_300img.csv \
    python bench.py --images 300 --feats 8192 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_ncu_launches.log 2>&1
B2M_K1_DIR1=skip $NCU --set full --clock-control none --import-source on -k regex:b2m_k1_filter_kernel -s 4 -c 2 \
    -o gpurun_out/r2_k1_split python bench.py --images 300 --feats 8192 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out | tail -n 12
