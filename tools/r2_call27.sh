#!/bin/bash
# Round 2, call 27 (1 GPU): ncu --set full of the FINAL code's RANSAC kernels (heavy batch) and guided kernel.
set -u
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
B2M_NO_OVERLAP=1 $NCU --set full --clock-control none --import-source on -k regex:b2m_ransac_kernel -s 18 -c 3 -o gpurun_out/r2c27_ransac_heavy \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2c27_ncu_ransac.log 2>&1
$NCU --set full --clock-control none --import-source on -k regex:b2m_k1_guided_kernel -s 2 -c 2 -o gpurun_out/r2c27_guided \
    python bench.py --config c5 --images 2000 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2c27_ncu_guided.log 2>&1
ls -la gpurun_out | grep r2c27 | awk '{print $5, $9}'
