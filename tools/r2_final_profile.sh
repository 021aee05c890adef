#!/bin/bash
# Round 2, final evidence call (1 GPU): launch list of a full c3 step, ncu --set full of the K1 GEMM launches (row
# direction + gathered column direction), the gather / resolve kernels and the three RANSAC kernels, compute-sanitizer
# passes, host-side profile of one step.  Everything lands in gpurun_out/; the summaries are copied to profiles/ afterwards.
set -u
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
# 1. launch list of one full step at 1000 x 8192 (my kernels only; the scene generator's torch kernels are filtered out)
$NCU --metrics gpu__time_duration.sum --clock-control none -k regex:b2m_ -c 6000 --csv --log-file gpurun_out/r2f_launches_c3_1000x8192.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2f_ncu_launches.log 2>&1
# 2. full captures: K1 GEMM (2nd batch: launches 3 and 4 of the kernel = row direction + gathered direction)
$NCU --set full --clock-control none --import-source on -k regex:b2m_k1_filter_kernel -s 2 -c 2 -o gpurun_out/r2f_k1_gather \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2f_ncu_k1.log 2>&1
# 3. full captures: RANSAC kernels (E, F, H of the 2nd batch), gather + resolve
$NCU --set full --clock-control none --import-source on -k regex:b2m_ransac_kernel -s 3 -c 3 -o gpurun_out/r2f_ransac \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2f_ncu_ransac.log 2>&1
$NCU --set full --clock-control none -k regex:"b2m_k1_gather_kernel|b2m_k1_resolve_kernel" -s 3 -c 3 -o gpurun_out/r2f_gather_resolve \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2f_ncu_gather.log 2>&1
# 4. host-side split of one step
B2M_HOSTPROF=1 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2f_hostprof.json 2> gpurun_out/r2f_hostprof.log
grep hostprof gpurun_out/r2f_hostprof.log | tail -n 4
# 5. sanitizer
tools/sanitize.sh
ls -la gpurun_out | tail -n 20
