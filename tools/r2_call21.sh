#!/bin/bash
# Round 2, call 21 (1 GPU): sustained int8 tcgen05 peak (the roofline denominator for a kernel timed inside a long step).
set -u
mkdir -p gpurun_out
(nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 250 > gpurun_out/r2c21_smi.log 2>&1 &) 
timeout 120 ./tools/microbench > gpurun_out/r2c21_microbench.json 2> gpurun_out/r2c21_microbench.err
cat gpurun_out/r2c21_microbench.json
sort gpurun_out/r2c21_smi.log | uniq -c | sort -rn | head -5
