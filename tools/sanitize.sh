#!/bin/bash
# compute-sanitizer passes over small shapes.  Run on a GPU box:
#   gpurun --timeout 1500 -- tools/sanitize.sh
# memcheck covers the paths written after round 1's last GPU session first (split cross-check schedule,
# undistortion kernel, batched estimator); racecheck / synccheck take the K1 GEMM kernel on tiny inputs.
set -u
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
T="timeout 600"
$T $S --tool memcheck --error-exitcode 1 python -m pytest -q -x -m gpu tests/test_zz_native_gpu.py \
    -k "column_direction_skip or distortion_models_two_view_geometry or batched_two_view" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit $?" | tee -a gpurun_out/sanitize_memcheck.log
$T $S --tool racecheck --error-exitcode 1 python -m pytest -q -x -m gpu tests/test_match_gpu.py -k "identity_and_reverse or empty_inputs" \
    > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck exit $?" | tee -a gpurun_out/sanitize_racecheck.log
$T $S --tool synccheck --error-exitcode 1 python -m pytest -q -x -m gpu tests/test_match_gpu.py -k "identity_and_reverse" \
    > gpurun_out/sanitize_synccheck.log 2>&1
echo "synccheck exit $?" | tee -a gpurun_out/sanitize_synccheck.log
tail -n 5 gpurun_out/sanitize_*.log
