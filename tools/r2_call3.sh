#!/bin/bash
# Round 2, GPU call 3: the whole GPU suite on the single host layer, smoke, the rewritten bench (c3, N=1).
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2c3_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r2c3_pytest_gpu.log
tail -n 60 gpurun_out/r2c3_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c3_smoke.log 2>&1; tail -n 3 gpurun_out/r2c3_smoke.log
python bench.py --steps 2 --warmup 3 > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.log
tail -n 5 gpurun_out/r2c3_bench.log; cat gpurun_out/r2c3_bench.json
