#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2c25_pytest.log 2>&1
tail -n 8 gpurun_out/r2c25_pytest.log
