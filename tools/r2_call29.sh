#!/bin/bash
# final tree: smoke() + the GPU suite exactly as the driver runs it
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -n 2
