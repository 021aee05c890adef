#!/bin/bash
# Round 2, call 26 (1 GPU): DB-inclusive match_exhaustive(database_path) with the final code (upstream's pragmas, host cache),
# on /dev/shm and on the box's file system.
set -u
mkdir -p gpurun_out
timeout 600 python tools/bench_db.py --images 50 --feats 4096 --out gpurun_out/r2c26_db_50x4096.json | tail -n 1 | cut -c1-600
timeout 900 python tools/bench_db.py --images 400 --feats 8192 --out gpurun_out/r2c26_db_400x8192.json | tail -n 1 | cut -c1-700
timeout 900 python tools/bench_db.py --images 400 --feats 8192 --dir /tmp --out gpurun_out/r2c26_db_400x8192_disk.json | tail -n 1 | cut -c1-700
