#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus N): N-rank bench of config c3 (+ c4 / c5 when asked), the 2-GPU database test.
#   tools/r2_multigpu.sh <N> [configs...]        e.g.  tools/r2_multigpu.sh 8 c3 c4 c5
set -u
N=${1:-2}; shift || true
CONFIGS=${@:-c3}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -n 8
if [ "$N" -ge 2 ]; then
  timeout 600 python -m pytest tests/test_zz_native_gpu.py -q -m gpu -p no:cacheprovider -k "multi_gpu or sharded" > gpurun_out/r2mg_pytest_n$N.log 2>&1
  tail -n 5 gpurun_out/r2mg_pytest_n$N.log
fi
for C in $CONFIGS; do
  STEPS=2; WARM=1
  EXTRA=""
  if [ "$C" = "c3" ]; then STEPS=3; WARM=2; fi
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps $STEPS --warmup $WARM --config $C --no-cpu $EXTRA > gpurun_out/r2mg_${C}_n$N.json 2> gpurun_out/r2mg_${C}_n$N.log
  echo "== $C N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2mg_${C}_n$N.json").read().strip().splitlines()[-1])
    print("$C N=$N", "value", round(d["value"]), "ms/step", round(d["ms_per_step"]), "k1", round(d["k1_ms_per_step"]), "verify", round(d["compact_verify_ms_per_step"]),
          "allgather", {k: (round(v,2) if isinstance(v,float) else v) for k,v in d["allgather"].items()}, "e2e", d["e2e"] and round(d["e2e"]["value"]), "pairs", d["config"]["pairs_per_step"], "verified frac", round(d["config"]["verified_pairs_fraction"],4))
except Exception as e:
    print("$C N=$N failed:", e)
PY
  tail -n 3 gpurun_out/r2mg_${C}_n$N.log
done
