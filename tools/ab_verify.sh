# A/B harness with verification on: args = variant names under pycolmap_b200/variants/
for v in "$@"; do
  export B2M_LIB=$PWD/pycolmap_b200/variants/$v.so
  timeout 300 python bench.py --images 300 --feats 8192 --steps 2 --warmup 1 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', {k:round(d[k],1) for k in ['value','ms_per_step','k1_ms_per_step','compact_verify_ms_per_step']}, d['clocks']['sm_mhz'])"
done
