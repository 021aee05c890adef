# A/B harness: bench several builds of libb200match.so (pycolmap_b200/variants/*.so, selected with B2M_LIB) on one box
for v in "$@"; do
  export B2M_LIB=$PWD/pycolmap_b200/variants/$v.so
  timeout 200 python bench.py --images 300 --feats 8192 --steps 3 --warmup 1 --no-cpu --no-e2e --verify 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', {k:round(d[k],1) for k in ['value','ms_per_step','k1_ms_per_step']}, round(d['roofline']['frac'],4), round(d['roofline']['avg_launch_ms'],3), d['clocks']['sm_mhz'])"
done
