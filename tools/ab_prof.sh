for v in "$@"; do
  export B2M_LIB=$PWD/pycolmap_b200/variants/$v.so
  echo "== $v"
  timeout 200 python bench.py --images 300 --feats 8192 --steps 1 --warmup 1 --no-cpu --no-e2e --verify 0 2>&1 >/dev/null | grep k1prof | tail -1
done
