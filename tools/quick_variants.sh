for lib in sg_pr_b200/libsgpr_b200.so tools/variants/lib_gw2.so tools/variants/lib_gw5.so; do
  export SGPR_B200_LIB=$PWD/$lib; echo "== $lib"
  timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider 2>&1 | tail -1
  timeout 120 python bench.py --steps 400 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done
