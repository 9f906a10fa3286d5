for lib in libpycwr_b200 libv_b5 libv_b6; do
  PYCWR_B200_LIB=$PWD/pycwr_b200/$lib.so python bench.py --steps 20 --warmup 3 --no-cpu --no-network 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', d['ms_per_step'])"
done
