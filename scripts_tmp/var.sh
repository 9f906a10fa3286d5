python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for lib in libpycwr_b200 libv_b4; do
  PYCWR_B200_LIB=$PWD/pycwr_b200/$lib.so python bench.py --steps 20 --warmup 3 --no-cpu --no-network 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', d['ms_per_step'])"
done
ncu --set full --clock-control none --import-source on -k regex:cappi3d_r2 -c 1 -s 3 -o gpurun_out/prof_r01_v7f -f python bench.py --steps 2 --warmup 3 --no-cpu --no-network > gpurun_out/ncu_v7f.log 2>&1
