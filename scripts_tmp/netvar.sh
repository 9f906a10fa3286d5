for lib in libpycwr_b200 libv_nopf; do
  PYCWR_B200_LIB=$PWD/pycwr_b200/$lib.so python bench.py --steps 4 --warmup 3 --no-cpu --network-scale 1.0 2>gpurun_out/err_$lib.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', d['ms_per_step'], d['network']['ms_per_step'])"
  tail -3 gpurun_out/err_$lib.log
done
python -m pytest tests/test_gpu_network.py -x -q 2>&1 | tail -2
