for lib in libv_nb3 libv_nb4; do
  PYCWR_B200_LIB=$PWD/pycwr_b200/$lib.so python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', d['network']['ms_per_step'])"
done
ncu --set full --clock-control none --import-source on -k regex:network_kernel -c 1 -s 3 -o gpurun_out/prof_r01_net_v3 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_net3.log 2>&1
