set -x
for v in default t384 t448; do
  if [ $v = default ]; then unset COFFE_B200_LIB; else export COFFE_B200_LIB=$PWD/build_variants/lib_$v.so; fi
  python bench.py --nz 4 --nsep 250 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_v31_$v.log 2>&1; tail -1 gpurun_out/bench_v31_$v.log | cut -c1-200
done
unset COFFE_B200_LIB
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_v31.log 2>&1; tail -3 gpurun_out/pytest_gpu_v31.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:di_items_smem -c 1 -f -o gpurun_out/items_v31_full python bench.py --nz 2 --nsep 120 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_v31.log 2>&1; tail -1 gpurun_out/ncu_full_v31.log
