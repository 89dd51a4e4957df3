for v in default t448 t512; do
  if [ $v = default ]; then unset COFFE_B200_LIB; else export COFFE_B200_LIB=$PWD/build_variants/lib_$v.so; fi
  python bench.py --nz 4 --nsep 250 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_v33_$v.log 2>&1; echo $v; tail -1 gpurun_out/bench_v33_$v.log | cut -c1-140
done
