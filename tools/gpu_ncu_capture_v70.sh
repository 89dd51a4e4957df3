# ncu evidence of the final tree of round 2: launch list of the bench command, --set full captures of the two integrated-piece kernels
# on the bench's own workload (cfg5), and compute-sanitizer memcheck over the smoke invocation
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_cfg5.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_cfg5.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:di_items_smem -c 1 -f -o gpurun_out/di_cfg5_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-covariance --no-extras > gpurun_out/ncu_full_cfg5.log 2>&1; tail -1 gpurun_out/ncu_full_cfg5.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:si_items_smem -c 1 -f -o gpurun_out/si_cfg5_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-covariance --no-extras > gpurun_out/ncu_full_si_cfg5.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_v70.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck_v70.log
ls -la gpurun_out/*.ncu-rep
