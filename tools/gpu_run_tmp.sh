set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_bench_cfg5.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-covariance > gpurun_out/ncu_launch_cfg5.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:di_items_smem -c 1 -f -o gpurun_out/di_cfg5_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-covariance --no-extras > gpurun_out/ncu_full_cfg5.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:si_items_smem -c 1 -f -o gpurun_out/si_cfg5_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-covariance --no-extras > gpurun_out/ncu_full_si_cfg5.log 2>&1
python bench.py > gpurun_out/bench_latest_full.log 2>&1; tail -1 gpurun_out/bench_latest_full.log | cut -c1-200
python bench.py --tables own --no-cpu-baseline > gpurun_out/bench_own_tables.log 2>&1; tail -1 gpurun_out/bench_own_tables.log | cut -c1-200
COFFE_B200_TRACE=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-covariance --no-extras 2> gpurun_out/trace.log > /dev/null
