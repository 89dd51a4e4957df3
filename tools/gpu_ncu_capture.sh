# ncu evidence for profiles/: (1) launch list of the bench command, (2) one --set full capture of the dominant kernel on the
# bench's own workload (cfg5), (3) the same on a small sample with source-level counters for tools/ncu_summary.py
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_cfg5.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_cfg5.log 2>&1
timeout 1200 ncu --set full --clock-control none -k regex:di_items_smem -c 1 -f -o gpurun_out/di_cfg5_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-covariance > gpurun_out/ncu_full_cfg5.log 2>&1; tail -1 gpurun_out/ncu_full_cfg5.log | cut -c1-200
