# ncu evidence for profiles/: (1) launch list of the bench command, (2) --set full captures of the dominant kernel and of the second
# one on the bench's own workload (cfg5) with source-level counters, (3) the covariance kernels of cfg4; summarised in the container by
# tools/ncu_summary.py (JSON read by bench.py) and tools/ncu_source_summary.py (opcode mix, shared-memory wavefronts per line)
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_cfg5.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_cfg5.log 2>&1
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:di_items_smem -c 1 -f -o gpurun_out/di_cfg5_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-covariance --no-extras > gpurun_out/ncu_full_cfg5.log 2>&1; tail -1 gpurun_out/ncu_full_cfg5.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:si_items_smem -c 1 -f -o gpurun_out/si_cfg5_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-covariance --no-extras > gpurun_out/ncu_full_si_cfg5.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:cov_ -c 6 -f -o gpurun_out/cov_cfg4_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_cov.log 2>&1
ls -la gpurun_out/*.ncu-rep
