python bench.py --no-cpu-baseline > gpurun_out/bench_v35b_full.log 2>&1; tail -1 gpurun_out/bench_v35b_full.log | cut -c1-200
python bench.py --no-cpu-baseline > gpurun_out/bench_v35c_full.log 2>&1; tail -1 gpurun_out/bench_v35c_full.log | cut -c1-200
nproc; uptime
