set -x
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_v4.log 2>&1; tail -5 gpurun_out/pytest_gpu_v4.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v32_nz4_nsep250.csv python bench.py --nz 4 --nsep 250 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_v32.log 2>&1
python bench.py > gpurun_out/bench_v32_full.log 2>&1; tail -1 gpurun_out/bench_v32_full.log | cut -c1-300
