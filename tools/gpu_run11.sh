set -x
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_v34.log 2>&1; tail -5 gpurun_out/pytest_gpu_v34.log
python bench.py --no-cpu-baseline > gpurun_out/bench_v34_full.log 2>&1; tail -1 gpurun_out/bench_v34_full.log | cut -c1-300
