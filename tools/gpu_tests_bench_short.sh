set -x
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_v35.log 2>&1; tail -5 gpurun_out/pytest_gpu_v35.log
python bench.py > gpurun_out/bench_v35_full.log 2>&1; tail -1 gpurun_out/bench_v35_full.log | cut -c1-300
