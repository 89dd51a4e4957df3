set -x
python bench.py --nz 4 --nsep 250 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_v32.log 2>&1; tail -1 gpurun_out/bench_v32.log | cut -c1-200
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_v32.log 2>&1; tail -3 gpurun_out/pytest_gpu_v32.log
