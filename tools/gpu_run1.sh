set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_v2.log 2>&1; tail -5 gpurun_out/pytest_gpu_v2.log
python bench.py --nz 4 --nsep 250 --steps 2 --warmup 1 --no-cpu-baseline --variant 1 > gpurun_out/bench_v2_small.log 2>&1; tail -2 gpurun_out/bench_v2_small.log
python bench.py --nz 4 --nsep 250 --steps 2 --warmup 1 --no-cpu-baseline --variant 0 > gpurun_out/bench_v1_small.log 2>&1; tail -2 gpurun_out/bench_v1_small.log
