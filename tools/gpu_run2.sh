set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_v2b.log 2>&1; tail -5 gpurun_out/pytest_gpu_v2b.log
python bench.py > gpurun_out/bench_v2_full.log 2>&1; tail -1 gpurun_out/bench_v2_full.log | cut -c1-600
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v2_nz4_nsep250.csv python bench.py --nz 4 --nsep 250 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:items_smem -c 1 -f -o gpurun_out/items_v2_full python bench.py --nz 2 --nsep 120 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
