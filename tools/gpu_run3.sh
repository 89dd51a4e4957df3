set -x
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_v3.log 2>&1; tail -15 gpurun_out/pytest_gpu_v3.log
python bench.py --nz 4 --nsep 250 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_v3_small.log 2>&1; tail -1 gpurun_out/bench_v3_small.log | cut -c1-300
python bench.py --no-cpu-baseline > gpurun_out/bench_v3_full.log 2>&1; tail -1 gpurun_out/bench_v3_full.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:di_items_smem -c 1 -f -o gpurun_out/items_v3_full python bench.py --nz 2 --nsep 120 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_v3.log 2>&1; tail -3 gpurun_out/ncu_full_v3.log
