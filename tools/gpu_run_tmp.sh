python tools/piece_sweep.py 2>&1 | tail -4
(time python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu_latest.log 2>&1; tail -6 gpurun_out/pytest_gpu_latest.log
python bench.py > gpurun_out/bench_latest_full.log 2>&1; tail -1 gpurun_out/bench_latest_full.log | cut -c1-300
