(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_latest.log 2>&1; tail -4 gpurun_out/pytest_gpu_latest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_latest_full.log 2>&1; tail -1 gpurun_out/bench_latest_full.log | cut -c1-200
