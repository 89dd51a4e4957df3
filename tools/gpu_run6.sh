set -x
(time python -m pytest tests/test_integrals.py tests/test_coffe_class.py -m gpu -x -q) > gpurun_out/pytest_gpu_v4a.log 2>&1; tail -30 gpurun_out/pytest_gpu_v4a.log
