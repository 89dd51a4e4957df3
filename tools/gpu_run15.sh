(time python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "average") 2>&1 | tail -15
