python -m pytest tests -m gpu -q -x -k "covariance or cov" 2>&1 | tail -2
python -m pytest tests/test_multigpu.py tests/test_coffe_class.py tests/test_dropin.py -m gpu -q 2>&1 | tail -2
