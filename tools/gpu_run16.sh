(time python -m pytest tests/test_gpu_parity.py -m gpu -q -k "across_the_grid or density_lensing_cross") 2>&1 | tail -25
