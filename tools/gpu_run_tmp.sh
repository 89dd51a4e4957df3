python tools/quad_sweep.py 2>&1 | tail -1
python -m pytest tests -m gpu -q -x -k "parity" 2>&1 | tail -2
