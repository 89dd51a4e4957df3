python -m pytest tests -m gpu -q -x -k "fftlog or covariance or cov" 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); c=d['covariance']; print(d['value'], c['value'], c['wall_ms'], c['kernel_ms'], c['fftlog'])"
