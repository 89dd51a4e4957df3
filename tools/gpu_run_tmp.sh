python bench.py --no-cpu-baseline --no-covariance 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['extras']['cfg1_flatsky_multipoles'])"
