python tools/dense_margin.py
python -m pytest tests -m gpu -q -x -k "rules_follow or fixture or dense or cfg3 or class or dropin" 2>&1 | tail -2
python bench.py --no-cpu-baseline --no-covariance --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['max_rel_err'])"
