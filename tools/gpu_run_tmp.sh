for i in 1 2 3; do python bench.py --tables own --no-cpu-baseline --no-covariance --no-extras 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['setup'])"; done
python -m pytest tests -m gpu -q -x -k "capi or multigpu or plan" 2>&1 | tail -2
