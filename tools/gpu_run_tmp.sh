python -m pytest tests -m gpu -q -x -k "rules_follow or quadrature" 2>&1 | tail -3
