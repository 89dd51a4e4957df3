python -m pytest tests -m gpu -q -x -k "rules_follow" 2>&1 | tail -5
