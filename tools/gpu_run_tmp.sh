python -m pytest tests -m gpu -q -x -k "rules_follow or cfg3 or fixture or two_pop or class" 2>&1 | tail -3
