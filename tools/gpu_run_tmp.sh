python -m pytest tests -m gpu -q -x -k "cfg2 or corrfunc or dropin or coffe_class" 2>&1 | tail -3
