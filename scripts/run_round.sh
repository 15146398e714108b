python -m pytest tests -m gpu -x -q -k "f64 or byproducts" 2>&1 | tail -8
