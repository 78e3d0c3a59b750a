set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "max rel diff\|listed" | tail -12
