set -x
timeout 300 python scripts/smoke_all_variants.py 2>&1 | tail -12
